# quick kernel check: adversarial random tiles + tile ABI parity, then the device-resident bench
L=${1:-q}; shift
mkdir -p gpurun_out /tmp/mdbench
( timeout 900 python -m pytest tests/test_random_tiles.py tests/test_gpu_parity.py -m gpu -x -q -k "random or tile_abi or noisy or deep or ultra" ) > gpurun_out/pytest_$L.log 2>&1; tail -3 gpurun_out/pytest_$L.log
python tools/kbench.py --variants ${VARIANTS:-1} --steps 20 "$@" > gpurun_out/kbench_$L.jsonl 2> gpurun_out/kbench_$L.err
cat gpurun_out/kbench_$L.jsonl; tail -3 gpurun_out/kbench_$L.err
