# kernel iteration: parity (random tiles + tile ABI + CLI) then the A/B bench of the count-kernel generators
L=${1:-iter}
mkdir -p gpurun_out /tmp/mdbench
( time timeout 900 python -m pytest tests/test_random_tiles.py tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/pytest_$L.log 2>&1; tail -5 gpurun_out/pytest_$L.log
python tools/kbench.py --variants ${VARIANTS:-1,2} --steps 10 --panel > gpurun_out/kbench_$L.jsonl 2> gpurun_out/kbench_$L.err
cat gpurun_out/kbench_$L.jsonl; tail -3 gpurun_out/kbench_$L.err
