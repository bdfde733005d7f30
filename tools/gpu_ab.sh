# A/B of env knobs on the device-resident pipeline: bash tools/gpu_ab.sh label "ENV=a;ENV=b" [extra kbench args]
L=${1:-ab}; E=${2:-}; shift; shift
mkdir -p gpurun_out /tmp/mdbench
python tools/kbench.py --variants 1 --steps 20 --envs "$E" "$@" > gpurun_out/kbench_$L.jsonl 2> gpurun_out/kbench_$L.err
cat gpurun_out/kbench_$L.jsonl; tail -3 gpurun_out/kbench_$L.err
