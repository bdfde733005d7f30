# run a pytest selection on the GPU box: bash tools/gpu_test.sh label <pytest args...>
L=${1:-t}; shift
mkdir -p gpurun_out
( time timeout 1500 python -m pytest "$@" ) > gpurun_out/pytest_$L.log 2>&1; tail -25 gpurun_out/pytest_$L.log
