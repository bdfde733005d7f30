# the large configurations of BASELINE.json through bench.py (driver-readable lines under gpurun_out/), GPU test tier first
L=${1:-big}
MBP=${2:-300}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_$L.log 2>&1; tail -6 gpurun_out/pytest_$L.log
for C in c3 c5; do
  python bench.py --config $C --mbp $MBP > gpurun_out/bench_${C}_$L.json 2> gpurun_out/bench_${C}_$L.err; tail -c 3500 gpurun_out/bench_${C}_$L.json; tail -2 gpurun_out/bench_${C}_$L.err
  python bench.py --config $C --mbp $MBP --impl reference --steps 2 --warmup 0 > gpurun_out/bench_${C}_ref_$L.json 2> gpurun_out/bench_${C}_ref_$L.err; tail -c 1200 gpurun_out/bench_${C}_ref_$L.json
done
python bench.py --config c4 > gpurun_out/bench_c4_$L.json 2> gpurun_out/bench_c4_$L.err; tail -c 3500 gpurun_out/bench_c4_$L.json; tail -2 gpurun_out/bench_c4_$L.err
python bench.py --config c4 --impl reference --steps 2 --warmup 0 > gpurun_out/bench_c4_ref_$L.json 2> gpurun_out/bench_c4_ref_$L.err; tail -c 1200 gpurun_out/bench_c4_ref_$L.json
MD_TIMING=1 methyldackel_b200/lib/MethylDackel extract --CHG --CHH --mergeContext -o /dev/shm/t_all /tmp/mdbench/h$MBP.fa /tmp/mdbench/h$MBP.bam 2>&1 | grep md-timing
