# after a host/driver change: GPU test tier, the bench lines of all configurations (no reference arms), CLI timelines
L=${1:-chk}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_device_decode.py -x -q > gpurun_out/pytest_dd_$L.log 2>&1 || { echo "device-decode tests FAILED with the piecewise push: falling back to MD_PUSH_PARTS=1"; tail -15 gpurun_out/pytest_dd_$L.log; export MD_PUSH_PARTS=1; }
tail -2 gpurun_out/pytest_dd_$L.log
echo "A/B of the piecewise push (tools/kprof.py, one 183 MB segment):"; KPROF_ONLY=decode KPROF_REPS=6 python tools/kprof.py 2>&1 | tail -1; MD_PUSH_PARTS=1 KPROF_ONLY=decode KPROF_REPS=6 python tools/kprof.py 2>&1 | tail -1
( time timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_device_decode.py ) > gpurun_out/pytest_$L.log 2>&1; tail -4 gpurun_out/pytest_$L.log
python bench.py --no-cpu-baseline > gpurun_out/bench_c2_$L.json 2> gpurun_out/bench_c2_$L.err; tail -c 600 gpurun_out/bench_c2_$L.json; tail -2 gpurun_out/bench_c2_$L.err
for C in c3 c5 c4; do
  M="--mbp 300"; [ $C = c4 ] && M=""
  MD_TIMING=1 python bench.py --config $C $M --no-cpu-baseline > gpurun_out/bench_${C}_$L.json 2> gpurun_out/bench_${C}_$L.err; tail -c 900 gpurun_out/bench_${C}_$L.json; grep -v md-timing gpurun_out/bench_${C}_$L.err | tail -2
done
G=/tmp/mdbench/h300; B=methyldackel_b200/lib/MethylDackel
{ for i in 1 2; do echo "== extract --CHG --CHH --mergeContext (run $i)"; ( time MD_TIMING=1 $B extract --CHG --CHH --mergeContext -o /dev/shm/t_all $G.fa $G.bam ) 2>&1 | grep -E "md-timing|real"; done
  echo "== extract (CpG)"; ( time MD_TIMING=1 $B extract -o /dev/shm/t_cpg $G.fa $G.bam ) 2>&1 | grep -E "md-timing|real"
  echo "== mbias"; ( time MD_TIMING=1 $B mbias --txt $G.fa $G.bam /dev/shm/t_mb > /dev/null ) 2>&1 | grep -E "md-timing|real"; } > gpurun_out/cli_timeline_$L.txt 2>&1
rm -f /dev/shm/t_all* /dev/shm/t_cpg* /dev/shm/t_mb*
grep -E "wall|real|contig to" gpurun_out/cli_timeline_$L.txt | cut -c1-250
