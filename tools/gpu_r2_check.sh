# after a host/driver change: GPU test tier, the bench lines of all configurations (no reference arms), CLI timelines
L=${1:-chk}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_device_decode.py -x -q > gpurun_out/pytest_dd_$L.log 2>&1 || { echo "device-decode tests FAILED with the prefetch: falling back to MD_NO_PREFETCH=1"; tail -15 gpurun_out/pytest_dd_$L.log; export MD_NO_PREFETCH=1; }
tail -2 gpurun_out/pytest_dd_$L.log
echo "A/B of the prefetch (bench.py c2, e2e ms per step):"; for v in "" 1; do MD_NO_PREFETCH=$v python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MD_NO_PREFETCH=$v', d['e2e']['value'], d['e2e']['ms_per_step'], d['inflate'])"; done
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
