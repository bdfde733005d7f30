# round 2 final evidence on one B200: smoke, GPU test tier, the bench lines of every configuration with their reference arms,
# CLI timelines, then the profile run (tools/gpu_r2_profile.sh) and the ncu launch list of the default bench command
L=${1:-r2f}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_$L.txt 2>&1; nproc >> gpurun_out/gpu_$L.txt; free -g | head -2 >> gpurun_out/gpu_$L.txt
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_$L.log 2>&1; tail -4 gpurun_out/smoke_$L.log
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_$L.log 2>&1; tail -6 gpurun_out/pytest_$L.log
python bench.py > gpurun_out/bench_c2_$L.json 2> gpurun_out/bench_c2_$L.err; tail -c 1500 gpurun_out/bench_c2_$L.json; tail -2 gpurun_out/bench_c2_$L.err
python bench.py --impl reference > gpurun_out/bench_c2_ref_$L.json 2> gpurun_out/bench_c2_ref_$L.err; tail -c 800 gpurun_out/bench_c2_ref_$L.json
for C in c3 c5 c4; do
  M="--mbp 300"; [ $C = c4 ] && M=""
  MD_TIMING=1 python bench.py --config $C $M > gpurun_out/bench_${C}_$L.json 2> gpurun_out/bench_${C}_$L.err; tail -c 1500 gpurun_out/bench_${C}_$L.json; grep -v md-timing gpurun_out/bench_${C}_$L.err | tail -2
  python bench.py --config $C $M --impl reference --steps 2 --warmup 0 > gpurun_out/bench_${C}_ref_$L.json 2> gpurun_out/bench_${C}_ref_$L.err; tail -c 800 gpurun_out/bench_${C}_ref_$L.json
done
G=/tmp/mdbench/h300; B=methyldackel_b200/lib/MethylDackel
{ for i in 1 2; do echo "== extract --CHG --CHH --mergeContext (run $i)"; ( time MD_TIMING=1 $B extract --CHG --CHH --mergeContext -o /dev/shm/t_all $G.fa $G.bam ) 2>&1 | grep -E "md-timing|real"; done
  echo "== extract (CpG)"; ( time MD_TIMING=1 $B extract -o /dev/shm/t_cpg $G.fa $G.bam ) 2>&1 | grep -E "md-timing|real"
  echo "== mbias"; ( time MD_TIMING=1 $B mbias --txt $G.fa $G.bam /dev/shm/t_mb > /dev/null ) 2>&1 | grep -E "md-timing|real"; } > gpurun_out/cli_timeline_$L.txt 2>&1
rm -f /dev/shm/t_all* /dev/shm/t_cpg* /dev/shm/t_mb*
bash tools/gpu_r2_profile.sh $L > gpurun_out/profile_$L.log 2>&1; tail -5 gpurun_out/profile_$L.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$L.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$L.log 2>&1
python tools/launch_shares.py gpurun_out/launches_$L.csv > gpurun_out/launch_shares_$L.txt 2>&1; head -8 gpurun_out/launch_shares_$L.txt
# roofline.traffic: DRAM bytes of the first timed device-resident count_warp launch inside the default bench command
timeout 600 ncu --set full --clock-control none --import-source on -k regex:count_warp --launch-skip 3 -c 1 -o gpurun_out/count_warp_bench_$L -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$L.log 2>&1
python tools/ncu_traffic.py gpurun_out/count_warp_bench_$L.ncu-rep "ncu --set full --clock-control none --launch-skip 3 -c 1 -k regex:count_warp: the first timed device-resident launch of count_warp inside \`python bench.py --steps 2 --warmup 3 --no-cpu-baseline\` (tools/gpu_r2_final.sh); summary in profiles/r2_ncu_summaries.md" > gpurun_out/traffic_$L.json 2>&1; cat gpurun_out/traffic_$L.json
python profiles/summarize.py gpurun_out/count_warp_bench_$L.ncu-rep "round 2, $L: count_warp<0,2> inside bench.py (config[1], device-resident)" > gpurun_out/summary_count_warp_bench_$L.md 2>&1; cat gpurun_out/summary_count_warp_bench_$L.md
python profiles/ncu_source.py ranges gpurun_out/count_warp_bench_$L.ncu-rep > gpurun_out/ranges_count_warp_bench_$L.txt 2>&1; head -12 gpurun_out/ranges_count_warp_bench_$L.txt
