# CLI timeline on the 300 Mbp set + bench lines (c2, c3)
mkdir -p gpurun_out /tmp/mdbench
B=methyldackel_b200/lib/MethylDackel; G=/tmp/mdbench/h300
[ -f $G.bam.bai ] || { methyldackel_b200/lib/mdsynth --out $G --human 300000000 --depth 30 --read-seed 77 > $G.n 2>/dev/null; }
cat $G.bam > /dev/null
for i in 1 2; do ( time MD_TIMING=1 $B extract --CHG --CHH --mergeContext -o /dev/shm/t_all $G.fa $G.bam ) 2>&1 | grep -E "md-timing|real"; done
( time MD_TIMING=1 $B extract -o /dev/shm/t_cpg $G.fa $G.bam ) 2>&1 | grep -E "md-timing|real"
python tools/kprof.py 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cli.json 2> gpurun_out/bench_cli.err; tail -c 2600 gpurun_out/bench_cli.json
python bench.py --config c3 --mbp 300 --no-cpu-baseline > gpurun_out/bench_c3_cli.json 2> gpurun_out/bench_c3_cli.err; tail -c 1400 gpurun_out/bench_c3_cli.json
