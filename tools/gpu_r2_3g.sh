# BASELINE.json configs[2] and configs[4] at FULL size: 3 Gbp x 30x, 24 contigs (600 M alignments, ~55 GB of BAM)
mkdir -p gpurun_out /tmp/mdbench
df -h /tmp /dev/shm | tail -2; free -g | head -2
( time python bench.py --config c3 --mbp 3000 --steps 1 --warmup 1 ) > gpurun_out/bench_c3_3g.json 2> gpurun_out/bench_c3_3g.err; tail -c 3500 gpurun_out/bench_c3_3g.json; tail -5 gpurun_out/bench_c3_3g.err
rm -rf /dev/shm/mdbench_out; free -g | head -2
( time python bench.py --config c5 --mbp 3000 --steps 1 --warmup 1 ) > gpurun_out/bench_c5_3g.json 2> gpurun_out/bench_c5_3g.err; tail -c 3000 gpurun_out/bench_c5_3g.json; tail -5 gpurun_out/bench_c5_3g.err
rm -rf /dev/shm/mdbench_out
( time MD_TIMING=1 methyldackel_b200/lib/MethylDackel extract --CHG --CHH --mergeContext -o /dev/shm/t3g /tmp/mdbench/h3000.fa /tmp/mdbench/h3000.bam ) 2>&1 | grep -E "md-timing|real" | tail -12; ls -la /dev/shm/ | head; rm -f /dev/shm/t3g*
