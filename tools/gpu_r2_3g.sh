# BASELINE.json configs[2] and configs[4] at FULL size: 3 Gbp x 30x, 24 contigs (600 M alignments, ~55 GB of BAM)
mkdir -p gpurun_out /tmp/mdbench
df -h /tmp /dev/shm | tail -2; free -g | head -2
( time MD_TIMING=1 python bench.py --config c3 --mbp 3000 --steps 1 --warmup 3 ) > gpurun_out/bench_c3_3g.json 2> gpurun_out/bench_c3_3g.err; tail -c 3500 gpurun_out/bench_c3_3g.json; grep -v md-timing gpurun_out/bench_c3_3g.err | tail -5
rm -rf /dev/shm/mdbench_out; free -g | head -2
( time MD_TIMING=1 python bench.py --config c5 --mbp 3000 --steps 1 --warmup 3 ) > gpurun_out/bench_c5_3g.json 2> gpurun_out/bench_c5_3g.err; tail -c 3000 gpurun_out/bench_c5_3g.json; grep -v md-timing gpurun_out/bench_c5_3g.err | tail -5
rm -rf /dev/shm/mdbench_out
