# iteration run: parity of the new count generator, A/B of the generators, the bench line, cold CLI timing on the 300 Mbp set
L=${1:-it}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_$L.log 2>&1; tail -6 gpurun_out/pytest_$L.log
python tools/kbench.py --variants 1,2 --steps 10 --panel > gpurun_out/kbench_$L.jsonl 2> gpurun_out/kbench_$L.err; cat gpurun_out/kbench_$L.jsonl
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$L.json 2> gpurun_out/bench_$L.err; tail -c 3200 gpurun_out/bench_$L.json; tail -3 gpurun_out/bench_$L.err
python bench.py --config c3 --mbp 300 --no-cpu-baseline > gpurun_out/bench_c3_$L.json 2> gpurun_out/bench_c3_$L.err; tail -c 1500 gpurun_out/bench_c3_$L.json
for i in 1 2; do ( time MD_TIMING=1 methyldackel_b200/lib/MethylDackel extract --CHG --CHH --mergeContext -o /dev/shm/t_all /tmp/mdbench/h300.fa /tmp/mdbench/h300.bam ) 2>&1 | grep -E "md-timing|real"; done
