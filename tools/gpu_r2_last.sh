# the default bench line and its reference arm from the last tree of the round
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_c2_last.json 2> gpurun_out/bench_c2_last.err; tail -c 1200 gpurun_out/bench_c2_last.json; tail -2 gpurun_out/bench_c2_last.err
python bench.py --impl reference > gpurun_out/bench_c2_ref_last.json 2> gpurun_out/bench_c2_ref_last.err; tail -c 600 gpurun_out/bench_c2_ref_last.json
