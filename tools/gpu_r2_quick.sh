# quick iteration: device-decode tests, the kernel driver's timings, one ncu capture of the inflate kernel, the bench line
L=${1:-q}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_device_decode.py -x -q ) > gpurun_out/pytest_$L.log 2>&1; tail -4 gpurun_out/pytest_$L.log
python tools/kprof.py > gpurun_out/kprof_$L.txt 2>&1; cat gpurun_out/kprof_$L.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:inflate_kernel --launch-skip 1 -c 1 -o gpurun_out/inflate_kernel_$L -f python tools/kprof.py > gpurun_out/ncu_inflate_$L.log 2>&1
python profiles/summarize.py gpurun_out/inflate_kernel_$L.ncu-rep "round 2, $L: inflate_kernel" > gpurun_out/summary_inflate_kernel_$L.md 2>&1; cat gpurun_out/summary_inflate_kernel_$L.md
python profiles/ncu_source.py hotspots gpurun_out/inflate_kernel_$L.ncu-rep 1.0 > gpurun_out/hotspots_inflate_kernel_$L.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$L.json 2> gpurun_out/bench_$L.err; tail -c 4000 gpurun_out/bench_$L.json; tail -3 gpurun_out/bench_$L.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$L.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$L.log 2>&1
python - "$L" <<'PY'
import csv, collections, sys
L = sys.argv[1]
rows = [r for r in csv.reader(open("gpurun_out/launches_%s.csv" % L)) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].split("(")[0]; a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
with open("gpurun_out/launch_shares_%s.txt" % L, "w") as f:
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        line = "%-64s n=%4d total %9.3f ms mean %8.1f us share %5.1f %%" % (k[:64], n, t / 1e6, t / n / 1e3, 100 * t / tot)
        print(line); f.write(line + "\n")
PY
