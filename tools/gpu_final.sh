# round-end measurement: full GPU test tier, the bench line, the reference arm, the launch list under ncu and one --set full
# capture of the count kernel inside bench.py; summaries are written under gpurun_out/ (copied to profiles/ by hand)
L=${1:-final}
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_$L.log 2>&1; tail -6 gpurun_out/pytest_$L.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$L.json 2> gpurun_out/bench_$L.err; tail -c 4500 gpurun_out/bench_$L.json; tail -3 gpurun_out/bench_$L.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$L.json 2> gpurun_out/bench_ref_$L.err; tail -c 1500 gpurun_out/bench_ref_$L.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$L.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$L.log 2>&1
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
timeout 600 ncu --set full --clock-control none --import-source on -k regex:count_warp --launch-skip 3 -c 1 -o gpurun_out/count_warp_$L -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$L.log 2>&1
python profiles/summarize.py gpurun_out/count_warp_$L.ncu-rep "round 1, $L (count_warp inside bench.py)" > gpurun_out/summary_$L.md 2>&1; cat gpurun_out/summary_$L.md
python profiles/ncu_hotspots.py gpurun_out/count_warp_$L.ncu-rep 1.0 > gpurun_out/hotspots_$L.txt 2>&1
python - "$L" <<'PY'
import csv, json, subprocess, sys
L = sys.argv[1]
txt = subprocess.run(["ncu", "-i", "gpurun_out/count_warp_%s.ncu-rep" % L, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines())); hdr, units, r = rows[0], rows[1], rows[2]
def val(k):
    v = float(r[hdr.index(k)].replace(",", "")); u = units[hdr.index(k)]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
out = {"kernel": r[hdr.index("Kernel Name")], "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
       "source": "ncu --set full --clock-control none, one launch of the kernel inside `bench.py --steps 2 --warmup 3` (tools/gpu_final.sh)"}
json.dump(out, open("gpurun_out/traffic_%s.json" % L, "w")); print(out)
PY
