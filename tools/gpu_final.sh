# round-end measurement: bench line (x2 for the decoders-per-warp choice), launch list under ncu, one --set full capture of the decode kernels
mkdir -p gpurun_out
MD_INFLATE_DPW=2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dpw2.json 2> gpurun_out/bench_dpw2.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_dpw2.json", "gpurun_out/bench_v7.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, "value", d["value"], "e2e", d["e2e"]["value"], "e2e_bam", d["e2e_bam"]["value"], d["e2e_bam"]["ms_per_step"], "cli", d.get("cli_from_bam", {}).get("seconds"))
    except Exception as e:
        print(f, "ERR", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_v7.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_v7.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].split("(")[0]; a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]): print("%-60s n=%4d total %.3f ms mean %.1f us" % (k[:60], n, t / 1e6, t / n / 1e3))
PY
