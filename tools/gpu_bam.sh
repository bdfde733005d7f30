# device-decode timing: per-stage times of the CLI (MD_TIMING) and the e2e_bam leg of the bench
L=${1:-bam}
mkdir -p gpurun_out /tmp/mdbench
B=/tmp/mdbench/c2_10mbp_r0
[ -f $B.bam.bai ] || methyldackel_b200/lib/mdsynth --out $B --contigs chr1:10000000 --depth 30 --read-seed 5678 >/dev/null 2>&1
for i in 1 2 3; do MD_TIMING=1 methyldackel_b200/lib/MethylDackel extract -o /tmp/d_$i $B.fa $B.bam 2>&1 | grep -E "device decode|wall"; done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$L.json 2> gpurun_out/bench_$L.err
python - "$L" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "e2e_bam", d["e2e_bam"]["value"], d["e2e_bam"]["ms_per_step"], "segments", d["e2e_bam"]["segments_per_step"], "cli", d.get("cli_from_bam", {}).get("seconds"))
PY
