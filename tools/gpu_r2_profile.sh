# round 2 profiles: every kernel of the path captured with ncu --set full from a library REBUILT on the box from the tree
# that travels with it (the source hash is checked by tools/kprof.py), summaries + per-line hot spots under gpurun_out/
L=${1:-r2}
mkdir -p gpurun_out
touch methyldackel_b200/csrc/mdgpu.cu
make -C methyldackel_b200/csrc gpu all > gpurun_out/build_$L.log 2>&1 || { tail -20 gpurun_out/build_$L.log; exit 1; }
python tools/kprof.py > gpurun_out/kprof_$L.txt 2>&1; cat gpurun_out/kprof_$L.txt
for K in "count_warp<\(int\)0" "count_warp<\(int\)1" "count_warp<\(int\)2" prep_kernel per_read_kernel inflate_kernel; do
  N=$(echo $K | tr -cd 'a-z_0-9')
  SKIP=1; [ "$N" = prep_kernel ] && SKIP=1
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$K" --launch-skip $SKIP -c 1 -o gpurun_out/${N}_$L -f python tools/kprof.py > gpurun_out/ncu_${N}_$L.log 2>&1
  python profiles/summarize.py gpurun_out/${N}_$L.ncu-rep "round 2, $L: $N (tools/kprof.py, config[1] data)" > gpurun_out/summary_${N}_$L.md 2>&1; cat gpurun_out/summary_${N}_$L.md
  python profiles/ncu_source.py hotspots gpurun_out/${N}_$L.ncu-rep 1.0 > gpurun_out/hotspots_${N}_$L.txt 2>&1; head -3 gpurun_out/hotspots_${N}_$L.txt
done
cuobjdump -sass methyldackel_b200/lib/libmdgpu.so 2>/dev/null | grep -E "Function :|UBLKCP|SYNCS|LDGSTS|UTMALDG" | grep -B1 -E "UBLKCP|SYNCS|LDGSTS|UTMALDG" | head -60 > gpurun_out/sass_tma_$L.txt
