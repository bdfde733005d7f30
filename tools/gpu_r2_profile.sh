# round 2 profiles: every kernel of the path captured with ncu --set full from a library REBUILT on the box from the tree
# that travels with it (the source hash is checked by tools/kprof.py), summaries + per-line hot spots under gpurun_out/
L=${1:-r2}
mkdir -p gpurun_out
touch methyldackel_b200/csrc/mdgpu.cu
make -C methyldackel_b200/csrc gpu all > gpurun_out/build_$L.log 2>&1 || { tail -20 gpurun_out/build_$L.log; exit 1; }
python tools/kprof.py > gpurun_out/kprof_$L.txt 2>&1; cat gpurun_out/kprof_$L.txt
# tools/kprof.py launches count_warp six times: extract defaults x2 (MODE 0), variant filter x2 (MODE 1), mbias x2 (MODE 2)
for SPEC in count_warp:1:count_warp_mode0 count_warp:3:count_warp_mode1 count_warp:5:count_warp_mode2 prep_kernel:1:prep_kernel per_read_kernel:1:per_read_kernel inflate_kernel:1:inflate_kernel; do
  K=${SPEC%%:*}; REST=${SPEC#*:}; SKIP=${REST%%:*}; N=${REST#*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^$K" --launch-skip $SKIP -c 1 -o gpurun_out/${N}_$L -f python tools/kprof.py > gpurun_out/ncu_${N}_$L.log 2>&1
  python profiles/summarize.py gpurun_out/${N}_$L.ncu-rep "round 2, $L: $N (tools/kprof.py, config[1] data)" > gpurun_out/summary_${N}_$L.md 2>&1; cat gpurun_out/summary_${N}_$L.md
  python profiles/ncu_source.py hotspots gpurun_out/${N}_$L.ncu-rep 1.0 > gpurun_out/hotspots_${N}_$L.txt 2>&1; head -3 gpurun_out/hotspots_${N}_$L.txt
done
cuobjdump -sass methyldackel_b200/lib/libmdgpu.so 2>/dev/null | grep -E "Function :|UBLKCP|SYNCS|LDGSTS|UTMALDG" | grep -B1 -E "UBLKCP|SYNCS|LDGSTS|UTMALDG" | head -60 > gpurun_out/sass_tma_$L.txt
