# one ncu --set full capture of the inflate kernel (second push of tools/kprof.py's device decode) + per-line hot spots
L=${1:-inf}
mkdir -p gpurun_out
KPROF_ONLY=decode timeout 600 ncu --set full --clock-control none --import-source on -k "regex:inflate_kernel" --launch-skip 1 -c 1 -o gpurun_out/inflate_kernel_$L -f python tools/kprof.py > gpurun_out/ncu_inflate_$L.log 2>&1
tail -2 gpurun_out/ncu_inflate_$L.log
python profiles/summarize.py gpurun_out/inflate_kernel_$L.ncu-rep "round 2, $L: inflate_kernel (tools/kprof.py, config[1] file in one segment)" > gpurun_out/summary_inflate_kernel_$L.md 2>&1; cat gpurun_out/summary_inflate_kernel_$L.md
python profiles/ncu_source.py hotspots gpurun_out/inflate_kernel_$L.ncu-rep 0.4 > gpurun_out/hotspots_inflate_kernel_$L.txt 2>&1; head -50 gpurun_out/hotspots_inflate_kernel_$L.txt
KPROF_ONLY=decode KPROF_REPS=6 python tools/kprof.py | tail -1
