# one --set full capture (with source) of the count kernel on config[1] (CpG-only), per-line hot spots
# usage (on the GPU box): bash tools/gpu_prof.sh [label] [cfg]
L=${1:-cur}; CFG=${2:-cpg}
mkdir -p gpurun_out /tmp/mdbench
timeout 600 ncu --set full --clock-control none --import-source on -k regex:count_warp --launch-skip 5 -c 1 -o gpurun_out/count_warp_$L -f \
    python tools/kbench.py --variants ${MD_GEN:-1} --steps 3 --only $CFG > gpurun_out/ncu_$L.log 2>&1
python profiles/summarize.py gpurun_out/count_warp_$L.ncu-rep "$L" > gpurun_out/summary_$L.md 2>&1; cat gpurun_out/summary_$L.md
python profiles/ncu_hotspots.py gpurun_out/count_warp_$L.ncu-rep 0.8 > gpurun_out/hotspots_$L.txt 2>&1; cat gpurun_out/hotspots_$L.txt | cut -c1-220
