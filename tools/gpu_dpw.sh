# inflate kernel: decoders per warp sweep (device stage times from MD_TIMING), then the bench line
L=methyldackel_b200/lib
mkdir -p /tmp/mdbench gpurun_out
B=/tmp/mdbench/c2_10mbp_r0
[ -f $B.bam.bai ] || $L/mdsynth --out $B --contigs chr1:10000000 --depth 30 --read-seed 5678 >/dev/null 2>&1
for d in 1 2 4 8 16 32; do echo "dpw=$d"; MD_INFLATE_DPW=$d MD_TIMING=1 $L/MethylDackel extract -o /tmp/d_$d $B.fa $B.bam 2>&1 | grep -E "device decode|wall"; done
cmp /tmp/d_1_CpG.bedGraph /tmp/d_8_CpG.bedGraph && cmp /tmp/d_1_CpG.bedGraph /tmp/d_32_CpG.bedGraph && echo SAME
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; tail -c 3000 gpurun_out/bench_v7.json; tail -5 gpurun_out/bench_v7.err
