# where does a cold `MethylDackel extract` on the 300 Mbp set spend its time, and what does the A/B build of the inflate kernel give
mkdir -p gpurun_out /tmp/mdbench
B=methyldackel_b200/lib/MethylDackel; G=/tmp/mdbench/h300
[ -f $G.bam.bai ] || methyldackel_b200/lib/mdsynth --out $G --human 300000000 --depth 30 --read-seed 77 > /dev/null 2>&1
cat $G.bam > /dev/null
run() { echo "== $*"; ( time env "$@" $B extract --CHG --CHH --mergeContext -o /dev/shm/t_all $G.fa $G.bam ) 2>&1 | grep -E "md-timing|real"; }
run A=1
run MD_TIMING=1
run MD_TIMING=1 MD_STAGE=0
run MD_TIMING=1 MD_FORMAT_THREADS=4
run MD_TIMING=1 CUDA_MODULE_LOADING=EAGER
echo "== CpG only"; ( time MD_TIMING=1 $B extract -o /dev/shm/t_cpg $G.fa $G.bam ) 2>&1 | grep -E "md-timing|real"
echo "== mbias"; ( time MD_TIMING=1 $B mbias --noSVG $G.fa $G.bam > /dev/null ) 2>&1 | grep -E "md-timing|real"
echo "== kprof default build"; python tools/kprof.py 2>&1 | tail -3
echo "== kprof 2 KB ring, 32 warps per SM"; KPROF_NOHASH=1 MD_LIBMDGPU=$PWD/methyldackel_b200/lib/libmdgpu_w2k.so python tools/kprof.py 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:inflate_kernel --launch-skip 1 -c 1 -o gpurun_out/inflate_kernel_cold -f python tools/kprof.py > gpurun_out/ncu_inflate_cold.log 2>&1
python profiles/summarize.py gpurun_out/inflate_kernel_cold.ncu-rep "inflate_kernel, rare paths out of line" 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^count_warp --launch-skip 1 -c 1 -o gpurun_out/count_warp_gen2 -f python tools/kprof.py > gpurun_out/ncu_count_gen2.log 2>&1
python profiles/summarize.py gpurun_out/count_warp_gen2.ncu-rep "count_warp<0,2>" 2>&1
