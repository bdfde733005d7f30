# CLI wall-clock study on the GPU box: host decode vs device decode vs the shim-built reference
set -e
L=methyldackel_b200/lib
mkdir -p /tmp/mdbench gpurun_out
B=/tmp/mdbench/c2_10mbp_r0
[ -f $B.bam.bai ] || $L/mdsynth --out $B --contigs chr1:10000000 --depth 30 --read-seed 5678 >/dev/null 2>&1
echo "== small, host decode"; for i in 1 2; do MD_TIMING=1 $L/MethylDackel extract -o /tmp/o_$i $B.fa $B.bam 2>&1 | grep -E "context ready|wall"; done
echo "== small, device decode"; for i in 1 2 3; do MD_DEVICE_DECODE=1 MD_TIMING=1 $L/MethylDackel extract -o /tmp/d_$i $B.fa $B.bam 2>&1 | grep -E "context ready|wall|device decode"; done
( time oracle/_ref/MethylDackel extract -@ 128 -o /tmp/oref $B.fa $B.bam ) 2>&1 | grep real
cmp <(tail -n +2 /tmp/d_1_CpG.bedGraph) <(tail -n +2 /tmp/oref_CpG.bedGraph) && echo SMALL-IDENTICAL
G=/tmp/mdbench/big4
( time $L/mdsynth --out $G --contigs chr1:12000000,chr2:10000000,chr3:10000000,chr4:8000000 --depth 30 --read-seed 99 ) 2>&1 | grep real
echo "== big, host decode"; MD_TIMING=1 $L/MethylDackel extract -o /tmp/g_1 $G.fa $G.bam 2>&1 | grep -E "context ready|wall"
echo "== big, device decode"; for s in 33554432 134217728 1073741824; do echo "segment $s"; MD_SEGMENT_BYTES=$s MD_DEVICE_DECODE=1 MD_TIMING=1 $L/MethylDackel extract -o /tmp/gd $G.fa $G.bam 2>&1 | grep -E "context ready|wall|device decode"; done
( time oracle/_ref/MethylDackel extract -@ 128 -o /tmp/gref $G.fa $G.bam ) 2>&1 | grep real
cmp <(tail -n +2 /tmp/gd_CpG.bedGraph) <(tail -n +2 /tmp/gref_CpG.bedGraph) && echo BIG-IDENTICAL
echo "== big all contexts, device decode"; MD_DEVICE_DECODE=1 MD_TIMING=1 $L/MethylDackel extract --CHG --CHH -o /tmp/ga $G.fa $G.bam 2>&1 | grep -E "context ready|wall|device decode"
( time oracle/_ref/MethylDackel extract -@ 128 --CHG --CHH -o /tmp/garef $G.fa $G.bam ) 2>&1 | grep real
cmp <(tail -n +2 /tmp/ga_CHH.bedGraph) <(tail -n +2 /tmp/garef_CHH.bedGraph) && echo ALLCTX-IDENTICAL
echo "== big mbias, device decode"; MD_DEVICE_DECODE=1 MD_TIMING=1 $L/MethylDackel mbias $G.fa $G.bam /tmp/gm 2>&1 | grep -i "context ready\|wall\|Suggested\|device decode"
( time oracle/_ref/MethylDackel mbias -@ 128 --txt $G.fa $G.bam /tmp/gmref ) 2>&1 | grep -i "real\|Suggested"
