# CLI wall-clock study on the GPU box: ours vs the shim-built reference, small (config[1]) and a 4x larger multi-contig input
set -e
L=methyldackel_b200/lib
mkdir -p /tmp/mdbench gpurun_out
B=/tmp/mdbench/c2_10mbp_r0
[ -f $B.bam.bai ] || $L/mdsynth --out $B --contigs chr1:10000000 --depth 30 --read-seed 5678 >/dev/null 2>&1
for i in 1 2 3; do MD_TIMING=1 $L/MethylDackel extract -o /tmp/o_$i $B.fa $B.bam 2>&1 | grep -E "context ready|device joined|last tile|destroyed|wall|calling"; done
( time oracle/_ref/MethylDackel extract -@ 128 -o /tmp/oref $B.fa $B.bam ) 2>&1 | grep real
cmp <(tail -n +2 /tmp/o_1_CpG.bedGraph) <(tail -n +2 /tmp/oref_CpG.bedGraph) && echo SMALL-IDENTICAL
G=/tmp/mdbench/big4
( time $L/mdsynth --out $G --contigs chr1:12000000,chr2:10000000,chr3:10000000,chr4:8000000 --depth 30 --read-seed 99 ) 2>&1 | grep real
ls -la $G.bam
for i in 1 2; do MD_TIMING=1 $L/MethylDackel extract -o /tmp/g_$i $G.fa $G.bam 2>&1 | grep -E "context ready|wall|calling"; done
( time oracle/_ref/MethylDackel extract -@ 128 -o /tmp/gref $G.fa $G.bam ) 2>&1 | grep real
cmp <(tail -n +2 /tmp/g_1_CpG.bedGraph) <(tail -n +2 /tmp/gref_CpG.bedGraph) && echo BIG-IDENTICAL
MD_TIMING=1 $L/MethylDackel extract --CHG --CHH -o /tmp/ga $G.fa $G.bam 2>&1 | grep -E "context ready|wall|calling"
( time oracle/_ref/MethylDackel extract -@ 128 --CHG --CHH -o /tmp/garef $G.fa $G.bam ) 2>&1 | grep real
cmp <(tail -n +2 /tmp/ga_CHH.bedGraph) <(tail -n +2 /tmp/garef_CHH.bedGraph) && echo ALLCTX-IDENTICAL
MD_TIMING=1 $L/MethylDackel mbias $G.fa $G.bam /tmp/gm 2>&1 | grep -i "context ready\|wall\|Suggested"
( time oracle/_ref/MethylDackel mbias -@ 128 --txt $G.fa $G.bam /tmp/gmref ) 2>&1 | grep -i "real\|Suggested"
