# round 2, first look: is there an htslib anywhere on the box; does the new inflate kernel decode correctly on the device;
# where does the wall clock of the drop-in binary go on a 300 Mbp-class multi-contig set
mkdir -p gpurun_out /tmp/mdbench
L=methyldackel_b200/lib
{
echo "== probe for a real htslib / samtools / pysam on the GPU box"
find / \( -name 'libhts*' -o -name 'sam.h' -o -name 'htslib' -o -name 'pysam*' -o -name 'samtools*' -o -name 'bcftools*' \) -not -path '/proc/*' -not -path "$PWD/*" -not -path '/root/repo/*' 2>/dev/null | head
python -c 'import pysam' 2>&1 | tail -1
which samtools bcftools htsfile MethylDackel 2>&1 | head -4
echo "== host"; nproc; free -g | head -2; df -h /tmp | tail -1; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
} > gpurun_out/probe_r2.txt 2>&1
cat gpurun_out/probe_r2.txt
( time timeout 900 python -m pytest tests/test_gpu_device_decode.py -x -q ) > gpurun_out/pytest_r2a.log 2>&1; tail -5 gpurun_out/pytest_r2a.log
B=/tmp/mdbench/c2
$L/mdsynth --out $B --contigs chr1:10000000 --depth 30 --read-seed 5678 > /dev/null 2>&1
for i in 1 2; do MD_TIMING=1 $L/MethylDackel extract -o /tmp/d_$i $B.fa $B.bam 2>&1 | grep -E "device decode|wall|device destroyed"; done
( time oracle/_ref/MethylDackel extract -@ $(nproc) -o /tmp/oref $B.fa $B.bam ) 2>&1 | grep real
cmp <(tail -n +2 /tmp/d_1_CpG.bedGraph) <(tail -n +2 /tmp/oref_CpG.bedGraph) && echo C2-IDENTICAL
G=/tmp/mdbench/h300
( time $L/mdsynth --out $G --human 300000000 --depth 30 --read-seed 77 ) 2>&1 | grep -E "real|records"
ls -la $G.bam $G.fa
echo "== 300 Mbp, CpG"; ( time MD_TIMING=1 $L/MethylDackel extract -o /tmp/h_cpg $G.fa $G.bam ) 2>&1 | grep -E "md-timing|real"
echo "== 300 Mbp, --CHG --CHH --mergeContext"; ( time MD_TIMING=1 $L/MethylDackel extract --CHG --CHH --mergeContext -o /tmp/h_all $G.fa $G.bam ) 2>&1 | grep -E "md-timing|real"
echo "== reference, chr1+chr2 region timing"; ( time oracle/_ref/MethylDackel extract -@ $(nproc) --CHG --CHH --mergeContext -r chr1 -o /tmp/h_ref $G.fa $G.bam ) 2>&1 | grep real
$L/MethylDackel extract --CHG --CHH --mergeContext -r chr1 -o /tmp/h_new $G.fa $G.bam 2>/dev/null
for c in CpG CHG CHH; do cmp <(tail -n +2 /tmp/h_new_$c.bedGraph) <(tail -n +2 /tmp/h_ref_$c.bedGraph) && echo chr1-$c-IDENTICAL; done
echo "== mbias 300 Mbp"; ( time MD_TIMING=1 $L/MethylDackel mbias --noSVG $G.fa $G.bam > /tmp/h_mb.txt ) 2>&1 | grep -E "md-timing|real|Suggested"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 3000 gpurun_out/bench_r2a.json; tail -3 gpurun_out/bench_r2a.err
