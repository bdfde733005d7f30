# what one shard of a 4-way job costs on its own (1 GPU): timeline of ranks 0 and 2 of 4, extract and mbias, then 4 shards
# concurrently on the one GPU (host contention like the 4-GPU box, GPU contention worse), then the merge copy alone
mkdir -p gpurun_out /tmp/mdbench
B=methyldackel_b200/lib/MethylDackel; G=/tmp/mdbench/h300
[ -f $G.bam.bai ] || { methyldackel_b200/lib/mdsynth --out $G --human 300000000 --depth 30 --read-seed 77 > $G.n 2>/dev/null; }
cat $G.bam $G.fa > /dev/null
{
nproc; free -g | head -2
echo "== whole job, one process"
( time MD_TIMING=1 $B extract --CHG --CHH --mergeContext -o /dev/shm/t_all $G.fa $G.bam ) 2>&1 | grep -E "md-timing|real"
for r in 0 2; do
  echo "== extract shard $r/4 alone"
  ( time MD_TIMING=1 $B extract --CHG --CHH --mergeContext -o /dev/shm/t_s $G.fa $G.bam --shardRank $r --shardWorld 4 ) 2>&1 | grep -E "md-timing|real"
done
echo "== mbias whole"
( time MD_TIMING=1 $B mbias --txt $G.fa $G.bam /dev/shm/t_mb > /dev/null ) 2>&1 | grep -E "md-timing|real"
echo "== mbias shard 2/4 alone"
( time MD_TIMING=1 $B mbias --txt $G.fa $G.bam /dev/shm/t_mb --shardRank 2 --shardWorld 4 --histOut /dev/shm/t_h2 > /dev/null ) 2>&1 | grep -E "md-timing|real"
echo "== 4 extract shards at once on one GPU"
( time ( for r in 0 1 2 3; do MD_TIMING=1 $B extract --CHG --CHH --mergeContext -o /dev/shm/t_p $G.fa $G.bam --shardRank $r --shardWorld 4 2> /tmp/shard$r.err & done; wait ) ) 2>&1 | grep real
for r in 0 1 2 3; do echo "-- rank $r"; grep -E "md-timing" /tmp/shard$r.err | tail -12; done
ls -l /dev/shm/t_p* | head
echo "== merge copy alone (python, copy_file_range of the CpG/CHG/CHH shards)"
python - <<'PY'
import os, time
t=time.time(); tot=0
for c in ("CpG","CHG","CHH"):
    dst=os.open("/dev/shm/t_m_%s"%c, os.O_WRONLY|os.O_CREAT|os.O_TRUNC, 0o644); off=0
    for r in range(4):
        p="/dev/shm/t_p_%s.bedGraph.shard%d"%(c,r)
        if not os.path.exists(p): print("missing",p); continue
        src=os.open(p, os.O_RDONLY); n=os.fstat(src).st_size; done=0
        while done<n:
            k=os.copy_file_range(src,dst,n-done,done,off+done)
            if k<=0: break
            done+=k
        off+=n; tot+=n; os.close(src)
    os.close(dst)
print("serial copy of %.2f GB: %.2f s"%(tot/1e9,time.time()-t))
PY
} > gpurun_out/shardtime.txt 2>&1
tail -5 gpurun_out/shardtime.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_head.txt 2>&1; tail -5 gpurun_out/pytest_head.txt
