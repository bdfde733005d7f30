# sharded product path on N GPUs of one box: the large configuration through api.extract_sharded / mbias_sharded (one process
# per GPU, torchrun, NCCL only for the barriers / reductions of the launcher), output compared with the single-process run
N=${1:-2}; MBP=${2:-300}; L=${3:-multi$N}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$L.txt 2>&1
methyldackel_b200/lib/mdsynth --out /tmp/mdbench/h${MBP}.tmp --human ${MBP}000000 --depth 30 --read-seed 77 > /tmp/mdbench_n.txt 2>/dev/null && for e in fa fa.fai bam bam.bai; do mv /tmp/mdbench/h${MBP}.tmp.$e /tmp/mdbench/h${MBP}.$e; done && tail -1 /tmp/mdbench_n.txt > /tmp/mdbench/h${MBP}.n
for C in c3 c5; do
  MD_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config $C --mbp $MBP --no-cpu-baseline > gpurun_out/bench_${C}_$L.json 2> gpurun_out/bench_${C}_$L.err; tail -c 2500 gpurun_out/bench_${C}_$L.json; grep -v md-timing gpurun_out/bench_${C}_$L.err | tail -3
done
# single-process output for the byte comparison
methyldackel_b200/lib/MethylDackel extract --CHG --CHH --mergeContext -o /dev/shm/one /tmp/mdbench/h${MBP}.fa /tmp/mdbench/h${MBP}.bam 2>/dev/null
O=$(python -c "import bench; print(bench.out_dir())")
for c in CpG CHG CHH; do cmp /dev/shm/one_$c.bedGraph <(sed "1s#$O/h${MBP}_c3_w$N#/dev/shm/one#" $O/h${MBP}_c3_w${N}_$c.bedGraph) && echo "SHARDED-$N-$c-IDENTICAL"; done 2>&1 | tee gpurun_out/cmp_$L.txt

python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_c2_$L.json 2> gpurun_out/bench_c2_$L.err; tail -c 1500 gpurun_out/bench_c2_$L.json
