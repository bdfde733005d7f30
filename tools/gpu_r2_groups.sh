# lanes-per-decoder sweep of the inflate kernel (variants built here by `make gpu OUT=lib/variants/<name> EXTRA=-D...`):
# correctness of each variant through the CLI device-decode tests, then its kernel time on the config[1] file
mkdir -p gpurun_out
V=methyldackel_b200/lib/variants
{
echo "== in-tree build (MD_INFLATE_GROUP default)"
timeout 420 python -m pytest tests/test_gpu_device_decode.py -x -q -k "flavours or multi_segment or noisy" > gpurun_out/groups_t0.txt 2>&1; RC=$?; tail -2 gpurun_out/groups_t0.txt
[ $RC = 124 ] && { echo "in-tree build hangs: giving up"; exit 1; }
KPROF_ONLY=decode KPROF_REPS=6 timeout 120 python tools/kprof.py 2>&1 | tail -1
for n in $(ls $V); do
  echo "== $n"
  LD_LIBRARY_PATH=$PWD/$V/$n MD_LIBMDGPU=$PWD/$V/$n/libmdgpu.so timeout 240 python -m pytest tests/test_gpu_device_decode.py -x -q -k "flavours or multi_segment or noisy" 2>&1 | tail -1
  MD_LIBMDGPU=$PWD/$V/$n/libmdgpu.so KPROF_ONLY=decode KPROF_REPS=6 timeout 120 python tools/kprof.py 2>&1 | tail -1
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_groups.err | tee gpurun_out/bench_groups.json | tail -c 1800
} > gpurun_out/groups.txt 2>&1
cat gpurun_out/groups.txt
