L=methyldackel_b200/lib
mkdir -p /tmp/mdbench
B=/tmp/mdbench/c2_10mbp_r0
[ -f $B.bam.bai ] || $L/mdsynth --out $B --contigs chr1:10000000 --depth 30 --read-seed 5678 >/dev/null 2>&1
run() { echo "== $*"; for i in 1 2 3; do env "$@" MD_TIMING=1 $L/MethylDackel extract -o /tmp/o_x $B.fa $B.bam 2>&1 | grep -E "context ready|wall" | tr '\n' ' '; echo; done; }
run MD_WARM_THREADS=0
run MD_WARM_THREADS=0 MD_NO_MALLOPT=1
run MD_WARM_THREADS=0 MD_DECODE_THREADS=4
run MD_WARM_THREADS=8
run MD_WARM_THREADS=64
python - <<'PY'
import time, ctypes
t=time.time(); l=ctypes.CDLL("libcudart.so.12"); p=ctypes.c_void_p(); l.cudaFree(0); print("bare cudaFree(0) in python: %.3f s" % (time.time()-t))
PY
