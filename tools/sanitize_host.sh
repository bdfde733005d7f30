#!/bin/bash
# The host side (lib/libmdhost.so: segment reader threads, hand-over thread, text stage, sharding) and the device-decode emulation under
# AddressSanitizer + UBSan (whole CPU test tier) and under ThreadSanitizer (the threaded drivers, run directly: make / g++ cannot run
# under a preloaded libtsan).  No GPU needed: the back end is the oracle port, as everywhere in the CPU tier.  Restores the normal builds.
set -e
cd "$(dirname "$0")/.."
G=$(dirname "$(gcc -print-file-name=libasan.so)")
build() { g++ -O1 -g -std=c++17 -fPIC $1 -fno-omit-frame-pointer -shared -pthread -o methyldackel_b200/lib/libmdhost.so methyldackel_b200/csrc/host/cli.cpp -lz
          g++ -O1 -g -std=c++17 -fPIC $1 -shared -o tests/native/libmdemu.so tests/native/mdemu.cpp; }
restore() { make -s -B -C methyldackel_b200/csrc host; make -s -B -C tests/native libmdemu.so; make -s -C methyldackel_b200/csrc all; }
trap restore EXIT
python -m pytest tests -q -m "not gpu" -x > /dev/null            # builds everything the fixtures need, unsanitised
build "-fsanitize=address,undefined"
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 LD_PRELOAD=$G/libasan.so:$G/libubsan.so \
    python -m pytest tests -x -q -m "not gpu" -p no:cacheprovider | tail -2
build "-fsanitize=thread"
T=$(mktemp -d); methyldackel_b200/lib/mdsynth --out $T/t --contigs chr1:60000,chr2:15000 --depth 25 > /dev/null 2>&1
for seg in 1 30000; do
  TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0" LD_PRELOAD=$G/libtsan.so MD_DEVICE_DECODE=1 MD_STAGE=1 MD_SEGMENT_BYTES=$seg python -c "
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import oracle_binding as ob
b = ob.OracleBackend(device_decode=True, overlapped=True, staging=True)
print('extract rc', ob.run_host_main('extract', ['--CHG', '--CHH', '--mergeContext', '$T/t.fa', '$T/t.bam', '-o', '$T/o$seg'], b), 'prefetched segments', b.state.get('prefetch_used'))
print('mbias rc', ob.run_host_main('mbias', ['--noSVG', '$T/t.fa', '$T/t.bam'], ob.OracleBackend(device_decode=True, overlapped=True, staging=True)))
print('host-decode rc', ob.run_host_main('extract', ['--CHG', '$T/t.fa', '$T/t.bam', '-o', '$T/h$seg', '-@', '4'], ob.OracleBackend()))
" > $T/log$seg 2>&1
  echo "segments of $seg bytes: $(grep -c 'WARNING: ThreadSanitizer' $T/log$seg) ThreadSanitizer reports; $(grep -c ' rc 0' $T/log$seg)/3 runs ok"
done
rm -rf $T
