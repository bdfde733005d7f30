// `MethylDackel` drop-in binary for the B200 path: same dispatcher surface as main.c:39-62 for the
// sub-commands this build accelerates; extract_main / mbias_main / perRead_main come from libMethylDackel.so (host/dropin.cpp),
// bound to libmdgpu and nothing else: without a usable CUDA device the program fails instead of computing on the CPU.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <chrono>
#include "../../../include/methyldackel.h"

static void usage_main() {
    fprintf(stderr, "MethylDackel (B200 build of the extract/mbias hot path): A tool for processing bisulfite sequencing alignments.\n"
                    "Usage: MethylDackel <command> [options]\n\nCommands:\n"
                    "    mbias    Determine the position-dependent methylation bias in a dataset.\n"
                    "    extract  Extract methylation metrics from an alignment file in BAM format.\n"
                    "    perRead  Generate a per-read methylation summary.\n"
                    "(mergeContext is not part of this build.)\n");
}

int main(int argc, char *argv[]) {
    if (argc == 1) { usage_main(); return 0; }
    if (!strcmp(argv[1], "-h") || !strcmp(argv[1], "--help")) { usage_main(); return 0; }
    if (!strcmp(argv[1], "-v") || !strcmp(argv[1], "--version")) { printf("0.6.1-b200 (B200 build; no HTSlib)\n"); return 0; }
    if (!strcmp(argv[1], "extract") || !strcmp(argv[1], "mbias")) {
        auto t0 = std::chrono::steady_clock::now();
        int rc = !strcmp(argv[1], "extract") ? extract_main(argc - 1, argv + 1) : mbias_main(argc - 1, argv + 1);
        if (getenv("MD_TIMING")) {      // where the wall clock went (stderr), for tuning
            mdh_run_stats st; mdh_last_run_stats(&st);
            double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            fprintf(stderr, "[md-timing] wall %.3f s: tiler (decode wait + assembly) %.3f, device calls %.3f, other %.3f on the calling thread; writer thread busy %.3f; %llu alignments in %llu tiles, %llu calls\n",
                    wall, st.t_decode_s, st.t_device_s, st.t_total_s - st.t_decode_s - st.t_device_s, st.t_format_s,
                    (unsigned long long) st.n_records, (unsigned long long) st.n_tiles, (unsigned long long) st.n_calls);
        }
        return rc;
    }
    if (!strcmp(argv[1], "perRead")) return perRead_main(argc - 1, argv + 1);
    if (!strcmp(argv[1], "mergeContext")) { fprintf(stderr, "The %s sub-command is not part of the B200 build.\n", argv[1]); return -1; }
    fprintf(stderr, "Unknown command!\n"); usage_main();
    return -1;
}
