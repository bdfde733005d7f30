// Host-side post-processing of the mbias histogram: inclusion-bound suggestions and the
// --txt table.  Follows svg.c:10-27 (CI), svg.c:240-296 (getThresholds), svg.c:423-425,435
// (the stderr suggestion line printed by makeSVGs) and svg.c:439-454 (makeTXT).
#pragma once
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cinttypes>
#include "../../../include/mdgpu.h"

namespace mdhost {

// View of one strand's counts inside the md_mbias_hist() layout.
struct StrandHist {
    const uint32_t *h; int strand; int l;
    uint32_t meth(int read, int i) const { return h[(((size_t) strand * 2 + (size_t)(read - 1)) * MD_MBIAS_MAXLEN + (size_t) i) * 2]; }
    uint32_t unmeth(int read, int i) const { return h[(((size_t) strand * 2 + (size_t)(read - 1)) * MD_MBIAS_MAXLEN + (size_t) i) * 2 + 1]; }
};

// Agresti-Coull 99.9 % interval, svg.c:10-27
inline double mbias_ci(uint32_t um, uint32_t m, int upper) {
    double X = (double) m, N = (double)(m + um), ZZ = 10.8275661707, Z = 3.2905267315;
    double N_dot = N + ZZ, P_dot = (1.0 / N_dot) * (X + 0.5 * ZZ), rv;
    if (upper) { rv = P_dot + Z * sqrt((P_dot / N_dot) * (1 - P_dot)); if (rv > 1.) rv = 1.0; }
    else { rv = P_dot - Z * sqrt((P_dot / N_dot) * (1 - P_dot)); if (rv < 0.) rv = 0.0; }
    return rv;
}

// svg.c:240-296.  Index i == l is read by the reference's middle-60 % loop when 0.8*l is integral;
// the histogram is zero there (the reference's arrays are zero-filled beyond l, MBias.c:32-37).
inline void mbias_thresholds(const StrandHist &s, int which, int *lthresh, int *rthresh) {
    int i, total = 0, middle = s.l / 2;
    double average = 0.0, minCI = 1.0, maxCI = 0.0, tmp, tmp2;
    for (i = (int)(0.2 * s.l); i <= (int)(0.8 * s.l); i++) {
        uint32_t m = s.meth(which, i), u = s.unmeth(which, i);
        if (m || u) {
            total++;
            average += ((double) m) / ((double)(m + u));
            tmp = mbias_ci(u, m, 1); if (minCI > tmp) minCI = tmp;
            tmp = mbias_ci(u, m, 0); if (maxCI < tmp) maxCI = tmp;
        }
    }
    if (total) average /= total;
    else { *lthresh = 0; *rthresh = 0; return; }
    for (i = middle; i >= 0; i--) {
        uint32_t m = s.meth(which, i), u = s.unmeth(which, i);
        if (m || u) {
            tmp = ((double) m) / ((double)(m + u));
            tmp2 = mbias_ci(u, m, 1);
            if (tmp2 < average && tmp < minCI && fabs(tmp - average) > 0.05) break;
            tmp2 = mbias_ci(u, m, 0);
            if (tmp2 > average && tmp > maxCI && fabs(tmp - average) > 0.05) break;
        }
    }
    *lthresh = (i >= 0) ? i + 2 : 0;
    for (i = middle + 1; i < s.l; i++) {
        uint32_t m = s.meth(which, i), u = s.unmeth(which, i);
        if (m || u) {
            tmp = ((double) m) / ((double)(m + u));
            tmp2 = mbias_ci(u, m, 1);
            if (tmp2 < average && tmp < minCI && fabs(tmp - average) > 0.05) break;
            tmp2 = mbias_ci(u, m, 0);
            if (tmp2 > average && tmp > maxCI && fabs(tmp - average) > 0.05) break;
        }
    }
    *rthresh = (i < s.l) ? i : 0;
}

// The "Suggested inclusion options:" line makeSVGs prints to stderr (svg.c:423-425,435)
inline void mbias_print_suggestions(FILE *err, const uint32_t *hist, const int32_t lens[4]) {
    static const char *abbrevs[4] = {"OT", "OB", "CTOT", "CTOB"};
    bool printing = false;
    for (int i = 0; i < 4; ++i) {
        if (!lens[i]) continue;
        StrandHist s{hist, i, lens[i]};
        int l1, r1, l2, r2;
        mbias_thresholds(s, 1, &l1, &r1);
        mbias_thresholds(s, 2, &l2, &r2);
        if (!printing) fprintf(err, "Suggested inclusion options:");
        fprintf(err, " --%s %i,%i,%i,%i", abbrevs[i], l1, r1, l2, r2);
        printing = true;
    }
    if (printing) fprintf(err, "\n");
}

// makeTXT, svg.c:439-454
inline void mbias_print_txt(FILE *out, const uint32_t *hist, const int32_t lens[4]) {
    static const char *abbrevs[4] = {"OT", "OB", "CTOT", "CTOB"};
    fprintf(out, "Strand\tRead\tPosition\tnMethylated\tnUnmethylated\n");
    for (int i = 0; i < 4; ++i) {
        if (!lens[i]) continue;
        StrandHist s{hist, i, lens[i]};
        for (int j = 0; j < lens[i]; ++j) {
            if (s.meth(1, j) || s.unmeth(1, j)) fprintf(out, "%s\t1\t%i\t%" PRIu32 "\t%" PRIu32 "\n", abbrevs[i], j + 1, s.meth(1, j), s.unmeth(1, j));
            if (s.meth(2, j) || s.unmeth(2, j)) fprintf(out, "%s\t2\t%i\t%" PRIu32 "\t%" PRIu32 "\n", abbrevs[i], j + 1, s.meth(2, j), s.unmeth(2, j));
        }
    }
}

}  // namespace mdhost
