/* TEST INFRASTRUCTURE — not product code.
 *
 * htslib-compatible shim, written from scratch for this repository so that the
 * reference's own C sources under /root/reference (common.c, overlaps.c,
 * extract.c, MBias.c, ...) can be compiled UNMODIFIED into oracle/_ref/ in a
 * container that has no htslib.  It restates, from the SAM/BAM/BAI/FAI format
 * specifications and htslib's documented API contract (htslib >= 1.11, which
 * is what MethylDackel.h:8-10 requires), only the entry points the reference
 * calls (SURVEY.md section 8c lists the call sites).  It is NOT htslib, shares
 * no code with it, and must never be linked into the product library.
 */
#ifndef MDSHIM_HTS_H
#define MDSHIM_HTS_H
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <limits.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HTS_VERSION 101700
#define MDSHIM 1

typedef int64_t hts_pos_t;
#define HTS_POS_MAX ((((int64_t)INT_MAX)<<32)|INT_MAX)
#define HTS_POS_MIN INT64_MIN
#define PRIhts_pos PRId64

#ifndef kroundup32
#define kroundup32(x) (--(x), (x)|=(x)>>1, (x)|=(x)>>2, (x)|=(x)>>4, (x)|=(x)>>8, (x)|=(x)>>16, ++(x))
#endif

struct mdshim_bgzf;
typedef struct htsFile {
    struct mdshim_bgzf *bgzf;
    char *fn;
    int is_write;
} htsFile;
typedef htsFile samFile;

struct mdshim_idx;
typedef struct mdshim_idx hts_idx_t;

typedef struct { uint64_t u, v; } hts_pair64_t;

typedef struct hts_itr_t {
    int tid;
    hts_pos_t beg, end;
    int n_off, i;
    hts_pair64_t *off;
    uint64_t curr_off;
    int finished;
} hts_itr_t;

const char *hts_version(void);
htsFile *hts_open(const char *fn, const char *mode);
int hts_close(htsFile *fp);
void hts_idx_destroy(hts_idx_t *idx);
void hts_itr_destroy(hts_itr_t *iter);
const char *hts_parse_reg(const char *str, int *beg, int *end);

#ifdef __cplusplus
}
#endif
#endif
