/* TEST INFRASTRUCTURE — htslib-compatible shim (see hts.h in this directory).
 * FASTA index subset used at extract.c:283,381, MBias.c:81,147,
 * common.c:477, mergeContext.c. Plain (uncompressed) FASTA only. */
#ifndef MDSHIM_FAIDX_H
#define MDSHIM_FAIDX_H
#include "hts.h"
#ifdef __cplusplus
extern "C" {
#endif
struct mdshim_faidx;
typedef struct mdshim_faidx faidx_t;

faidx_t *fai_load(const char *fn);
void fai_destroy(faidx_t *fai);
/* 0-based, end inclusive; returns malloc'd NUL-terminated bases, *len = count
 * (or -2 if the sequence is unknown, -1 on I/O error). */
char *faidx_fetch_seq(const faidx_t *fai, const char *c_name, int p_beg_i, int p_end_i, int *len);
int faidx_seq_len(const faidx_t *fai, const char *seq);
int faidx_nseq(const faidx_t *fai);
const char *faidx_iseq(const faidx_t *fai, int i);
#ifdef __cplusplus
}
#endif
#endif
