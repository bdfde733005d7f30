/* TEST INFRASTRUCTURE — compile-only stand-in for libBigWig's bigWig.h.
 * The reference only touches libBigWig when -M/--mappability is given
 * (extract.c:1066-1233), which SURVEY.md section 2 marks out of scope;
 * bwOpen() here always fails, so that option reports "Couldn't open". */
#ifndef MDSHIM_BIGWIG_H
#define MDSHIM_BIGWIG_H
#include <stdint.h>
typedef struct { int64_t nKeys; char **chrom; uint32_t *len; } chromList_t;
typedef struct { chromList_t *cl; } bigWigFile_t;
typedef struct { uint32_t l, m; uint32_t *start, *end; float *value; } bwOverlappingIntervals_t;
static inline bigWigFile_t *bwOpen(char *fname, void *cb, const char *mode) { (void) fname; (void) cb; (void) mode; return 0; }
static inline void bwClose(bigWigFile_t *fp) { (void) fp; }
static inline bwOverlappingIntervals_t *bwGetValues(bigWigFile_t *fp, char *chrom, uint32_t start, uint32_t end, int includeNA) {
    (void) fp; (void) chrom; (void) start; (void) end; (void) includeNA; return 0; }
static inline void bwDestroyOverlappingIntervals(bwOverlappingIntervals_t *o) { (void) o; }
#endif
