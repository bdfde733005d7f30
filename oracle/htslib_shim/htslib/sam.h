/* TEST INFRASTRUCTURE — htslib-compatible shim (see hts.h in this directory).
 * BAM record, header, index-iterator and pileup API subset used by the
 * reference (call sites: common.c:85,413-433; overlaps.c:27-147;
 * extract.c:283-295,379-399,494; MBias.c:145-218; perRead.c). */
#ifndef MDSHIM_SAM_H
#define MDSHIM_SAM_H
#include <stdint.h>
#include <stdlib.h>
#include "hts.h"

#ifdef __cplusplus
extern "C" {
#endif

#define BAM_CMATCH      0
#define BAM_CINS        1
#define BAM_CDEL        2
#define BAM_CREF_SKIP   3
#define BAM_CSOFT_CLIP  4
#define BAM_CHARD_CLIP  5
#define BAM_CPAD        6
#define BAM_CEQUAL      7
#define BAM_CDIFF       8
#define BAM_CBACK       9
#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK  0xf
#define BAM_CIGAR_TYPE  0x3C1A7

#define bam_cigar_op(c) ((c)&BAM_CIGAR_MASK)
#define bam_cigar_oplen(c) ((c)>>BAM_CIGAR_SHIFT)
#define bam_cigar_type(o) (BAM_CIGAR_TYPE>>((o)<<1)&3)

#define BAM_FPAIRED        1
#define BAM_FPROPER_PAIR   2
#define BAM_FUNMAP         4
#define BAM_FMUNMAP        8
#define BAM_FREVERSE      16
#define BAM_FMREVERSE     32
#define BAM_FREAD1        64
#define BAM_FREAD2       128
#define BAM_FSECONDARY   256
#define BAM_FQCFAIL      512
#define BAM_FDUP        1024
#define BAM_FSUPPLEMENTARY 2048

typedef struct sam_hdr_t {
    int32_t n_targets;
    uint32_t *target_len;
    char **target_name;
    size_t l_text;
    char *text;
} sam_hdr_t;
typedef sam_hdr_t bam_hdr_t;

typedef struct bam1_core_t {
    hts_pos_t pos;
    int32_t tid;
    uint16_t bin;
    uint8_t qual;
    uint8_t l_extranul;
    uint16_t flag;
    uint16_t l_qname;
    uint32_t n_cigar;
    int32_t l_qseq;
    int32_t mtid;
    hts_pos_t mpos;
    hts_pos_t isize;
} bam1_core_t;

typedef struct bam1_t {
    bam1_core_t core;
    uint64_t id;
    uint8_t *data;
    int l_data;
    uint32_t m_data;
} bam1_t;

#define bam_is_rev(b) (((b)->core.flag&BAM_FREVERSE) != 0)
#define bam_is_mrev(b) (((b)->core.flag&BAM_FMREVERSE) != 0)
#define bam_get_qname(b) ((char*)(b)->data)
#define bam_get_cigar(b) ((uint32_t*)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b)   ((b)->data + ((b)->core.n_cigar<<2) + (b)->core.l_qname)
#define bam_get_qual(b)  ((b)->data + ((b)->core.n_cigar<<2) + (b)->core.l_qname + (((b)->core.l_qseq + 1)>>1))
#define bam_get_aux(b)   ((b)->data + ((b)->core.n_cigar<<2) + (b)->core.l_qname + (((b)->core.l_qseq + 1)>>1) + (b)->core.l_qseq)
#define bam_get_l_aux(b) ((b)->l_data - ((b)->core.n_cigar<<2) - (b)->core.l_qname - (b)->core.l_qseq - (((b)->core.l_qseq + 1)>>1))
#define bam_seqi(s, i) ((s)[(i)>>1] >> ((~(i)&1)<<2) & 0xf)

bam1_t *bam_init1(void);
void bam_destroy1(bam1_t *b);
bam1_t *bam_copy1(bam1_t *bdst, const bam1_t *bsrc);
hts_pos_t bam_cigar2rlen(int n_cigar, const uint32_t *cigar);
hts_pos_t bam_endpos(const bam1_t *b);
uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]);
int64_t bam_aux2i(const uint8_t *s);

sam_hdr_t *sam_hdr_read(samFile *fp);
void sam_hdr_destroy(sam_hdr_t *h);
#define bam_hdr_destroy(h) sam_hdr_destroy(h)
int sam_hdr_name2tid(sam_hdr_t *h, const char *ref);
#define bam_name2id(h, ref) sam_hdr_name2tid((h), (ref))

int sam_read1(samFile *fp, sam_hdr_t *h, bam1_t *b);
hts_idx_t *sam_index_load(htsFile *fp, const char *fn);
int sam_index_build(const char *fn, int min_shift);
#define bam_index_build(fn, min_shift) (sam_index_build((fn), (min_shift)))
hts_itr_t *sam_itr_queryi(const hts_idx_t *idx, int tid, hts_pos_t beg, hts_pos_t end);
int sam_itr_next(htsFile *htsfp, hts_itr_t *itr, bam1_t *r);
#define sam_itr_destroy(iter) hts_itr_destroy(iter)

/* ---- pileup ---- */
typedef union { void *p; int64_t i; double f; } bam_pileup_cd;

typedef struct bam_pileup1_t {
    bam1_t *b;
    int32_t qpos;
    int indel, level;
    uint32_t is_del:1, is_head:1, is_tail:1, is_refskip:1, :1, aux:27;
    bam_pileup_cd cd;
    int cigar_ind;
} bam_pileup1_t;

typedef int (*bam_plp_auto_f)(void *data, bam1_t *b);

struct mdshim_plp;
typedef struct mdshim_plp *bam_plp_t;
struct mdshim_mplp;
typedef struct mdshim_mplp *bam_mplp_t;

bam_mplp_t bam_mplp_init(int n, bam_plp_auto_f func, void **data);
void bam_mplp_destroy(bam_mplp_t iter);
void bam_mplp_set_maxcnt(bam_mplp_t iter, int maxcnt);
int bam_mplp_auto(bam_mplp_t iter, int *_tid, int *_pos, int *n_plp, const bam_pileup1_t **plp);
int bam_mplp64_auto(bam_mplp_t iter, int *_tid, hts_pos_t *_pos, int *n_plp, const bam_pileup1_t **plp);
void bam_mplp_constructor(bam_mplp_t iter, int (*func)(void *data, const bam1_t *b, bam_pileup_cd *cd));
void bam_mplp_destructor(bam_mplp_t iter, int (*func)(void *data, const bam1_t *b, bam_pileup_cd *cd));

#ifdef __cplusplus
}
#endif
#endif
