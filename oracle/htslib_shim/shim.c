/* TEST INFRASTRUCTURE — not product code.
 *
 * Implementation of the htslib-compatible shim declared in the htslib/ headers (this
 * directory).  Written from scratch against the public format specifications
 * (SAM/BAM v1 sections 4.1-4.2 for BGZF+BAM records, section 5 for the BAI
 * index; the samtools `faidx` five-column .fai format) and htslib's documented
 * API contract, so that the reference's own sources compile unmodified into
 * oracle/_ref/.  htslib itself is NOT in /root/reference (no submodule,
 * SURVEY.md section 8c) and not in this image, hence the restatement.
 *
 * The pileup engine (bam_mplp_*) restates htslib's documented behaviour:
 *   - records come from the user callback (common.c:407 filter_func);
 *   - a record with tid<0 or FUNMAP is dropped at push;
 *   - the constructor callback (overlaps.c:121) fires at push time, on the
 *     engine's own copy of the record, for every record whose reference end is
 *     beyond the current column;
 *   - a column p is emitted only once a record starting beyond p (or EOF) has
 *     been seen, columns ascend, and for each buffered record covering p the
 *     entry carries qpos (index of the aligned query base) or is_del /
 *     is_refskip inside D / N operations;
 *   - the destructor callback (overlaps.c:141) fires when a record whose end
 *     is <= the column being built is dropped from the buffer.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include <ctype.h>
#include <assert.h>
#include <zlib.h>
#include "htslib/hts.h"
#include "htslib/sam.h"
#include "htslib/faidx.h"

/* =====================================================================
 * BGZF reader
 * ===================================================================== */
#define BGZF_MAX_BLOCK 65536

typedef struct mdshim_bgzf {
    FILE *fp;
    uint8_t *cbuf;      /* compressed block */
    uint8_t *ubuf;      /* inflated block */
    int ulen, uoff;     /* inflated length / read cursor */
    int64_t block_addr; /* file offset of the block in ubuf */
    int64_t next_addr;  /* file offset of the following block */
    int at_eof;
    z_stream zs;
    int zs_init;
} mdshim_bgzf;

static mdshim_bgzf *bgzf_open_r(const char *fn) {
    FILE *fp = fopen(fn, "rb");
    if (!fp) return NULL;
    mdshim_bgzf *z = (mdshim_bgzf *) calloc(1, sizeof(*z));
    z->fp = fp;
    z->cbuf = (uint8_t *) malloc(BGZF_MAX_BLOCK);
    z->ubuf = (uint8_t *) malloc(BGZF_MAX_BLOCK);
    setvbuf(fp, NULL, _IOFBF, 1 << 20);
    return z;
}

static void bgzf_close_r(mdshim_bgzf *z) {
    if (!z) return;
    if (z->zs_init) inflateEnd(&z->zs);
    if (z->fp) fclose(z->fp);
    free(z->cbuf); free(z->ubuf); free(z);
}

/* Load the block starting at file offset addr. 0 ok, 1 EOF, -1 error */
static int bgzf_load_block(mdshim_bgzf *z, int64_t addr) {
    uint8_t hdr[12];
    if (fseeko(z->fp, addr, SEEK_SET) != 0) return -1;
    size_t got = fread(hdr, 1, 12, z->fp);
    if (got == 0) { z->at_eof = 1; z->ulen = z->uoff = 0; z->block_addr = addr; z->next_addr = addr; return 1; }
    if (got != 12 || hdr[0] != 31 || hdr[1] != 139 || hdr[2] != 8 || !(hdr[3] & 4)) return -1;
    int xlen = hdr[10] | (hdr[11] << 8);
    if (xlen > 1024) return -1;
    uint8_t extra[1024];
    if (fread(extra, 1, (size_t) xlen, z->fp) != (size_t) xlen) return -1;
    int bsize = -1, off = 0;
    while (off + 4 <= xlen) {
        int slen = extra[off + 2] | (extra[off + 3] << 8);
        if (extra[off] == 'B' && extra[off + 1] == 'C' && slen == 2) bsize = extra[off + 4] | (extra[off + 5] << 8);
        off += 4 + slen;
    }
    if (bsize < 0) return -1;
    int clen = bsize + 1 - 12 - xlen; /* deflate payload + 8-byte trailer */
    if (clen < 8 || clen > BGZF_MAX_BLOCK) return -1;
    if (fread(z->cbuf, 1, (size_t) clen, z->fp) != (size_t) clen) return -1;
    uint32_t isize = (uint32_t) z->cbuf[clen - 4] | ((uint32_t) z->cbuf[clen - 3] << 8) |
                     ((uint32_t) z->cbuf[clen - 2] << 16) | ((uint32_t) z->cbuf[clen - 1] << 24);
    if (isize > BGZF_MAX_BLOCK) return -1;
    if (!z->zs_init) {
        memset(&z->zs, 0, sizeof(z->zs));
        if (inflateInit2(&z->zs, -15) != Z_OK) return -1;
        z->zs_init = 1;
    } else inflateReset(&z->zs);
    z->zs.next_in = z->cbuf; z->zs.avail_in = (uInt)(clen - 8);
    z->zs.next_out = z->ubuf; z->zs.avail_out = BGZF_MAX_BLOCK;
    int rc = inflate(&z->zs, Z_FINISH);
    if (rc != Z_STREAM_END || z->zs.total_out != isize) return -1;
    z->ulen = (int) isize; z->uoff = 0;
    z->block_addr = addr; z->next_addr = addr + bsize + 1;
    return 0;
}

static int64_t bgzf_tell_r(mdshim_bgzf *z) { return (z->block_addr << 16) | (int64_t)(z->uoff & 0xFFFF); }

static int bgzf_seek_r(mdshim_bgzf *z, uint64_t voff) {
    int64_t addr = (int64_t)(voff >> 16);
    int within = (int)(voff & 0xFFFF);
    z->at_eof = 0;
    if (addr != z->block_addr || z->ulen == 0) {
        int rc = bgzf_load_block(z, addr);
        if (rc < 0) return -1;
    }
    if (within > z->ulen) return -1;
    z->uoff = within;
    return 0;
}

/* Returns bytes read (may be short at EOF) or -1 */
static ssize_t bgzf_read_r(mdshim_bgzf *z, void *dst_, size_t n) {
    uint8_t *dst = (uint8_t *) dst_;
    size_t done = 0;
    while (done < n) {
        if (z->uoff >= z->ulen) {
            if (z->at_eof) break;
            int rc = bgzf_load_block(z, z->ulen == 0 && z->next_addr == 0 ? z->block_addr : z->next_addr);
            if (rc < 0) return -1;
            if (rc == 1) break;
            if (z->ulen == 0) continue; /* empty (EOF marker) block */
        }
        size_t take = (size_t)(z->ulen - z->uoff);
        if (take > n - done) take = n - done;
        memcpy(dst + done, z->ubuf + z->uoff, take);
        z->uoff += (int) take; done += take;
        if (z->uoff == z->ulen) { /* report the next block as the position, like a virtual offset should */
            z->block_addr = z->next_addr; z->ulen = 0; z->uoff = 0;
        }
    }
    return (ssize_t) done;
}

/* =====================================================================
 * hts file handle
 * ===================================================================== */
const char *hts_version(void) { return "shim-1.17-api (oracle/htslib_shim, not htslib)"; }

htsFile *hts_open(const char *fn, const char *mode) {
    if (!mode || mode[0] != 'r') { fprintf(stderr, "[shim] hts_open: only read mode is supported\n"); return NULL; }
    mdshim_bgzf *z = bgzf_open_r(fn);
    if (!z) return NULL;
    /* sanity: must be BGZF */
    if (bgzf_load_block(z, 0) != 0) { bgzf_close_r(z); fprintf(stderr, "[shim] %s is not a BGZF/BAM file (CRAM/SAM unsupported by the shim)\n", fn); return NULL; }
    htsFile *fp = (htsFile *) calloc(1, sizeof(*fp));
    fp->bgzf = z; fp->fn = strdup(fn);
    return fp;
}

int hts_close(htsFile *fp) {
    if (!fp) return 0;
    bgzf_close_r(fp->bgzf); free(fp->fn); free(fp);
    return 0;
}

static int64_t parse_decimal_commas(const char *s, const char **endp) {
    int64_t v = 0; int neg = 0;
    while (isspace((unsigned char) *s)) ++s;
    if (*s == '+') ++s; else if (*s == '-') { neg = 1; ++s; }
    while (isdigit((unsigned char) *s) || *s == ',') { if (*s != ',') v = v * 10 + (*s - '0'); ++s; }
    *endp = s;
    return neg ? -v : v;
}

const char *hts_parse_reg(const char *s, int *beg, int *end) {
    const char *colon = strrchr(s, ':');
    if (!colon) { *beg = 0; *end = INT_MAX; return s + strlen(s); }
    const char *p;
    int64_t b = parse_decimal_commas(colon + 1, &p) - 1, e;
    if (b < 0) {
        if (b != -1 && *p == '-' && colon[1] != '\0') { fprintf(stderr, "[shim] Coordinates must be > 0\n"); return NULL; }
        if (isdigit((unsigned char) *p) || *p == '\0' || *p == ',') {
            e = (b == -1) ? INT_MAX : -(b + 1);
            *beg = 0; *end = (int) e;
            return colon;
        } else if (b < -1) return NULL;
    }
    if (*p == '\0' || *p == ',') e = INT_MAX;
    else if (*p == '-') {
        const char *q;
        e = parse_decimal_commas(p + 1, &q);
        if (*q != '\0' && *q != ',') return NULL;
    } else return NULL;
    if (b >= e) return NULL;
    if (b > INT_MAX) return NULL;
    if (e > INT_MAX) e = INT_MAX;
    *beg = (int) b; *end = (int) e;
    return colon;
}

/* =====================================================================
 * BAM header + records
 * ===================================================================== */
static int rd_i32(mdshim_bgzf *z, int32_t *v) {
    uint8_t b[4];
    if (bgzf_read_r(z, b, 4) != 4) return -1;
    *v = (int32_t)((uint32_t) b[0] | ((uint32_t) b[1] << 8) | ((uint32_t) b[2] << 16) | ((uint32_t) b[3] << 24));
    return 0;
}

sam_hdr_t *sam_hdr_read(samFile *fp) {
    mdshim_bgzf *z = fp->bgzf;
    if (bgzf_seek_r(z, 0) < 0) return NULL;
    char magic[4];
    if (bgzf_read_r(z, magic, 4) != 4 || memcmp(magic, "BAM\1", 4) != 0) return NULL;
    int32_t l_text, n_ref;
    if (rd_i32(z, &l_text) < 0 || l_text < 0) return NULL;
    sam_hdr_t *h = (sam_hdr_t *) calloc(1, sizeof(*h));
    h->l_text = (size_t) l_text;
    h->text = (char *) malloc((size_t) l_text + 1);
    if (bgzf_read_r(z, h->text, (size_t) l_text) != l_text) { sam_hdr_destroy(h); return NULL; }
    h->text[l_text] = 0;
    if (rd_i32(z, &n_ref) < 0 || n_ref < 0) { sam_hdr_destroy(h); return NULL; }
    h->n_targets = n_ref;
    h->target_name = (char **) calloc((size_t) n_ref ? (size_t) n_ref : 1, sizeof(char *));
    h->target_len = (uint32_t *) calloc((size_t) n_ref ? (size_t) n_ref : 1, sizeof(uint32_t));
    for (int i = 0; i < n_ref; ++i) {
        int32_t l_name, l_ref;
        if (rd_i32(z, &l_name) < 0 || l_name <= 0) { sam_hdr_destroy(h); return NULL; }
        h->target_name[i] = (char *) malloc((size_t) l_name + 1);
        if (bgzf_read_r(z, h->target_name[i], (size_t) l_name) != l_name) { sam_hdr_destroy(h); return NULL; }
        h->target_name[i][l_name] = 0;
        if (rd_i32(z, &l_ref) < 0) { sam_hdr_destroy(h); return NULL; }
        h->target_len[i] = (uint32_t) l_ref;
    }
    return h;
}

void sam_hdr_destroy(sam_hdr_t *h) {
    if (!h) return;
    if (h->target_name) for (int i = 0; i < h->n_targets; ++i) free(h->target_name[i]);
    free(h->target_name); free(h->target_len); free(h->text); free(h);
}

int sam_hdr_name2tid(sam_hdr_t *h, const char *ref) {
    for (int i = 0; i < h->n_targets; ++i) if (strcmp(h->target_name[i], ref) == 0) return i;
    return -1;
}

bam1_t *bam_init1(void) { return (bam1_t *) calloc(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t *b) { if (b) { free(b->data); free(b); } }

static int bam_reserve(bam1_t *b, size_t n) {
    if (b->m_data < n) {
        size_t m = n; --m; m |= m >> 1; m |= m >> 2; m |= m >> 4; m |= m >> 8; m |= m >> 16; ++m;
        uint8_t *d = (uint8_t *) realloc(b->data, m);
        if (!d) return -1;
        b->data = d; b->m_data = (uint32_t) m;
    }
    return 0;
}

bam1_t *bam_copy1(bam1_t *dst, const bam1_t *src) {
    if (bam_reserve(dst, (size_t) src->l_data) < 0) return NULL;
    memcpy(dst->data, src->data, (size_t) src->l_data);
    dst->l_data = src->l_data;
    dst->core = src->core;
    dst->id = src->id;
    return dst;
}

/* returns >=0 on success, -1 on EOF, < -1 on error */
static int bam_read1_r(mdshim_bgzf *z, bam1_t *b) {
    int32_t block_size;
    uint8_t x[32];
    ssize_t r = bgzf_read_r(z, x, 4);
    if (r == 0) return -1;
    if (r != 4) return -2;
    block_size = (int32_t)((uint32_t) x[0] | ((uint32_t) x[1] << 8) | ((uint32_t) x[2] << 16) | ((uint32_t) x[3] << 24));
    if (block_size < 32) return -4;
    if (bgzf_read_r(z, x, 32) != 32) return -3;
#define LE32(p) ((uint32_t)(p)[0] | ((uint32_t)(p)[1] << 8) | ((uint32_t)(p)[2] << 16) | ((uint32_t)(p)[3] << 24))
    bam1_core_t *c = &b->core;
    c->tid = (int32_t) LE32(x);
    c->pos = (int32_t) LE32(x + 4);
    uint32_t bmq = LE32(x + 8), fnc = LE32(x + 12);
    c->bin = (uint16_t)(bmq >> 16); c->qual = (uint8_t)(bmq >> 8 & 0xff); c->l_qname = (uint16_t)(bmq & 0xff);
    c->l_extranul = 0;
    c->flag = (uint16_t)(fnc >> 16); c->n_cigar = fnc & 0xffff;
    c->l_qseq = (int32_t) LE32(x + 16);
    c->mtid = (int32_t) LE32(x + 20);
    c->mpos = (int32_t) LE32(x + 24);
    c->isize = (int32_t) LE32(x + 28);
#undef LE32
    size_t l_data = (size_t) block_size - 32;
    if (bam_reserve(b, l_data ? l_data : 1) < 0) return -4;
    if (bgzf_read_r(z, b->data, l_data) != (ssize_t) l_data) return -4;
    b->l_data = (int) l_data;
    return 4 + block_size;
}

int sam_read1(samFile *fp, sam_hdr_t *h, bam1_t *b) { (void) h; return bam_read1_r(fp->bgzf, b); }

hts_pos_t bam_cigar2rlen(int n_cigar, const uint32_t *cigar) {
    hts_pos_t l = 0;
    for (int k = 0; k < n_cigar; ++k) if (bam_cigar_type(bam_cigar_op(cigar[k])) & 2) l += bam_cigar_oplen(cigar[k]);
    return l;
}

hts_pos_t bam_endpos(const bam1_t *b) {
    hts_pos_t rlen = (b->core.flag & BAM_FUNMAP) ? 0 : bam_cigar2rlen((int) b->core.n_cigar, bam_get_cigar(b));
    if (rlen == 0) rlen = 1;
    return b->core.pos + rlen;
}

static int aux_type_size(int t) {
    switch (t) {
        case 'A': case 'c': case 'C': return 1;
        case 's': case 'S': return 2;
        case 'i': case 'I': case 'f': return 4;
        case 'd': return 8;
        default: return 0;
    }
}

/* returns pointer to the TYPE byte of the matching field (so p+1 is the value) */
uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]) {
    uint8_t *s = bam_get_aux(b), *end = b->data + b->l_data;
    while (s + 3 <= end) {
        int match = (s[0] == (uint8_t) tag[0] && s[1] == (uint8_t) tag[1]);
        uint8_t *tp = s + 2;
        int t = *tp;
        uint8_t *v = tp + 1;
        if (t == 'Z' || t == 'H') {
            uint8_t *q = v;
            while (q < end && *q) ++q;
            if (q >= end) return NULL;
            if (match) return tp;
            s = q + 1;
        } else if (t == 'B') {
            if (v + 5 > end) return NULL;
            int sz = aux_type_size(v[0]);
            uint32_t n = (uint32_t) v[1] | ((uint32_t) v[2] << 8) | ((uint32_t) v[3] << 16) | ((uint32_t) v[4] << 24);
            if (sz == 0) return NULL;
            if (match) return tp;
            s = v + 5 + (size_t) sz * n;
        } else {
            int sz = aux_type_size(t);
            if (sz == 0 || v + sz > end) return NULL;
            if (match) return tp;
            s = v + sz;
        }
    }
    return NULL;
}

int64_t bam_aux2i(const uint8_t *s) {
    int t = *s++;
    switch (t) {
        case 'c': return (int8_t) s[0];
        case 'C': return s[0];
        case 's': return (int16_t)(s[0] | (s[1] << 8));
        case 'S': return (uint16_t)(s[0] | (s[1] << 8));
        case 'i': return (int32_t)((uint32_t) s[0] | ((uint32_t) s[1] << 8) | ((uint32_t) s[2] << 16) | ((uint32_t) s[3] << 24));
        case 'I': return (uint32_t)((uint32_t) s[0] | ((uint32_t) s[1] << 8) | ((uint32_t) s[2] << 16) | ((uint32_t) s[3] << 24));
        default: errno = EINVAL; return 0;
    }
}

/* =====================================================================
 * BAI index + region iterator
 * ===================================================================== */
typedef struct { uint32_t bin; int n; hts_pair64_t *chunks; } bai_bin_t;
typedef struct { int n_bin; bai_bin_t *bins; int n_intv; uint64_t *ioff; } bai_ref_t;
struct mdshim_idx { int n_ref; bai_ref_t *refs; };

static int fr_u32(FILE *f, uint32_t *v) { uint8_t b[4]; if (fread(b, 1, 4, f) != 4) return -1; *v = (uint32_t) b[0] | ((uint32_t) b[1] << 8) | ((uint32_t) b[2] << 16) | ((uint32_t) b[3] << 24); return 0; }
static int fr_u64(FILE *f, uint64_t *v) { uint32_t lo, hi; if (fr_u32(f, &lo) < 0 || fr_u32(f, &hi) < 0) return -1; *v = ((uint64_t) hi << 32) | lo; return 0; }

static hts_idx_t *bai_load(const char *fnidx) {
    FILE *f = fopen(fnidx, "rb");
    if (!f) return NULL;
    char magic[4];
    uint32_t n_ref;
    if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "BAI\1", 4) != 0 || fr_u32(f, &n_ref) < 0) { fclose(f); return NULL; }
    hts_idx_t *idx = (hts_idx_t *) calloc(1, sizeof(*idx));
    idx->n_ref = (int) n_ref;
    idx->refs = (bai_ref_t *) calloc(n_ref ? n_ref : 1, sizeof(bai_ref_t));
    for (uint32_t r = 0; r < n_ref; ++r) {
        bai_ref_t *R = &idx->refs[r];
        uint32_t n_bin, n_intv;
        if (fr_u32(f, &n_bin) < 0) goto fail;
        R->n_bin = (int) n_bin;
        R->bins = (bai_bin_t *) calloc(n_bin ? n_bin : 1, sizeof(bai_bin_t));
        for (uint32_t k = 0; k < n_bin; ++k) {
            uint32_t bin, n_chunk;
            if (fr_u32(f, &bin) < 0 || fr_u32(f, &n_chunk) < 0) goto fail;
            R->bins[k].bin = bin; R->bins[k].n = (int) n_chunk;
            R->bins[k].chunks = (hts_pair64_t *) calloc(n_chunk ? n_chunk : 1, sizeof(hts_pair64_t));
            for (uint32_t c = 0; c < n_chunk; ++c)
                if (fr_u64(f, &R->bins[k].chunks[c].u) < 0 || fr_u64(f, &R->bins[k].chunks[c].v) < 0) goto fail;
        }
        if (fr_u32(f, &n_intv) < 0) goto fail;
        R->n_intv = (int) n_intv;
        R->ioff = (uint64_t *) calloc(n_intv ? n_intv : 1, sizeof(uint64_t));
        for (uint32_t k = 0; k < n_intv; ++k) if (fr_u64(f, &R->ioff[k]) < 0) goto fail;
    }
    fclose(f);
    return idx;
fail:
    fclose(f);
    hts_idx_destroy(idx);
    return NULL;
}

void hts_idx_destroy(hts_idx_t *idx) {
    if (!idx) return;
    for (int r = 0; r < idx->n_ref; ++r) {
        bai_ref_t *R = &idx->refs[r];
        if (R->bins) for (int k = 0; k < R->n_bin; ++k) free(R->bins[k].chunks);
        free(R->bins); free(R->ioff);
    }
    free(idx->refs); free(idx);
}

hts_idx_t *sam_index_load(htsFile *fp, const char *fn) {
    (void) fp;
    size_t n = strlen(fn);
    char *p = (char *) malloc(n + 8);
    sprintf(p, "%s.bai", fn);
    hts_idx_t *idx = bai_load(p);
    if (!idx && n > 4 && strcmp(fn + n - 4, ".bam") == 0) {
        strcpy(p, fn); strcpy(p + n - 4, ".bai");
        idx = bai_load(p);
    }
    free(p);
    return idx;
}

int sam_index_build(const char *fn, int min_shift) {
    (void) min_shift;
    fprintf(stderr, "[shim] building an index for %s is not supported by the oracle shim; supply a .bai\n", fn);
    return -1;
}

static int cmp_pair(const void *a, const void *b) {
    const hts_pair64_t *x = (const hts_pair64_t *) a, *y = (const hts_pair64_t *) b;
    return (x->u > y->u) - (x->u < y->u);
}

hts_itr_t *sam_itr_queryi(const hts_idx_t *idx, int tid, hts_pos_t beg, hts_pos_t end) {
    hts_itr_t *it = (hts_itr_t *) calloc(1, sizeof(*it));
    it->tid = tid; it->beg = beg; it->end = end; it->i = -1;
    if (tid < 0 || tid >= idx->n_ref || beg >= end) { it->finished = 1; return it; }
    if (beg < 0) beg = 0;
    const bai_ref_t *R = &idx->refs[tid];
    /* lower bound on file offset from the 16 kb linear index */
    uint64_t min_off = 0;
    if (R->n_intv > 0) {
        hts_pos_t w = beg >> 14;
        min_off = (w < R->n_intv) ? R->ioff[w] : R->ioff[R->n_intv - 1];
    }
    /* candidate bins (BAI scheme: min_shift 14, 5 levels) */
    hts_pos_t e = end - 1;
    if (e >= (1LL << 29)) e = (1LL << 29) - 1;
    int cap = 0, n = 0;
    hts_pair64_t *off = NULL;
    static const int lvl_first[6] = {0, 1, 9, 73, 585, 4681};
    for (int k = 0; k < R->n_bin; ++k) {
        uint32_t bin = R->bins[k].bin;
        if (bin >= 37449) continue; /* 37450 is the metadata pseudo-bin */
        int l = 5;
        while (l > 0 && bin < (uint32_t) lvl_first[l]) --l;
        int shift = 14 + 3 * (5 - l);
        hts_pos_t b0 = ((hts_pos_t)(bin - (uint32_t) lvl_first[l])) << shift, b1 = b0 + (1LL << shift);
        if (b1 <= beg || b0 > e) continue;
        for (int c = 0; c < R->bins[k].n; ++c) {
            if (R->bins[k].chunks[c].v <= min_off) continue;
            if (n == cap) { cap = cap ? cap * 2 : 16; off = (hts_pair64_t *) realloc(off, (size_t) cap * sizeof(*off)); }
            off[n++] = R->bins[k].chunks[c];
        }
    }
    if (n == 0) { free(off); it->finished = 1; return it; }
    qsort(off, (size_t) n, sizeof(*off), cmp_pair);
    /* merge overlapping / abutting chunks */
    int m = 0;
    for (int k = 1; k < n; ++k) {
        if (off[k].u <= off[m].v) { if (off[k].v > off[m].v) off[m].v = off[k].v; }
        else off[++m] = off[k];
    }
    it->n_off = m + 1; it->off = off;
    return it;
}

void hts_itr_destroy(hts_itr_t *it) { if (it) { free(it->off); free(it); } }

int sam_itr_next(htsFile *fp, hts_itr_t *it, bam1_t *b) {
    mdshim_bgzf *z = fp->bgzf;
    if (!it || it->finished) return -1;
    for (;;) {
        if (it->curr_off == 0 || it->curr_off >= it->off[it->i].v) {
            if (it->i == it->n_off - 1) { it->finished = 1; return -1; }
            if (it->i < 0 || it->off[it->i].v != it->off[it->i + 1].u) {
                if (bgzf_seek_r(z, it->off[it->i + 1].u) < 0) { it->finished = 1; return -2; }
                it->curr_off = (uint64_t) bgzf_tell_r(z);
            }
            ++it->i;
        }
        int ret = bam_read1_r(z, b);
        if (ret < 0) { it->finished = 1; return ret; }
        it->curr_off = (uint64_t) bgzf_tell_r(z);
        if (b->core.tid != it->tid || b->core.pos >= it->end) { it->finished = 1; return -1; }
        if (bam_endpos(b) > it->beg) return ret;
    }
}

/* =====================================================================
 * faidx (plain FASTA + .fai)
 * ===================================================================== */
typedef struct { char *name; int64_t len, offset; int line_blen, line_len; } fai_rec_t;
struct mdshim_faidx { FILE *fp; int n; fai_rec_t *recs; };

static int fai_build_in_memory(const char *fn, faidx_t *fai) {
    FILE *f = fopen(fn, "rb");
    if (!f) return -1;
    int c, cap = 0;
    int64_t off = 0;
    fai_rec_t *cur = NULL;
    int64_t line_b = 0, line_l = 0;
    int in_name = 0, name_done = 0, first_line = 1;
    size_t nl = 0, ncap = 0;
    char *nm = NULL;
    while ((c = fgetc(f)) != EOF) {
        ++off;
        if (in_name) {
            if (c == '\n') {
                in_name = 0;
                if (fai->n == cap) { cap = cap ? cap * 2 : 8; fai->recs = (fai_rec_t *) realloc(fai->recs, (size_t) cap * sizeof(fai_rec_t)); }
                cur = &fai->recs[fai->n++];
                memset(cur, 0, sizeof(*cur));
                nm = (char *) realloc(nm, nl + 1); nm[nl] = 0;
                cur->name = strdup(nm);
                cur->offset = off;
                first_line = 1; line_b = line_l = 0;
            } else if (!name_done) {
                if (isspace(c)) name_done = 1;
                else { if (nl + 1 >= ncap) { ncap = ncap ? ncap * 2 : 64; nm = (char *) realloc(nm, ncap); } nm[nl++] = (char) c; }
            }
            continue;
        }
        if (c == '>' && line_l == 0) { in_name = 1; name_done = 0; nl = 0; continue; }
        if (!cur) continue;
        ++line_l;
        if (c == '\n') {
            if (first_line && line_b > 0) { cur->line_blen = (int) line_b; cur->line_len = (int) line_l; first_line = 0; }
            line_b = line_l = 0;
        } else if (c != '\r') { ++line_b; ++cur->len; }
    }
    if (cur && first_line && line_b > 0) { cur->line_blen = (int) line_b; cur->line_len = (int) line_b + 1; }
    free(nm);
    fclose(f);
    return 0;
}

faidx_t *fai_load(const char *fn) {
    faidx_t *fai = (faidx_t *) calloc(1, sizeof(*fai));
    size_t n = strlen(fn);
    char *p = (char *) malloc(n + 8);
    sprintf(p, "%s.fai", fn);
    FILE *fi = fopen(p, "r");
    if (fi) {
        char line[4096];
        int cap = 0;
        while (fgets(line, sizeof line, fi)) {
            char name[2048]; long long len, off; int lb, ll;
            if (sscanf(line, "%2047[^\t]\t%lld\t%lld\t%d\t%d", name, &len, &off, &lb, &ll) != 5) continue;
            if (fai->n == cap) { cap = cap ? cap * 2 : 8; fai->recs = (fai_rec_t *) realloc(fai->recs, (size_t) cap * sizeof(fai_rec_t)); }
            fai_rec_t *r = &fai->recs[fai->n++];
            r->name = strdup(name); r->len = len; r->offset = off; r->line_blen = lb; r->line_len = ll;
        }
        fclose(fi);
    } else {
        if (fai_build_in_memory(fn, fai) < 0) { free(p); free(fai); return NULL; }
        FILE *fo = fopen(p, "w"); /* best effort, like `samtools faidx`; read-only dirs are fine */
        if (fo) {
            for (int i = 0; i < fai->n; ++i)
                fprintf(fo, "%s\t%lld\t%lld\t%d\t%d\n", fai->recs[i].name, (long long) fai->recs[i].len, (long long) fai->recs[i].offset, fai->recs[i].line_blen, fai->recs[i].line_len);
            fclose(fo);
        }
    }
    free(p);
    fai->fp = fopen(fn, "rb");
    if (!fai->fp) { fai_destroy(fai); return NULL; }
    return fai;
}

void fai_destroy(faidx_t *fai) {
    if (!fai) return;
    for (int i = 0; i < fai->n; ++i) free(fai->recs[i].name);
    free(fai->recs);
    if (fai->fp) fclose(fai->fp);
    free(fai);
}

static const fai_rec_t *fai_find(const faidx_t *fai, const char *name) {
    for (int i = 0; i < fai->n; ++i) if (strcmp(fai->recs[i].name, name) == 0) return &fai->recs[i];
    return NULL;
}

int faidx_seq_len(const faidx_t *fai, const char *seq) { const fai_rec_t *r = fai_find(fai, seq); return r ? (int) r->len : -1; }
int faidx_nseq(const faidx_t *fai) { return fai->n; }
const char *faidx_iseq(const faidx_t *fai, int i) { return fai->recs[i].name; }

char *faidx_fetch_seq(const faidx_t *fai, const char *c_name, int p_beg_i, int p_end_i, int *len) {
    const fai_rec_t *r = fai_find(fai, c_name);
    if (!r) { *len = -2; fprintf(stderr, "[shim] The sequence \"%s\" was not found\n", c_name); return NULL; }
    int64_t beg = p_beg_i, end = p_end_i;
    if (end < beg) beg = end;
    if (beg < 0) beg = 0; else if (r->len <= beg) beg = r->len;
    if (end < 0) end = 0; else if (r->len <= end) end = r->len - 1;
    int64_t e1 = end + 1; /* exclusive */
    if (e1 < beg) e1 = beg;
    int64_t want = e1 - beg;
    char *s = (char *) malloc((size_t) want + 1);
    if (!s) { *len = -1; return NULL; }
    int64_t got = 0;
    if (want > 0) {
        int64_t foff = r->offset + beg / r->line_blen * r->line_len + beg % r->line_blen;
        if (fseeko(fai->fp, foff, SEEK_SET) != 0) { free(s); *len = -1; return NULL; }
        int c;
        while (got < want && (c = fgetc(fai->fp)) != EOF) if (isgraph(c)) s[got++] = (char) c;
    }
    s[got] = 0;
    *len = (int) got;
    return s;
}

/* =====================================================================
 * pileup
 * ===================================================================== */
typedef struct { int k; hts_pos_t x, y, end; } cstate_t; /* op index, its ref start, its query start, last ref pos */

typedef struct lbnode {
    bam1_t b;
    hts_pos_t beg, end;
    cstate_t s;
    bam_pileup_cd cd;
    struct lbnode *next;
} lbnode_t;

struct mdshim_plp {
    lbnode_t *head, *tail, *free_list;
    int is_eof, error, maxcnt, cnt;
    int32_t tid, max_tid;
    hts_pos_t pos, max_pos;
    bam_pileup1_t *plp; int max_plp;
    bam1_t *b;
    bam_plp_auto_f func; void *data;
    int (*ctor)(void *, const bam1_t *, bam_pileup_cd *);
    int (*dtor)(void *, const bam1_t *, bam_pileup_cd *);
};

static lbnode_t *node_alloc(struct mdshim_plp *it) {
    lbnode_t *p;
    ++it->cnt;
    if (it->free_list) { p = it->free_list; it->free_list = p->next; p->next = NULL; return p; }
    return (lbnode_t *) calloc(1, sizeof(lbnode_t));
}
static void node_free(struct mdshim_plp *it, lbnode_t *p) {
    --it->cnt;
    p->next = it->free_list; it->free_list = p; /* keeps p->b.data for reuse */
}

static struct mdshim_plp *plp_init(bam_plp_auto_f func, void *data) {
    struct mdshim_plp *it = (struct mdshim_plp *) calloc(1, sizeof(*it));
    it->head = it->tail = node_alloc(it);
    it->max_tid = it->tid = -1; it->max_pos = it->pos = -1; /* nothing seen yet */
    it->tid = 0; it->pos = 0;
    it->maxcnt = 8000;
    it->func = func; it->data = data;
    if (func) it->b = bam_init1();
    return it;
}

static void plp_destroy(struct mdshim_plp *it) {
    lbnode_t *p, *q;
    for (p = it->head; p; p = q) {
        q = p->next;
        if (p != it->tail && it->dtor) it->dtor(it->data, &p->b, &p->cd);
        free(p->b.data); free(p);
    }
    for (p = it->free_list; p; p = q) { q = p->next; free(p->b.data); free(p); }
    free(it->plp);
    if (it->b) bam_destroy1(it->b);
    free(it);
}

static int is_ref_op(int op) { return op == BAM_CMATCH || op == BAM_CDEL || op == BAM_CREF_SKIP || op == BAM_CEQUAL || op == BAM_CDIFF; }
static int is_match_op(int op) { return op == BAM_CMATCH || op == BAM_CEQUAL || op == BAM_CDIFF; }

/* Locate column `pos` inside the record's CIGAR, resuming from the cached op. */
static int resolve_column(bam_pileup1_t *p, hts_pos_t pos, cstate_t *s) {
    bam1_t *b = p->b;
    const bam1_core_t *c = &b->core;
    const uint32_t *cig = bam_get_cigar(b);
    int k;
    if (s->k == -1) {
        if (c->n_cigar == 1) {
            if (is_match_op((int) bam_cigar_op(cig[0]))) { s->k = 0; s->x = c->pos; s->y = 0; }
        } else {
            s->x = c->pos; s->y = 0;
            for (k = 0; k < (int) c->n_cigar; ++k) {
                int op = (int) bam_cigar_op(cig[k]);
                if (is_ref_op(op)) break;
                if (op == BAM_CINS || op == BAM_CSOFT_CLIP) s->y += bam_cigar_oplen(cig[k]);
            }
            assert(k < (int) c->n_cigar);
            s->k = k;
        }
    } else {
        hts_pos_t l = bam_cigar_oplen(cig[s->k]);
        if (pos - s->x >= l) {
            if (is_match_op((int) bam_cigar_op(cig[s->k]))) s->y += l;
            s->x += l;
            for (k = s->k + 1; k < (int) c->n_cigar; ++k) {
                int op = (int) bam_cigar_op(cig[k]);
                if (is_ref_op(op)) break;
                if (op == BAM_CINS || op == BAM_CSOFT_CLIP) s->y += bam_cigar_oplen(cig[k]);
            }
            assert(k < (int) c->n_cigar);
            s->k = k;
        }
    }
    {
        int op = (int) bam_cigar_op(cig[s->k]);
        hts_pos_t l = bam_cigar_oplen(cig[s->k]);
        p->is_del = p->indel = p->is_refskip = 0;
        if (s->x + l - 1 == pos && s->k + 1 < (int) c->n_cigar) { /* look ahead for an indel right after this column */
            int op2 = (int) bam_cigar_op(cig[s->k + 1]);
            hts_pos_t l2 = bam_cigar_oplen(cig[s->k + 1]);
            if (op2 == BAM_CDEL && op != BAM_CDEL) p->indel = -(int) l2;
            else if (op2 == BAM_CINS) p->indel = (int) l2;
        }
        if (is_match_op(op)) p->qpos = (int32_t)(s->y + (pos - s->x));
        else if (op == BAM_CDEL || op == BAM_CREF_SKIP) { p->is_del = 1; p->qpos = (int32_t) s->y; p->is_refskip = (op == BAM_CREF_SKIP); }
        p->is_head = (pos == c->pos); p->is_tail = (pos == s->end);
        p->cigar_ind = s->k;
    }
    return 1;
}

static int plp_push(struct mdshim_plp *it, const bam1_t *b) {
    if (it->error) return -1;
    if (!b) { it->is_eof = 1; return 0; }
    if (b->core.tid < 0) return 0;
    if (b->core.flag & BAM_FUNMAP) return 0;
    if (it->tid == b->core.tid && it->pos == b->core.pos && it->cnt > it->maxcnt) return 0;
    if (!bam_copy1(&it->tail->b, b)) return -1;
    it->tail->beg = b->core.pos;
    it->tail->end = b->core.pos + bam_cigar2rlen((int) b->core.n_cigar, bam_get_cigar(b));
    it->tail->s.k = -1; it->tail->s.x = it->tail->s.y = 0; it->tail->s.end = it->tail->end - 1;
    if (b->core.tid < it->max_tid) { fprintf(stderr, "[shim] The input is not sorted (chromosomes out of order)\n"); it->error = 1; return -1; }
    if (b->core.tid == it->max_tid && it->tail->beg < it->max_pos) { fprintf(stderr, "[shim] The input is not sorted (reads out of order)\n"); it->error = 1; return -1; }
    it->max_tid = b->core.tid; it->max_pos = it->tail->beg;
    if (it->tail->end > it->pos || it->tail->b.core.tid > it->tid) {
        lbnode_t *next = node_alloc(it);
        if (it->ctor && it->ctor(it->data, &it->tail->b, &it->tail->cd) < 0) { it->error = 1; return -1; }
        it->tail->next = next;
        it->tail = next;
    }
    return 0;
}

static const bam_pileup1_t *plp_next(struct mdshim_plp *it, int *_tid, hts_pos_t *_pos, int *_n_plp) {
    if (it->error) { *_n_plp = -1; return NULL; }
    *_n_plp = 0;
    if (it->is_eof && it->head == it->tail) return NULL;
    while (it->is_eof || it->max_tid > it->tid || (it->max_tid == it->tid && it->max_pos > it->pos)) {
        int n_plp = 0;
        lbnode_t **pptr = &it->head;
        while (*pptr != it->tail) {
            lbnode_t *p = *pptr;
            if (p->b.core.tid < it->tid || (p->b.core.tid == it->tid && p->end <= it->pos)) {
                if (it->dtor) it->dtor(it->data, &p->b, &p->cd);
                *pptr = p->next;
                node_free(it, p);
            } else {
                if (p->b.core.tid == it->tid && p->beg <= it->pos) {
                    if (n_plp == it->max_plp) {
                        it->max_plp = it->max_plp ? it->max_plp << 1 : 256;
                        it->plp = (bam_pileup1_t *) realloc(it->plp, sizeof(bam_pileup1_t) * (size_t) it->max_plp);
                    }
                    it->plp[n_plp].b = &p->b;
                    it->plp[n_plp].cd = p->cd;
                    if (resolve_column(it->plp + n_plp, it->pos, &p->s)) ++n_plp;
                }
                pptr = &(*pptr)->next;
            }
        }
        *_n_plp = n_plp; *_tid = it->tid; *_pos = it->pos;
        if (it->head != it->tail && it->tid > it->head->b.core.tid) {
            fprintf(stderr, "[shim] Unsorted input. Pileup aborts\n");
            it->error = 1; *_n_plp = -1;
            return NULL;
        }
        if (it->tid < it->head->b.core.tid) { it->tid = it->head->b.core.tid; it->pos = it->head->beg; }
        else if (it->pos < it->head->beg) it->pos = it->head->beg;
        else ++it->pos;
        if (n_plp) return it->plp;
        if (it->is_eof && it->head == it->tail) break;
    }
    return NULL;
}

static const bam_pileup1_t *plp_auto(struct mdshim_plp *it, int *_tid, hts_pos_t *_pos, int *_n_plp) {
    const bam_pileup1_t *plp;
    if (it->func == 0 || it->error) { *_n_plp = -1; return 0; }
    if ((plp = plp_next(it, _tid, _pos, _n_plp)) != 0) return plp;
    *_n_plp = 0;
    if (it->is_eof) return 0;
    int ret;
    while ((ret = it->func(it->data, it->b)) >= 0) {
        if (plp_push(it, it->b) < 0) { *_n_plp = -1; return 0; }
        if ((plp = plp_next(it, _tid, _pos, _n_plp)) != 0) return plp;
    }
    if (ret < -1) { it->error = ret; *_n_plp = -1; return 0; }
    if (plp_push(it, 0) < 0) { *_n_plp = -1; return 0; }
    if ((plp = plp_next(it, _tid, _pos, _n_plp)) != 0) return plp;
    return 0;
}

struct mdshim_mplp {
    int n;
    struct mdshim_plp **iter;
    int32_t *tid; hts_pos_t *pos; int *n_plp;
    const bam_pileup1_t **plp;
    int32_t min_tid; hts_pos_t min_pos;
};

bam_mplp_t bam_mplp_init(int n, bam_plp_auto_f func, void **data) {
    bam_mplp_t m = (bam_mplp_t) calloc(1, sizeof(*m));
    m->n = n;
    m->iter = (struct mdshim_plp **) calloc((size_t) n, sizeof(*m->iter));
    m->tid = (int32_t *) calloc((size_t) n, sizeof(int32_t));
    m->pos = (hts_pos_t *) calloc((size_t) n, sizeof(hts_pos_t));
    m->n_plp = (int *) calloc((size_t) n, sizeof(int));
    m->plp = (const bam_pileup1_t **) calloc((size_t) n, sizeof(*m->plp));
    m->min_tid = -1; m->min_pos = -1; /* sentinel: every iterator is due */
    for (int i = 0; i < n; ++i) { m->iter[i] = plp_init(func, data[i]); m->pos[i] = -1; m->tid[i] = -1; }
    return m;
}

void bam_mplp_destroy(bam_mplp_t m) {
    for (int i = 0; i < m->n; ++i) plp_destroy(m->iter[i]);
    free(m->iter); free(m->tid); free(m->pos); free(m->n_plp); free((void *) m->plp); free(m);
}

void bam_mplp_set_maxcnt(bam_mplp_t m, int maxcnt) { for (int i = 0; i < m->n; ++i) m->iter[i]->maxcnt = maxcnt; }
void bam_mplp_constructor(bam_mplp_t m, int (*func)(void *, const bam1_t *, bam_pileup_cd *)) { for (int i = 0; i < m->n; ++i) m->iter[i]->ctor = func; }
void bam_mplp_destructor(bam_mplp_t m, int (*func)(void *, const bam1_t *, bam_pileup_cd *)) { for (int i = 0; i < m->n; ++i) m->iter[i]->dtor = func; }

int bam_mplp64_auto(bam_mplp_t m, int *_tid, hts_pos_t *_pos, int *n_plp, const bam_pileup1_t **plp) {
    int ret = 0;
    int32_t new_tid = INT32_MAX; hts_pos_t new_pos = HTS_POS_MAX;
    for (int i = 0; i < m->n; ++i) {
        if (m->pos[i] == m->min_pos && m->tid[i] == m->min_tid) {
            int tid; hts_pos_t pos;
            m->plp[i] = plp_auto(m->iter[i], &tid, &pos, &m->n_plp[i]);
            if (m->iter[i]->error) return -1;
            if (m->plp[i]) { m->tid[i] = tid; m->pos[i] = pos; }
            else { m->tid[i] = INT32_MAX; m->pos[i] = HTS_POS_MAX; }
        }
        if (m->plp[i]) {
            if (m->tid[i] < new_tid) { new_tid = m->tid[i]; new_pos = m->pos[i]; }
            else if (m->tid[i] == new_tid && m->pos[i] < new_pos) new_pos = m->pos[i];
        }
    }
    m->min_tid = new_tid; m->min_pos = new_pos;
    if (new_pos == HTS_POS_MAX) return 0;
    *_tid = new_tid; *_pos = new_pos;
    for (int i = 0; i < m->n; ++i) {
        if (m->pos[i] == m->min_pos && m->tid[i] == m->min_tid) { n_plp[i] = m->n_plp[i]; plp[i] = m->plp[i]; ++ret; }
        else { n_plp[i] = 0; plp[i] = 0; }
    }
    return ret;
}

int bam_mplp_auto(bam_mplp_t m, int *_tid, int *_pos, int *n_plp, const bam_pileup1_t **plp) {
    hts_pos_t pos64 = 0;
    int ret = bam_mplp64_auto(m, _tid, &pos64, n_plp, plp);
    if (ret >= 0) *_pos = (pos64 < INT_MAX) ? (int) pos64 : INT_MAX;
    return ret;
}
