/* TEST INFRASTRUCTURE — htslib-compatible shim (see hts.h in this directory).
 * Buffered line/token reader generated per stream type, as instantiated at
 * bed.c:6 (gzFile/gzread) and mergeContext.c:15 (FILE pointer + fgets2). */
#ifndef MDSHIM_KSEQ_H
#define MDSHIM_KSEQ_H
#include <ctype.h>
#include <string.h>
#include <stdlib.h>
#include "kstring.h"

#define KS_SEP_SPACE 0
#define KS_SEP_TAB   1
#define KS_SEP_LINE  2
#define KS_SEP_MAX   2

#define KSTREAM_INIT(type_t, __read, __bufsize)                                   \
    typedef struct __kstream_t {                                                  \
        unsigned char *buf;                                                       \
        int begin, end, is_eof;                                                   \
        type_t f;                                                                 \
    } kstream_t;                                                                  \
    static inline kstream_t *ks_init(type_t f) {                                  \
        kstream_t *ks = (kstream_t *) calloc(1, sizeof(kstream_t));               \
        ks->f = f;                                                                \
        ks->buf = (unsigned char *) malloc(__bufsize);                            \
        return ks;                                                                \
    }                                                                             \
    static inline void ks_destroy(kstream_t *ks) {                                \
        if (ks) { free(ks->buf); free(ks); }                                      \
    }                                                                             \
    static inline int mdshim_ks_fill(kstream_t *ks) {                             \
        if (ks->is_eof) return 0;                                                 \
        ks->begin = 0;                                                            \
        ks->end = __read(ks->f, ks->buf, __bufsize);                              \
        if (ks->end <= 0) { ks->end = 0; ks->is_eof = 1; return 0; }              \
        return 1;                                                                 \
    }                                                                             \
    static inline int ks_getuntil(kstream_t *ks, int delimiter, kstring_t *str, int *dret) { \
        int got = 0;                                                              \
        if (dret) *dret = 0;                                                      \
        str->l = 0;                                                               \
        for (;;) {                                                                \
            int i;                                                                \
            if (ks->begin >= ks->end) {                                           \
                if (!mdshim_ks_fill(ks)) break;                                   \
            }                                                                     \
            for (i = ks->begin; i < ks->end; ++i) {                               \
                int c = ks->buf[i];                                               \
                if (delimiter == KS_SEP_LINE) { if (c == '\n') break; }           \
                else if (delimiter == KS_SEP_SPACE) { if (isspace(c)) break; }    \
                else if (delimiter == KS_SEP_TAB) { if (isspace(c) && c != ' ') break; } \
                else if (c == delimiter) break;                                   \
            }                                                                     \
            got = 1;                                                              \
            kputsn((const char *) ks->buf + ks->begin, (size_t)(i - ks->begin), str); \
            if (i < ks->end) {                                                    \
                if (dret) *dret = ks->buf[i];                                     \
                ks->begin = i + 1;                                                \
                break;                                                            \
            }                                                                     \
            ks->begin = ks->end;                                                  \
        }                                                                         \
        if (!got && ks->is_eof) return -1;                                        \
        if (str->s == NULL) { str->m = 1; str->s = (char *) calloc(1, 1); }       \
        if (delimiter == KS_SEP_LINE && str->l > 0 && str->s[str->l - 1] == '\r') \
            str->s[--str->l] = 0;                                                 \
        str->s[str->l] = 0;                                                       \
        return (int) str->l;                                                      \
    }
#endif
