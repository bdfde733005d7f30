/* TEST INFRASTRUCTURE — htslib-compatible shim (see hts.h in this directory).
 * Growable string used by extract.c:98 (kputs) and the kstream readers. */
#ifndef MDSHIM_KSTRING_H
#define MDSHIM_KSTRING_H
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#ifndef KSTRING_T
#define KSTRING_T kstring_t
typedef struct kstring_t {
    size_t l, m;
    char *s;
} kstring_t;
#endif

static inline int ks_resize(kstring_t *s, size_t size) {
    if (s->m < size) {
        size_t m = s->m ? s->m : 64;
        while (m < size) m += (m >> 1) + 16;
        char *t = (char *) realloc(s->s, m);
        if (!t) return -1;
        s->s = t; s->m = m;
    }
    return 0;
}
static inline int kputsn(const char *p, size_t l, kstring_t *s) {
    if (ks_resize(s, s->l + l + 2) < 0) return EOF;
    memcpy(s->s + s->l, p, l);
    s->l += l;
    s->s[s->l] = 0;
    return (int) l;
}
static inline int kputs(const char *p, kstring_t *s) { return kputsn(p, strlen(p), s); }
static inline int kputc(int c, kstring_t *s) {
    if (ks_resize(s, s->l + 2) < 0) return EOF;
    s->s[s->l++] = (char) c;
    s->s[s->l] = 0;
    return (unsigned char) c;
}
#endif
