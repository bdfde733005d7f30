/* TEST INFRASTRUCTURE — htslib-compatible shim (see hts.h in this directory).
 * String-keyed hash map exposing the macro surface overlaps.c:12-23,125-145
 * uses (KHASH_MAP_INIT_STR / kh_init / kh_get / kh_put / kh_value / kh_del /
 * kh_end / kh_destroy).  Keys are borrowed pointers, exactly as the reference
 * relies on (the qname lives inside the pileup's copy of the record).
 * Own implementation: open addressing, linear probing, tombstones. */
#ifndef MDSHIM_KHASH_H
#define MDSHIM_KHASH_H
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

typedef uint32_t khint_t;
typedef khint_t khiter_t;

static inline uint32_t mdshim_strhash(const char *s) {
    uint32_t h = 2166136261u;
    for (; *s; ++s) { h ^= (unsigned char) *s; h *= 16777619u; }
    return h;
}

#define khash_t(name) mdshim_kh_##name##_t

#define KHASH_MAP_INIT_STR(name, khval_t)                                          \
    typedef struct mdshim_kh_##name##_s {                                          \
        khint_t n_buckets, size, n_occupied;                                       \
        uint8_t *state; /* 0 empty, 1 live, 2 deleted */                           \
        const char **keys;                                                         \
        khval_t *vals;                                                             \
    } mdshim_kh_##name##_t;                                                        \
    static inline mdshim_kh_##name##_t *mdshim_kh_init_##name(void) {              \
        return (mdshim_kh_##name##_t *) calloc(1, sizeof(mdshim_kh_##name##_t));   \
    }                                                                              \
    static inline void mdshim_kh_destroy_##name(mdshim_kh_##name##_t *h) {         \
        if (h) { free(h->state); free((void *) h->keys); free(h->vals); free(h); } \
    }                                                                              \
    static inline khint_t mdshim_kh_get_##name(const mdshim_kh_##name##_t *h, const char *key) { \
        if (!h->n_buckets) return 0;                                               \
        khint_t mask = h->n_buckets - 1, i = mdshim_strhash(key) & mask, n = 0;    \
        while (h->state[i] != 0 && n < h->n_buckets) {                             \
            if (h->state[i] == 1 && strcmp(h->keys[i], key) == 0) return i;        \
            i = (i + 1) & mask; ++n;                                               \
        }                                                                          \
        return h->n_buckets;                                                       \
    }                                                                              \
    static inline void mdshim_kh_resize_##name(mdshim_kh_##name##_t *h, khint_t nb) { \
        uint8_t *os = h->state; const char **ok = h->keys; khval_t *ov = h->vals;  \
        khint_t onb = h->n_buckets, j;                                             \
        h->state = (uint8_t *) calloc(nb, 1);                                      \
        h->keys = (const char **) calloc(nb, sizeof(char *));                      \
        h->vals = (khval_t *) calloc(nb, sizeof(khval_t));                         \
        h->n_buckets = nb; h->n_occupied = 0; h->size = 0;                         \
        for (j = 0; j < onb; ++j) if (os[j] == 1) {                                \
            khint_t mask = nb - 1, i = mdshim_strhash(ok[j]) & mask;               \
            while (h->state[i]) i = (i + 1) & mask;                                \
            h->state[i] = 1; h->keys[i] = ok[j]; h->vals[i] = ov[j];               \
            ++h->size; ++h->n_occupied;                                            \
        }                                                                          \
        free(os); free((void *) ok); free(ov);                                     \
    }                                                                              \
    static inline khint_t mdshim_kh_put_##name(mdshim_kh_##name##_t *h, const char *key, int *ret) { \
        if ((h->n_occupied + 1) * 4 >= h->n_buckets * 3 || h->n_buckets == 0) {    \
            khint_t nb = h->n_buckets ? h->n_buckets : 16;                         \
            if ((h->size + 1) * 2 >= nb) nb <<= 1;                                 \
            mdshim_kh_resize_##name(h, nb);                                        \
        }                                                                          \
        khint_t mask = h->n_buckets - 1, i = mdshim_strhash(key) & mask, tomb = h->n_buckets; \
        while (h->state[i] != 0) {                                                 \
            if (h->state[i] == 1 && strcmp(h->keys[i], key) == 0) { *ret = 0; return i; } \
            if (h->state[i] == 2 && tomb == h->n_buckets) tomb = i;                \
            i = (i + 1) & mask;                                                    \
        }                                                                          \
        if (tomb != h->n_buckets) { i = tomb; *ret = 2; }                          \
        else { ++h->n_occupied; *ret = 1; }                                        \
        h->state[i] = 1; h->keys[i] = key; ++h->size;                              \
        return i;                                                                  \
    }                                                                              \
    static inline void mdshim_kh_del_##name(mdshim_kh_##name##_t *h, khint_t x) {  \
        if (x < h->n_buckets && h->state[x] == 1) { h->state[x] = 2; --h->size; }  \
    }

#define kh_init(name) mdshim_kh_init_##name()
#define kh_destroy(name, h) mdshim_kh_destroy_##name(h)
#define kh_get(name, h, k) mdshim_kh_get_##name(h, k)
#define kh_put(name, h, k, r) mdshim_kh_put_##name(h, k, r)
#define kh_del(name, h, k) mdshim_kh_del_##name(h, k)
#define kh_exist(h, x) ((h)->state[(x)] == 1)
#define kh_key(h, x) ((h)->keys[x])
#define kh_val(h, x) ((h)->vals[x])
#define kh_value(h, x) ((h)->vals[x])
#define kh_begin(h) (khint_t)(0)
#define kh_end(h) ((h)->n_buckets)
#define kh_size(h) ((h)->size)
#endif
