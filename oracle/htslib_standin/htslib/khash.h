/* Stand-in for htslib/khash.h — see kstring.h.  The reference expands KHASH_MAP_INIT_STR in
 * hts_utils.h:72 and walks header dictionaries in helpers that the PLP path never reaches
 * (hts_utils.cpp:404-447,557-562,984-1000,1174-1190); this is a plain linear-scan table with the
 * same macro surface, sufficient to compile and link them. */
#ifndef STANDIN_KHASH_H
#define STANDIN_KHASH_H
#include <stdint.h>
#include <string.h>
typedef uint32_t khint_t;
typedef khint_t khiter_t;
#define khash_t(name) kh_##name##_t
#define STANDIN_KHASH_DECL(name, khkey_t, khval_t, eq)                                             \
  typedef struct kh_##name##_s { khint_t n_buckets, size; unsigned char* used; khkey_t* keys; khval_t* vals; } kh_##name##_t; \
  static inline khint_t kh_get_##name(const kh_##name##_t* h, khkey_t key) {                     \
    if (!h) return 0;                                                                             \
    for (khint_t i = 0; i < h->n_buckets; ++i)                                                    \
      if (h->used[i] && eq(h->keys[i], key)) return i;                                            \
    return h->n_buckets;                                                                          \
  }
#define standin_kh_str_eq(a, b) (strcmp((a), (b)) == 0)
#define standin_kh_int_eq(a, b) ((a) == (b))
#define KHASH_MAP_INIT_STR(name, khval_t) STANDIN_KHASH_DECL(name, const char*, khval_t, standin_kh_str_eq)
#define KHASH_MAP_INIT_INT(name, khval_t) STANDIN_KHASH_DECL(name, int32_t, khval_t, standin_kh_int_eq)
#define KHASH_SET_INIT_STR(name) STANDIN_KHASH_DECL(name, const char*, char, standin_kh_str_eq)
#define kh_get(name, h, k) kh_get_##name((h), (k))
#define kh_size(h) ((h)->size)
#define kh_begin(h) ((khint_t)0)
#define kh_end(h) ((h)->n_buckets)
#define kh_exist(h, x) ((h)->used[(x)])
#define kh_val(h, x) ((h)->vals[(x)])
#define kh_value(h, x) ((h)->vals[(x)])
#define kh_key(h, x) ((h)->keys[(x)])
#endif
