/* Stand-in for htslib/kstring.h — TEST INFRASTRUCTURE (oracle/_ref build only).
 * htslib (samtools/htslib; un-vendored and un-pinned by the reference, Dockerfile:26) is absent
 * from this image.  These headers declare the slice of its public API that the reference's
 * demuxlet / freemuxlet translation units mention, so that those units compile UNMODIFIED from
 * /root/reference; oracle/htslib_standin/standin.cpp implements the calls the PLP path reaches
 * (gzip text I/O, VCF text records) and aborts in the rest.  No likelihood arithmetic lives here. */
#ifndef STANDIN_KSTRING_H
#define STANDIN_KSTRING_H
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifndef KSTRING_T
#define KSTRING_T kstring_t
typedef struct kstring_t { size_t l, m; char* s; } kstring_t;
#endif
#ifdef __cplusplus
extern "C" {
#endif
int kvsprintf(kstring_t* s, const char* fmt, va_list ap);
int ksprintf(kstring_t* s, const char* fmt, ...);
int ksplit_core(char* s, int delimiter, int* _max, int** _offsets);
#ifdef __cplusplus
}
#endif
static inline int ks_resize(kstring_t* s, size_t size) {
  if (s->m < size) {
    size_t m = size < 16 ? 16 : size;
    m += m >> 1;
    char* t = (char*)realloc(s->s, m);
    if (!t) return -1;
    s->s = t; s->m = m;
  }
  return 0;
}
static inline int kputsn(const char* p, int l, kstring_t* s) {
  if (ks_resize(s, s->l + l + 2) < 0) return EOF;
  memcpy(s->s + s->l, p, l); s->l += l; s->s[s->l] = 0;
  return l;
}
static inline int kputs(const char* p, kstring_t* s) { return kputsn(p, (int)strlen(p), s); }
static inline int kputc(int c, kstring_t* s) {
  if (ks_resize(s, s->l + 2) < 0) return EOF;
  s->s[s->l++] = (char)c; s->s[s->l] = 0;
  return c;
}
static inline int kputw(int c, kstring_t* s) { char b[16]; int n = snprintf(b, sizeof b, "%d", c); return kputsn(b, n, s); }
static inline int kputl(long c, kstring_t* s) { char b[32]; int n = snprintf(b, sizeof b, "%ld", c); return kputsn(b, n, s); }
static inline int kputuw(unsigned c, kstring_t* s) { char b[16]; int n = snprintf(b, sizeof b, "%u", c); return kputsn(b, n, s); }
static inline char* ks_release(kstring_t* s) { char* r = s->s; s->l = s->m = 0; s->s = NULL; return r; }
static inline int* ksplit(kstring_t* s, int delimiter, int* n) {
  int max = 0, *offsets = 0;
  *n = ksplit_core(s->s, delimiter, &max, &offsets);
  return offsets;
}
#endif
