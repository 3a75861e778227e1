/* Stand-in for htslib/faidx.h — see kstring.h. */
#ifndef STANDIN_FAIDX_H
#define STANDIN_FAIDX_H
#ifdef __cplusplus
extern "C" {
#endif
struct __faidx_t;
typedef struct __faidx_t faidx_t;
int fai_build(const char* fn);
void fai_destroy(faidx_t* fai);
faidx_t* fai_load(const char* fn);
char* fai_fetch(const faidx_t* fai, const char* reg, int* len);
int faidx_nseq(const faidx_t* fai);
char* faidx_fetch_seq(const faidx_t* fai, const char* c_name, int p_beg_i, int p_end_i, int* len);
int faidx_has_seq(const faidx_t* fai, const char* seq);
const char* faidx_iseq(const faidx_t* fai, int i);
int faidx_seq_len(const faidx_t* fai, const char* seq);
#ifdef __cplusplus
}
#endif
#endif
