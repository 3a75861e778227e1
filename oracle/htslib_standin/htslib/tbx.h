/* Stand-in for htslib/tbx.h — see kstring.h. */
#ifndef STANDIN_TBX_H
#define STANDIN_TBX_H
#include "hts.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct tbx_conf_t { int32_t preset, sc, bc, ec, meta_char, line_skip; } tbx_conf_t;
typedef struct tbx_t { tbx_conf_t conf; hts_idx_t* idx; void* dict; } tbx_t;
tbx_t* tbx_index_load(const char* fn);
void tbx_destroy(tbx_t* tbx);
int tbx_name2id(tbx_t* tbx, const char* ss);
hts_itr_t* tbx_itr_querys(tbx_t* tbx, const char* reg);
hts_itr_t* tbx_itr_queryi(tbx_t* tbx, int tid, int beg, int end);
int tbx_itr_next(htsFile* fp, tbx_t* tbx, hts_itr_t* itr, void* r);
#define tbx_itr_destroy(iter) hts_itr_destroy(iter)
const char** tbx_seqnames(tbx_t* tbx, int* n);
#ifdef __cplusplus
}
#endif
#endif
