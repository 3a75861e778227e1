/* Stand-in for htslib/bgzf.h — see kstring.h. */
#ifndef STANDIN_BGZF_H
#define STANDIN_BGZF_H
#include <sys/types.h>
#include "hts.h"
#ifdef __cplusplus
extern "C" {
#endif
BGZF* bgzf_open(const char* path, const char* mode);
int bgzf_close(BGZF* fp);
ssize_t bgzf_read(BGZF* fp, void* data, size_t length);
ssize_t bgzf_write(BGZF* fp, const void* data, size_t length);
int bgzf_getline(BGZF* fp, int delim, kstring_t* str);
int64_t bgzf_seek(BGZF* fp, int64_t pos, int whence);
int bgzf_flush(BGZF* fp);
int bgzf_is_bgzf(const char* fn);
int bgzf_mt(BGZF* fp, int n_threads, int n_sub_blks);
ssize_t hwrite(hFILE* fp, const void* buffer, size_t nbytes);
#ifdef __cplusplus
}
#endif
#endif
