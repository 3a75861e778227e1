/* Stand-in for htslib/hfile.h — see kstring.h. */
#ifndef STANDIN_HFILE_H
#define STANDIN_HFILE_H
#include <sys/types.h>
#include "hts.h"
#ifdef __cplusplus
extern "C" {
#endif
hFILE* hopen(const char* filename, const char* mode, ...);
int hclose(hFILE* fp);
ssize_t hread(hFILE* fp, void* buffer, size_t nbytes);
ssize_t hwrite(hFILE* fp, const void* buffer, size_t nbytes);
int hflush(hFILE* fp);
#ifdef __cplusplus
}
#endif
#endif
