/* Stand-in for htslib/hts.h — see kstring.h. */
#ifndef STANDIN_HTS_H
#define STANDIN_HTS_H
#include <stddef.h>
#include <stdint.h>
#include "kstring.h"
#ifdef __cplusplus
extern "C" {
#endif
struct BGZF;
struct hFILE;
typedef struct BGZF BGZF;
typedef struct hFILE hFILE;
enum htsFormatCategory { unknown_category, sequence_data, variant_data, index_file, region_list };
enum htsExactFormat { unknown_format, binary_format, text_format, sam, bam, bai, cram, crai, vcf, bcf, csi, gzi, tbi, bed };
enum htsCompression { no_compression, gzip, bgzf, custom };
typedef struct htsFormat {
  enum htsFormatCategory category;
  enum htsExactFormat format;
  struct { short major, minor; } version;
  enum htsCompression compression;
  short compression_level;
  void* specific;
} htsFormat;
typedef struct htsFile {
  uint32_t is_bin : 1, is_write : 1, is_be : 1, is_cram : 1, is_bgzf : 1, dummy : 27;
  int64_t lineno;
  kstring_t line;
  char *fn, *fn_aux;
  union { BGZF* bgzf; struct cram_fd* cram; hFILE* hfile; void* voidp; } fp;
  htsFormat format;
  void* standin;  /* stand-in private state (gzFile) */
} htsFile;
htsFile* hts_open(const char* fn, const char* mode);
int hts_close(htsFile* fp);
int hts_getline(htsFile* fp, int delimiter, kstring_t* str);
int hts_set_threads(htsFile* fp, int n);
int hts_set_fai_filename(htsFile* fp, const char* fn_aux);
const htsFormat* hts_get_format(htsFile* fp);
const char* hts_version(void);
char** hts_readlines(const char* fn, int* _n);
char** hts_readlist(const char* fn, int is_file, int* _n);

struct hts_idx_t;
typedef struct hts_idx_t hts_idx_t;
typedef struct hts_itr_t {
  uint32_t read_rest : 1, finished : 1, is_cram : 1, dummy : 29;
  int tid, beg, end, n_off, i;
  int curr_tid, curr_beg, curr_end;
  uint64_t curr_off;
} hts_itr_t;
#define HTS_IDX_NOCOOR (-2)
#define HTS_IDX_START (-3)
#define HTS_IDX_REST (-4)
#define HTS_IDX_NONE (-5)
#define HTS_FMT_CSI 0
#define HTS_FMT_BAI 1
#define HTS_FMT_TBI 2
#define HTS_FMT_CRAI 3
void hts_idx_destroy(hts_idx_t* idx);
void hts_itr_destroy(hts_itr_t* iter);
const char* hts_parse_reg(const char* str, int* beg, int* end);
typedef int (*hts_name2id_f)(void*, const char*);
typedef const char* (*hts_id2name_f)(void*, int);
#define hts_expand(type_t, n, m, ptr) if ((n) > (m)) { (m) = (n); (m) += (m) >> 1; (ptr) = (type_t*)realloc((ptr), (m) * sizeof(type_t)); }
#define hts_expand0(type_t, n, m, ptr) if ((n) > (m)) { int t_ = (m); (m) = (n); (m) += (m) >> 1; (ptr) = (type_t*)realloc((ptr), (m) * sizeof(type_t)); memset(((type_t*)ptr) + t_, 0, sizeof(type_t) * ((m) - t_)); }
#ifdef __cplusplus
}
#endif
#endif
