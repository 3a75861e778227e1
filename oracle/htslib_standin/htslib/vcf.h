/* Stand-in for htslib/vcf.h — see kstring.h.  The structs keep htslib's public field names
 * because the reference reads them directly (e.g. bcf1_t::rid/pos/n_allele/d.allele,
 * bcf_hdr_t::id[BCF_DT_CTG][rid].key, bcf_fmt_t, bcf_info_t). */
#ifndef STANDIN_VCF_H
#define STANDIN_VCF_H
#include <assert.h>
#include <limits.h>
#include <stdint.h>
#include "bgzf.h"
#include "hts.h"
#include "kstring.h"
#ifdef __cplusplus
extern "C" {
#endif
#define BCF_HL_FLT 0
#define BCF_HL_INFO 1
#define BCF_HL_FMT 2
#define BCF_HL_CTG 3
#define BCF_HL_STR 4
#define BCF_HL_GEN 5
#define BCF_HT_FLAG 0
#define BCF_HT_INT 1
#define BCF_HT_REAL 2
#define BCF_HT_STR 3
#define BCF_VL_FIXED 0
#define BCF_VL_VAR 1
#define BCF_VL_A 2
#define BCF_VL_G 3
#define BCF_VL_R 4
#define BCF_DT_ID 0
#define BCF_DT_CTG 1
#define BCF_DT_SAMPLE 2
typedef struct bcf_hrec_t { int type; char *key, *value; int nkeys; char **keys, **vals; } bcf_hrec_t;
typedef struct bcf_idinfo_t { uint32_t info[3]; bcf_hrec_t* hrec[3]; int id; } bcf_idinfo_t;
typedef struct bcf_idpair_t { const char* key; const bcf_idinfo_t* val; } bcf_idpair_t;
typedef struct bcf_hdr_t {
  int32_t n[3];
  bcf_idpair_t* id[3];
  void* dict[3];
  char** samples;
  bcf_hrec_t** hrec;
  int nhrec, dirty;
  int ntransl, *transl[2];
  int nsamples_ori;
  uint8_t* keep_samples;
  kstring_t mem;
} bcf_hdr_t;
extern uint8_t bcf_type_shift[];
#define BCF_BT_NULL 0
#define BCF_BT_INT8 1
#define BCF_BT_INT16 2
#define BCF_BT_INT32 3
#define BCF_BT_FLOAT 5
#define BCF_BT_CHAR 7
#define VCF_REF 0
#define VCF_SNP 1
#define VCF_MNP 2
#define VCF_INDEL 4
#define VCF_OTHER 8
#define VCF_BND 16
typedef struct variant_t { int type, n; } variant_t;
typedef struct bcf_fmt_t { int id, n, size, type; uint8_t* p; uint32_t p_len; uint32_t p_off : 31, p_free : 1; } bcf_fmt_t;
typedef struct bcf_info_t {
  int key, type, len;
  union { int32_t i; float f; } v1;
  uint8_t* vptr;
  uint32_t vptr_len;
  uint32_t vptr_off : 31, vptr_free : 1;
} bcf_info_t;
#define BCF1_DIRTY_ID 1
#define BCF1_DIRTY_ALS 2
#define BCF1_DIRTY_FLT 4
#define BCF1_DIRTY_INF 8
typedef struct bcf_dec_t {
  int m_fmt, m_info, m_id, m_als, m_allele, m_flt;
  int n_flt;
  int* flt;
  char *id, *als;
  char** allele;
  bcf_info_t* info;
  bcf_fmt_t* fmt;
  variant_t* var;
  int n_var, var_type;
  int shared_dirty, indiv_dirty;
} bcf_dec_t;
#define BCF_ERR_CTG_UNDEF 1
#define BCF_ERR_TAG_UNDEF 2
#define BCF_ERR_NCOLS 4
#define BCF_ERR_LIMITS 8
typedef struct bcf1_t {
  int32_t rid, pos, rlen;
  float qual;
  uint32_t n_info : 16, n_allele : 16;
  uint32_t n_fmt : 8, n_sample : 24;
  kstring_t shared, indiv;
  bcf_dec_t d;
  int max_unpack, unpacked, unpack_size[3], errcode;
  void* standin;  /* stand-in private state (parsed text record) */
} bcf1_t;
typedef htsFile vcfFile;
#define bcf_init1() bcf_init()
#define bcf_read1(fp, h, v) bcf_read((fp), (h), (v))
#define vcf_read1(fp, h, v) vcf_read((fp), (h), (v))
#define bcf_write1(fp, h, v) bcf_write((fp), (h), (v))
#define vcf_write1(fp, h, v) vcf_write((fp), (h), (v))
#define bcf_destroy1(v) bcf_destroy(v)
#define bcf_empty1(v) bcf_empty(v)
#define vcf_parse1(s, h, v) vcf_parse((s), (h), (v))
#define bcf_clear1(v) bcf_clear(v)
#define vcf_format1(h, v, s) vcf_format((h), (v), (s))
bcf_hdr_t* bcf_hdr_init(const char* mode);
void bcf_hdr_destroy(bcf_hdr_t* h);
bcf1_t* bcf_init(void);
void bcf_destroy(bcf1_t* v);
void bcf_empty(bcf1_t* v);
void bcf_clear(bcf1_t* v);
#define bcf_open(fn, mode) hts_open((fn), (mode))
#define vcf_open(fn, mode) hts_open((fn), (mode))
#define bcf_close(fp) hts_close(fp)
#define vcf_close(fp) hts_close(fp)
bcf_hdr_t* bcf_hdr_read(htsFile* fp);
int bcf_hdr_set_samples(bcf_hdr_t* hdr, const char* samples, int is_file);
int bcf_subset_format(const bcf_hdr_t* hdr, bcf1_t* rec);
int bcf_hdr_write(htsFile* fp, bcf_hdr_t* h);
int vcf_parse(kstring_t* s, const bcf_hdr_t* h, bcf1_t* v);
int vcf_format(const bcf_hdr_t* h, const bcf1_t* v, kstring_t* s);
int bcf_read(htsFile* fp, const bcf_hdr_t* h, bcf1_t* v);
#define BCF_UN_STR 1
#define BCF_UN_FLT 2
#define BCF_UN_INFO 4
#define BCF_UN_SHR (BCF_UN_STR | BCF_UN_FLT | BCF_UN_INFO)
#define BCF_UN_FMT 8
#define BCF_UN_IND BCF_UN_FMT
#define BCF_UN_ALL (BCF_UN_SHR | BCF_UN_FMT)
int bcf_unpack(bcf1_t* b, int which);
bcf1_t* bcf_dup(bcf1_t* src);
bcf1_t* bcf_copy(bcf1_t* dst, bcf1_t* src);
int bcf_write(htsFile* fp, bcf_hdr_t* h, bcf1_t* v);
bcf_hdr_t* vcf_hdr_read(htsFile* fp);
int vcf_hdr_write(htsFile* fp, const bcf_hdr_t* h);
int vcf_read(htsFile* fp, const bcf_hdr_t* h, bcf1_t* v);
int vcf_write(htsFile* fp, const bcf_hdr_t* h, bcf1_t* v);
bcf_hdr_t* bcf_hdr_dup(const bcf_hdr_t* hdr);
int bcf_hdr_combine(bcf_hdr_t* dst, const bcf_hdr_t* src);
int bcf_hdr_add_sample(bcf_hdr_t* hdr, const char* sample);
int bcf_hdr_set(bcf_hdr_t* hdr, const char* fname);
char* bcf_hdr_fmt_text(const bcf_hdr_t* hdr, int is_bcf, int* len);
int bcf_hdr_append(bcf_hdr_t* h, const char* line);
int bcf_hdr_printf(bcf_hdr_t* h, const char* format, ...);
const char* bcf_hdr_get_version(const bcf_hdr_t* hdr);
void bcf_hdr_remove(bcf_hdr_t* h, int type, const char* key);
bcf_hdr_t* bcf_hdr_subset(const bcf_hdr_t* h0, int n, char* const* samples, int* imap);
const char** bcf_hdr_seqnames(const bcf_hdr_t* h, int* nseqs);
#define bcf_hdr_nsamples(hdr) (hdr)->n[BCF_DT_SAMPLE]
int bcf_hdr_parse(bcf_hdr_t* hdr, char* htxt);
int bcf_hdr_sync(bcf_hdr_t* h);
bcf_hrec_t* bcf_hdr_parse_line(const bcf_hdr_t* h, const char* line, int* len);
void bcf_hrec_format(const bcf_hrec_t* hrec, kstring_t* str);
int bcf_hdr_add_hrec(bcf_hdr_t* hdr, bcf_hrec_t* hrec);
bcf_hrec_t* bcf_hdr_get_hrec(const bcf_hdr_t* hdr, int type, const char* key, const char* value, const char* str_class);
bcf_hrec_t* bcf_hrec_dup(bcf_hrec_t* hrec);
void bcf_hrec_add_key(bcf_hrec_t* hrec, const char* str, int len);
void bcf_hrec_set_val(bcf_hrec_t* hrec, int i, const char* str, int len, int is_quoted);
int bcf_hrec_find_key(bcf_hrec_t* hrec, const char* key);
void bcf_hrec_destroy(bcf_hrec_t* hrec);
int bcf_subset(const bcf_hdr_t* h, bcf1_t* v, int n, int* imap);
int bcf_translate(const bcf_hdr_t* dst_hdr, bcf_hdr_t* src_hdr, bcf1_t* src_line);
int bcf_get_variant_types(bcf1_t* rec);
int bcf_get_variant_type(bcf1_t* rec, int ith_allele);
int bcf_is_snp(bcf1_t* v);
int bcf_update_filter(const bcf_hdr_t* hdr, bcf1_t* line, int* flt_ids, int n);
int bcf_add_filter(const bcf_hdr_t* hdr, bcf1_t* line, int flt_id);
int bcf_remove_filter(const bcf_hdr_t* hdr, bcf1_t* line, int flt_id, int pass);
int bcf_has_filter(const bcf_hdr_t* hdr, bcf1_t* line, char* filter);
int bcf_update_alleles(const bcf_hdr_t* hdr, bcf1_t* line, const char** alleles, int nals);
int bcf_update_alleles_str(const bcf_hdr_t* hdr, bcf1_t* line, const char* alleles_string);
int bcf_update_id(const bcf_hdr_t* hdr, bcf1_t* line, const char* id);
int bcf_add_id(const bcf_hdr_t* hdr, bcf1_t* line, const char* id);
#define bcf_update_info_int32(hdr, line, key, values, n) bcf_update_info((hdr), (line), (key), (values), (n), BCF_HT_INT)
#define bcf_update_info_float(hdr, line, key, values, n) bcf_update_info((hdr), (line), (key), (values), (n), BCF_HT_REAL)
#define bcf_update_info_flag(hdr, line, key, string, n) bcf_update_info((hdr), (line), (key), (string), (n), BCF_HT_FLAG)
#define bcf_update_info_string(hdr, line, key, string) bcf_update_info((hdr), (line), (key), (string), 1, BCF_HT_STR)
int bcf_update_info(const bcf_hdr_t* hdr, bcf1_t* line, const char* key, const void* values, int n, int type);
#define bcf_update_format_int32(hdr, line, key, values, n) bcf_update_format((hdr), (line), (key), (values), (n), BCF_HT_INT)
#define bcf_update_format_float(hdr, line, key, values, n) bcf_update_format((hdr), (line), (key), (values), (n), BCF_HT_REAL)
#define bcf_update_format_char(hdr, line, key, values, n) bcf_update_format((hdr), (line), (key), (values), (n), BCF_HT_STR)
#define bcf_update_genotypes(hdr, line, gts, n) bcf_update_format((hdr), (line), "GT", (gts), (n), BCF_HT_INT)
int bcf_update_format_string(const bcf_hdr_t* hdr, bcf1_t* line, const char* key, const char** values, int n);
int bcf_update_format(const bcf_hdr_t* hdr, bcf1_t* line, const char* key, const void* values, int n, int type);
#define bcf_gt_phased(idx) (((idx) + 1) << 1 | 1)
#define bcf_gt_unphased(idx) (((idx) + 1) << 1)
#define bcf_gt_missing 0
#define bcf_gt_is_missing(val) ((val) >> 1 ? 0 : 1)
#define bcf_gt_is_phased(idx) ((idx)&1)
#define bcf_gt_allele(val) (((val) >> 1) - 1)
#define bcf_alleles2gt(a, b) ((a) > (b) ? ((a) * ((a) + 1) / 2 + (b)) : ((b) * ((b) + 1) / 2 + (a)))
static inline void bcf_gt2alleles(int igt, int* a, int* b) {
  int k = 0, dk = 1;
  while (k < igt) { dk++; k += dk; }
  *b = dk - 1; *a = igt - k + *b;
}
bcf_fmt_t* bcf_get_fmt(const bcf_hdr_t* hdr, bcf1_t* line, const char* key);
bcf_info_t* bcf_get_info(const bcf_hdr_t* hdr, bcf1_t* line, const char* key);
bcf_fmt_t* bcf_get_fmt_id(bcf1_t* line, const int id);
bcf_info_t* bcf_get_info_id(bcf1_t* line, const int id);
#define bcf_get_info_int32(hdr, line, tag, dst, ndst) bcf_get_info_values(hdr, line, tag, (void**)(dst), ndst, BCF_HT_INT)
#define bcf_get_info_float(hdr, line, tag, dst, ndst) bcf_get_info_values(hdr, line, tag, (void**)(dst), ndst, BCF_HT_REAL)
#define bcf_get_info_string(hdr, line, tag, dst, ndst) bcf_get_info_values(hdr, line, tag, (void**)(dst), ndst, BCF_HT_STR)
#define bcf_get_info_flag(hdr, line, tag, dst, ndst) bcf_get_info_values(hdr, line, tag, (void**)(dst), ndst, BCF_HT_FLAG)
int bcf_get_info_values(const bcf_hdr_t* hdr, bcf1_t* line, const char* tag, void** dst, int* ndst, int type);
#define bcf_get_format_int32(hdr, line, tag, dst, ndst) bcf_get_format_values(hdr, line, tag, (void**)(dst), ndst, BCF_HT_INT)
#define bcf_get_format_float(hdr, line, tag, dst, ndst) bcf_get_format_values(hdr, line, tag, (void**)(dst), ndst, BCF_HT_REAL)
#define bcf_get_format_char(hdr, line, tag, dst, ndst) bcf_get_format_values(hdr, line, tag, (void**)(dst), ndst, BCF_HT_STR)
#define bcf_get_genotypes(hdr, line, dst, ndst) bcf_get_format_values(hdr, line, "GT", (void**)(dst), ndst, BCF_HT_INT)
int bcf_get_format_string(const bcf_hdr_t* hdr, bcf1_t* line, const char* tag, char*** dst, int* ndst);
int bcf_get_format_values(const bcf_hdr_t* hdr, bcf1_t* line, const char* tag, void** dst, int* ndst, int type);
int bcf_hdr_id2int(const bcf_hdr_t* hdr, int type, const char* id);
#define bcf_hdr_int2id(hdr, type, int_id) ((hdr)->id[type][int_id].key)
static inline int bcf_hdr_name2id(const bcf_hdr_t* hdr, const char* id) { return bcf_hdr_id2int(hdr, BCF_DT_CTG, id); }
static inline const char* bcf_hdr_id2name(const bcf_hdr_t* hdr, int rid) { return hdr->id[BCF_DT_CTG][rid].key; }
static inline const char* bcf_seqname(const bcf_hdr_t* hdr, bcf1_t* rec) { return hdr->id[BCF_DT_CTG][rec->rid].key; }
#define bcf_hdr_id2length(hdr, type, int_id) ((hdr)->id[BCF_DT_ID][int_id].val->info[type] >> 8 & 0xf)
#define bcf_hdr_id2number(hdr, type, int_id) ((hdr)->id[BCF_DT_ID][int_id].val->info[type] >> 12)
#define bcf_hdr_id2type(hdr, type, int_id) ((hdr)->id[BCF_DT_ID][int_id].val->info[type] >> 4 & 0xf)
#define bcf_hdr_id2coltype(hdr, type, int_id) ((hdr)->id[BCF_DT_ID][int_id].val->info[type] & 0xf)
#define bcf_hdr_idinfo_exists(hdr, type, int_id) ((int_id < 0 || bcf_hdr_id2coltype(hdr, type, int_id) == 0xf) ? 0 : 1)
#define bcf_hdr_id2hrec(hdr, dict_type, col_type, int_id) ((hdr)->id[(dict_type) == BCF_DT_CTG ? BCF_DT_CTG : BCF_DT_ID][int_id].val->hrec[(dict_type) == BCF_DT_CTG ? 0 : (col_type)])
void bcf_fmt_array(kstring_t* s, int n, int type, void* data);
uint8_t* bcf_fmt_sized_array(kstring_t* s, uint8_t* ptr);
void bcf_enc_vchar(kstring_t* s, int l, const char* a);
void bcf_enc_vint(kstring_t* s, int n, int32_t* a, int wsize);
void bcf_enc_vfloat(kstring_t* s, int n, float* a);
#define bcf_itr_destroy(iter) hts_itr_destroy(iter)
hts_itr_t* bcf_itr_queryi(const hts_idx_t* idx, int tid, int beg, int end);
hts_itr_t* bcf_itr_querys(const hts_idx_t* idx, const bcf_hdr_t* hdr, const char* s);
int bcf_itr_next(htsFile* htsfp, hts_itr_t* itr, void* r);
hts_idx_t* bcf_index_load(const char* fn);
hts_idx_t* bcf_index_load2(const char* fn, const char* fnidx);
int bcf_index_build(const char* fn, int min_shift);
#define bcf_int8_vector_end (INT8_MIN + 1)
#define bcf_int16_vector_end (INT16_MIN + 1)
#define bcf_int32_vector_end (INT32_MIN + 1)
#define bcf_str_vector_end 0
#define bcf_int8_missing INT8_MIN
#define bcf_int16_missing INT16_MIN
#define bcf_int32_missing INT32_MIN
#define bcf_str_missing 0x07
extern uint32_t bcf_float_vector_end;
extern uint32_t bcf_float_missing;
static inline void bcf_float_set(float* ptr, uint32_t value) { union { uint32_t i; float f; } u; u.i = value; *ptr = u.f; }
#define bcf_float_set_vector_end(x) bcf_float_set(&(x), bcf_float_vector_end)
#define bcf_float_set_missing(x) bcf_float_set(&(x), bcf_float_missing)
static inline int bcf_float_is_missing(float f) { union { uint32_t i; float f; } u; u.f = f; return u.i == bcf_float_missing ? 1 : 0; }
static inline int bcf_float_is_vector_end(float f) { union { uint32_t i; float f; } u; u.f = f; return u.i == bcf_float_vector_end ? 1 : 0; }
#ifdef __cplusplus
}
#endif
#endif
