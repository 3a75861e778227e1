/* Stand-in for htslib/sam.h — see kstring.h.  Declarations only: the --sam mode of demuxlet and
 * dsc-pileup are outside the PLP path; every function here aborts if reached. */
#ifndef STANDIN_SAM_H
#define STANDIN_SAM_H
#include <stdint.h>
#include "hts.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct bam_hdr_t {
  int32_t n_targets, ignore_sam_err;
  uint32_t l_text;
  uint32_t* target_len;
  int8_t* cigar_tab;
  char** target_name;
  char* text;
  void* sdict;
} bam_hdr_t;
typedef bam_hdr_t sam_hdr_t;
#define BAM_CMATCH 0
#define BAM_CINS 1
#define BAM_CDEL 2
#define BAM_CREF_SKIP 3
#define BAM_CSOFT_CLIP 4
#define BAM_CHARD_CLIP 5
#define BAM_CPAD 6
#define BAM_CEQUAL 7
#define BAM_CDIFF 8
#define BAM_CBACK 9
#define BAM_CIGAR_STR "MIDNSHP=XB"
#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK 0xf
#define BAM_CIGAR_TYPE 0x3C1A7
#define bam_cigar_op(c) ((c)&BAM_CIGAR_MASK)
#define bam_cigar_oplen(c) ((c) >> BAM_CIGAR_SHIFT)
#define bam_cigar_opchr(c) (BAM_CIGAR_STR "??????"[bam_cigar_op(c)])
#define bam_cigar_gen(l, o) ((l) << BAM_CIGAR_SHIFT | (o))
#define bam_cigar_type(o) (BAM_CIGAR_TYPE >> ((o) << 1) & 3)
#define BAM_FPAIRED 1
#define BAM_FPROPER_PAIR 2
#define BAM_FUNMAP 4
#define BAM_FMUNMAP 8
#define BAM_FREVERSE 16
#define BAM_FMREVERSE 32
#define BAM_FREAD1 64
#define BAM_FREAD2 128
#define BAM_FSECONDARY 256
#define BAM_FQCFAIL 512
#define BAM_FDUP 1024
#define BAM_FSUPPLEMENTARY 2048
typedef struct bam1_core_t {
  int32_t tid, pos;
  uint16_t bin;
  uint8_t qual, l_qname;
  uint16_t flag, n_cigar;
  int32_t l_qseq, mtid, mpos, isize;
} bam1_core_t;
typedef struct bam1_t {
  bam1_core_t core;
  int l_data;
  uint32_t m_data;
  uint8_t* data;
  uint64_t id;
} bam1_t;
#define bam_is_rev(b) (((b)->core.flag & BAM_FREVERSE) != 0)
#define bam_is_mrev(b) (((b)->core.flag & BAM_FMREVERSE) != 0)
#define bam_get_qname(b) ((char*)(b)->data)
#define bam_get_cigar(b) ((uint32_t*)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname)
#define bam_get_qual(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1))
#define bam_get_aux(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1) + (b)->core.l_qseq)
#define bam_get_l_aux(b) ((b)->l_data - ((b)->core.n_cigar << 2) - (b)->core.l_qname - (b)->core.l_qseq - (((b)->core.l_qseq + 1) >> 1))
#define bam_seqi(s, i) ((s)[(i) >> 1] >> ((~(i)&1) << 2) & 0xf)
extern const char seq_nt16_str[];
extern const unsigned char seq_nt16_table[256];
extern const int seq_nt16_int[];
typedef htsFile samFile;
bam_hdr_t* bam_hdr_init(void);
bam_hdr_t* bam_hdr_read(BGZF* fp);
int bam_hdr_write(BGZF* fp, const bam_hdr_t* h);
void bam_hdr_destroy(bam_hdr_t* h);
int bam_name2id(bam_hdr_t* h, const char* ref);
bam_hdr_t* bam_hdr_dup(const bam_hdr_t* h0);
bam1_t* bam_init1(void);
void bam_destroy1(bam1_t* b);
int bam_read1(BGZF* fp, bam1_t* b);
int bam_write1(BGZF* fp, const bam1_t* b);
bam1_t* bam_copy1(bam1_t* bdst, const bam1_t* bsrc);
bam1_t* bam_dup1(const bam1_t* bsrc);
int bam_cigar2qlen(int n_cigar, const uint32_t* cigar);
int bam_cigar2rlen(int n_cigar, const uint32_t* cigar);
int32_t bam_endpos(const bam1_t* b);
hts_idx_t* sam_index_load(htsFile* fp, const char* fn);
hts_idx_t* sam_index_load2(htsFile* fp, const char* fn, const char* fnidx);
#define bam_itr_destroy(iter) hts_itr_destroy(iter)
#define sam_itr_destroy(iter) hts_itr_destroy(iter)
hts_itr_t* sam_itr_queryi(const hts_idx_t* idx, int tid, int beg, int end);
hts_itr_t* sam_itr_querys(const hts_idx_t* idx, bam_hdr_t* hdr, const char* region);
int sam_itr_next(htsFile* htsfp, hts_itr_t* itr, bam1_t* r);
#define sam_open(fn, mode) (hts_open((fn), (mode)))
#define sam_close(fp) hts_close(fp)
bam_hdr_t* sam_hdr_parse(int l_text, const char* text);
bam_hdr_t* sam_hdr_read(samFile* fp);
int sam_hdr_write(samFile* fp, const bam_hdr_t* h);
int sam_parse1(kstring_t* s, bam_hdr_t* h, bam1_t* b);
int sam_format1(const bam_hdr_t* h, const bam1_t* b, kstring_t* str);
int sam_read1(samFile* fp, bam_hdr_t* h, bam1_t* b);
int sam_write1(samFile* fp, const bam_hdr_t* h, const bam1_t* b);
uint8_t* bam_aux_get(const bam1_t* b, const char tag[2]);
int32_t bam_aux2i(const uint8_t* s);
double bam_aux2f(const uint8_t* s);
char bam_aux2A(const uint8_t* s);
char* bam_aux2Z(const uint8_t* s);
void bam_aux_append(bam1_t* b, const char tag[2], char type, int len, const uint8_t* data);
int bam_aux_del(bam1_t* b, uint8_t* s);
#ifdef __cplusplus
}
#endif
#endif
