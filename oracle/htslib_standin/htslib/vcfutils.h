/* Stand-in for htslib/vcfutils.h — see kstring.h. */
#ifndef STANDIN_VCFUTILS_H
#define STANDIN_VCFUTILS_H
#include "vcf.h"
#ifdef __cplusplus
extern "C" {
#endif
struct kbitset_t;
int bcf_trim_alleles(const bcf_hdr_t* header, bcf1_t* line);
void bcf_remove_alleles(const bcf_hdr_t* header, bcf1_t* line, int mask);
int bcf_calc_ac(const bcf_hdr_t* header, bcf1_t* line, int* ac, int which);
#define BCF_UN_AC_INFO 1
#define BCF_UN_AC_FMT 2
#define GT_HOM_RR 0
#define GT_HOM_AA 1
#define GT_HET_RA 2
#define GT_HET_AA 3
#define GT_HAPL_R 4
#define GT_HAPL_A 5
#define GT_UNKN 6
int bcf_gt_type(bcf_fmt_t* fmt_ptr, int isample, int* ial, int* jal);
static inline int bcf_acgt2int(char c) {
  if ((int)c > 96) c -= 32;
  if (c == 'A') return 0;
  if (c == 'C') return 1;
  if (c == 'G') return 2;
  if (c == 'T') return 3;
  return -1;
}
#define bcf_int2acgt(i) "ACGT"[i]
#ifdef __cplusplus
}
#endif
#endif
