/* Stand-in for htslib/kseq.h — see kstring.h.  Only the separator constants are used
 * (tsv_reader.cpp:31 passes KS_SEP_LINE to hts_getline). */
#ifndef STANDIN_KSEQ_H
#define STANDIN_KSEQ_H
#include "kstring.h"
#define KS_SEP_SPACE 0
#define KS_SEP_TAB 1
#define KS_SEP_LINE 2
#define KS_SEP_MAX 2
#endif
