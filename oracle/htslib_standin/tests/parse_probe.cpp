// parse_probe.cpp — TEST INFRASTRUCTURE.  Prints, for every record of a text VCF, the bits of the float32 values the htslib
// stand-in hands to the reference for a FORMAT tag (bcf_get_format_float, what bcf_filtered_reader.cpp:418 calls) and the
// int32 values of another (bcf_get_format_int32, :257), so that tests can pin the stand-in's number parsing against the
// product loaders and against (float)strtod — htslib's own conversion (vcf_parse_format stores `strtod(...)` into a float).
//   g++ -std=c++14 -I oracle/htslib_standin parse_probe.cpp ../standin.cpp -lz -o parse_probe ; ./parse_probe in.vcf.gz GP PL
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "htslib/vcf.h"
int main(int argc, char** argv) {
  if (argc < 4) return 2;
  htsFile* fp = hts_open(argv[1], "r");
  if (!fp) return 3;
  bcf_hdr_t* h = bcf_hdr_read(fp);
  bcf1_t* v = bcf_init();
  float* f = NULL; int nf = 0;
  int32_t* q = NULL; int nq = 0;
  while (bcf_read(fp, h, v) >= 0) {
    bcf_unpack(v, BCF_UN_ALL);
    int n = bcf_get_format_float(h, v, argv[2], &f, &nf);
    printf("F");
    for (int i = 0; i < n; ++i) { uint32_t b; memcpy(&b, &f[i], 4); printf(" %08x", b); }
    n = bcf_get_format_int32(h, v, argv[3], &q, &nq);
    printf("\nI");
    for (int i = 0; i < n; ++i) printf(" %d", q[i]);
    printf("\n");
  }
  free(f); free(q);
  bcf_destroy(v); bcf_hdr_destroy(h); hts_close(fp);
  return 0;
}
