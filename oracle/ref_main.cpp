// ref_main.cpp — entry point of oracle/_ref/popscle_ref (TEST INFRASTRUCTURE).
// The reference's own main (cramore.cpp:37-72) registers every sub-command, which would pull in
// dsc-pileup, the GTF tools etc.; this one dispatches only the three commands of the
// demuxlet / freemuxlet path to the reference's unmodified command bodies (cramore.cpp:43-45).
#include <cstdint>
#include <cstdio>
#include <cstring>
int32_t cmdCramDemuxlet(int32_t argc, char** argv);    // cmd_cram_demuxlet.cpp:6
int32_t cmdCramFreemux2(int32_t argc, char** argv);    // cmd_cram_freemux2.cpp:11   (`popscle freemuxlet`)
int32_t cmdCramFreemuxlet(int32_t argc, char** argv);  // cmd_cram_freemuxlet.cpp:11 (`popscle freemuxlet-old`)
int main(int argc, char** argv) {
  if (argc >= 2) {
    if (!strcmp(argv[1], "demuxlet")) return cmdCramDemuxlet(argc - 1, argv + 1);
    if (!strcmp(argv[1], "freemuxlet")) return cmdCramFreemux2(argc - 1, argv + 1);
    if (!strcmp(argv[1], "freemuxlet-old")) return cmdCramFreemuxlet(argc - 1, argv + 1);
  }
  fprintf(stderr, "usage: popscle_ref demuxlet|freemuxlet|freemuxlet-old [options]\n");
  return 2;
}
