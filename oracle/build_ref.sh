#!/bin/bash
# Builds oracle/_ref/popscle_ref: the reference's OWN translation units of the demuxlet /
# freemuxlet path, compiled unmodified from where they lie under /root/reference (nothing is
# copied into this repo), with the reference's flags (CMakeLists.txt:4-5: C++14, -O3 -pthread),
# linked against oracle/htslib_standin (text-I/O stand-in for the absent htslib) and zlib.
# TEST INFRASTRUCTURE: used to pin oracle/popscle_oracle.c and as bench.py's reference arm.
# Where /root/reference does not exist (the GPU box) the prebuilt binary is kept as is.
set -euo pipefail
REF=${REFERENCE_DIR:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -d "$REF" ]; then
  echo "build_ref.sh: no reference tree at $REF; keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
# command bodies + everything they link to inside the reference tree
UNITS="cmd_cram_demuxlet cmd_cram_freemux2 cmd_cram_freemuxlet sc_drop_seq PhredHelper tsv_reader
       bcf_filtered_reader bcf_chunked_reader sam_filtered_reader bam_ordered_reader genomeChunk
       genome_interval interval_tree interval reference_sequence utils hts_utils Error params"
CXX=${CXX:-g++}
FLAGS="-std=c++14 -O3 -pthread -w -I $HERE/htslib_standin -I $REF"
stamp() { stat -c %Y "$1" 2>/dev/null || echo 0; }
newest_hdr=$(ls -t "$HERE"/htslib_standin/htslib/*.h | head -1)
pids=()
for u in $UNITS; do
  o="$OUT/obj/$u.o"
  if [ "$(stamp "$o")" -lt "$(stamp "$REF/$u.cpp")" ] || [ "$(stamp "$o")" -lt "$(stamp "$newest_hdr")" ]; then
    $CXX $FLAGS -c "$REF/$u.cpp" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
$CXX $FLAGS -c "$HERE/htslib_standin/standin.cpp" -o "$OUT/obj/standin.o"
$CXX $FLAGS -c "$HERE/ref_main.cpp" -o "$OUT/obj/ref_main.o"
$CXX -pthread -o "$OUT/popscle_ref" "$OUT"/obj/*.o -lz
echo "built $OUT/popscle_ref"
