// loader.h — dsc-pileup CEL/VAR/PLP (+ text VCF) -> flat cell-major pileup for the C ABI.
// Reproduces sc_dropseq_lib_t::load_from_plp (reference sc_drop_seq.cpp:103-384) and the genotype
// preparation of BCFFilteredReader (bcf_filtered_reader.cpp:367-461, :505-581) for text VCF.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/popscle_b200.h"

namespace pscl_host {

struct LoadOptions {
  std::string plp_prefix, vcf;           // vcf empty = no genotypes (freemuxlet)
  std::string field = "GP";
  double geno_error_offset = 0.1;
  double geno_error_coeff = 0.0;   // --geno-error-coeff: err = offset + (1-offset)(1-R2)*coeff (sc_drop_seq.cpp:299-306)
  std::string r2_info = "R2";      // --r2-info
  std::vector<std::string> sm;           // --sm / --sm-list (held in a std::set by the reference: sorted)
  int min_bq = 1, cap_bq = 60;           // library defaults (sc_drop_seq.h:181); the commands pass 13 / 20
  int min_read = 0, min_umi = 0, min_snp = 0;
  std::vector<std::string> group_list;   // empty = all barcodes
  bool has_group_list = false;
  int min_mac = 1;
  double min_callrate = 0.5;
  int max_alleles = 2;
};

struct Loaded {
  // pileup (cell-major CSR)
  int32_t n_cells = 0, n_snps = 0;
  std::vector<int64_t> cell_ptr, pair_read_ptr;
  std::vector<int32_t> pair_snp;
  std::vector<uint8_t> read_allele, read_qual;
  // sites
  std::vector<std::string> chrom, rid2chr;
  std::vector<int32_t> pos;
  std::vector<char> ref, alt;
  std::vector<double> af;
  // droplets
  std::vector<std::string> barcodes;
  std::vector<int64_t> cell_uniq_reads, cell_totl_reads;
  // genotypes (when a VCF was given)
  std::vector<std::string> samples;
  std::vector<double> gp;       // [V][nv][3] after the geno-error mixing (host copy: checksums, debugging)
  std::vector<uint8_t> has_gp;  // [V]
  // raw forms of ABI 4 (what crosses PCIe; the library mixes on the device): the reader's float posteriors, the per-SNP
  // error rate, and hard calls when every posterior row is one-hot (--field GT without missing calls)
  std::vector<float> gp_f32;    // [V][nv][3]
  std::vector<double> err_snp;  // [V]
  std::vector<uint8_t> gt8;     // [V][nv], valid while gt8_ok
  bool gt8_ok = true;
  pscl_geno geno_view() const {
    pscl_geno g;
    g.n_samples = (int32_t)samples.size();
    g.gp = nullptr; g.has_gp = has_gp.data();
    g.gp_f32 = gt8_ok ? nullptr : gp_f32.data();
    g.gt8 = gt8_ok ? gt8.data() : nullptr;
    g.geno_err_snp = err_snp.data(); g.geno_err = 0.0;
    return g;
  }

  // compact arrays of ABI 2 (filled by view() on first use): 32-bit read offsets, allele<<6|qual
  mutable std::vector<uint32_t> pair_read_ptr32;
  mutable std::vector<uint8_t> read_aq;
  // delta-coded pair arrays of ABI 3: first SNP per cell, 16-bit SNP gaps, 8-bit base-call counts
  mutable std::vector<int32_t> cell_first_snp;
  mutable std::vector<uint16_t> pair_snp_delta16;
  mutable std::vector<uint8_t> pair_nreads8;
  // ABI 6: 8-bit gaps / 2-bit counts with the large values on the side (1.25 B per pair)
  mutable std::vector<uint8_t> pair_snp_delta8, pair_nreads2, nreads_big;
  mutable std::vector<uint32_t> snp_gap_big;
  mutable std::vector<int64_t> cell_gap_big_ptr, nreads_big_ptr;
  mutable std::vector<uint8_t> read_packed, read_palette;  // ABI 6: base-calls as 4/5/6-bit palette indices
  mutable int read_bits = 0, packed_state = 0;             // packed_state: 0 not tried, 1 usable, -1 more than 64 distinct bytes
  mutable int tiny_state = 0;   // 0 not tried, 1 usable, -1 a pair without base-calls or with >= 256 of them
  mutable int delta_state = 0;  // 0 not tried, 1 usable, -1 a gap or a count does not fit
  mutable int compact_state = 0;  // the same for read_aq (an allele code > 2 or a quality > 63 rules the packed form out)

  pscl_pileup view() const {
    pscl_pileup p;
    p.pair_read_ptr32 = nullptr; p.read_aq = nullptr;
    p.cell_first_snp = nullptr; p.pair_snp_delta16 = nullptr; p.pair_nreads8 = nullptr;
    p.pair_snp_delta8 = nullptr; p.snp_gap_big = nullptr; p.cell_gap_big_ptr = nullptr;
    p.pair_nreads2 = nullptr; p.nreads_big = nullptr; p.nreads_big_ptr = nullptr; p.n_gap_big = p.n_nreads_big = 0;
    p.read_packed = nullptr; p.read_palette = nullptr; p.read_bits = 0; p.reserved_ = 0;
    if (read_allele.size() < (1ull << 32)) {  // halves the bytes pscl_plp_upload sends over PCIe
      if (pair_read_ptr32.size() != pair_read_ptr.size()) pair_read_ptr32.assign(pair_read_ptr.begin(), pair_read_ptr.end());
      if (compact_state == 0) {  // decided once per loaded pileup (a later call must not forget a failed check)
        bool fits = true;
        read_aq.resize(read_allele.size());
        for (size_t i = 0; i < read_aq.size(); ++i) {
          fits = fits && read_allele[i] <= 2 && read_qual[i] <= 63;  // e.g. --cap-BQ above 63
          read_aq[i] = (uint8_t)((read_allele[i] << 6) | (read_qual[i] & 63));
        }
        compact_state = fits ? 1 : -1;
      }
      const bool ok = compact_state == 1;
      if (ok) { p.pair_read_ptr32 = pair_read_ptr32.data(); p.read_aq = read_aq.data(); }
      if (ok && delta_state == 0) {  // 3 B per pair instead of 8: SNP ids rise within a cell, counts are small
        const size_t P = pair_snp.size();
        cell_first_snp.assign((size_t)n_cells, 0);
        pair_snp_delta16.assign(P, 0);
        pair_nreads8.assign(P, 0);
        bool fits = true;
        for (int32_t c = 0; c < n_cells && fits; ++c) {
          const int64_t b = cell_ptr[c], e = cell_ptr[c + 1];
          if (e > b) cell_first_snp[c] = pair_snp[b];
          for (int64_t i = b; i < e; ++i) {
            const int64_t d = i > b ? (int64_t)pair_snp[i] - pair_snp[i - 1] : 0, n = pair_read_ptr[i + 1] - pair_read_ptr[i];
            if (d < 0 || d > 65535 || n < 0 || n > 255) { fits = false; break; }
            pair_snp_delta16[i] = (uint16_t)d; pair_nreads8[i] = (uint8_t)n;
          }
        }
        delta_state = fits ? 1 : -1;
      }
      if (ok && delta_state == 1) {
        p.cell_first_snp = cell_first_snp.data(); p.pair_snp_delta16 = pair_snp_delta16.data(); p.pair_nreads8 = pair_nreads8.data();
      }
      if (ok && tiny_state == 0) {  // ABI 6 (the library prefers it): one byte per gap, two bits per count
        const size_t P = pair_snp.size();
        if (cell_first_snp.size() != (size_t)n_cells) cell_first_snp.assign((size_t)n_cells, 0);
        pair_snp_delta8.assign(P, 0); pair_nreads2.assign((P + 3) / 4, 0);
        snp_gap_big.clear(); nreads_big.clear();
        cell_gap_big_ptr.assign((size_t)n_cells + 1, 0); nreads_big_ptr.assign(P / 1024 + 2, 0);
        bool fits = true;
        for (int32_t c = 0; c < n_cells && fits; ++c) {
          const int64_t b = cell_ptr[c], e = cell_ptr[c + 1];
          cell_gap_big_ptr[c] = (int64_t)snp_gap_big.size();
          if (e > b) cell_first_snp[c] = pair_snp[b];
          for (int64_t i = b; i < e; ++i) {
            const int64_t d = i > b ? (int64_t)pair_snp[i] - pair_snp[i - 1] : 0, n = pair_read_ptr[i + 1] - pair_read_ptr[i];
            if (d < 0 || n < 1 || n > 255) { fits = false; break; }
            if (i % 1024 == 0) nreads_big_ptr[(size_t)(i / 1024)] = (int64_t)nreads_big.size();
            if (d >= 255) { pair_snp_delta8[i] = 255; snp_gap_big.push_back((uint32_t)d); } else pair_snp_delta8[i] = (uint8_t)d;
            if (n >= 4) nreads_big.push_back((uint8_t)n); else pair_nreads2[(size_t)(i >> 2)] |= (uint8_t)(n << (2 * (i & 3)));
          }
        }
        cell_gap_big_ptr[(size_t)n_cells] = (int64_t)snp_gap_big.size();
        for (size_t k = P / 1024 + (P % 1024 ? 1 : 0); k < nreads_big_ptr.size(); ++k) nreads_big_ptr[k] = (int64_t)nreads_big.size();
        if (P % 1024 == 0 && P > 0) nreads_big_ptr[P / 1024] = (int64_t)nreads_big.size();
        tiny_state = fits ? 1 : -1;
      }
      if (ok && packed_state == 0) {  // few distinct allele<<6|qual bytes survive --min-BQ / --cap-BQ: index them
        int slot[256];
        for (int i = 0; i < 256; ++i) slot[i] = -1;
        bool seen[256] = {false};
        for (uint8_t b : read_aq) seen[b] = true;
        read_palette.clear();
        for (int i = 0; i < 256; ++i) if (seen[i]) { slot[i] = (int)read_palette.size(); read_palette.push_back((uint8_t)i); }  // ascending, as np.unique
        if (read_palette.size() <= 64) {
          read_bits = read_palette.size() <= 16 ? 4 : read_palette.size() <= 32 ? 5 : 6;
          read_palette.resize((size_t)1 << read_bits, 0);
          read_packed.assign((read_aq.size() * read_bits + 7) / 8 + 1, 0);
          for (size_t r = 0; r < read_aq.size(); ++r) {
            const size_t o = r * read_bits;
            const unsigned v = (unsigned)slot[read_aq[r]] << (o & 7);
            read_packed[o >> 3] |= (uint8_t)v; read_packed[(o >> 3) + 1] |= (uint8_t)(v >> 8);
          }
          packed_state = 1;
        } else packed_state = -1;
      }
      if (ok && tiny_state == 1 && packed_state == 1) { p.read_packed = read_packed.data(); p.read_palette = read_palette.data(); p.read_bits = read_bits; }
      if (ok && tiny_state == 1) {
        p.cell_first_snp = cell_first_snp.data(); p.pair_snp_delta8 = pair_snp_delta8.data(); p.cell_gap_big_ptr = cell_gap_big_ptr.data();
        p.snp_gap_big = snp_gap_big.empty() ? nullptr : snp_gap_big.data(); p.n_gap_big = (int64_t)snp_gap_big.size();
        p.pair_nreads2 = pair_nreads2.data(); p.nreads_big_ptr = nreads_big_ptr.data();
        p.nreads_big = nreads_big.empty() ? nullptr : nreads_big.data(); p.n_nreads_big = (int64_t)nreads_big.size();
      }
    }
    p.n_cells = n_cells; p.n_snps = n_snps;
    p.n_pairs = (int64_t)pair_snp.size(); p.n_reads = (int64_t)read_allele.size();
    p.cell_ptr = cell_ptr.data(); p.pair_snp = pair_snp.data(); p.pair_read_ptr = pair_read_ptr.data();
    p.read_allele = read_allele.data(); p.read_qual = read_qual.data(); p.snp_af = af.data();
    return p;
  }
};

void load_plp(const LoadOptions& o, Loaded& out);
std::vector<std::string> read_first_column(const std::string& path);

}  // namespace pscl_host
