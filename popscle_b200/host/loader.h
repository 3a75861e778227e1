// loader.h — dsc-pileup CEL/VAR/PLP (+ text VCF) -> flat cell-major pileup for the C ABI.
// Reproduces sc_dropseq_lib_t::load_from_plp (reference sc_drop_seq.cpp:103-384) and the genotype
// preparation of BCFFilteredReader (bcf_filtered_reader.cpp:367-461, :505-581) for text VCF.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/popscle_b200.h"

namespace pscl_host {

// host threads of the ingest (PSCL_LOADER_THREADS; default: the machine's, at most 32)
inline int loader_threads() {
  if (const char* t = getenv("PSCL_LOADER_THREADS")) return std::max(1, atoi(t));
  const unsigned hc = std::thread::hardware_concurrency();
  return (int)std::min<unsigned>(hc ? hc : 4u, 32u);
}

template <typename F>
void parallel_for(int n, int threads, F fn) {  // fn(i) for i in [0, n) on up to `threads` host threads; the first exception wins
  if (threads <= 1 || n <= 1) { for (int i = 0; i < n; ++i) fn(i); return; }
  std::vector<std::thread> th;
  std::exception_ptr err;
  std::mutex mu;
  std::atomic<int> next(0);
  for (int t = 0; t < std::min(threads, n); ++t)
    th.emplace_back([&] {
      try {
        for (int i = next++; i < n; i = next++) fn(i);
      } catch (...) { std::lock_guard<std::mutex> lk(mu); if (!err) err = std::current_exception(); }
    });
  for (auto& t : th) t.join();
  if (err) std::rethrow_exception(err);
}

// PSCL_TRACE=1: wall-clock since the previous call on stderr
inline void trace_lap(const char* who, const char* what) {
  static const bool on = getenv("PSCL_TRACE") != nullptr;
  static auto prev = std::chrono::steady_clock::now();
  if (!on) return;
  const auto t = std::chrono::steady_clock::now();
  if (what) fprintf(stderr, "[%s] %-34s %8.1f ms\n", who, what, std::chrono::duration<double, std::milli>(t - prev).count());
  prev = t;
}

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
  mutable std::vector<int64_t> cell_read_ptr;              // ABI 7: first base-call of every cell
  mutable int read_bits = 0, packed_state = 0;             // packed_state: 0 not tried, 1 usable, -1 more than 64 distinct bytes
  mutable int tiny_state = 0;   // 0 not tried, 1 usable, -1 a pair without base-calls or with >= 256 of them
  mutable int delta_state = 0;  // 0 not tried, 1 usable, -1 a gap or a count does not fit
  mutable int compact_state = 0;  // the same for read_aq (an allele code > 2 or a quality > 63 rules the packed form out)

  // The compact forms that cross PCIe, built on first use on several host threads.  ABI 6 (8-bit SNP gaps, 2-bit counts,
  // palette-indexed base-calls) is what the library prefers; the ABI-3 arrays (16-bit gaps, 8-bit counts) are built only when
  // a pileup does not fit ABI 6 or when `all_forms` asks for them (--dry-run checksums every form).
  pscl_pileup view(bool all_forms = false) const {
    pscl_pileup p;
    p.pair_read_ptr32 = nullptr; p.read_aq = nullptr;
    p.cell_first_snp = nullptr; p.pair_snp_delta16 = nullptr; p.pair_nreads8 = nullptr;
    p.pair_snp_delta8 = nullptr; p.snp_gap_big = nullptr; p.cell_gap_big_ptr = nullptr;
    p.pair_nreads2 = nullptr; p.nreads_big = nullptr; p.nreads_big_ptr = nullptr; p.n_gap_big = p.n_nreads_big = 0;
    p.read_packed = nullptr; p.read_palette = nullptr; p.read_bits = 0; p.reserved_ = 0; p.cell_read_ptr = nullptr;
    const int T = loader_threads();
    const size_t P = pair_snp.size(), N = read_allele.size();
    auto chunks = [&](size_t n, size_t align, auto fn) {  // fn(begin, end) over [0, n) in pieces whose starts are multiples of `align`
      const size_t per = std::max<size_t>(align, ((n / (size_t)std::max(T, 1) / 4 + align) / align) * align);
      const int nc = (int)((n + per - 1) / per);
      parallel_for(nc, T, [&](int i) { fn((size_t)i * per, std::min(n, ((size_t)i + 1) * per)); });
    };
    if (N < (1ull << 32)) {  // halves the bytes pscl_plp_upload sends over PCIe
      if (pair_read_ptr32.size() != pair_read_ptr.size()) {
        pair_read_ptr32.resize(pair_read_ptr.size());
        chunks(pair_read_ptr.size(), 1, [&](size_t b, size_t e) { for (size_t i = b; i < e; ++i) pair_read_ptr32[i] = (uint32_t)pair_read_ptr[i]; });
      }
      if (compact_state == 0) {  // decided once per loaded pileup (a later call must not forget a failed check)
        std::atomic<bool> fits(true);
        read_aq.resize(N);
        chunks(N, 1, [&](size_t b, size_t e) {
          bool ok = true;
          for (size_t i = b; i < e; ++i) {
            ok = ok && read_allele[i] <= 2 && read_qual[i] <= 63;  // e.g. --cap-BQ above 63
            read_aq[i] = (uint8_t)((read_allele[i] << 6) | (read_qual[i] & 63));
          }
          if (!ok) fits = false;
        });
        compact_state = fits ? 1 : -1;
      }
      const bool ok = compact_state == 1;
      if (ok) { p.pair_read_ptr32 = pair_read_ptr32.data(); p.read_aq = read_aq.data(); }
      if (ok && tiny_state == 0) {  // ABI 6: one byte per gap, two bits per count, the large values on the side
        std::atomic<bool> fits(true);
        cell_first_snp.assign((size_t)n_cells, 0);
        pair_snp_delta8.resize(P); pair_nreads2.assign((P + 3) / 4, 0);
        cell_gap_big_ptr.assign((size_t)n_cells + 1, 0); nreads_big_ptr.assign(P / 1024 + 2, 0);
        // counts: blocks of 1024 pairs (the granule of nreads_big_ptr) -> number of large counts per block, then the values
        const size_t NBK = (P + 1023) / 1024;
        chunks(NBK, 1, [&](size_t b, size_t e) {
          bool good = true;
          for (size_t k = b; k < e; ++k) {
            int64_t big = 0;
            const size_t i1 = std::min(P, (k + 1) * 1024);
            for (size_t i = k * 1024; i < i1; ++i) {
              const int64_t n = pair_read_ptr[i + 1] - pair_read_ptr[i];
              good = good && n >= 1 && n <= 255;
              if (n >= 4) ++big; else pair_nreads2[i >> 2] |= (uint8_t)(n << (2 * (i & 3)));
            }
            nreads_big_ptr[k + 1] = big;  // shifted by one: the running sum below turns it into the block's start
          }
          if (!good) fits = false;
        });
        for (size_t k = 0; k < NBK; ++k) nreads_big_ptr[k + 1] += nreads_big_ptr[k];
        for (size_t k = NBK + 1; k < nreads_big_ptr.size(); ++k) nreads_big_ptr[k] = nreads_big_ptr[NBK];
        // gaps: per droplet
        chunks((size_t)n_cells, 1, [&](size_t cb, size_t ce) {
          bool good = true;
          for (size_t c = cb; c < ce; ++c) {
            const int64_t b = cell_ptr[c], e = cell_ptr[c + 1];
            if (e > b) cell_first_snp[c] = pair_snp[b];
            int64_t big = 0;
            for (int64_t i = b + 1; i < e; ++i) {
              const int64_t d = (int64_t)pair_snp[i] - pair_snp[i - 1];
              good = good && d >= 0;
              big += d >= 255;
            }
            cell_gap_big_ptr[c + 1] = big;
          }
          if (!good) fits = false;
        });
        for (int32_t c = 0; c < n_cells; ++c) cell_gap_big_ptr[(size_t)c + 1] += cell_gap_big_ptr[c];
        if (fits) {
          nreads_big.resize((size_t)nreads_big_ptr[NBK]);
          snp_gap_big.resize((size_t)cell_gap_big_ptr[(size_t)n_cells]);
          chunks(NBK, 1, [&](size_t b, size_t e) {
            for (size_t k = b; k < e; ++k) {
              size_t w = (size_t)nreads_big_ptr[k];
              const size_t i1 = std::min(P, (k + 1) * 1024);
              for (size_t i = k * 1024; i < i1; ++i) {
                const int64_t n = pair_read_ptr[i + 1] - pair_read_ptr[i];
                if (n >= 4) nreads_big[w++] = (uint8_t)n;
              }
            }
          });
          chunks((size_t)n_cells, 1, [&](size_t cb, size_t ce) {
            for (size_t c = cb; c < ce; ++c) {
              const int64_t b = cell_ptr[c], e = cell_ptr[c + 1];
              size_t w = (size_t)cell_gap_big_ptr[c];
              if (e > b) pair_snp_delta8[b] = 0;
              for (int64_t i = b + 1; i < e; ++i) {
                const int64_t d = (int64_t)pair_snp[i] - pair_snp[i - 1];
                if (d >= 255) { pair_snp_delta8[i] = 255; snp_gap_big[w++] = (uint32_t)d; } else pair_snp_delta8[i] = (uint8_t)d;
              }
            }
          });
        }
        tiny_state = fits ? 1 : -1;
      }
      if (ok && delta_state == 0 && (all_forms || tiny_state != 1)) {  // ABI 3: 3 B per pair instead of 8
        std::atomic<bool> fits(true);
        if (cell_first_snp.size() != (size_t)n_cells) cell_first_snp.assign((size_t)n_cells, 0);
        pair_snp_delta16.assign(P, 0);
        pair_nreads8.assign(P, 0);
        chunks((size_t)n_cells, 1, [&](size_t cb, size_t ce) {
          for (size_t c = cb; c < ce; ++c) {
            const int64_t b = cell_ptr[c], e = cell_ptr[c + 1];
            if (e > b) cell_first_snp[c] = pair_snp[b];
            for (int64_t i = b; i < e; ++i) {
              const int64_t d = i > b ? (int64_t)pair_snp[i] - pair_snp[i - 1] : 0, n = pair_read_ptr[i + 1] - pair_read_ptr[i];
              if (d < 0 || d > 65535 || n < 0 || n > 255) { fits = false; return; }
              pair_snp_delta16[i] = (uint16_t)d; pair_nreads8[i] = (uint8_t)n;
            }
          }
        });
        delta_state = fits ? 1 : -1;
      }
      if (ok && delta_state == 1) {
        p.cell_first_snp = cell_first_snp.data(); p.pair_snp_delta16 = pair_snp_delta16.data(); p.pair_nreads8 = pair_nreads8.data();
      }
      if (ok && packed_state == 0) {  // few distinct allele<<6|qual bytes survive --min-BQ / --cap-BQ: index them
        std::atomic<uint64_t> seen_bits[4];
        for (auto& w : seen_bits) w = 0;
        chunks(N, 1, [&](size_t b, size_t e) {
          uint64_t loc[4] = {0, 0, 0, 0};
          for (size_t i = b; i < e; ++i) loc[read_aq[i] >> 6] |= 1ull << (read_aq[i] & 63);
          for (int k = 0; k < 4; ++k) seen_bits[k] |= loc[k];
        });
        int slot[256];
        read_palette.clear();
        for (int i = 0; i < 256; ++i) {  // ascending, as np.unique
          slot[i] = -1;
          if ((seen_bits[i >> 6] >> (i & 63)) & 1) { slot[i] = (int)read_palette.size(); read_palette.push_back((uint8_t)i); }
        }
        if (read_palette.size() <= 64) {
          read_bits = read_palette.size() <= 16 ? 4 : read_palette.size() <= 32 ? 5 : 6;
          read_palette.resize((size_t)1 << read_bits, 0);
          read_packed.assign((N * read_bits + 7) / 8 + 1, 0);
          const int rb = read_bits;
          chunks(N, 8, [&](size_t b, size_t e) {  // a piece starts at a multiple of 8 base-calls = on a byte boundary
            for (size_t r = b; r < e; ++r) {
              const size_t o = r * rb;
              const unsigned v = (unsigned)slot[read_aq[r]] << (o & 7);
              read_packed[o >> 3] |= (uint8_t)v;
              if (v >> 8) read_packed[(o >> 3) + 1] |= (uint8_t)(v >> 8);  // never past the piece's last byte
            }
          });
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
    if (cell_read_ptr.size() != (size_t)n_cells + 1) {
      cell_read_ptr.resize((size_t)n_cells + 1);
      for (int32_t c = 0; c <= n_cells; ++c) cell_read_ptr[(size_t)c] = pair_read_ptr[(size_t)cell_ptr[(size_t)c]];
    }
    p.cell_read_ptr = cell_read_ptr.data();
    p.n_cells = n_cells; p.n_snps = n_snps;
    p.n_pairs = (int64_t)P; p.n_reads = (int64_t)N;
    p.cell_ptr = cell_ptr.data(); p.pair_snp = pair_snp.data(); p.pair_read_ptr = pair_read_ptr.data();
    p.read_allele = read_allele.data(); p.read_qual = read_qual.data(); p.snp_af = af.data();
    return p;
  }
};

void load_plp(const LoadOptions& o, Loaded& out);
std::vector<std::string> read_first_column(const std::string& path);

}  // namespace pscl_host
