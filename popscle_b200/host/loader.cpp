// loader.cpp — see loader.h
#include "loader.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>
#include <numeric>
#include <set>

#include "tsv.h"

namespace pscl_host {

std::vector<std::string> read_first_column(const std::string& path) {
  LineReader r(path);
  std::string line;
  std::vector<char*> f;
  std::vector<std::string> out;
  while (next_row(r, line, f)) out.emplace_back(f[0]);
  return out;
}

namespace {

std::vector<char> slurp_text(const std::string& path, int threads, bool* was_bgzf);

// ---- text VCF: records that pass the demuxlet site filters, with float posteriors ------------------------------------------
// The file is inflated whole (BGZF blocks on several threads: what bgzip / bcftools write), its header read, and the data
// lines cut into line-aligned pieces that are parsed in parallel (GT decoding, the site filters, GT / GP / PL to posteriors
// are all per record).  The cursor then hands the records out in file order; a record's error (or a contig the header does
// not list) is kept with it and only raised / numbered when the merge-join with the VAR table reaches it, as the reference's
// record-at-a-time reader would (bcf_filtered_reader.cpp:505-581, :748-762).
struct VcfHeader {
  std::map<std::string, int> contig;  // header order
  std::vector<std::string> samples;   // selected sample names, in reference order
  std::vector<int> cols;              // their VCF columns
};
struct VcfRec {
  int rid = -1;        // >= 0: contig id; < 0: -1 - index into the piece's `names` (not in the header)
  int pos = 0;
  char ref = 0, alt = 0;
  uint8_t state = 0;   // 0 record, 1 dropped by a site filter, 2 error (message in text)
  size_t gps_off = 0;  // into the piece's gps
  std::string text;    // INFO column (kept when --geno-error-coeff reads it) or the error message
};
struct VcfPiece {
  std::vector<VcfRec> recs;
  std::vector<float> gps;            // nv * 3 floats per record of state 0
  std::vector<std::string> names;    // contigs the header does not list, in order of first appearance in the piece
};
struct VcfLineParser {  // one per thread: the scratch vectors keep their capacity from record to record
  const LoadOptions& o;
  const VcfHeader& H;
  const std::string& path;
  std::vector<std::string> c, alleles, fmt, parts, vals;
  std::vector<int> g1, g2, acs;
  std::vector<std::vector<std::string>> smp;
  std::vector<float> gps;
  VcfLineParser(const LoadOptions& opt, const VcfHeader& h, const std::string& p) : o(opt), H(h), path(p) {}

  // 0 = record (fields in r, posteriors in this->gps), 1 = dropped by a site filter; throws host_error on a malformed line
  int parse(const std::string& line, VcfRec& r) {
    const std::vector<int>& cols = H.cols;
    const int nv = (int)cols.size();
    split_char(line, '\t', c);
    if (c.size() < 10) throw host_error("VCF data line with fewer than 10 columns in " + path);
    alleles.clear();
    alleles.push_back(c[3]);
    if (c[4] != ".") { split_char(c[4], ',', parts); alleles.insert(alleles.end(), parts.begin(), parts.end()); }
    const int nal = (int)alleles.size();
    if (nal > o.max_alleles) return 1;
    split_char(c[8], ':', fmt);
    int gi = -1, fi = -1;
    for (size_t i = 0; i < fmt.size(); ++i) {
      if (fmt[i] == "GT") gi = (int)i;
      if (fmt[i] == o.field) fi = (int)i;
    }
    // GT is parsed (and required) only when something reads it: the site filters minMAC / minCallRate
    // (require_GT, bcf_filter_arg.h:110-113) or --field GT itself
    const bool need_gt = o.min_mac > 0 || o.min_callrate > 0 || o.field == "GT";
    if (gi < 0 && need_gt) throw host_error("Cannot find the field GT from the VCF file at position " + c[0] + ":" + c[1]);
    g1.assign(nv, -1); g2.assign(nv, -1); acs.assign(nal, 0);  // members: no allocation per record
    int an = 0;
    if ((int)smp.size() != nv) smp.resize(nv);
    for (int i = 0; i < nv; ++i) {
      split_char(c[9 + cols[i]], ':', smp[i]);
      if (gi < 0) continue;
      const std::string& gt = gi < (int)smp[i].size() ? smp[i][gi] : std::string(".");
      size_t sep = gt.find_first_of("/|");
      std::string a = gt.substr(0, sep), b = sep == std::string::npos ? std::string(".") : gt.substr(sep + 1);
      g1[i] = (a == "." || a.empty()) ? -1 : atoi(a.c_str());
      g2[i] = (b == "." || b.empty()) ? -1 : atoi(b.c_str());
      if (g1[i] >= nal || g2[i] >= nal)  // an allele number the record does not define ("0/3" at a biallelic site)
        throw host_error("GT " + gt + " names an allele beyond the ALT list at position " + c[0] + ":" + c[1]);
      if (g1[i] >= 0) { ++an; ++acs[g1[i]]; }
      if (g2[i] >= 0) { ++an; ++acs[g2[i]]; }
    }
    if (need_gt) {
      if (nv > 0 && o.min_callrate > (double)an / (2.0 * nv)) return 1;
      const int ac = an - acs[0];
      if (ac < o.min_mac || an - ac < o.min_mac) return 1;
    }
    // ---- posteriors with gt_error = 0 (load_from_plp passes 0, sc_drop_seq.cpp:113,285) ----
    const int ngen = nal * (nal + 1) / 2;
    gps.assign((size_t)nv * ngen, 0.f);
    if (o.field == "GT") {  // :379-412
      for (int i = 0; i < nv; ++i) {
        if (g1[i] < 0 || g2[i] < 0) {
          int l = 0;
          for (int j = 0; j < nal; ++j)
            for (int k = 0; k <= j; ++k, ++l)
              gps[(size_t)i * ngen + l] = (float)((j == k ? 1.0 : 2.0) * (acs[j] + 1.0 / nal) / (an + 1.0) * (acs[k] + 1.0 / nal) / (an + 1.0));
        } else {
          const int a = std::max(g1[i], g2[i]), b = std::min(g1[i], g2[i]);
          gps[(size_t)i * ngen + a * (a + 1) / 2 + b] = 1.0f;
        }
      }
    } else if (o.field == "PL") {  // :250-327, ploidy 2
      if (fi < 0) throw host_error("Cannot parse posterior probability at " + c[0] + ":" + c[1]);
      std::vector<double> pls((size_t)nv * ngen), af(nal, 1.0 / nal), gp(ngen), post((size_t)nv * ngen);
      for (int i = 0; i < nv; ++i) {
        split_char(fi < (int)smp[i].size() ? smp[i][fi] : std::string("."), ',', vals);
        for (int l = 0; l < ngen; ++l) pls[(size_t)i * ngen + l] = l < (int)vals.size() ? atoi(vals[l].c_str()) : 0;
      }
      for (int iter = 0; iter < 10; ++iter) {
        std::vector<double> nw(nal, 0.0);
        for (int i = 0; i < nv; ++i) {
          double sum = 0;
          int l = 0;
          for (int j = 0; j < nal; ++j)
            for (int k = 0; k <= j; ++k, ++l) sum += (gp[l] = (j == k ? 1 : 2) * af[j] * af[k] * std::pow(0.1, pls[(size_t)i * ngen + l] / 10.0));
          l = 0;
          for (int j = 0; j < nal; ++j)
            for (int k = 0; k <= j; ++k, ++l) { gp[l] /= sum; nw[j] += gp[l]; nw[k] += gp[l]; post[(size_t)i * ngen + l] = gp[l]; }
        }
        for (int j = 0; j < nal; ++j) af[j] = nw[j] / (2.0 * nv);
      }
      for (size_t i = 0; i < post.size(); ++i) gps[i] = (float)post[i];
    } else {  // GP-like float field, :434-458 with gt_error = 0
      if (fi < 0) throw host_error("Cannot parse posterior probability at " + c[0] + ":" + c[1]);
      for (int i = 0; i < nv; ++i) {
        split_char(fi < (int)smp[i].size() ? smp[i][fi] : std::string("."), ',', vals);
        float sum = 0.f;
        for (int l = 0; l < ngen; ++l) { float x = l < (int)vals.size() ? (float)atof(vals[l].c_str()) : 0.f; gps[(size_t)i * ngen + l] = x; sum += x; }
        for (int l = 0; l < ngen; ++l) gps[(size_t)i * ngen + l] /= sum;
      }
    }
    r.pos = atoi(c[1].c_str());
    r.ref = alleles[0][0];
    r.alt = nal > 1 ? alleles[1][0] : '.';
    if (nal != 2) gps.resize((size_t)nv * 3, 0.f);
    if (o.geno_error_coeff > 0) r.text = c[7];
    return 0;
  }
};

struct VcfCursor {
  const LoadOptions& o;
  VcfHeader H;
  std::vector<std::string>& samples = H.samples;
  std::map<std::string, int> contig;  // header order, then first appearance (numbered as the cursor reaches them)
  std::vector<VcfPiece> pieces;
  size_t pi = 0, ri = 0;
  bool eof = false;
  int rid = -1, pos = 0;
  char ref = 0, alt = 0;
  const float* gps = nullptr;         // [nv*3] of the current record
  std::string info;                   // INFO column of the current record
  std::string path;

  VcfCursor(const std::string& p, const LoadOptions& opt) : o(opt), path(p) {
    const int T = loader_threads();
    const bool trace = getenv("PSCL_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    const std::vector<char> text = slurp_text(path, T, nullptr);
    const auto t1 = std::chrono::steady_clock::now();
    const char* tb = text.data();
    const char* const te = tb + text.size();
    // ---- header ----
    bool saw = false;
    const char* lb = tb;
    std::string line;
    std::vector<std::string> c;
    while (lb < te) {
      const char* nl = (const char*)memchr(lb, '\n', (size_t)(te - lb));
      const char* le = nl ? nl : te;
      line.assign(lb, le);
      if (!line.empty() && line.back() == '\r') line.pop_back();
      lb = nl ? nl + 1 : te;
      if (line.compare(0, 2, "##") == 0) {
        if (line.compare(0, 13, "##contig=<ID=") == 0) {
          size_t e = line.find_first_of(",>", 13);
          std::string id = line.substr(13, e == std::string::npos ? std::string::npos : e - 13);
          if (!H.contig.count(id)) { int k = (int)H.contig.size(); H.contig[id] = k; }
        }
        continue;
      }
      if (line.compare(0, 6, "#CHROM") == 0) {
        split_char(line, '\t', c);
        std::vector<std::string> all(c.begin() + std::min<size_t>(9, c.size()), c.end());
        if (!o.sm.empty()) {  // std::set iteration = sorted, duplicates dropped (bcf_filtered_reader.cpp:107-125)
          std::set<std::string> want(o.sm.begin(), o.sm.end());
          for (const auto& s : want) {
            auto it = std::find(all.begin(), all.end(), s);
            if (it == all.end()) throw host_error("Cannot find sample ID " + s + " from the BCF file");
            H.samples.push_back(s);
            H.cols.push_back((int)(it - all.begin()));
          }
        } else {
          H.samples = all;
          H.cols.resize(all.size());
          std::iota(H.cols.begin(), H.cols.end(), 0);
        }
        saw = true;
      }
      break;
    }
    if (!saw) throw host_error("Failed reading the BCF/VCF header from " + path + " (text VCF, plain or gzip, is supported)");
    contig = H.contig;
    // ---- data lines, in line-aligned pieces ----
    const char* body = lb;
    const size_t len = (size_t)(te - body);
    const int NP = (int)std::max<size_t>(1, std::min<size_t>((size_t)T * 4, len / (1 << 18) + 1));
    std::vector<const char*> cut(NP + 1);
    cut[0] = body; cut[NP] = te;
    for (int i = 1; i < NP; ++i) {
      const char* q = body + len * (size_t)i / NP;
      const char* nl = (const char*)memchr(q, '\n', (size_t)(te - q));
      cut[i] = nl ? nl + 1 : te;
      if (cut[i] < cut[i - 1]) cut[i] = cut[i - 1];
    }
    pieces.resize(NP);
    const size_t nv3 = H.cols.size() * 3;
    parallel_for(NP, T, [&](int k) {
      VcfPiece P;  // on this thread's stack while it grows
      VcfLineParser ps(o, H, path);
      std::string ln;
      auto name_of = [&](const std::string& nm) {  // -1 - index of a contig the header does not list
        size_t j = 0;
        while (j < P.names.size() && P.names[j] != nm) ++j;
        if (j == P.names.size()) P.names.push_back(nm);
        return -1 - (int)j;
      };
      const char* b = cut[k];
      while (b < cut[k + 1]) {
        const char* nl = (const char*)memchr(b, '\n', (size_t)(cut[k + 1] - b));
        const char* e = nl ? nl : cut[k + 1];
        ln.assign(b, e);
        b = nl ? nl + 1 : cut[k + 1];
        if (!ln.empty() && ln.back() == '\r') ln.pop_back();
        if (ln.empty()) continue;
        VcfRec r;
        try {
          r.state = (uint8_t)ps.parse(ln, r);
          // the reference's reader looks the contig up before it filters: a dropped record numbers its contig too
          auto it = H.contig.find(ps.c[0]);
          r.rid = it != H.contig.end() ? it->second : name_of(ps.c[0]);
          if (r.state == 0) {
            r.gps_off = P.gps.size();
            P.gps.insert(P.gps.end(), ps.gps.begin(), ps.gps.begin() + nv3);
          }
        } catch (const host_error& err) {
          r.state = 2;
          r.text = err.what();
          r.rid = 0;
          if (ps.c.size() >= 10) {  // the line was split: its contig is known (and numbered, as above) before anything fails
            auto it = H.contig.find(ps.c[0]);
            r.rid = it != H.contig.end() ? it->second : name_of(ps.c[0]);
          }
        }
        P.recs.push_back(std::move(r));
      }
      pieces[k] = std::move(P);
    });
    if (trace)
      fprintf(stderr, "[load_plp] vcf: inflate %.1f ms, records parsed on %d threads %.1f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count(), T,
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
  }

  // advance to the next record passing the site filters (bcf_filtered_reader.cpp:505-581, :748-762)
  bool read() {
    for (;;) {
      while (pi < pieces.size() && ri >= pieces[pi].recs.size()) { ++pi; ri = 0; }
      if (pi >= pieces.size()) { eof = true; return false; }
      const VcfPiece& P = pieces[pi];
      const VcfRec& r = P.recs[ri++];
      int id = r.rid;
      if (id < 0) {  // a contig the header does not list: numbered at its first appearance
        const std::string& name = P.names[(size_t)(-1 - id)];
        auto it = contig.find(name);
        if (it == contig.end()) { int k = (int)contig.size(); it = contig.emplace(name, k).first; }
        id = it->second;
      }
      if (r.state == 2) throw host_error(r.text);
      if (r.state == 1) continue;
      rid = id; pos = r.pos; ref = r.ref; alt = r.alt;
      gps = P.gps.data() + r.gps_off;
      info = r.text;
      return true;
    }
  }
};

void expect_header(LineReader& r, const char* const* names, int n, const std::string& what, const char* expecting) {
  std::string line;
  std::vector<char*> f;
  if (!next_row(r, line, f)) throw host_error("Cannot read the first line of " + what);
  bool ok = (int)f.size() == n;
  for (int i = 0; ok && i < n; ++i) ok = strcmp(f[i], names[i]) == 0;
  if (!ok) throw host_error("The header line of " + what + " is malformed or outdated. Expecting " + expecting);
}


// ---- whole-file ingest for the big table (.plp.gz): inflate and parse on several host threads ---------------------------
// The text of a gzip / BGZF / plain file.  dsc-pileup writes its tables through htslib's BGZF layer: gzip members of at most
// 64 KB that carry their own compressed size in a 'BC' extra field, so every block inflates independently — on `threads`
// host threads here.  Any other gzip stream is inflated by one thread; a file without the gzip magic is returned as is.
std::vector<char> slurp_text(const std::string& path, int threads, bool* was_bgzf) {
  if (was_bgzf) *was_bgzf = false;
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw host_error("Cannot open file " + path + " for reading");
  std::vector<unsigned char> raw;
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  raw.resize(sz > 0 ? (size_t)sz : 0);
  if (sz > 0 && fread(raw.data(), 1, raw.size(), f) != raw.size()) { fclose(f); throw host_error("Cannot read file " + path); }
  fclose(f);
  if (raw.size() < 18 || raw[0] != 0x1f || raw[1] != 0x8b) return std::vector<char>(raw.begin(), raw.end());
  // BGZF? walk the members by their BSIZE fields
  struct Blk { size_t in, in_len, out, out_len; };
  std::vector<Blk> blk;
  bool bgzf = true;
  size_t o = 0, total = 0;
  while (o < raw.size()) {
    if (o + 18 > raw.size() || raw[o] != 0x1f || raw[o + 1] != 0x8b || raw[o + 2] != 8 || !(raw[o + 3] & 4)) { bgzf = false; break; }
    const size_t xlen = raw[o + 10] | (raw[o + 11] << 8);
    size_t x = o + 12, bsize = 0;
    const size_t xend = x + xlen;
    if (xend > raw.size()) { bgzf = false; break; }
    while (x + 4 <= xend) {
      const size_t slen = raw[x + 2] | (raw[x + 3] << 8);
      if (raw[x] == 'B' && raw[x + 1] == 'C' && slen == 2 && x + 6 <= xend) bsize = (size_t)(raw[x + 4] | (raw[x + 5] << 8)) + 1;
      x += 4 + slen;
    }
    if (!bsize || o + bsize > raw.size() || bsize < xlen + 12 + 8) { bgzf = false; break; }
    const size_t isize = raw[o + bsize - 4] | (raw[o + bsize - 3] << 8) | (raw[o + bsize - 2] << 16) | ((size_t)raw[o + bsize - 1] << 24);
    blk.push_back(Blk{xend, o + bsize - 8 - xend, total, isize});
    total += isize;
    o += bsize;
  }
  std::vector<char> text;
  if (bgzf && !blk.empty()) {
    if (was_bgzf) *was_bgzf = true;
    text.resize(total);
    parallel_for((int)blk.size(), threads, [&](int i) {
      const Blk& b = blk[i];
      if (b.out_len == 0) return;
      z_stream z;
      memset(&z, 0, sizeof z);
      if (inflateInit2(&z, -15) != Z_OK) throw host_error("zlib: inflateInit2 failed");
      z.next_in = raw.data() + b.in; z.avail_in = (uInt)b.in_len;
      z.next_out = (Bytef*)text.data() + b.out; z.avail_out = (uInt)b.out_len;
      const int rc = inflate(&z, Z_FINISH);
      inflateEnd(&z);
      if (rc != Z_STREAM_END || z.avail_out != 0) throw host_error("Corrupt BGZF block in " + path);
    });
    return text;
  }
  // one gzip stream (possibly several concatenated members): a single inflater
  z_stream z;
  memset(&z, 0, sizeof z);
  if (inflateInit2(&z, 15 + 32) != Z_OK) throw host_error("zlib: inflateInit2 failed");
  z.next_in = raw.data(); z.avail_in = (uInt)std::min<size_t>(raw.size(), 1u << 30);
  size_t in_done = 0, out_done = 0;
  text.resize(std::max<size_t>(raw.size() * 4, 1 << 16));
  for (;;) {
    if (out_done == text.size()) text.resize(text.size() * 2);
    z.next_out = (Bytef*)text.data() + out_done;
    z.avail_out = (uInt)std::min<size_t>(text.size() - out_done, 1u << 30);
    const uInt before_out = z.avail_out, before_in = z.avail_in;
    const int rc = inflate(&z, Z_NO_FLUSH);
    out_done += before_out - z.avail_out;
    in_done += before_in - z.avail_in;
    if (rc == Z_STREAM_END) {
      if (in_done >= raw.size()) break;
      inflateReset(&z);  // the next member
    } else if (rc != Z_OK && rc != Z_BUF_ERROR) { inflateEnd(&z); throw host_error("Corrupt gzip stream in " + path); }
    else if (rc == Z_BUF_ERROR && z.avail_in == 0 && in_done >= raw.size()) break;  // truncated input: keep what there is
    if (z.avail_in == 0 && in_done < raw.size()) { z.next_in = raw.data() + in_done; z.avail_in = (uInt)std::min<size_t>(raw.size() - in_done, 1u << 30); }
  }
  inflateEnd(&z);
  text.resize(out_done);
  return text;
}

// ---- VAR, merge-joined with the VCF cursor (sc_drop_seq.cpp:206-330) ----
void load_var_vcf(const LoadOptions& o, Loaded& L) {
  std::string line;
  std::vector<char*> f;
  {
    LineReader r(o.plp_prefix + ".var.gz");
    static const char* H[] = {"#SNP_ID", "CHROM", "POS", "REF", "ALT", "AF"};
    expect_header(r, H, 6, o.plp_prefix + ".var.gz", "#SNP_ID CHROM POS REF ALT AF");
    std::unique_ptr<VcfCursor> vc;
    int nv = 0;
    if (!o.vcf.empty()) {
      vc.reset(new VcfCursor(o.vcf, o));
      if (!vc->read()) throw host_error("Cannot read any single variant from " + o.vcf);
      L.samples = vc->samples;
      nv = (int)L.samples.size();
    }
    std::map<std::string, int> chr2rid;
    std::vector<const float*> hit;       // per VAR row: the posteriors of its VCF record (inside vc->pieces), or null
    std::vector<std::string> hit_info;   // ... and its INFO column, when --geno-error-coeff reads it
    while (next_row(r, line, f)) {
      if (f.size() < 6) throw host_error("Cannot access field at 5 >= " + std::to_string(f.size()));
      auto it = chr2rid.find(f[1]);
      if (it == chr2rid.end()) { int k = (int)chr2rid.size(); it = chr2rid.emplace(f[1], k).first; L.rid2chr.emplace_back(f[1]); }
      const int rid = it->second, pos = atoi(f[2]);
      const char ref = f[3][0], alt = f[4][0];
      L.chrom.emplace_back(f[1]); L.pos.push_back(pos); L.ref.push_back(ref); L.alt.push_back(alt); L.af.push_back(atof(f[5]));
      if (!vc) continue;
      bool found = false;
      for (;;) {  // VCF header rids are compared NUMERICALLY with the VAR first-appearance rids (:262-264)
        if (vc->eof || vc->rid > rid) break;
        if (vc->rid == rid) {
          if (vc->pos > pos) break;
          if (vc->pos == pos) { found = (vc->ref == ref && vc->alt == alt); break; }
        }
        vc->read();
      }
      L.has_gp.push_back(found ? 1 : 0);
      hit.push_back(found ? vc->gps : nullptr);
      if (o.geno_error_coeff > 0) hit_info.push_back(found ? vc->info : std::string());
    }
    // ---- the genotype tables of the matched SNPs, filled in parallel (the join above only recorded where each SNP's record is) ----
    const size_t V = L.chrom.size();
    if (vc) {
      L.gp.assign(V * (size_t)nv * 3, 0.0);
      L.gp_f32.assign(V * (size_t)nv * 3, 0.f);
      L.gt8.assign(V * (size_t)nv, 0);
      L.err_snp.assign(V, 0.0);
      const int T = loader_threads();
      const int NC = (int)std::max<size_t>(1, std::min<size_t>((size_t)T * 4, V / 2048 + 1));
      std::atomic<bool> all_hard(true);
      std::mutex mu;
      size_t bad_v = V;  // the first SNP whose INFO field cannot be used: its error is the one a serial reader would raise
      parallel_for(NC, T, [&](int k) {
        bool hard = true;
        for (size_t v = V * (size_t)k / NC; v < V * (size_t)(k + 1) / NC; ++v) {
          const float* g0 = hit[v];
          if (!g0) continue;
          const size_t base = v * (size_t)nv * 3;
          for (int j = 0; j < nv; ++j) {  // raw forms for the device-side mixing
            const float* g = g0 + (size_t)j * 3;
            L.gp_f32[base + 3 * j] = g[0]; L.gp_f32[base + 3 * j + 1] = g[1]; L.gp_f32[base + 3 * j + 2] = g[2];
            const int code = (g[0] == 1.f && g[1] == 0.f && g[2] == 0.f) ? 0 : (g[0] == 0.f && g[1] == 1.f && g[2] == 0.f) ? 1
                           : (g[0] == 0.f && g[1] == 0.f && g[2] == 1.f) ? 2 : -1;
            if (code < 0) hard = false; else L.gt8[base / 3 + j] = (uint8_t)code;
          }
          double avg[3] = {1e-10, 1e-10, 1e-10};  // :288-292
          for (int i = 0; i < nv * 3; ++i) avg[i % 3] += (L.gp[base + i] = (double)g0[i]);
          const double sum = avg[0] + avg[1] + avg[2];
          avg[0] /= sum; avg[1] /= sum; avg[2] /= sum;
          double err = o.geno_error_offset;
          if (o.geno_error_coeff > 0) {  // look for the R2 INFO field: exactly one float (sc_drop_seq.cpp:301-306)
            const std::string& info = hit_info[v];
            bool ok = false;
            float r2 = 0.f;
            const std::string key = o.r2_info + "=";
            size_t b2 = 0;
            while (b2 <= info.size()) {
              size_t e2 = info.find(';', b2);
              if (e2 == std::string::npos) e2 = info.size();
              if (info.compare(b2, key.size(), key) == 0) {
                const std::string val = info.substr(b2 + key.size(), e2 - b2 - key.size());
                if (!val.empty() && val != "." && val.find(',') == std::string::npos) { r2 = (float)atof(val.c_str()); ok = true; }
                break;
              }
              b2 = e2 + 1;
            }
            if (!ok) { std::lock_guard<std::mutex> lk(mu); if (v < bad_v) bad_v = v; break; }
            err += (1 - o.geno_error_offset) * (1 - r2) * o.geno_error_coeff;
          }
          L.err_snp[v] = err;  // the library clamps it the same way
          if (err > 0.999) err = 0.999;
          if (err < 0) err = 0;
          if (err > 0)
            for (int i = 0; i < nv * 3; ++i) L.gp[base + i] = (1 - err) * L.gp[base + i] + err * avg[i % 3];
        }
        if (!hard) all_hard = false;
      });
      if (bad_v < V)
        throw host_error("Cannot extract " + o.r2_info + " (1 float value) from INFO field at " + L.chrom[bad_v] + ":" + std::to_string(L.pos[bad_v]) + ". Cannot use --geno-error-coeff");
      if (!all_hard) L.gt8_ok = false;
    }
  }
}

}  // namespace

void load_plp(const LoadOptions& o, Loaded& L) {
  // PSCL_TRACE=1: wall-clock of the loader's phases on stderr
  const bool trace = getenv("PSCL_TRACE") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    const auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[load_plp] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count());
    t_prev = t;
  };
  std::string line;
  std::vector<char*> f;
  // ---- CEL (sc_drop_seq.cpp:124-203) ----
  std::vector<int32_t> index_bcs;
  std::vector<int64_t> tmp_totl, tmp_uniq, tmp_nsnp;
  {
    LineReader r(o.plp_prefix + ".cel.gz");
    static const char* H[] = {"#DROPLET_ID", "BARCODE", "NUM.READ", "NUM.UMI", "NUM.UMIwSNP", "NUM.SNP"};
    expect_header(r, H, 6, o.plp_prefix + ".cel.gz", "#DROPLET_ID BARCODE NUM.READ NUM.UMI NUM.UMIwSNP NUM.SNP");
    std::set<std::string> valid(o.group_list.begin(), o.group_list.end());
    int nskip = 0;
    while (next_row(r, line, f)) {
      if (f.size() < 6) throw host_error("Cannot access field at 5 >= " + std::to_string(f.size()));
      if (o.has_group_list && !valid.count(f[1])) { ++nskip; index_bcs.push_back(-1); continue; }
      const int n_reads = atoi(f[2]), n_umis = atoi(f[3]), n_uws = atoi(f[4]), n_snps = atoi(f[5]);
      if (n_reads < o.min_read || n_umis < o.min_umi || n_snps < o.min_snp) { index_bcs.push_back(-1); ++nskip; continue; }
      const int new_id = (int)L.barcodes.size();
      if (new_id + nskip != atoi(f[0]))
        throw host_error("Observed DROPLET_ID " + std::string(f[0]) + " is different from expected DROPLET_ID. Did you modify the digital pileup files by yourself?");
      L.barcodes.emplace_back(f[1]);
      index_bcs.push_back(new_id);
      tmp_totl.push_back(n_reads); tmp_uniq.push_back(n_uws); tmp_nsnp.push_back(n_snps);
    }
  }
  const int32_t C = (int32_t)L.barcodes.size();
  lap("cel.gz");
  // ---- VAR + VCF (:206-330) on a thread of its own, beside the PLP table ----
  std::exception_ptr var_err, plp_err;
  std::thread var_thread([&] {
    try {
      const auto t0 = std::chrono::steady_clock::now();
      load_var_vcf(o, L);
      if (trace) fprintf(stderr, "[load_plp] %-28s %8.1f ms (own thread)\n", "var.gz + vcf", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    } catch (...) { var_err = std::current_exception(); }
  });
  // ---- PLP (:335-372): rows -> (cell, snp, reads); then cell-major, SNP ascending ----
  // The table is inflated whole (BGZF blocks on several threads) and cut at line ends into pieces that are parsed in
  // parallel.  A piece files its rows under the block of droplets they belong to; afterwards one thread per droplet block
  // walks the pieces in file order — so it sees every droplet's rows in file order — first counting, then writing the
  // cell-major image through one cursor per droplet.  dsc-pileup writes SNP-major, the reference's std::map makes it
  // cell-major with SNPs ascending: a droplet's rows already ascend in SNP id in either order, so there is no sort.
  // (Rows are whitespace-separated like tsv_reader's ksplit, numbers read with atoi's rules, and the table ends at the
  // first empty line, tsv_reader.cpp:37-41.  A (cell,SNP) listed on several rows is one pair.)
  struct Row { int32_t cell, snp; uint32_t beg, n; };  // beg: into the piece's al / bq
  struct Piece { std::vector<std::vector<Row>> bk; std::vector<uint8_t> al, bq; int32_t max_snp = -1; };
  const int T = loader_threads();
  const int NB = (int)std::max<int64_t>(1, std::min<int64_t>(T > 1 ? 4 * (int64_t)T : 1, C));  // droplet blocks
  const int32_t cpb = std::max<int32_t>(1, (C + NB - 1) / NB);                                   // droplets per block
  std::vector<Piece> pc;
  std::vector<int64_t> npair((size_t)C, 0), nread((size_t)C, 0);
  std::vector<int32_t> last((size_t)C, -1);
  std::atomic<bool> unsorted(false);
  try {
    bool bg = false;
    const std::vector<char> text = slurp_text(o.plp_prefix + ".plp.gz", T, &bg);
    lap(bg ? "plp.gz inflate (BGZF, parallel)" : "plp.gz inflate");
    const char* const tb = text.data();
    const char* te = tb + text.size();
    auto is_ws = [](char ch) { return ch == ' ' || ch == '\t' || ch == '\v' || ch == '\f' || ch == '\r' || ch == '\n'; };
    // header line
    const char* hl = (const char*)memchr(tb, '\n', text.size());
    if (text.empty()) throw host_error("Cannot read the first line of " + o.plp_prefix + ".plp.gz");
    {
      std::string head(tb, hl ? hl : te);
      std::vector<char*> hf;
      split_ws(head, hf);
      static const char* H[] = {"#DROPLET_ID", "SNP_ID", "ALLELES", "BASEQS"};
      bool ok = hf.size() == 4;
      for (int i = 0; ok && i < 4; ++i) ok = strcmp(hf[i], H[i]) == 0;
      if (head.empty()) throw host_error("Cannot read the first line of " + o.plp_prefix + ".plp.gz");
      if (!ok) throw host_error("The header line of " + o.plp_prefix + ".plp.gz is malformed or outdated. Expecting #DROPLET_ID SNP_ID ALLELES BASEQS");
    }
    const char* body = hl ? hl + 1 : te;
    // the first empty line ends the table
    for (const char* p = body; p < te;) {
      const char* nl = (const char*)memchr(p, '\n', (size_t)(te - p));
      const char* le = nl ? nl : te;
      if (le == p || (le == p + 1 && *p == '\r')) { te = p; break; }
      if (!nl) break;
      // only line starts matter: jump to the next "\n\n" / "\n\r\n" candidate
      const char* dbl = (const char*)memmem(nl, (size_t)(te - nl), "\n\n", 2);
      const char* dbr = (const char*)memmem(nl, (size_t)(te - nl), "\n\r\n", 3);
      const char* cand = dbl && dbr ? std::min(dbl, dbr) : dbl ? dbl : dbr;
      if (!cand) break;
      p = cand + 1;
    }
    const size_t len = (size_t)(te - body);
    const int NP = (int)std::max<size_t>(1, std::min<size_t>((size_t)T * 4, len / (1 << 20) + 1));
    std::vector<const char*> cut(NP + 1);
    cut[0] = body; cut[NP] = te;
    for (int i = 1; i < NP; ++i) {
      const char* p = body + len * (size_t)i / NP;
      const char* nl = (const char*)memchr(p, '\n', (size_t)(te - p));
      cut[i] = nl ? nl + 1 : te;
      if (cut[i] < cut[i - 1]) cut[i] = cut[i - 1];
    }
    pc.resize(NP);
    const int n_index = (int)index_bcs.size();
    parallel_for(NP, T, [&](int pi) {
      Piece P;  // on this thread's stack while it grows: neighbours in pc[] would share cache lines
      P.bk.resize(NB);
      const size_t guess = (size_t)(cut[pi + 1] - cut[pi]) / 12 + 16;
      for (auto& v : P.bk) v.reserve(guess / NB + guess / (4 * NB) + 16);
      P.al.reserve(guess * 2); P.bq.reserve(guess * 2);
      const char* lb = cut[pi];
      while (lb < cut[pi + 1]) {
        const char* nl = (const char*)memchr(lb, '\n', (size_t)(cut[pi + 1] - lb));
        const char* const end = nl ? nl : cut[pi + 1];
        const char* p = lb;
        lb = nl ? nl + 1 : cut[pi + 1];
        const char* fb[4]; const char* fe[4];
        int nf = 0;
        while (p < end) {
          while (p < end && is_ws(*p)) ++p;
          if (p >= end) break;
          const char* q = p;
          while (q < end && !is_ws(*q)) ++q;
          if (nf < 4) { fb[nf] = p; fe[nf] = q; }
          ++nf;
          p = q;
        }
        if (nf < 4) throw host_error("Cannot access field at 3 >= " + std::to_string(nf));
        auto to_int = [](const char* b2, const char* e2) {  // atoi: optional sign, leading digits
          bool neg = false;
          if (b2 < e2 && (*b2 == '-' || *b2 == '+')) { neg = *b2 == '-'; ++b2; }
          long v = 0;
          while (b2 < e2 && *b2 >= '0' && *b2 <= '9') { v = v * 10 + (*b2 - '0'); ++b2; }
          return (int)(neg ? -v : v);
        };
        const int drop = to_int(fb[0], fe[0]);
        if (drop < 0 || drop >= n_index) throw host_error("DROPLET_ID " + std::string(fb[0], fe[0]) + " of .plp.gz is not in .cel.gz");
        const int ibc = index_bcs[drop];
        if (ibc < 0) continue;
        const int snp = to_int(fb[1], fe[1]);
        if (snp < 0) throw host_error("SNP_ID " + std::string(fb[1], fe[1]) + " of .plp.gz is not in .var.gz");
        if (snp > P.max_snp) P.max_snp = snp;  // checked against the VAR table once that thread is done
        const char *pa = fb[2], *pq = fb[3];
        const size_t l = (size_t)(fe[3] - fb[3]), la = (size_t)(fe[2] - fb[2]);
        const size_t b0 = P.al.size();
        for (size_t i = 0; i < l && i < la; ++i) {
          int q = (int)(signed char)(pq[i] - 33);
          if (q >= o.min_bq) {
            if (q > o.cap_bq) q = o.cap_bq;
            P.al.push_back((uint8_t)(pa[i] - '0'));
            P.bq.push_back((uint8_t)q);
          }
        }
        if (P.al.size() == b0) continue;
        if (P.al.size() >= (1ull << 32)) throw host_error("a piece of .plp.gz holds 2^32 base-calls or more");
        P.bk[ibc / cpb].push_back(Row{ibc, snp, (uint32_t)b0, (uint32_t)(P.al.size() - b0)});
      }
      pc[pi] = std::move(P);
    });
    lap("plp.gz parse (parallel)");
    // pass 1: pairs and base-calls per droplet; is every droplet's stream ascending in SNP id?
    parallel_for(NB, T, [&](int b) {
      bool uns = false;
      for (const Piece& P : pc)
        for (const Row& rw : P.bk[b]) {
          if (rw.snp != last[rw.cell]) { uns = uns || rw.snp < last[rw.cell]; ++npair[rw.cell]; last[rw.cell] = rw.snp; }
          nread[rw.cell] += rw.n;
        }
      if (uns) unsorted = true;
    });
    lap("plp.gz count (parallel)");
  } catch (...) { plp_err = std::current_exception(); }
  var_thread.join();
  if (var_err) std::rethrow_exception(var_err);  // the reference reads the VAR / VCF tables first: their errors win
  if (plp_err) std::rethrow_exception(plp_err);
  const int32_t V = (int32_t)L.chrom.size();
  for (const Piece& P : pc)
    if (P.max_snp >= V) throw host_error("SNP_ID " + std::to_string(P.max_snp) + " of .plp.gz is not in .var.gz");
  lap("var.gz + vcf (wait)");
  L.n_cells = C; L.n_snps = V;
  L.cell_ptr.assign((size_t)C + 1, 0);
  L.cell_uniq_reads.assign(C, 0);
  if (!unsorted) {
    std::vector<int64_t> ppos(C), rpos(C);  // next pair / next base-call of each droplet
    int64_t pp = 0, rr = 0;
    for (int32_t c = 0; c < C; ++c) {
      ppos[c] = pp; rpos[c] = rr; L.cell_ptr[c] = pp;
      pp += npair[c]; rr += nread[c];
      L.cell_uniq_reads[c] = nread[c];
    }
    L.cell_ptr[C] = pp;
    L.pair_snp.resize((size_t)pp);
    L.pair_read_ptr.resize((size_t)pp + 1);
    L.pair_read_ptr[0] = 0;
    L.read_allele.resize((size_t)rr); L.read_qual.resize((size_t)rr);
    std::fill(last.begin(), last.end(), -1);
    parallel_for(NB, T, [&](int b) {
      for (const Piece& P : pc)
        for (const Row& rw : P.bk[b]) {
          const int32_t c = rw.cell;
          if (rw.snp != last[c]) { L.pair_snp[(size_t)ppos[c]++] = rw.snp; last[c] = rw.snp; }
          const int64_t w = rpos[c];
          memcpy(&L.read_allele[(size_t)w], &P.al[rw.beg], rw.n);
          memcpy(&L.read_qual[(size_t)w], &P.bq[rw.beg], rw.n);
          rpos[c] = w + rw.n;
          L.pair_read_ptr[(size_t)ppos[c]] = w + rw.n;  // end of the pair the row belongs to (= start of the next one)
        }
    });
    lap("flat image (parallel scatter)");
  } else {
    // some droplet's rows do not ascend in SNP id (not a dsc-pileup file): all rows in one list, stable sort by (cell, SNP)
    struct GRow { int32_t cell, snp; int64_t beg; };
    std::vector<GRow> rows;
    std::vector<uint8_t> al, bq;
    for (int b = 0; b < NB; ++b)      // block by block, pieces in file order: a droplet's rows keep their file order
      for (const Piece& P : pc)
        for (const Row& rw : P.bk[b]) {
          rows.push_back(GRow{rw.cell, rw.snp, (int64_t)al.size()});
          al.insert(al.end(), P.al.begin() + rw.beg, P.al.begin() + rw.beg + rw.n);
          bq.insert(bq.end(), P.bq.begin() + rw.beg, P.bq.begin() + rw.beg + rw.n);
        }
    rows.push_back(GRow{-1, -1, (int64_t)al.size()});  // sentinel: end of the last row's reads
    const size_t R = rows.size() - 1;
    L.read_allele.resize(al.size()); L.read_qual.resize(bq.size());
    std::vector<uint32_t> order(R);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
      return rows[a].cell != rows[b].cell ? rows[a].cell < rows[b].cell : rows[a].snp < rows[b].snp;
    });
    lap("cell-major order");
    L.pair_read_ptr.assign(1, 0);
    L.pair_snp.reserve(R); L.pair_read_ptr.reserve(R + 1);
    int prev_c = -1, prev_s = -1;
    size_t w = 0;  // reads written so far
    for (size_t k = 0; k < R; ++k) {
      const uint32_t i = order[k];
      const GRow rw = rows[i];
      const bool same = rw.cell == prev_c && rw.snp == prev_s;  // a (cell,SNP) listed on several rows is one pair
      const int64_t rb = rw.beg, re = rows[i + 1].beg;
      for (int64_t r = rb; r < re; ++r, ++w) { L.read_allele[w] = al[r]; L.read_qual[w] = bq[r]; }
      if (same) L.pair_read_ptr.back() = (int64_t)w;
      else {
        L.pair_snp.push_back(rw.snp);
        L.pair_read_ptr.push_back((int64_t)w);
        ++L.cell_ptr[(size_t)rw.cell + 1];
      }
      L.cell_uniq_reads[rw.cell] += re - rb;
      prev_c = rw.cell; prev_s = rw.snp;
    }
    for (int32_t c = 0; c < C; ++c) L.cell_ptr[c + 1] += L.cell_ptr[c];
    lap("flat image");
  }
  // sanity check on the observed counts (:375-381): NUM.READ replaces the pass count where the rest agrees
  L.cell_totl_reads = L.cell_uniq_reads;
  for (int32_t c = 0; c < C; ++c)
    if (L.cell_uniq_reads[c] == tmp_uniq[c] && tmp_nsnp[c] == L.cell_ptr[c + 1] - L.cell_ptr[c]) L.cell_totl_reads[c] = tmp_totl[c];
}

}  // namespace pscl_host
