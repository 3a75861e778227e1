// standin.cpp — stand-in for the slice of htslib that the reference's demuxlet / freemuxlet
// PLP path calls.  TEST INFRASTRUCTURE (linked only into oracle/_ref/popscle_ref).
//
// htslib (samtools/htslib, which the reference neither vendors nor pins: Dockerfile:26 clones
// HEAD) is absent from this image and there is no network.  The reference's likelihood code does
// no arithmetic inside htslib; it uses it for (a) line-wise reading of gzip text
// (tsv_reader.cpp:14,31,45), (b) VCF header + record access (bcf_chunked_reader.cpp:162-318,
// bcf_filtered_reader.cpp:193-461) and (c) formatted writing (hts_utils.cpp:1013-1034).  This
// file implements exactly those calls for TEXT VCF (plain or gzip) with zlib, written from the
// documented public behaviour of the API.  Everything else the linked translation units
// mention (BAM/CRAM, BCF binary, tabix/CSI random access, FASTA index, the bcftools filter
// engine) aborts with a message: those paths are out of scope and a silent stub
// would be worse than a crash.
#include <ctype.h>
#include <limits.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <map>
#include <string>
#include <vector>

#include "htslib/bgzf.h"
#include "htslib/faidx.h"
#include "htslib/hfile.h"
#include "htslib/hts.h"
#include "htslib/kseq.h"
#include "htslib/kstring.h"
#include "htslib/sam.h"
#include "htslib/tbx.h"
#include "htslib/vcf.h"
#include "htslib/vcfutils.h"

struct hFILE { FILE* f; };
struct BGZF { gzFile gz; };

namespace {
struct FileState {
  gzFile in = nullptr;
  hFILE hf{nullptr};
  BGZF bz{nullptr};
  bool has_pending = false;  // one line read ahead by the format sniffer
  std::string pending;
};

[[noreturn]] void unsupported(const char* what) {
  fprintf(stderr, "\n[htslib stand-in] %s is not implemented: only the CEL/VAR/PLP + text-VCF path of\n"
                  "demuxlet / freemuxlet is supported by oracle/_ref (SURVEY.md §8c)\n", what);
  abort();
}

bool read_line(FileState* st, std::string& out) {
  if (st->has_pending) { out.swap(st->pending); st->has_pending = false; return true; }
  out.clear();
  char buf[1 << 16];
  bool any = false;
  while (gzgets(st->in, buf, sizeof buf)) {
    any = true;
    size_t n = strlen(buf);
    if (n && buf[n - 1] == '\n') { out.append(buf, n - 1); if (!out.empty() && out.back() == '\r') out.pop_back(); return true; }
    out.append(buf, n);
  }
  return any;
}
}  // namespace

// ---------------------------------------------------------------------------------------------
// kstring
// ---------------------------------------------------------------------------------------------
extern "C" int kvsprintf(kstring_t* s, const char* fmt, va_list ap) {
  va_list args;
  va_copy(args, ap);
  int l = vsnprintf(s->s ? s->s + s->l : nullptr, s->s ? s->m - s->l : 0, fmt, args);
  va_end(args);
  if (l < 0) return -1;
  if ((size_t)l + 1 > s->m - s->l || !s->s) {
    if (ks_resize(s, s->l + l + 2) < 0) return -1;
    va_copy(args, ap);
    l = vsnprintf(s->s + s->l, s->m - s->l, fmt, args);
    va_end(args);
  }
  s->l += l;
  return l;
}
extern "C" int ksprintf(kstring_t* s, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  int l = kvsprintf(s, fmt, ap);
  va_end(ap);
  return l;
}
// Splits in place: delimiter 0 = any run of white space; otherwise runs of `delimiter`.  Tokens
// are the maximal runs of other characters; offsets are malloc'ed (grown through *_max).
extern "C" int ksplit_core(char* s, int delimiter, int* _max, int** _offsets) {
  int n = 0, max = *_max, *offsets = *_offsets;
  const int l = (int)strlen(s);
  auto is_delim = [&](int c) { return delimiter == 0 ? (isspace(c) != 0) : (c == delimiter); };
  int i = 0;
  while (i < l) {
    while (i < l && is_delim((unsigned char)s[i])) ++i;
    if (i >= l) break;
    const int start = i;
    while (i < l && !is_delim((unsigned char)s[i])) ++i;
    if (i < l) s[i++] = 0;
    if (n == max) {
      max = max ? max << 1 : 2;
      offsets = (int*)realloc(offsets, sizeof(int) * max);
    }
    offsets[n++] = start;
  }
  *_max = max;
  *_offsets = offsets;
  return n;
}

// ---------------------------------------------------------------------------------------------
// files
// ---------------------------------------------------------------------------------------------
extern "C" htsFile* hts_open(const char* fn, const char* mode) {
  htsFile* fp = (htsFile*)calloc(1, sizeof(htsFile));
  FileState* st = new FileState();
  fp->standin = st;
  fp->fn = strdup(fn);
  if (strchr(mode, 'w')) {
    fp->is_write = 1;
    fp->format.category = unknown_category;
    fp->format.format = text_format;
    if (strchr(mode, 'z') || strchr(mode, 'g') || strchr(mode, 'b')) {  // "wz" bgzf / "wg" gzip: gzip members either way
      st->bz.gz = gzopen(fn, "wb");
      if (!st->bz.gz) { free(fp->fn); delete st; free(fp); return nullptr; }
      fp->format.compression = strchr(mode, 'g') ? gzip : bgzf;
      fp->fp.bgzf = &st->bz;
    } else {
      st->hf.f = fopen(fn, "w");
      if (!st->hf.f) { free(fp->fn); delete st; free(fp); return nullptr; }
      fp->format.compression = no_compression;
      fp->fp.hfile = &st->hf;
    }
    return fp;
  }
  FILE* raw = fopen(fn, "rb");
  if (!raw) { free(fp->fn); delete st; free(fp); return nullptr; }
  unsigned char magic[16] = {0};
  size_t got = fread(magic, 1, sizeof magic, raw);
  fclose(raw);
  fp->format.compression = no_compression;
  if (got >= 2 && magic[0] == 0x1f && magic[1] == 0x8b)
    fp->format.compression = (got >= 14 && (magic[3] & 4) && magic[12] == 'B' && magic[13] == 'C') ? bgzf : gzip;
  st->in = gzopen(fn, "rb");  // transparent for plain text
  if (!st->in) { free(fp->fn); delete st; free(fp); return nullptr; }
  gzbuffer(st->in, 1 << 20);
  std::string first;
  if (read_line(st, first)) { st->pending = first; st->has_pending = true; }
  if (first.compare(0, 16, "##fileformat=VCF") == 0) { fp->format.category = variant_data; fp->format.format = vcf; }
  else if (first.compare(0, 3, "BCF") == 0) unsupported("reading binary BCF");
  else { fp->format.category = unknown_category; fp->format.format = text_format; }
  return fp;
}

extern "C" int hts_close(htsFile* fp) {
  if (!fp) return -1;
  FileState* st = (FileState*)fp->standin;
  int rc = 0;
  if (st->in) rc |= gzclose(st->in);
  if (st->bz.gz) rc |= gzclose(st->bz.gz);
  if (st->hf.f) rc |= fclose(st->hf.f);
  delete st;
  free(fp->fn);
  free(fp->line.s);
  free(fp);
  return rc ? -1 : 0;
}

extern "C" int hts_getline(htsFile* fp, int delimiter, kstring_t* str) {
  if (delimiter != KS_SEP_LINE) unsupported("hts_getline with a delimiter other than KS_SEP_LINE");
  FileState* st = (FileState*)fp->standin;
  std::string line;
  if (!read_line(st, line)) return -1;
  str->l = 0;
  kputsn(line.data(), (int)line.size(), str);
  ++fp->lineno;
  return (int)str->l;
}

extern "C" ssize_t hwrite(hFILE* fp, const void* buffer, size_t nbytes) {
  return (ssize_t)fwrite(buffer, 1, nbytes, fp->f);
}
extern "C" ssize_t bgzf_write(BGZF* fp, const void* data, size_t length) {
  if (length == 0) return 0;
  int n = gzwrite(fp->gz, data, (unsigned)length);
  return n <= 0 ? -1 : n;
}

// ---------------------------------------------------------------------------------------------
// VCF text: header
// ---------------------------------------------------------------------------------------------
namespace {
struct HdrState {
  std::vector<std::string> ctg, smp, ids;  // dictionaries in first-appearance order
  std::map<std::string, int> ctg_ix, smp_ix, id_ix;
  std::map<std::string, int> info_type, fmt_type;  // BCF_HT_* of declared INFO / FORMAT tags
  std::vector<bcf_idpair_t> pairs[3];
};
std::map<const bcf_hdr_t*, HdrState*> g_hdr;

HdrState* hs(const bcf_hdr_t* h) {
  auto it = g_hdr.find(h);
  if (it == g_hdr.end()) unsupported("a bcf_hdr_t that did not come from bcf_hdr_read");
  return it->second;
}

std::string attr(const std::string& line, const char* key) {  // value of key= inside <...>
  std::string k = std::string(key) + "=";
  size_t p = line.find("<");
  while (p != std::string::npos) {
    size_t q = line.find(k, p);
    if (q == std::string::npos) return "";
    if (line[q - 1] == '<' || line[q - 1] == ',') {
      size_t b = q + k.size(), e = b;
      if (b < line.size() && line[b] == '"') { e = line.find('"', b + 1); return line.substr(b + 1, e - b - 1); }
      while (e < line.size() && line[e] != ',' && line[e] != '>') ++e;
      return line.substr(b, e - b);
    }
    p = q + 1;
  }
  return "";
}
int type_of(const std::string& t) {
  if (t == "Integer") return BCF_HT_INT;
  if (t == "Float") return BCF_HT_REAL;
  if (t == "Flag") return BCF_HT_FLAG;
  return BCF_HT_STR;
}
int add_id(HdrState* s, const std::string& id) {
  auto it = s->id_ix.find(id);
  if (it != s->id_ix.end()) return it->second;
  int k = (int)s->ids.size();
  s->ids.push_back(id);
  s->id_ix[id] = k;
  return k;
}
void publish(bcf_hdr_t* h, HdrState* s) {  // (re)build the public id[] / n[] / samples views
  const std::vector<std::string>* src[3] = {&s->ids, &s->ctg, &s->smp};
  for (int t = 0; t < 3; ++t) {
    s->pairs[t].resize(src[t]->size());
    for (size_t i = 0; i < src[t]->size(); ++i) { s->pairs[t][i].key = (*src[t])[i].c_str(); s->pairs[t][i].val = nullptr; }
    h->n[t] = (int32_t)src[t]->size();
    h->id[t] = s->pairs[t].data();
  }
  free(h->samples);
  h->samples = (char**)malloc(sizeof(char*) * (s->smp.size() + 1));
  for (size_t i = 0; i < s->smp.size(); ++i) h->samples[i] = const_cast<char*>(s->smp[i].c_str());
}
}  // namespace

extern "C" bcf_hdr_t* bcf_hdr_read(htsFile* fp) {
  if (fp->format.format != vcf) unsupported("bcf_hdr_read on a non-VCF-text file");
  FileState* st = (FileState*)fp->standin;
  bcf_hdr_t* h = (bcf_hdr_t*)calloc(1, sizeof(bcf_hdr_t));
  HdrState* s = new HdrState();
  g_hdr[h] = s;
  add_id(s, "PASS");  // htslib always defines PASS as id 0
  std::string line;
  bool saw_chrom = false;
  while (read_line(st, line)) {
    if (line.compare(0, 2, "##") == 0) {
      if (line.compare(0, 9, "##contig=") == 0) {
        std::string id = attr(line, "ID");
        if (!id.empty() && !s->ctg_ix.count(id)) { s->ctg_ix[id] = (int)s->ctg.size(); s->ctg.push_back(id); }
      } else if (line.compare(0, 7, "##INFO=") == 0) {
        std::string id = attr(line, "ID"); add_id(s, id); s->info_type[id] = type_of(attr(line, "Type"));
      } else if (line.compare(0, 9, "##FORMAT=") == 0) {
        std::string id = attr(line, "ID"); add_id(s, id); s->fmt_type[id] = type_of(attr(line, "Type"));
      } else if (line.compare(0, 9, "##FILTER=") == 0) {
        add_id(s, attr(line, "ID"));
      }
      continue;
    }
    if (line.compare(0, 6, "#CHROM") == 0) {
      saw_chrom = true;
      std::vector<char> buf(line.begin(), line.end());
      buf.push_back(0);
      int mx = 0, *off = nullptr;
      int n = ksplit_core(buf.data(), '\t', &mx, &off);
      for (int i = 9; i < n; ++i) {
        std::string id(buf.data() + off[i]);
        if (s->smp_ix.count(id)) unsupported("duplicate sample names in the VCF header");
        s->smp_ix[id] = (int)s->smp.size();
        s->smp.push_back(id);
      }
      free(off);
      break;
    }
    // a data line before #CHROM: malformed
    st->pending = line; st->has_pending = true;
    break;
  }
  if (!saw_chrom) { g_hdr.erase(h); delete s; free(h); return nullptr; }
  publish(h, s);
  return h;
}

extern "C" void bcf_hdr_destroy(bcf_hdr_t* h) {
  if (!h) return;
  auto it = g_hdr.find(h);
  if (it != g_hdr.end()) { delete it->second; g_hdr.erase(it); }
  free(h->samples);
  free(h);
}

extern "C" int bcf_hdr_id2int(const bcf_hdr_t* h, int which, const char* id) {
  HdrState* s = hs(h);
  const std::map<std::string, int>& m = which == BCF_DT_CTG ? s->ctg_ix : which == BCF_DT_SAMPLE ? s->smp_ix : s->id_ix;
  auto it = m.find(id);
  return it == m.end() ? -1 : it->second;
}

// ---------------------------------------------------------------------------------------------
// VCF text: records
// ---------------------------------------------------------------------------------------------
namespace {
struct RecState {
  std::string id, als_buf;
  std::vector<std::string> alleles, flt, info_kv, fmt_keys, smp;  // raw text columns
  std::vector<char*> allele_ptr;
  std::vector<int> flt_ids;
};
RecState* rs(bcf1_t* v) { return (RecState*)v->standin; }

std::vector<std::string> split(const std::string& s, char d) {
  std::vector<std::string> out;
  size_t b = 0;
  for (;;) {
    size_t e = s.find(d, b);
    out.push_back(s.substr(b, e == std::string::npos ? std::string::npos : e - b));
    if (e == std::string::npos) break;
    b = e + 1;
  }
  return out;
}
}  // namespace

uint32_t bcf_float_missing = 0x7F800001;
uint32_t bcf_float_vector_end = 0x7F800002;

extern "C" bcf1_t* bcf_init(void) {
  bcf1_t* v = (bcf1_t*)calloc(1, sizeof(bcf1_t));
  v->standin = new RecState();
  return v;
}
extern "C" void bcf_destroy(bcf1_t* v) {
  if (!v) return;
  delete rs(v);
  free(v);
}

extern "C" int bcf_read(htsFile* fp, const bcf_hdr_t* h, bcf1_t* v) {
  FileState* st = (FileState*)fp->standin;
  HdrState* s = hs(h);
  RecState* r = rs(v);
  std::string line;
  do {
    if (!read_line(st, line)) return -1;
  } while (line.empty());
  std::vector<std::string> c = split(line, '\t');
  if (c.size() < 8) { fprintf(stderr, "[htslib stand-in] VCF data line with %zu columns\n", c.size()); abort(); }
  auto it = s->ctg_ix.find(c[0]);
  if (it == s->ctg_ix.end()) {  // htslib adds undeclared contigs on the fly (with a warning)
    fprintf(stderr, "[W::vcf_parse] contig '%s' is not defined in the header\n", c[0].c_str());
    s->ctg_ix[c[0]] = (int)s->ctg.size();
    s->ctg.push_back(c[0]);
    publish(const_cast<bcf_hdr_t*>(h), s);
    it = s->ctg_ix.find(c[0]);
  }
  v->rid = it->second;
  v->pos = atoi(c[1].c_str()) - 1;
  r->id = c[2];
  r->alleles.clear();
  r->alleles.push_back(c[3]);
  if (c[4] != ".")
    for (auto& a : split(c[4], ',')) r->alleles.push_back(a);
  v->n_allele = (uint32_t)r->alleles.size();
  v->rlen = (int32_t)c[3].size();
  r->allele_ptr.clear();
  for (auto& a : r->alleles) r->allele_ptr.push_back(const_cast<char*>(a.c_str()));
  v->d.allele = r->allele_ptr.data();
  v->d.id = const_cast<char*>(r->id.c_str());
  if (c[5] == ".") bcf_float_set_missing(v->qual); else v->qual = (float)atof(c[5].c_str());
  r->flt_ids.clear();
  if (c[6] != ".")
    for (auto& f : split(c[6], ';')) r->flt_ids.push_back(add_id(s, f));
  v->d.n_flt = (int)r->flt_ids.size();
  v->d.flt = r->flt_ids.data();
  r->info_kv.clear();
  if (c[7] != ".") r->info_kv = split(c[7], ';');
  v->n_info = (uint32_t)r->info_kv.size();
  r->fmt_keys.clear();
  r->smp.clear();
  if (c.size() > 8) {
    r->fmt_keys = split(c[8], ':');
    r->smp.assign(c.begin() + 9, c.end());
    if (r->smp.size() != s->smp.size()) { fprintf(stderr, "[htslib stand-in] %zu sample columns, header has %zu\n", r->smp.size(), s->smp.size()); abort(); }
  }
  v->n_fmt = (uint32_t)r->fmt_keys.size();
  v->n_sample = (uint32_t)r->smp.size();
  v->unpacked = BCF_UN_ALL;
  v->errcode = 0;
  return 0;
}

extern "C" int bcf_unpack(bcf1_t*, int) { return 0; }  // text records are decoded by bcf_read

extern "C" int bcf_is_snp(bcf1_t* v) {
  RecState* r = rs(v);
  for (auto& a : r->alleles)
    if (a.size() != 1) return 0;
  return 1;
}

// FORMAT values.  GT (BCF_HT_INT): `ploidy` ints per sample, (allele+1)<<1 | phased, 0 = missing,
// padded with bcf_int32_vector_end.  Other tags: per-sample comma-separated vectors, padded.
// Return: number of values written (n_sample * per-sample width); -1 tag not in header,
// -2 type clash, -3 tag absent from this record.
extern "C" int bcf_get_format_values(const bcf_hdr_t* h, bcf1_t* v, const char* tag, void** dst, int* ndst, int type) {
  HdrState* s = hs(h);
  RecState* r = rs(v);
  const bool is_gt = strcmp(tag, "GT") == 0;
  auto ft = s->fmt_type.find(tag);
  if (ft == s->fmt_type.end()) return -1;
  if (!is_gt && ft->second != type) return -2;
  int col = -1;
  for (size_t i = 0; i < r->fmt_keys.size(); ++i)
    if (r->fmt_keys[i] == tag) col = (int)i;
  if (col < 0) return -3;
  const int ns = (int)r->smp.size();
  std::vector<std::vector<std::string>> vals(ns);
  int width = 0;
  for (int i = 0; i < ns; ++i) {
    std::vector<std::string> f = split(r->smp[i], ':');
    std::string cell = col < (int)f.size() ? f[col] : ".";
    if (is_gt) {
      std::string cur;
      for (char ch : cell) {
        if (ch == '/' || ch == '|') { vals[i].push_back(cur); vals[i].push_back(std::string(1, ch)); cur.clear(); }
        else cur += ch;
      }
      vals[i].push_back(cur);
      width = std::max(width, (int)(vals[i].size() + 1) / 2);
    } else {
      vals[i] = split(cell, ',');
      width = std::max(width, (int)vals[i].size());
    }
  }
  const int total = ns * width;
  if (type == BCF_HT_STR) unsupported("string FORMAT fields");
  if (*ndst < total || !*dst) { *dst = realloc(*dst, sizeof(int32_t) * (total ? total : 1)); *ndst = total; }
  for (int i = 0; i < ns; ++i) {
    if (is_gt) {
      int32_t* out = (int32_t*)*dst + (size_t)i * width;
      int k = 0;
      bool phased = false;
      for (size_t t = 0; t < vals[i].size(); ++t) {
        if (t & 1) { phased = vals[i][t] == "|"; continue; }
        const std::string& a = vals[i][t];
        int32_t code = (a == "." || a.empty()) ? 0 : ((atoi(a.c_str()) + 1) << 1);
        out[k++] = code | (phased ? 1 : 0);
      }
      for (; k < width; ++k) out[k] = bcf_int32_vector_end;
    } else if (type == BCF_HT_INT) {
      int32_t* out = (int32_t*)*dst + (size_t)i * width;
      int k = 0;
      for (auto& a : vals[i]) out[k++] = (a == "." || a.empty()) ? bcf_int32_missing : atoi(a.c_str());
      for (; k < width; ++k) out[k] = bcf_int32_vector_end;
    } else {
      float* out = (float*)*dst + (size_t)i * width;
      int k = 0;
      for (auto& a : vals[i]) { if (a == "." || a.empty()) bcf_float_set_missing(out[k]); else out[k] = (float)atof(a.c_str()); ++k; }
      for (; k < width; ++k) bcf_float_set_vector_end(out[k]);
    }
  }
  return total;
}

extern "C" int bcf_get_info_values(const bcf_hdr_t* h, bcf1_t* v, const char* tag, void** dst, int* ndst, int type) {
  HdrState* s = hs(h);
  RecState* r = rs(v);
  auto it = s->info_type.find(tag);
  if (it == s->info_type.end()) return -1;
  if (it->second != type) return -2;
  const std::string key = std::string(tag) + "=";
  for (auto& kv : r->info_kv) {
    if (kv.compare(0, key.size(), key) != 0) continue;
    std::vector<std::string> f = split(kv.substr(key.size()), ',');
    const int n = (int)f.size();
    if (type == BCF_HT_STR || type == BCF_HT_FLAG) unsupported("string / flag INFO fields");
    if (*ndst < n || !*dst) { *dst = realloc(*dst, sizeof(int32_t) * n); *ndst = n; }
    for (int i = 0; i < n; ++i) {
      if (type == BCF_HT_INT) ((int32_t*)*dst)[i] = f[i] == "." ? bcf_int32_missing : atoi(f[i].c_str());
      else if (f[i] == ".") bcf_float_set_missing(((float*)*dst)[i]);
      else ((float*)*dst)[i] = (float)atof(f[i].c_str());
    }
    return n;
  }
  return -3;
}

// ---------------------------------------------------------------------------------------------
// everything below is mentioned by the linked reference units but never reached on the PLP path
// ---------------------------------------------------------------------------------------------
#define UNREACHED(ret, name, args) extern "C" ret name args { unsupported(#name); }
UNREACHED(int, hts_set_fai_filename, (htsFile*, const char*))
UNREACHED(void, hts_idx_destroy, (hts_idx_t*))
extern "C" void hts_itr_destroy(hts_itr_t* it) { if (it) unsupported("hts_itr_destroy on a live iterator"); }
UNREACHED(int, bcf_hdr_append, (bcf_hdr_t*, const char*))
UNREACHED(int, bcf_hrec_find_key, (bcf_hrec_t*, const char*))
UNREACHED(int, vcf_parse, (kstring_t*, const bcf_hdr_t*, bcf1_t*))
UNREACHED(int, vcf_format, (const bcf_hdr_t*, const bcf1_t*, kstring_t*))
UNREACHED(hts_itr_t*, bcf_itr_querys, (const hts_idx_t*, const bcf_hdr_t*, const char*))
UNREACHED(int, bcf_itr_next, (htsFile*, hts_itr_t*, void*))
UNREACHED(hts_idx_t*, bcf_index_load, (const char*))
UNREACHED(tbx_t*, tbx_index_load, (const char*))
extern "C" void tbx_destroy(tbx_t* t) { if (t) unsupported("tbx_destroy"); }
UNREACHED(hts_itr_t*, tbx_itr_querys, (tbx_t*, const char*))
UNREACHED(int, tbx_itr_next, (htsFile*, tbx_t*, hts_itr_t*, void*))
UNREACHED(int32_t, bam_endpos, (const bam1_t*))
// SAMFilteredReader's constructor / destructor allocate and free record buffers even when no BAM
// is ever opened (sam_filtered_reader.h), so these two are real
extern "C" bam1_t* bam_init1(void) { return (bam1_t*)calloc(1, sizeof(bam1_t)); }
extern "C" void bam_destroy1(bam1_t* b) { if (b) { free(b->data); free(b); } }
UNREACHED(hts_itr_t*, sam_itr_querys, (const hts_idx_t*, bam_hdr_t*, const char*))
UNREACHED(int, sam_itr_next, (htsFile*, hts_itr_t*, bam1_t*))
UNREACHED(hts_idx_t*, sam_index_load, (htsFile*, const char*))
UNREACHED(bam_hdr_t*, sam_hdr_read, (samFile*))
UNREACHED(int, sam_read1, (samFile*, bam_hdr_t*, bam1_t*))
UNREACHED(int, bam_name2id, (bam_hdr_t*, const char*))
UNREACHED(void, bam_hdr_destroy, (bam_hdr_t*))
UNREACHED(uint8_t*, bam_aux_get, (const bam1_t*, const char*))
UNREACHED(char*, bam_aux2Z, (const uint8_t*))
UNREACHED(faidx_t*, fai_load, (const char*))
UNREACHED(int, fai_build, (const char*))
UNREACHED(int, faidx_nseq, (const faidx_t*))
UNREACHED(const char*, faidx_iseq, (const faidx_t*, int))
UNREACHED(int, faidx_seq_len, (const faidx_t*, const char*))
UNREACHED(char*, faidx_fetch_seq, (const faidx_t*, const char*, int, int, int*))

// bcftools' -i/-e expression engine (filter.cpp, 3.4k lines over htslib internals): demuxlet and
// freemuxlet expose no flag that sets an expression, so vfilt.filt stays NULL (bcf_filter_arg.h:79-113)
struct _filter_t;
_filter_t* filter_init(bcf_hdr_t*, const char*) { unsupported("filter_init (bcftools expressions)"); }
int filter_test(_filter_t*, bcf1_t*, const uint8_t**) { unsupported("filter_test (bcftools expressions)"); }
