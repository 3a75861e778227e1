// main.cpp — `popscle demuxlet | freemuxlet | freemuxlet-old`: the kept command line
// (reference cramore.cpp:37-72, cmd_cram_demuxlet.cpp:6-123, cmd_cram_freemux2.cpp:11-104,
// cmd_cram_freemuxlet.cpp:11-83) over the C ABI of libpopscle_b200.so.  Same flag names, defaults
// and required-option errors; the likelihood work happens on the GPU and there is no CPU fallback.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "loader.h"
#include "tsv.h"
#include "writers.h"

using namespace pscl_host;

namespace {

void notice(const char* fmt, ...) {  // Error.cpp:79-94
  char ts[64];
  time_t t = time(NULL);
  strftime(ts, sizeof ts, "%Y/%m/%d %H:%M:%S", localtime(&t));
  fprintf(stderr, "NOTICE [%s] - ", ts);
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fprintf(stderr, "\n");
}

// params.cpp in miniature: `--name value`; flags take no value; repeats are rejected except for
// multi-valued options, which append (params.cpp:114-180, :449-487)
struct Opt {
  enum Kind { INT, DBL, STR, FLAG, MULTI_DBL, MULTI_STR } kind;
  void* target;
};
struct Parser {
  std::map<std::string, Opt> spec;
  void add(const char* n, int* t) { spec[n] = {Opt::INT, t}; }
  void add(const char* n, double* t) { spec[n] = {Opt::DBL, t}; }
  void add(const char* n, std::string* t) { spec[n] = {Opt::STR, t}; }
  void add(const char* n, bool* t) { spec[n] = {Opt::FLAG, t}; }
  void add(const char* n, std::vector<double>* t) { spec[n] = {Opt::MULTI_DBL, t}; }
  void add(const char* n, std::vector<std::string>* t) { spec[n] = {Opt::MULTI_STR, t}; }
  void read(int argc, char** argv) {
    std::map<std::string, int> seen;
    for (int i = 1; i < argc; ++i) {
      std::string a = argv[i];
      if (a.compare(0, 2, "--") != 0 || !spec.count(a.substr(2))) throw host_error("Cannot recognize the option " + a);
      Opt& o = spec[a.substr(2)];
      if (o.kind == Opt::FLAG) { *(bool*)o.target = true; continue; }
      if (i + 1 >= argc) throw host_error("Missing argument for option " + a);
      const char* v = argv[++i];
      if (o.kind != Opt::MULTI_DBL && o.kind != Opt::MULTI_STR && seen[a]++) throw host_error("Option " + a + " was specified more than once");
      char* end = nullptr;
      switch (o.kind) {
        case Opt::INT: { long x = strtol(v, &end, 10); if (*end) throw host_error("Invalid value " + std::string(v) + " for option " + a); *(int*)o.target = (int)x; break; }
        case Opt::DBL: { double x = strtod(v, &end); if (*end) throw host_error("Invalid value " + std::string(v) + " for option " + a); *(double*)o.target = x; break; }
        case Opt::MULTI_DBL: { double x = strtod(v, &end); if (*end) throw host_error("Invalid value " + std::string(v) + " for option " + a); ((std::vector<double>*)o.target)->push_back(x); break; }
        case Opt::STR: *(std::string*)o.target = v; break;
        case Opt::MULTI_STR: ((std::vector<std::string>*)o.target)->push_back(v); break;
        default: break;
      }
    }
  }
};

// The engine behind a command, with the reference's error convention: one GPU (pscl_ctx) or, with `--gpus N`
// (N > 1; 0 = every visible GPU) or PSCL_GPU_IDS=0,1,..., several GPUs of this box (pscl_multi: barcode-sharded
// demuxlet, SNP-sharded freemuxlet; the reference's own advice is to split the barcodes by hand with --group-list,
// cmd_cram_demuxlet.cpp:75).
struct Engine {
  pscl_ctx* h = nullptr;
  pscl_multi* m = nullptr;
  explicit Engine(int gpus) {
    char err[512] = {0};
    std::vector<int> ids;
    if (const char* e = getenv("PSCL_GPU_IDS")) {
      for (const char* c = e; *c;) { ids.push_back(atoi(c)); while (*c && *c != ',') ++c; if (*c == ',') ++c; }
    }
    if (ids.size() > 1 || (ids.empty() && gpus != 1)) {
      if (pscl_multi_create(ids.empty() ? NULL : ids.data(), ids.empty() ? gpus : (int)ids.size(), &m, err, sizeof err) != PSCL_OK) throw host_error(err);
    } else if (pscl_create(ids.empty() ? 0 : ids[0], &h, err, sizeof err) != PSCL_OK) throw host_error(err);
  }
  ~Engine() { pscl_destroy(h); pscl_multi_destroy(m); }
  int n_gpus() const { return m ? pscl_multi_size(m) : 1; }
  void chk(int rc) { if (rc != PSCL_OK) throw host_error(m ? pscl_multi_last_error(m) : pscl_last_error(h)); }
  void demux_run(const pscl_pileup* v, const pscl_geno* g, const pscl_demux_opts* o, pscl_demux_cell* out) {
    chk(m ? pscl_multi_demux_run(m, v, g, o, out, NULL) : pscl_demux_run(h, v, g, o, out, NULL));
  }
  void fmx_run(const pscl_pileup* v, const pscl_fmx_opts* o, const int32_t* init, pscl_fmx_cell* out, double* gl, int32_t* cnt, pscl_fmx_result* res,
               double* gl0 = nullptr, int32_t* cnt0 = nullptr) {
    if (m && (gl0 || cnt0)) throw host_error("--aux-files needs a single GPU (--gpus 1)");
    chk(m ? pscl_multi_fmx_run(m, v, o, init, out, gl, cnt, res) : pscl_fmx_run_aux(h, v, o, init, out, gl, cnt, res, gl0, cnt0));
  }
};

// The CUDA context(s) take a few hundred milliseconds to come up: they are created on a thread of their own while the
// loader reads the files.
struct EngineStart {
  std::unique_ptr<Engine> eng;
  std::exception_ptr err;
  std::thread th;
  explicit EngineStart(int gpus) : th([this, gpus] { try { eng.reset(new Engine(gpus)); } catch (...) { err = std::current_exception(); } }) {}
  ~EngineStart() { if (th.joinable()) th.join(); }
  Engine& get() {
    if (th.joinable()) th.join();
    if (err) std::rethrow_exception(err);
    return *eng;
  }
};

// test hook: what the loader produced, without touching the GPU
int dry_run(const Loaded& L) {
  unsigned long long h = 1469598103934665603ull;
  auto mix = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; } };
  mix(L.cell_ptr.data(), L.cell_ptr.size() * 8); mix(L.pair_snp.data(), L.pair_snp.size() * 4);
  mix(L.pair_read_ptr.data(), L.pair_read_ptr.size() * 8); mix(L.read_allele.data(), L.read_allele.size());
  mix(L.read_qual.data(), L.read_qual.size()); mix(L.af.data(), L.af.size() * 8);
  unsigned long long hg = 1469598103934665603ull;
  std::swap(h, hg); mix(L.gp.data(), L.gp.size() * 8); mix(L.has_gp.data(), L.has_gp.size()); std::swap(h, hg);
  // the compact forms that actually cross the ABI (view(): ABI 2/3 arrays; geno_view(): ABI 4 raw genotypes)
  unsigned long long hc = 1469598103934665603ull, hr = 1469598103934665603ull;
  const pscl_pileup v = L.view(true);
  const int form = v.pair_snp_delta16 ? 3 : v.pair_read_ptr32 ? 2 : 0;
  std::swap(h, hc);
  if (v.pair_read_ptr32) { mix(v.pair_read_ptr32, (size_t)(v.n_pairs + 1) * 4); mix(v.read_aq, (size_t)v.n_reads); }
  if (v.pair_snp_delta16) { mix(v.cell_first_snp, (size_t)v.n_cells * 4); mix(v.pair_snp_delta16, (size_t)v.n_pairs * 2); mix(v.pair_nreads8, (size_t)v.n_pairs); }
  std::swap(h, hc);
  // ABI 6 forms (8-bit gaps, 2-bit counts, palette-indexed base-calls), when the pileup fits them
  unsigned long long ht = 1469598103934665603ull;
  const int tiny = v.pair_snp_delta8 ? 1 : 0, rbits = v.read_bits;
  std::swap(h, ht);
  if (v.pair_snp_delta8) {
    mix(v.cell_first_snp, (size_t)v.n_cells * 4); mix(v.pair_snp_delta8, (size_t)v.n_pairs); mix(v.snp_gap_big, (size_t)v.n_gap_big * 4);
    mix(v.cell_gap_big_ptr, ((size_t)v.n_cells + 1) * 8); mix(v.pair_nreads2, (size_t)(v.n_pairs + 3) / 4); mix(v.nreads_big, (size_t)v.n_nreads_big);
    mix(v.nreads_big_ptr, ((size_t)v.n_pairs / 1024 + 2) * 8);
  }
  if (v.read_packed) { mix(v.read_packed, ((size_t)v.n_reads * v.read_bits + 7) / 8 + 1); mix(v.read_palette, (size_t)1 << v.read_bits); }
  std::swap(h, ht);
  int gt8 = -1;
  if (!L.samples.empty()) {
    const pscl_geno g = L.geno_view();
    gt8 = g.gt8 ? 1 : 0;
    std::swap(h, hr);
    if (g.gt8) mix(g.gt8, L.gt8.size()); else mix(g.gp_f32, L.gp_f32.size() * 4);
    mix(g.geno_err_snp, L.err_snp.size() * 8);
    std::swap(h, hr);
  }
  printf("{\"cells\": %d, \"snps\": %d, \"pairs\": %zu, \"reads\": %zu, \"samples\": %zu, \"has_gp\": %zu, \"pileup_fnv1a\": \"%016llx\", \"geno_fnv1a\": \"%016llx\", "
         "\"compact_form\": %d, \"compact_fnv1a\": \"%016llx\", \"gt8\": %d, \"raw_geno_fnv1a\": \"%016llx\", "
         "\"tiny_form\": %d, \"read_bits\": %d, \"tiny_fnv1a\": \"%016llx\"}\n",
         L.n_cells, L.n_snps, L.pair_snp.size(), L.read_allele.size(), L.samples.size(),
         (size_t)std::count(L.has_gp.begin(), L.has_gp.end(), 1), h, hg, form, hc, gt8, hr, tiny, rbits, ht);
  return 0;
}

int cmd_demuxlet(int argc, char** argv) {
  std::string sam, tagGroup = "CB", tagUMI = "UB", plp, vcf, field = "GP", r2Info = "R2", smList, out, groupList;
  double genoErrorOffset = 0.1, genoErrorCoeff = 0.0, minCallRate = 0.5, doubletPrior = 0.5;
  int minMAC = 1, samVerbose = 1000000, vcfVerbose = 10000, capBQ = 20, minBQ = 13, minMQ = 20, minTD = 0, exclFlag = 3844;
  int minTotal = 0, minUMI = 0, minSNP = 0, gpus = 1;
  std::vector<std::string> sm;
  std::vector<double> alphas;
  bool dry = false;
  Parser p;
  p.add("sam", &sam); p.add("tag-group", &tagGroup); p.add("tag-UMI", &tagUMI); p.add("plp", &plp); p.add("vcf", &vcf);
  p.add("field", &field); p.add("geno-error-offset", &genoErrorOffset); p.add("geno-error-coeff", &genoErrorCoeff);
  p.add("r2-info", &r2Info); p.add("min-mac", &minMAC); p.add("min-callrate", &minCallRate); p.add("sm", &sm);
  p.add("sm-list", &smList); p.add("out", &out); p.add("alpha", &alphas); p.add("doublet-prior", &doubletPrior);
  p.add("sam-verbose", &samVerbose); p.add("vcf-verbose", &vcfVerbose); p.add("cap-BQ", &capBQ); p.add("min-BQ", &minBQ);
  p.add("min-MQ", &minMQ); p.add("min-TD", &minTD); p.add("excl-flag", &exclFlag); p.add("group-list", &groupList);
  p.add("min-total", &minTotal); p.add("min-umi", &minUMI); p.add("min-snp", &minSNP); p.add("dry-run", &dry);
  p.add("gpus", &gpus);  // extension: GPUs of this box to shard the barcodes over (1; 0 = all)
  p.read(argc, argv);
  if (alphas.empty()) { alphas.push_back(0.0); alphas.push_back(0.5); }  // cmd_cram_demuxlet.cpp:85-89
  if (!sam.empty()) throw host_error("--sam (BAM/CRAM pileup on the fly) needs htslib; run `popscle dsc-pileup` first and pass --plp");
  if (plp.empty() || vcf.empty() || out.empty()) throw host_error("Missing required option(s) : --plp (or --sam), --vcf, --out");
  LoadOptions lo;
  lo.plp_prefix = plp; lo.vcf = vcf; lo.field = field; lo.geno_error_offset = genoErrorOffset; lo.geno_error_coeff = genoErrorCoeff; lo.r2_info = r2Info; lo.sm = sm;
  if (!smList.empty()) for (auto& s : read_first_column(smList)) lo.sm.push_back(s);
  lo.min_bq = minBQ; lo.cap_bq = capBQ; lo.min_read = minTotal; lo.min_umi = minUMI; lo.min_snp = minSNP;
  if (!groupList.empty()) { lo.group_list = read_first_column(groupList); lo.has_group_list = true; }
  lo.min_mac = minMAC; lo.min_callrate = minCallRate;
  notice("Loading pileup information with prefix %s", plp.c_str());
  std::unique_ptr<EngineStart> start(dry ? nullptr : new EngineStart(gpus));
  trace_lap("popscle", nullptr);
  Loaded L;
  load_plp(lo, L);
  trace_lap("popscle", "load_plp");
  notice("Finished loading %d droplets, %d variants, %zu UMIs in total..", L.n_cells, L.n_snps, L.read_allele.size());
  if (dry) return dry_run(L);
  notice("Starting to identify best matching individual IDs");
  pscl_pileup view = L.view();
  pscl_geno geno = L.geno_view();  // raw posteriors / hard calls + error rates: the library mixes on the device
  trace_lap("popscle", "compact forms");
  Engine& eng = start->get();
  trace_lap("popscle", "engine (wait)");
  pscl_demux_opts opts = {(int32_t)alphas.size(), alphas.data(), doubletPrior};
  std::vector<pscl_demux_cell> cells((size_t)L.n_cells);
  if (eng.n_gpus() > 1) notice("Sharding %d droplets over %d GPUs by pair count", L.n_cells, eng.n_gpus());
  eng.demux_run(&view, &geno, &opts, cells.data());
  trace_lap("popscle", "demux_run");
  write_best(out + ".best", L, cells, alphas, minTotal, minUMI, minSNP);
  trace_lap("popscle", "write_best");
  notice("Finished writing output files");
  return 0;
}

int cmd_freemux(int argc, char** argv, bool old_mode) {
  std::string plp, initClusterFile, out, groupList;
  int nSamples = 0, verbose = 100, initIteration = 10, capBQ = old_mode ? 40 : 20, minBQ = 13, minTotal = 0, minUMI = 0, minSNP = 0, seed = 0, gpus = 1;
  double doubletPrior = 0.5, genoError = old_mode ? 0.0 : 0.1, bfThres = 5.41, fracInitClust = 1.0;
  bool auxFiles = false, keepInitMissing = false, randomize = false, dry = false;
  Parser p;
  p.add("plp", &plp); p.add("init-cluster", &initClusterFile); p.add("out", &out); p.add("nsample", &nSamples);
  p.add("aux-files", &auxFiles); p.add("verbose", &verbose); p.add("doublet-prior", &doubletPrior); p.add("geno-error", &genoError);
  p.add("bf-thres", &bfThres); p.add("frac-init-clust", &fracInitClust); p.add("iter-init", &initIteration);
  p.add("keep-init-missing", &keepInitMissing); p.add("cap-BQ", &capBQ); p.add("min-BQ", &minBQ); p.add("group-list", &groupList);
  p.add("min-total", &minTotal); p.add("min-snp", &minSNP); p.add("dry-run", &dry);
  p.add("gpus", &gpus);  // extension: GPUs of this box to shard the SNPs over (1; 0 = all)
  if (old_mode) p.add("min-uniq", &minUMI);
  else { p.add("min-umi", &minUMI); p.add("randomize-singlet-score", &randomize); p.add("seed", &seed); }
  p.read(argc, argv);
  if (plp.empty() || out.empty() || nSamples == 0) throw host_error("Missing required option(s) : --plp, --out, --nsample");
  // --aux-files: .clust0.samples.gz / .clust0.vcf.gz (cmd_cram_freemux2.cpp:265-347).  freemuxlet-old's aux files are dumps of its
  // pairwise distance matrix (.ldist.gz, cmd_cram_freemuxlet.cpp:176-181, :348-430), which this engine never materialises
  if (auxFiles && old_mode) throw host_error("--aux-files is not supported by freemuxlet-old here (its .ldist dumps need the pairwise distance matrix)");
  LoadOptions lo;
  lo.plp_prefix = plp;
  if (old_mode) {
    // cmd_cram_freemuxlet.cpp:83 never copies the filter flags into the loader: library defaults apply
  } else {
    lo.min_bq = minBQ; lo.cap_bq = capBQ; lo.min_read = minTotal; lo.min_umi = minUMI; lo.min_snp = minSNP;
    if (!groupList.empty()) { lo.group_list = read_first_column(groupList); lo.has_group_list = true; }
  }
  std::unique_ptr<EngineStart> start(dry ? nullptr : new EngineStart(gpus));
  trace_lap("popscle", nullptr);
  Loaded L;
  load_plp(lo, L);
  trace_lap("popscle", "load_plp");
  notice("Finished loading %d droplets, %d variants, %zu UMIs in total..", L.n_cells, L.n_snps, L.read_allele.size());
  std::vector<int32_t> init;
  if (!initClusterFile.empty()) {  // cmd_cram_freemux2.cpp:92-104, :198-216
    std::map<std::string, int32_t> m;
    LineReader r(initClusterFile);
    std::string line;
    std::vector<char*> f;
    while (next_row(r, line, f)) {
      if (f.size() != 2) throw host_error("Initial clustering file " + initClusterFile + " has to have 2 columnes");
      const int k = atoi(f[1]);
      if (k >= 0) {
        if (k >= nSamples) throw host_error("--nsample " + std::to_string(nSamples) + " parameter was set. The cluster ID must be between 0 to " + std::to_string(nSamples - 1));
        m[f[0]] = k;
      }
    }
    init.assign((size_t)L.n_cells, -1);
    int nmiss = 0;
    for (int32_t i = 0; i < L.n_cells; ++i) {
      auto it = m.find(L.barcodes[i]);
      if (it == m.end()) ++nmiss; else init[i] = it->second;
    }
    if (nmiss > 0) fprintf(stderr, "WARNING: %d of %d droplets do not have initial cluster assignment\n", nmiss, L.n_cells);
  }
  if (dry) return dry_run(L);
  pscl_pileup view = L.view();
  trace_lap("popscle", "compact forms");
  Engine& eng = start->get();
  trace_lap("popscle", "engine (wait)");
  pscl_fmx_opts o = {nSamples, doubletPrior, genoError, 10, 1, fracInitClust, -1e300, old_mode ? 1 : 0, randomize ? 1 : 0, seed,
                     bfThres, old_mode ? initIteration : 0, keepInitMissing ? 1 : 0};
  std::vector<pscl_fmx_cell> cells((size_t)L.n_cells);
  std::vector<double> gl((size_t)L.n_snps * nSamples * 9), gl0(auxFiles ? gl.size() : 0);
  std::vector<int32_t> cnt((size_t)L.n_snps * nSamples * 3), cnt0(auxFiles ? cnt.size() : 0);
  pscl_fmx_result res;
  if (eng.n_gpus() > 1) notice("Sharding %d variants over %d GPUs by pair count", L.n_snps, eng.n_gpus());
  eng.fmx_run(&view, &o, init.empty() ? NULL : init.data(), cells.data(), gl.data(), cnt.data(), &res, auxFiles ? gl0.data() : NULL,
              auxFiles ? cnt0.data() : NULL);
  trace_lap("popscle", "fmx_run");
  if (auxFiles) {
    write_clust0_samples(out + ".clust0.samples.gz", L, cells);
    write_clust_vcf(out + ".clust0.vcf.gz", L, nSamples, gl0, cnt0, true);
  }
  notice("Finished %d EM iterations: %d singlets, %d doublets, %d ambiguous, and %d changed", res.n_iter, res.n_singlet, res.n_doublet,
         res.n_ambiguous, res.n_changed);
  write_lmix(out + ".lmix", L, cells, old_mode);
  write_clust_vcf(out + ".clust1.vcf.gz", L, nSamples, gl, cnt);
  write_clust_samples(out + ".clust1.samples.gz", L, cells);
  trace_lap("popscle", "writers");
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  try {
    if (argc >= 2) {
      if (!strcmp(argv[1], "demuxlet")) return cmd_demuxlet(argc - 1, argv + 1);
      if (!strcmp(argv[1], "freemuxlet")) return cmd_freemux(argc - 1, argv + 1, false);
      if (!strcmp(argv[1], "freemuxlet-old")) return cmd_freemux(argc - 1, argv + 1, true);
    }
    fprintf(stderr, "usage: popscle demuxlet|freemuxlet|freemuxlet-old [options]  (B200 engine; see DESIGN.md)\n");
    return 2;
  } catch (const std::exception& e) {
    fprintf(stderr, "\nFATAL ERROR - \n%s\n\n", e.what());  // Error.cpp:29-42, then the reference aborts
    return 134;
  }
}
