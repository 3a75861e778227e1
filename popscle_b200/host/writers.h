// writers.h — the reference's output files, formatted with the reference's own printf formats:
// .best (cmd_cram_demuxlet.cpp:629,:993-1013), .lmix (cmd_cram_freemux2.cpp:111,:161),
// .clust1.samples.gz (:660-665), .clust1.vcf.gz (:608-658), and --aux-files' .clust0.samples.gz (:265-274) / .clust0.vcf.gz (:291-347).
#pragma once
#include <string>
#include <vector>

#include "loader.h"

namespace pscl_host {

void write_best(const std::string& path, const Loaded& L, const std::vector<pscl_demux_cell>& cells,
                const std::vector<double>& alphas, int min_total, int min_umi, int min_snp);
void write_lmix(const std::string& path, const Loaded& L, const std::vector<pscl_fmx_cell>& cells, bool old_mode);
void write_clust_samples(const std::string& path, const Loaded& L, const std::vector<pscl_fmx_cell>& cells);
// initial = the rows of .clust0.vcf.gz: (prior * GL) / maxGL instead of prior * (GL / maxGL), GQ = -0.1 * log10 (sic, :341)
void write_clust_vcf(const std::string& path, const Loaded& L, int n_clusters, const std::vector<double>& clust_gl,
                     const std::vector<int32_t>& clust_cnt, bool initial = false);
void write_clust0_samples(const std::string& path, const Loaded& L, const std::vector<pscl_fmx_cell>& cells);

}  // namespace pscl_host
