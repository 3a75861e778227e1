#!/usr/bin/env python
"""Generates tests/golden/<case>/: seeded dsc-pileup files (+ VCF) and the outputs the REFERENCE
ITSELF writes for them — oracle/_ref/popscle_ref, i.e. the reference's unmodified command bodies
compiled by oracle/build_ref.sh (only possible where /root/reference exists).

    python tests/golden/make_golden.py            # regenerate every case

Each case directory holds the inputs, `cmd.json` (the argv the reference ran with) and the
reference's output files (`ref.*`).  tests/test_golden.py replays the same argv through
popscle_b200.cli with the oracle (CPU) and with the CUDA library (GPU) and compares.
"""
import gzip
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np

from popscle_b200 import plpio, synth

REF = os.path.join(ROOT, "oracle", "_ref", "popscle_ref")


def run_ref(case_dir, argv):
    with open(os.path.join(case_dir, "cmd.json"), "w") as f:
        json.dump({"argv": argv}, f, indent=1)
    r = subprocess.run([REF] + argv, cwd=case_dir, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"reference failed in {case_dir}: {r.stderr[-2000:]}")
    # the date header of .clust1.vcf.gz changes daily: store VCFs uncompressed with that line dropped
    for fn in sorted(os.listdir(case_dir)):
        if fn.startswith("ref.") and fn.endswith(".gz"):
            txt = gzip.open(os.path.join(case_dir, fn), "rt").read()
            txt = "".join(l for l in txt.splitlines(True) if not l.startswith("##fileDate="))
            with open(os.path.join(case_dir, fn[:-3]), "w") as f:
                f.write(txt)
            os.remove(os.path.join(case_dir, fn))


def fresh(name):
    d = os.path.join(HERE, name)
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    return d


def base(C, nv, V, kbar, seed, d, allele2=0.0):
    s = synth.make_pileup(C=C, nv=nv, V=V, kbar=kbar, seed=seed, cap_bq=40)
    if allele2 > 0:
        rng = np.random.default_rng(seed)
        s.plp.read_allele[rng.random(s.plp.n_reads) < allele2] = 2
    sites = plpio.default_sites(V, s.af, seed=seed)
    rng = np.random.default_rng(seed + 1)
    bcs = ["".join(rng.choice(list("ACGT"), 16)) + "-1" for _ in range(C)]  # unsorted barcodes: .best is barcode-sorted
    plpio.write_plp(os.path.join(d, "p"), s.plp, sites, bcs)
    return s, sites, bcs


def main():
    if not os.path.exists(REF):
        sys.exit("oracle/_ref/popscle_ref missing: run oracle/build_ref.sh where /root/reference exists")
    only = set(sys.argv[1:])  # optional: regenerate only the named cases
    if only:
        return main_only(only)
    main_all()
    main_only({"demux_r2", "fmx_random", "demux_gt8", "demux_gp8", "demux_64x21", "fmx_old_seed", "fmx_old_refine", "fmx_old_frac", "fmx_aux"})


def main_only(only):
    # 17. freemuxlet --aux-files: the initial clusters (.clust0.samples.gz, -1 for the droplets beyond --frac-init-clust) and the
    # cluster pileups built from them (.clust0.vcf.gz) beside the usual outputs (cmd_cram_freemux2.cpp:265-347)
    if "fmx_aux" in only:
        d = fresh("fmx_aux")
        base(130, 4, 800, 210, 23, d, allele2=0.02)
        run_ref(d, ["freemuxlet", "--plp", "p", "--nsample", "4", "--aux-files", "--frac-init-clust", "0.7", "--out", "ref"])
    # 9. demuxlet with --geno-error-coeff: per-site INFO/R2 scales the genotype error (sc_drop_seq.cpp:299-306)
    if "demux_r2" in only:
        d = fresh("demux_r2")
        s, sites, bcs = base(80, 4, 600, 150, 19, d)
        rng = np.random.default_rng(9)
        r2 = np.round(rng.uniform(0.3, 1.0, 600), 3)
        plpio.write_vcf(os.path.join(d, "g.vcf.gz"), sites, [f"S{j}" for j in range(4)], geno=s.geno, r2=r2)
        run_ref(d, ["demuxlet", "--plp", "p", "--vcf", "g.vcf.gz", "--field", "GT", "--geno-error-offset", "0.02", "--geno-error-coeff", "0.6",
                    "--out", "ref"])
    # 10. freemuxlet with --randomize-singlet-score: the singlet scores are shuffled with libc rand() seeded by --seed
    # (cmd_cram_freemux2.cpp:164-181) before the droplets are sorted for the greedy seeding of 60 % of them
    if "fmx_random" in only:
        d = fresh("fmx_random")
        base(110, 3, 700, 200, 21, d)
        run_ref(d, ["freemuxlet", "--plp", "p", "--nsample", "3", "--randomize-singlet-score", "--seed", "11", "--frac-init-clust", "0.6",
                    "--out", "ref"])
    # 11. the benchmark's own kernel shape through the reference: 8 samples, hard calls (k_demux_default on dictionary-coded
    # genotypes) ...
    if "demux_gt8" in only:
        d = fresh("demux_gt8")
        s, sites, bcs = base(140, 8, 1500, 220, 23, d, allele2=0.02)
        plpio.write_vcf(os.path.join(d, "g.vcf.gz"), sites, [f"D{j}" for j in range(8)], geno=s.geno)
        run_ref(d, ["demuxlet", "--plp", "p", "--vcf", "g.vcf.gz", "--field", "GT", "--out", "ref"])
    # 12. ... and 8 samples with soft posteriors (k_demux_default on gathered genotype rows), SNPs missing from the VCF
    if "demux_gp8" in only:
        d = fresh("demux_gp8")
        s, sites, bcs = base(120, 8, 1400, 220, 24, d)
        rng = np.random.default_rng(12)
        gp = 0.85 * np.eye(3)[s.geno.T] + 0.15 * rng.dirichlet([1, 1, 1], size=(1400, 8))
        keep = rng.random(1400) > 0.1
        plpio.write_vcf(os.path.join(d, "g.vcf.gz"), sites, [f"D{j}" for j in range(8)], geno=s.geno, gp=gp, keep=keep)
        run_ref(d, ["demuxlet", "--plp", "p", "--vcf", "g.vcf.gz", "--field", "GP", "--out", "ref"])
    # 13. configs[3]'s shape through the reference: 64 samples, the 21-point alpha grid (k_demux_poly), a few droplets
    if "demux_64x21" in only:
        d = fresh("demux_64x21")
        s, sites, bcs = base(24, 64, 1500, 260, 25, d)
        plpio.write_vcf(os.path.join(d, "g.vcf.gz"), sites, [f"P{j:02d}" for j in range(64)], geno=s.geno)
        argv = ["demuxlet", "--plp", "p", "--vcf", "g.vcf.gz", "--field", "GT"]
        for i in range(21):
            argv += ["--alpha", "%g" % (0.025 * i)]
        run_ref(d, argv + ["--out", "ref"])
    # 14. freemuxlet-old with ITS OWN seeding: pairwise Bayes-factor matrix, votes on the un-seeded libc rand() stream, ten
    # refinement sweeps in std::random_shuffle order (cmd_cram_freemuxlet.cpp:165-346)
    if "fmx_old_seed" in only:
        d = fresh("fmx_old_seed")
        base(110, 3, 700, 200, 26, d)
        run_ref(d, ["freemuxlet-old", "--plp", "p", "--nsample", "3", "--geno-error", "0.05", "--out", "ref"])
    # 15. ... the refinement sweeps alone, from a partial --init-cluster file, missing droplets kept missing, a looser
    # Bayes-factor threshold, half of the droplets in the first round
    if "fmx_old_refine" in only:
        d = fresh("fmx_old_refine")
        s, sites, bcs = base(100, 4, 700, 200, 27, d)
        rng = np.random.default_rng(15)
        with open(os.path.join(d, "init.tsv"), "w") as f:
            for c in range(100):
                if rng.random() < 0.7:
                    f.write(f"{bcs[c]}\t{int(rng.integers(0, 4) if rng.random() < 0.3 else s.truth_d1[c])}\n")
        run_ref(d, ["freemuxlet-old", "--plp", "p", "--nsample", "4", "--init-cluster", "init.tsv", "--keep-init-missing", "--bf-thres", "3.0",
                    "--out", "ref"])
    if "fmx_old_frac" in only:
        d = fresh("fmx_old_frac")
        base(90, 3, 600, 180, 28, d)
        run_ref(d, ["freemuxlet-old", "--plp", "p", "--nsample", "3", "--frac-init-clust", "0.5", "--iter-init", "0", "--out", "ref"])
    unknown = only - {"demux_r2", "fmx_random", "demux_gt8", "demux_gp8", "demux_64x21", "fmx_old_seed", "fmx_old_refine", "fmx_old_frac", "fmx_aux"}
    if unknown:
        sys.exit(f"cases {sorted(unknown)} are generated by the full run only")


def main_all():

    # 1. demuxlet, hard genotypes, all defaults (tutorial-like shape: BASELINE.json configs[0], scaled down)
    d = fresh("demux_gt")
    s, sites, bcs = base(120, 4, 800, 150, 11, d)
    plpio.write_vcf(os.path.join(d, "g.vcf.gz"), sites, [f"S{j}" for j in range(4)], geno=s.geno)
    run_ref(d, ["demuxlet", "--plp", "p", "--vcf", "g.vcf.gz", "--field", "GT", "--out", "ref"])

    # 2. demuxlet, soft GP field, 3-point alpha grid, SNPs missing from the VCF, non-default priors
    d = fresh("demux_gp")
    s, sites, bcs = base(90, 5, 700, 160, 12, d, allele2=0.05)
    rng = np.random.default_rng(2)
    onehot = np.eye(3)[s.geno.T]                                        # [V][nv][3]
    gp = 0.9 * onehot + 0.1 * rng.dirichlet([1, 1, 1], size=(700, 5))
    keep = rng.random(700) > 0.2
    names = ["d_" + x for x in "EACBD"]                                # VCF column order != sorted order
    plpio.write_vcf(os.path.join(d, "g.vcf.gz"), sites, names, geno=s.geno, gp=gp, keep=keep)
    run_ref(d, ["demuxlet", "--plp", "p", "--vcf", "g.vcf.gz", "--field", "GP", "--alpha", "0", "--alpha", "0.25", "--alpha", "0.5",
                "--doublet-prior", "0.3", "--geno-error-offset", "0.05", "--out", "ref"])

    # 3. demuxlet, PL field, missing genotype calls, other BQ window, sample subset (--sm: alphabetical order quirk)
    d = fresh("demux_pl")
    s, sites, bcs = base(70, 4, 600, 150, 13, d)
    rng = np.random.default_rng(3)
    geno = s.geno.astype(int).copy()
    geno[rng.random(geno.shape) < 0.08] = -1
    pl = np.full((600, 4, 3), 60)
    g = np.where(geno.T < 0, 0, geno.T)
    pl[np.arange(600)[:, None], np.arange(4)[None, :], g] = 0
    pl[:, :, 1] = np.where(g == 1, 0, np.minimum(pl[:, :, 1], 25))
    names = ["zeta", "alpha", "mid", "beta"]
    plpio.write_vcf(os.path.join(d, "g.vcf.gz"), sites, names, geno=geno, pl=pl)
    run_ref(d, ["demuxlet", "--plp", "p", "--vcf", "g.vcf.gz", "--field", "PL", "--sm", "zeta", "--sm", "alpha", "--sm", "beta",
                "--cap-BQ", "30", "--min-BQ", "15", "--out", "ref"])

    # 4. demuxlet with droplet filters: --group-list subset and --min-snp (INT_ID counts skipped droplets)
    d = fresh("demux_filters")
    s, sites, bcs = base(100, 3, 600, 120, 14, d)
    plpio.write_vcf(os.path.join(d, "g.vcf.gz"), sites, ["A", "B", "C"], geno=s.geno)
    with open(os.path.join(d, "groups.txt"), "w") as f:
        f.write("\n".join(bcs[::2]) + "\n")
    run_ref(d, ["demuxlet", "--plp", "p", "--vcf", "g.vcf.gz", "--field", "GT", "--group-list", "groups.txt", "--min-snp", "100", "--out", "ref"])

    # 5. freemuxlet, defaults (greedy seeding + EM with early stop)
    d = fresh("fmx_default")
    s, sites, bcs = base(150, 4, 900, 220, 15, d, allele2=0.03)
    run_ref(d, ["freemuxlet", "--plp", "p", "--nsample", "4", "--out", "ref", "--seed", "1"])

    # 6. freemuxlet from --init-cluster (with unassigned droplets), non-default priors, partial seeding fraction
    d = fresh("fmx_init")
    s, sites, bcs = base(120, 3, 800, 200, 16, d)
    rng = np.random.default_rng(6)
    with open(os.path.join(d, "init.tsv"), "w") as f:
        for c in range(120):
            if rng.random() < 0.85:
                f.write(f"{bcs[c]}\t{int(s.truth_d1[c])}\n")
    run_ref(d, ["freemuxlet", "--plp", "p", "--nsample", "3", "--init-cluster", "init.tsv", "--doublet-prior", "0.3", "--geno-error", "0.05",
                "--out", "ref", "--seed", "1"])

    # 7. freemuxlet, greedy seeding of half the droplets only
    d = fresh("fmx_frac")
    s, sites, bcs = base(100, 3, 700, 200, 17, d)
    run_ref(d, ["freemuxlet", "--plp", "p", "--nsample", "3", "--frac-init-clust", "0.5", "--out", "ref", "--seed", "1"])

    # 8. freemuxlet-old EM (cmd_cram_freemuxlet.cpp) from --init-cluster, vote refinement off
    d = fresh("fmx_old")
    s, sites, bcs = base(100, 3, 700, 200, 18, d)
    with open(os.path.join(d, "init.tsv"), "w") as f:
        for c in range(100):
            f.write(f"{bcs[c]}\t{int(s.truth_d1[c])}\n")
    run_ref(d, ["freemuxlet-old", "--plp", "p", "--nsample", "3", "--init-cluster", "init.tsv", "--iter-init", "0", "--geno-error", "0.05", "--out", "ref"])
    sz = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(HERE) for f in fs)
    print("golden fixtures written, total bytes:", sz)


if __name__ == "__main__":
    main()
