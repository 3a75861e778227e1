"""`popscle demuxlet | freemuxlet | freemuxlet-old` — host side of the kept command line.

Same flag names, defaults and required-option errors as the reference command bodies
(cmd_cram_demuxlet.cpp:6-83, cmd_cram_freemux2.cpp:11-71, cmd_cram_freemuxlet.cpp:11-66): parse,
load the dsc-pileup files, hand the flat pileup to the CUDA library through the C ABI, write the
reference's output files.  The engine argument exists so the parity tests can drive the same
host code with the CPU oracle; the product default is the CUDA Context and there is no fallback.

    python -m popscle_b200.cli demuxlet --plp P --vcf V.vcf.gz --field GT --out O
    python -m popscle_b200.cli freemuxlet --plp P --nsample 8 --out O
"""
from __future__ import annotations

import sys
import time

import numpy as np

from . import plpio, report


class UsageError(Exception):
    """the reference's error(): "FATAL ERROR - ..." then abort (Error.cpp:29-42)"""


def _notice(msg):
    sys.stderr.write("NOTICE [%s] - %s\n" % (time.strftime("%Y/%m/%d %H:%M:%S"), msg))


def _parse(argv, spec):
    """params.cpp semantics in miniature: `--name value`, flags without a value, repeats rejected
    except for multi-valued options which append (params.cpp:114-180, :449-487)."""
    out = {k: (list(v[1]) if v[0].startswith("multi") else v[1]) for k, v in spec.items()}
    seen = set()
    i = 0
    while i < len(argv):
        a = argv[i]
        if not a.startswith("--") or a[2:] not in spec:
            raise UsageError(f"Cannot recognize the option {a}")
        name = a[2:]
        kind = spec[name][0]
        if kind == "flag":
            out[name] = True
            i += 1
            continue
        if i + 1 >= len(argv):
            raise UsageError(f"Missing argument for option {a}")
        val = argv[i + 1]
        if not kind.startswith("multi") and name in seen:
            raise UsageError(f"Option {a} was specified more than once")
        seen.add(name)
        conv = {"int": int, "float": float, "str": str, "multi_float": float, "multi_str": str}[kind]
        try:
            x = conv(val)
        except ValueError:
            raise UsageError(f"Invalid value {val} for option {a}")
        if kind.startswith("multi"):
            out[name].append(x)
        else:
            out[name] = x
        i += 2
    return out


DEMUXLET_SPEC = {
    "sam": ("str", ""), "tag-group": ("str", "CB"), "tag-UMI": ("str", "UB"),
    "plp": ("str", ""), "vcf": ("str", ""), "field": ("str", "GP"),
    "geno-error-offset": ("float", 0.1), "geno-error-coeff": ("float", 0.0), "r2-info": ("str", "R2"),
    "min-mac": ("int", 1), "min-callrate": ("float", 0.5), "sm": ("multi_str", []), "sm-list": ("str", ""),
    "out": ("str", ""), "alpha": ("multi_float", []), "doublet-prior": ("float", 0.5),
    "sam-verbose": ("int", 1000000), "vcf-verbose": ("int", 10000),
    "cap-BQ": ("int", 20), "min-BQ": ("int", 13), "min-MQ": ("int", 20), "min-TD": ("int", 0), "excl-flag": ("int", 3844),
    "group-list": ("str", ""), "min-total": ("int", 0), "min-umi": ("int", 0), "min-snp": ("int", 0),
    "gpus": ("int", 1),  # extension: GPUs of this box to shard the barcodes over (0 = all)
}

FREEMUX_COMMON = {
    "plp": ("str", ""), "init-cluster": ("str", ""), "out": ("str", ""), "nsample": ("int", 0),
    "aux-files": ("flag", False), "verbose": ("int", 100), "doublet-prior": ("float", 0.5),
    "bf-thres": ("float", 5.41), "frac-init-clust": ("float", 1.0), "iter-init": ("int", 10),
    "keep-init-missing": ("flag", False), "min-BQ": ("int", 13), "group-list": ("str", ""),
    "min-total": ("int", 0), "min-snp": ("int", 0),
    "gpus": ("int", 1),  # extension: GPUs of this box to shard the SNPs over (0 = all)
}
FREEMUXLET_SPEC = dict(FREEMUX_COMMON, **{"geno-error": ("float", 0.1), "cap-BQ": ("int", 20), "min-umi": ("int", 0),
                                         "randomize-singlet-score": ("flag", False), "seed": ("int", 0)})
FREEMUXLET_OLD_SPEC = dict(FREEMUX_COMMON, **{"geno-error": ("float", 0.0), "cap-BQ": ("int", 40), "min-uniq": ("int", 0)})


def _read_list(path):
    return [r[0] for r in plpio._rows(path)]


def _default_engine(gpus=1):
    """One GPU (Context), or `--gpus N` / PSCL_GPU_IDS=0,1,... GPUs of this box (Multi: pscl_multi_* of the C ABI).
    Raises without libpopscle_b200.so or an sm_100 GPU: no CPU fallback."""
    import os
    from .capi import Context, Multi
    ids = [int(x) for x in os.environ.get("PSCL_GPU_IDS", "").split(",") if x != ""]
    if len(ids) > 1:
        return Multi(gpu_ids=ids)
    if not ids and gpus != 1:
        return Multi(n_gpu=gpus)
    return Context(ids[0] if ids else 0)


def demuxlet(argv, engine=None):
    o = _parse(argv, DEMUXLET_SPEC)
    alphas = o["alpha"] or [0.0, 0.5]  # cmd_cram_demuxlet.cpp:85-89 (the help text says 0.1..0.5; the code wins)
    if o["sam"]:
        raise UsageError("--sam (BAM/CRAM pileup on the fly) needs htslib; run `popscle dsc-pileup` first and pass --plp")
    if not o["plp"] or not o["vcf"] or not o["out"]:
        raise UsageError("Missing required option(s) : --plp (or --sam), --vcf, --out")
    sm = list(o["sm"]) + (_read_list(o["sm-list"]) if o["sm-list"] else [])
    _notice(f"Loading pileup information with prefix {o['plp']}")
    L = plpio.load_plp(o["plp"], o["vcf"], field=o["field"], geno_error_offset=o["geno-error-offset"], geno_error_coeff=o["geno-error-coeff"],
                       r2_info=o["r2-info"], sm_list=sm or None,
                       min_bq=o["min-BQ"], cap_bq=o["cap-BQ"], min_read=o["min-total"], min_umi=o["min-umi"], min_snp=o["min-snp"],
                       group_list=_read_list(o["group-list"]) if o["group-list"] else None,
                       min_mac=o["min-mac"], min_callrate=o["min-callrate"])
    _notice("Starting to identify best matching individual IDs")
    own = engine is None
    eng = engine or _default_engine(o["gpus"])
    try:
        # the compact boundary, as the C++ host uses it: 8-bit SNP gaps / 2-bit counts (ABI 6) when they fit, raw posteriors or
        # hard calls + error rates (ABI 4) mixed on the device
        if getattr(eng, "accepts_compact", False):
            cells = eng.demux_run(L.plp, L.geno.raw(), L.geno.has_gp, alphas, o["doublet-prior"], compact=4)
        else:  # an engine with the plain interface (the tests' CPU stand-in)
            cells = eng.demux_run(L.plp, L.geno.gp, L.geno.has_gp, alphas, o["doublet-prior"])
    finally:
        if own:
            eng.close()
    report.write_best(o["out"] + ".best", cells, L.barcodes, L.cell_uniq_reads, L.geno.samples, alphas,
                      min_total=o["min-total"], min_umi=o["min-umi"], min_snp=o["min-snp"], totl_reads=L.cell_totl_reads)
    _notice("Finished writing output files")
    return 0


def _freemux(argv, old, engine=None):
    o = _parse(argv, FREEMUXLET_OLD_SPEC if old else FREEMUXLET_SPEC)
    if not o["plp"] or not o["out"] or o["nsample"] == 0:
        raise UsageError("Missing required option(s) : --plp, --out, --nsample")
    nS = o["nsample"]
    if old:
        # cmd_cram_freemuxlet.cpp:83 never copies the filter flags into the loader: library defaults apply
        L = plpio.load_plp(o["plp"], None)
    else:
        L = plpio.load_plp(o["plp"], None, min_bq=o["min-BQ"], cap_bq=o["cap-BQ"], min_read=o["min-total"], min_umi=o["min-umi"],
                           min_snp=o["min-snp"], group_list=_read_list(o["group-list"]) if o["group-list"] else None)
    init = None
    if o["init-cluster"]:
        m = {}
        for r in plpio._rows(o["init-cluster"]):
            if len(r) != 2:
                raise UsageError(f"Initial clustering file {o['init-cluster']} has to have 2 columnes")
            k = plpio._atoi(r[1])
            if k >= 0:
                if k >= nS:
                    raise UsageError(f"--nsample {nS} parameter was set. The cluster ID must be between 0 to {nS - 1}")
                m[r[0]] = k
        init = np.array([m.get(b, -1) for b in L.barcodes], dtype=np.int32)
    own = engine is None
    eng = engine or _default_engine(o["gpus"])
    try:
        opts = eng.fmx_opts(nS, doublet_prior=o["doublet-prior"], geno_error=o["geno-error"], max_iter=10, early_stop=True,
                            frac_init_clust=o["frac-init-clust"], singlet_score_thres=-1e300, mode_old=old,
                            randomize_singlet_score=bool(o.get("randomize-singlet-score", False)), seed=int(o.get("seed", 0) or 0),
                            **(dict(bf_thres=o["bf-thres"], iter_init=o["iter-init"], keep_init_missing=o["keep-init-missing"]) if old else {}))
        gl0 = cnt0 = None
        if o["aux-files"]:
            # .clust0.samples.gz / .clust0.vcf.gz (cmd_cram_freemux2.cpp:265-347).  freemuxlet-old's aux files are dumps of its
            # pairwise distance matrix (.ldist.gz, cmd_cram_freemuxlet.cpp:176-181, :348-430), which this engine never
            # materialises (two bits per droplet pair on the device)
            if old:
                raise UsageError("--aux-files is not supported by freemuxlet-old here (its .ldist dumps need the pairwise distance matrix)")
            if not hasattr(eng, "fmx_run_aux"):
                raise UsageError("--aux-files needs a single GPU (--gpus 1)")
            cells, res, gl, cnt, gl0, cnt0 = eng.fmx_run_aux(L.plp, opts, init)
        else:
            cells, res, gl, cnt = eng.fmx_run(L.plp, opts, init, want_clusters=True)
    finally:
        if own:
            eng.close()
    if gl0 is not None:
        report.write_clust0_samples(o["out"] + ".clust0.samples.gz", cells, L.barcodes)
        report.write_clust_vcf(o["out"] + ".clust0.vcf.gz", L.sites, L.rid2chr, gl0, cnt0, report.observed_snps(L.plp), initial=True)
    report.write_lmix(o["out"] + ".lmix", cells, L.barcodes, old=old)
    report.write_clust_samples(o["out"] + ".clust1.samples.gz", cells, L.barcodes)
    report.write_clust_vcf(o["out"] + ".clust1.vcf.gz", L.sites, L.rid2chr, gl, cnt, report.observed_snps(L.plp))
    return 0


def freemuxlet(argv, engine=None):
    return _freemux(argv, False, engine)


def freemuxlet_old(argv, engine=None):
    return _freemux(argv, True, engine)


COMMANDS = {"demuxlet": demuxlet, "freemuxlet": freemuxlet, "freemuxlet-old": freemuxlet_old}


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if not argv or argv[0] not in COMMANDS:
        sys.stderr.write("usage: popscle demuxlet|freemuxlet|freemuxlet-old [options]\n")
        return 2
    try:
        return COMMANDS[argv[0]](argv[1:])
    except UsageError as e:
        sys.stderr.write("\nFATAL ERROR - \n%s\n\n" % e)
        return 134  # the reference aborts (uncaught pexception)


if __name__ == "__main__":
    sys.exit(main())
