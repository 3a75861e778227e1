#!/usr/bin/env python
"""Times the C++ host's file ingest (N1: CEL/VAR/PLP + VCF -> flat image) and, with a GPU, files -> files:
python tools/time_loader.py <cells> [--gz] [--run] [--threads T]
Writes a config-2-shaped pileup (8 samples x 100 k SNPs) under /tmp as BGZF (what dsc-pileup writes) or one gzip stream."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from popscle_b200 import plpio, synth

cells = int(sys.argv[1])
bgzf = "--gz" not in sys.argv
run = "--run" in sys.argv
thr = sys.argv[sys.argv.index("--threads") + 1] if "--threads" in sys.argv else None
d = f"/tmp/pscl_loader_{cells}_{'bgzf' if bgzf else 'gz'}"
os.makedirs(d, exist_ok=True)
pre = os.path.join(d, "pileup")
if not os.path.exists(pre + ".plp.gz"):
    t0 = time.perf_counter()
    s = synth.make_config(2, cells=cells)
    sites = plpio.default_sites(s.plp.n_snps, s.plp.snp_af)
    plpio.write_plp(pre, s.plp, sites, bgzf=bgzf)
    plpio.write_vcf(os.path.join(d, "ref.vcf.gz"), sites, [f"S{j}" for j in range(s.geno.shape[0])], geno=s.geno)
    print(f"wrote {pre}.* ({s.plp.n_pairs} pairs, {s.plp.n_reads} reads, .plp.gz {os.path.getsize(pre + '.plp.gz') >> 20} MB) in {time.perf_counter() - t0:.1f} s", flush=True)
env = dict(os.environ, PSCL_TRACE="1")
if thr:
    env["PSCL_LOADER_THREADS"] = thr
exe = os.path.join(ROOT, "popscle_b200", "popscle")
for rep in range(2):
    for mode in (["--out", os.path.join(d, "dry"), "--dry-run"],) + ((["--out", os.path.join(d, "out")],) if run else ()):
        t0 = time.perf_counter()
        r = subprocess.run([exe, "demuxlet", "--plp", pre, "--vcf", os.path.join(d, "ref.vcf.gz"), "--field", "GT"] + mode,
                           env=env, capture_output=True, text=True)
        dt = time.perf_counter() - t0
        print(f"== {' '.join(mode)}: {dt:.3f} s wall, exit {r.returncode}")
        print("\n".join(l for l in r.stderr.splitlines() if " ms" in l))
        if r.returncode:
            print(r.stderr[-2000:])
