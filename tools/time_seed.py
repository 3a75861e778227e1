#!/usr/bin/env python
"""Times the freemuxlet set-up phases (SNP-major view, stage 1, greedy seeding, first M-step) and one EM iteration on a
synthetic pileup: python tools/time_seed.py <cells> <nS> <snps> <kbar> [--check]  (needs a GPU)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import torch
from popscle_b200 import Context, synth

C, nS, V, kbar = (int(x) for x in sys.argv[1:5])
s = synth.make_pileup(C, nS, V, kbar, 20260104)
plp = s.plp
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream()
ctx = Context(0, stream=stream.cuda_stream)
npairs = nS * (nS + 1) // 2
st = torch.zeros(4 * C, dtype=torch.float64, device=dev)
llk = torch.zeros(C * npairs, dtype=torch.float64, device=dev)
cl = torch.zeros(C, dtype=torch.int32, device=dev)


def timed(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"{name:28s} {1e3 * (time.perf_counter() - t0):10.3f} ms", flush=True)
    return r


print(f"cells {C} nS {nS} snps {V} pairs {plp.n_pairs} reads {plp.n_reads}")
dplp = timed("upload", lambda: ctx.upload(plp, compact=True))
o = ctx.fmx_opts(nS, early_stop=False, max_iter=10)
timed("fmx_init (SNP-major view)", lambda: ctx.fmx_init(dplp, o))
timed("stage1", lambda: ctx.fmx_stage1(st.data_ptr()))
timed("seed (greedy)", lambda: ctx.fmx_seed(st.data_ptr(), None, cl.data_ptr()))
timed("mstep (initial)", lambda: ctx.fmx_mstep(cl.data_ptr()))
for it in range(3):
    timed(f"estep {it}", lambda: ctx.fmx_estep(it, llk.data_ptr()))
    timed(f"classify {it}", lambda: ctx.fmx_classify(llk.data_ptr(), cl.data_ptr()))
    timed(f"mstep {it}", lambda: ctx.fmx_mstep(None))
seeds = cl.cpu().numpy()
cells, _, _ = ctx.fmx_fetch()
print("seed cluster sizes", np.bincount(cells["init_clust"][cells["init_clust"] >= 0], minlength=nS).tolist())
ctx.close()
