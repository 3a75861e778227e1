#!/usr/bin/env python
"""Per-call wall time of the end-to-end pscl_demux_run call (configs[1], pinned compact arrays), with and without Python's
cyclic GC, to tell host-side jitter of the harness from the library's own time."""
import gc, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, torch
from popscle_b200 import Context, Pileup, RawGeno, synth
from popscle_b200.capi import DEMUX_CELL_DTYPE, bind_to_device
bind_to_device(0)
s = synth.make_config(2)
plp = s.plp
keep = []
def pin(a):
    t = torch.from_numpy(a).pin_memory(); keep.append(t); return t.numpy()
p32, aq = plp.compact()
c4 = tuple(pin(x) for x in plp.compact4())
pr = plp.packed_reads()
h = Pileup(plp.n_cells, plp.n_snps, pin(plp.cell_ptr), plp.pair_snp, plp.pair_read_ptr, plp.read_allele, plp.read_qual, None)
h._compact = (p32, aq); h._compact4 = c4; h._packed_reads = (pin(pr[0]), pin(pr[1]), pr[2])
g = RawGeno(gt8=pin(np.ascontiguousarray(s.geno.T.astype(np.uint8))), err=0.1)
out_t = torch.empty(plp.n_cells * DEMUX_CELL_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
out = out_t.numpy().view(DEMUX_CELL_DTYPE)
ctx = Context(0)
for _ in range(3):
    ctx.demux_run(h, g, None, [0.0, 0.5], 0.5, compact=4, out=out)
for mode in ("gc on", "gc off", "gc on", "gc off"):
    if mode == "gc off":
        gc.collect(); gc.disable()
    else:
        gc.enable()
    ts = []
    for _ in range(30):
        t0 = time.perf_counter(); ctx.demux_run(h, g, None, [0.0, 0.5], 0.5, compact=4, out=out); ts.append(round(1e3 * (time.perf_counter() - t0), 2))
    print(mode, "median %.3f mean %.3f" % (np.median(ts), np.mean(ts)), ts)
