"""Device-resident kernel time of k_demux_default's variants on bench.py's workload (tools only; numbers for profiles/)."""
import argparse, sys
import numpy as np
import torch

sys.path.insert(0, ".")
import bench
from popscle_b200 import Context, _build

ap = argparse.ArgumentParser()
ap.add_argument("--kernels", default="dict,cls,dict,cls")
ap.add_argument("--steps", type=int, default=15)
a = ap.parse_args()
_build.build_cuda()
sys.argv = sys.argv[:1]
args = bench.parse_args()
cfg, s, gp = bench.make_workload(args, 0)
torch.cuda.set_device(0)
stream = torch.cuda.current_stream()
ctx = Context(0, stream=stream.cuda_stream)
d = ctx.upload(s.plp)
ctx.demux_set_geno(gp, None, s.plp.n_snps)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
names = {"lane": 1, "dict": 6, "cls": 7, "auto": 0}
ref = None
for k in a.kernels.split(","):
    ctx.demux_select_kernel(names[k])
    for _ in range(3):
        ctx.demux_score(d, bench.ALPHAS, 0.5)
    torch.cuda.synchronize()
    ms, tot = [], []
    for _ in range(a.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.demux_score(d, bench.ALPHAS, 0.5)
        e1.record(stream)
        torch.cuda.synchronize()
        ms.append(ctx.demux_last_kernel_ms()[0]); tot.append(e0.elapsed_time(e1))
    rec = ctx.demux_fetch()
    if ref is None:
        ref = rec.tobytes()
    print("%-5s ran %d  kernel %.4f ms (min %.4f)  step %.4f ms  identical %s" % (k, ctx.demux_last_kernel(), np.mean(ms), np.min(ms), np.mean(tot),
                                                                                  rec.tobytes() == ref), flush=True)
