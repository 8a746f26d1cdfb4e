"""K1 experiment: accuracy of the warp grids against an fp64 evaluation (random and strongly correlated features) and
the time of corr_warp alone at bs=32, n=3.  Variants are selected through environment variables read by the library
(e.g. TSNET_K1_CHUNK_KB).  Test infrastructure: uses the oracle as the checker."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wacv23_tsnet_b200 import ops
from oracle import tsnet_oracle as O

m = ops.MathMode("fp16x3")
coord = torch.cat([torch.linspace(-1, 1, 32), torch.linspace(-1, 1, 32)]).cuda()


def inputs(kind, B, n, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "random":
        tar = torch.relu(torch.randn(B, 512, 32, 32, generator=g))
        srcs = [torch.randn(B, 512, 32, 32, generator=g) * 3 for _ in range(n)]
    else:  # correlated: every source is a spatially shifted copy of the target plus noise -> cos ~ 0.95 at the match
        tar = torch.randn(B, 512, 32, 32, generator=g)
        srcs = [torch.roll(tar, shifts=(2 * i + 1, -(i + 2)), dims=(2, 3)) + 0.3 * torch.randn(B, 512, 32, 32, generator=g)
                for i in range(n)]
    tb = torch.zeros(B, 1, 256, 256, dtype=torch.uint8)
    tb[:, :, 40:200, 30:220] = 1
    sbs = []
    for i in range(n):
        sb = torch.zeros(B, 1, 256, 256, dtype=torch.uint8)
        sb[:, :, 20 + 10 * i:180, 50:230 - 10 * i] = 1
        sbs.append(sb)
    return tar, srcs, tb, sbs


def run(tar, srcs, tb, sbs):
    B, n = tar.shape[0], len(srcs)
    tar_d = tar.permute(0, 2, 3, 1).contiguous().cuda()
    src_d = torch.stack([s.permute(0, 2, 3, 1).contiguous() for s in srcs]).cuda()
    return ops.corr_grids(tar_d.view(B, 1024, 512), src_d.view(n, B, 1024, 512), tb.squeeze(1).contiguous().cuda(),
                          [s.squeeze(1).contiguous().cuda() for s in sbs], coord, m)


res = {"env": {k: v for k, v in os.environ.items() if k.startswith("TSNET_")}}
for kind in ("random", "correlated"):
    tar, srcs, tb, sbs = inputs(kind, 2, 3, 7)
    _, ref = O.corr_warp(tar, srcs, tb, sbs)
    _, tru = O.corr_warp(tar.double(), [s.double() for s in srcs], tb, sbs)
    grids = run(tar, srcs, tb, sbs)
    torch.cuda.synchronize()
    k_err = max(float((grids[i].cpu().double() - tru[i]).abs().max()) for i in range(3))
    r_err = max(float((ref[i].double() - tru[i]).abs().max()) for i in range(3))
    k_ref = max(float((grids[i].cpu() - ref[i]).abs().max()) for i in range(3))
    res[kind] = {"kernel_vs_fp64": k_err, "fp32ref_vs_fp64": r_err, "kernel_vs_fp32ref": k_ref}

B, n = 32, 3
torch.manual_seed(0)
tar = torch.relu(torch.randn(B, 1024, 512, device="cuda"))
src = torch.randn(n, B, 1024, 512, device="cuda") * 3
g = torch.Generator().manual_seed(3)
def rect():
    out = torch.zeros(B, 256, 256, dtype=torch.uint8)
    u = torch.rand(B, 4, generator=g)
    for b in range(B):
        hh, ww = int(256 * (0.55 + 0.30 * u[b, 0])), int(256 * (0.55 + 0.30 * u[b, 1]))
        y0, x0 = int((256 - hh) * u[b, 2]), int((256 - ww) * u[b, 3])
        out[b, y0:y0 + hh, x0:x0 + ww] = 1
    return out.cuda()
tb, sbs = rect(), [rect() for _ in range(n)]
def timed(fn, tag):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    res[tag + "_chain_us"] = e0.elapsed_time(e1) / 20 * 1e3
    ops.PROFILE = {}
    for _ in range(5): fn()
    torch.cuda.synchronize()
    for k, evs in ops.PROFILE.items():
        res[tag + "_" + k[0] + "_us"] = sum(a.elapsed_time(b) for a, b in evs) / 5 * 1e3
    ops.PROFILE = None

dec_hi = torch.empty((B, 32, 32, 1024), dtype=torch.int16, device="cuda"); dec_lo = torch.empty_like(dec_hi)
full = lambda tb_, sbs_, sort: ops.corr_chain(tar, src, tb_, sbs_, coord, m, want_grids=False, want_mean=False,
                                              taps=(dec_hi, dec_lo), sort=sort)
ones = torch.ones(B, 256, 256, dtype=torch.uint8, device="cuda")
timed(lambda: full(tb, sbs, True), "rect_sorted")
timed(lambda: full(tb, sbs, False), "rect_raster")
timed(lambda: full(ones, [ones] * n, True), "allone")
print(json.dumps(res))
