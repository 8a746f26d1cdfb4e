"""Compare the partial softmax states of the 2-CTA tile kernel (TSNET_K1_2CTA=1) with the 1-CTA kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wacv23_tsnet_b200 import ops
m = ops.MathMode(os.environ.get("MATH", "fp16x3"))
B, n, hw, C = 1, 1, 1024, 512
g = torch.Generator().manual_seed(3)
tar = torch.randn(B, hw, C, generator=g).cuda()
src = (torch.randn(n, B, hw, C, generator=g) * 2).cuda()
ones = torch.ones(B, 256, 256, dtype=torch.uint8, device="cuda")
coord = torch.cat([torch.linspace(-1, 1, 32), torch.linspace(-1, 1, 32)]).cuda()
al = lambda x: (x + 255) & ~255
NM, runs, chunks = (n + 1) * B, hw // 128, hw // 256
o = 0
for sz in (NM * hw * 2, NM * hw * 4, n * B * hw * 4, n * B * hw * 4, NM * runs, NM * runs * 8, B * n * runs * chunks * 4, B * 4):
    o = al(o + sz)
state_off = o
res = {}
for mode in ("0", "1"):
    os.environ["TSNET_K1_2CTA"] = mode
    plan = ops.corr_prepare(ones, [ones] * n, coord, B, C, 32, 32, m, sort=False)
    t_ops = ops.l2norm_split(tar, m, rank=plan.rank_t)
    s_ops = ops.l2norm_split(src.view(n * B, hw, C), m, rank=plan.rank_s)
    plan.ws[state_off:].zero_()
    ops.corr_warp(plan, t_ops, s_ops, [src[i] for i in range(n)], m, want_grids=True, want_mean=False)
    torch.cuda.synchronize()
    res[mode] = plan.ws[state_off:state_off + n * B * hw * 8 * 16].view(torch.float32).view(hw, 8, 4).clone()
a, b = res["0"], res["1"]
d = (a - b).abs().amax(-1)   # [row, state]
print("rows x states wrong (per 128-row tile, per state slot):")
print((d > 1e-3 * a.abs().amax()).view(8, 128, 8).sum(1).tolist())
for r in (0, 1, 127, 128, 300):
    print("row", r, "ref", [round(float(x), 3) for x in a[r, :, 0]], "| 2cta", [round(float(x), 3) for x in b[r, :, 0]])
print("ref row0 state0", a[0, 0].tolist(), "2cta", b[0, 0].tolist())
