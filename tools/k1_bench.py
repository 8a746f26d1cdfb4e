"""Time corr_warp alone (bs=32, n=3) -- used with experimental builds selected through TSNET_LIB_PATH."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wacv23_tsnet_b200 import ops
B, n = 32, 3
m = ops.MathMode("fp16x3")
torch.manual_seed(0)
tar = torch.relu(torch.randn(B, 1024, 512, device="cuda"))
src = torch.randn(n, B, 1024, 512, device="cuda") * 3
tb = torch.zeros(B, 256, 256, dtype=torch.uint8, device="cuda"); tb[:, 40:200, 30:220] = 1
sbs = [tb.clone() for _ in range(n)]
coord = torch.cat([torch.linspace(-1, 1, 32), torch.linspace(-1, 1, 32)]).cuda()
plan = ops.corr_prepare(tb, sbs, coord, B, 512, 32, 32, m)
tar_ops = ops.l2norm_split(tar, m, rank=plan.rank_t); src_ops = ops.l2norm_split(src.view(n * B, 1024, 512), m, rank=plan.rank_s)
run = lambda: ops.corr_warp(plan, tar_ops, src_ops, [src[i] for i in range(n)], m, want_grids=True, want_mean=False)
for _ in range(3): run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print(f"{os.environ.get('TSNET_LIB_PATH', 'default')}: corr_warp {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
