"""Time the kw-folded 7x7 stem conv (96 x 256 x 256, Cout 64): vertical-reuse kernel vs the plain implicit GEMM."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wacv23_tsnet_b200 import ops
m = ops.MathMode(os.environ.get("MATH", "fp16x3"))
B = int(os.environ.get("B", 96))
torch.manual_seed(0)
w = torch.randn(64, 8, 7, 7, device="cuda") * 0.02
b = torch.zeros(64, device="cuda")
pc = ops.PackedConv(w, b, m, fold_kw=True)
img = torch.rand(B, 3, 256, 256, device="cuda") * 255 - 100
lbl = (torch.rand(B, 2, 256, 256, device="cuda") > 0.7).float()
hi, lo, g = ops.stem_taps(img, 255.0, lbl, pc.Cp, m)
y = torch.empty((B, 256, 256, 64), device="cuda"); st = torch.empty((B * 2048, 64, 2), device="cuda")
NOSTATS = os.environ.get("NOSTATS") == "1"
def run(): ops.conv_gemm(hi, lo, g, pc, "7x1", B, 256, 256, m, m.act_scale, y=y, stats=None if NOSTATS else st, want_stats=not NOSTATS)
for tag, env in (("vr", None), ("plain", "1")):
    if env: os.environ["TSNET_NO_VR"] = env
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    print(f"{tag}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
