"""Micro-benchmark of the Winograd GEMM (tsnet_wino_gemm_fwd) and its transform passes on the dominant layer shape
(512 -> 512 3x3 over 96 samples): time and accuracy as a function of `chunk_kb` (K blocks accumulated in TMEM before the
promotion to fp32 registers).  Run under gpurun:  python tools/wino_bench.py"""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wacv23_tsnet_b200 import lib as L, ops

m = ops.MathMode("fp16x3")
torch.manual_seed(0)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (B, Cin, Cout) in ((96, 512, 512), (96, 1024, 1024), (32, 512, 512)):
    x = torch.randn(B, 32, 32, Cin, device="cuda")
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.02
    b = torch.randn(Cout, device="cuda") * 0.1
    pw = ops.PackedWino(w, b, m)
    taps = ops.build_taps(x, m, L.TAPS_WINO)
    xs = x[:2].permute(0, 3, 1, 2)
    ref = F.conv2d(F.pad(xs, (1, 1, 1, 1), mode="reflect").double(), w.double(), b.double()).permute(0, 2, 3, 1)
    mbuf = torch.empty(16 * B * 256 * Cout, dtype=torch.float32, device="cuda")
    for ck in (2, 3, 4, 8):
        y, _ = ops.wino_conv(taps, pw, B, 32, 32, m, m.act_scale, chunk_kb=ck, m_buf=mbuf)
        err = float((y[:2].double() - ref).abs().max() / ref.abs().max())
        d = L.WinoGemmDesc()
        d.B, d.TH, d.TW, d.C, d.Cout = B, 16, 16, Cin, Cout
        d.split, d.fmt, d.out_scale, d.chunk_kb, d.flags = m.split, m.fmt, 1.0 / (pw.scale * m.act_scale), ck, 0
        import ctypes as C
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = lambda t: C.c_void_p(t.data_ptr())
        t_ms = timed(lambda: L.check(L.load().tsnet_wino_gemm_fwd(C.byref(d), p(taps[0]), p(taps[1]), p(pw.u_hi),
                                                                   p(pw.u_lo), p(mbuf), st)))
        print(f"wino_gemm B={B} {Cin}->{Cout} chunk_kb={ck}: {t_ms:.3f} ms, rel err vs fp64 conv {err:.2e}", flush=True)
    t_in = timed(lambda: ops.build_taps(x, m, L.TAPS_WINO, taps=(taps[0], taps[1])))
    print(f"  input pass {t_in:.3f} ms", flush=True)
    del x, w, pw, taps, mbuf
    torch.cuda.empty_cache()
