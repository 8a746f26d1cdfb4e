"""CPU suite, part 3: the Winograd F(2x2, 3x3) transform passes (wacv23_tsnet_b200/csrc/wino_passes.cuh) executed by
the host emulation (oracle/wino_emul.cu: the SAME kernel bodies, run by loops over (block, thread)) against torch.

Covers the weight transform, the input pass "T" (InstanceNorm + ReLU + residual + reflect pad + B^T d B + hi/lo split,
act_out), the output pass "I" (A^T M A + bias + addend, InstanceNorm partial statistics) and the whole chain
T -> 16 plane GEMMs (emulated in fp64 with the layouts tsnet_wino_gemm_fwd assumes) -> I against an fp64
ReflectionPad2d(1) + Conv2d of the reference's ResnetBlock (model/TSNet.py:19-27)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BT = torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=torch.float64)
G = torch.tensor([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], dtype=torch.float64)
AT = torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=torch.float64)


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libwino_emul.so"))
    vp = C.c_void_p
    lib.wino_emul_weight.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.wino_emul_input.argtypes = [vp, vp, vp, vp, vp, vp] + [C.c_int] * 10 + [C.c_float, C.c_int]
    lib.wino_emul_output.argtypes = [vp, vp, vp, C.c_longlong, vp, vp] + [C.c_int] * 5
    return lib


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _recon(hi, lo, fmt=0):
    dt = torch.float16 if fmt == 0 else torch.bfloat16
    return hi.view(dt).double() + lo.view(dt).double()


def v_logical(t):
    """V as stored, [B, 16, Cp/64, TH, TW, 64] (K-block-major), -> [B, 16, TH, TW, Cp]"""
    B, P, KB, TH, TW, _ = t.shape
    return t.permute(0, 1, 3, 4, 2, 5).reshape(B, P, TH, TW, KB * 64)


def m_slab(m_logical, B, Cc):
    """M logical [16, B * T, C] -> as stored, slab-major [16, B, C/32, T, 32]"""
    T = m_logical.shape[1] // B
    return m_logical.view(16, B, T, Cc // 32, 32).permute(0, 1, 3, 2, 4).contiguous()


def run_input(emul, x, mean_rstd=None, relu=False, residual=None, want_act=False, scale=16.0, fmt=0, nthreads=64,
              Cp_total=None, c_off=0, act_extra=0):
    B, H, W, Cc = x.shape
    Cp_total = Cp_total or (Cc + 63) // 64 * 64
    hi = torch.zeros((B, 16, Cp_total // 64, H // 2, W // 2, 64), dtype=torch.int16)
    lo = torch.zeros_like(hi)
    act = torch.zeros((B, H, W, Cc + act_extra), dtype=torch.float32) if want_act else None
    emul.wino_emul_input(_p(x), _p(mean_rstd), _p(residual), _p(act), _p(hi), _p(lo), B, H, W, Cc, int(relu), Cp_total,
                         c_off, fmt, Cc + act_extra, act_extra, scale, nthreads)
    return v_logical(hi), v_logical(lo), act


def run_output(emul, m, B, H, W, Cc, bias=None, addend=None, want_stats=True, nthreads=64):
    y = torch.zeros((B, H, W, Cc), dtype=torch.float32)
    stats = torch.zeros((B * H * W // 32, Cc, 2), dtype=torch.float32) if want_stats else None
    rows = 0 if addend is None else addend.numel() // Cc
    ms = m_slab(m, B, Cc)
    emul.wino_emul_output(_p(ms), _p(bias), _p(addend), rows, _p(y), _p(stats), B, H, W, Cc, nthreads)
    return y, stats


def ref_input_transform(v_nhwc):
    """B^T d B of every 4x4 tile (stride 2) of the reflect-padded activation -> [B, 16, H/2, W/2, C] (fp64)."""
    xp = F.pad(v_nhwc.permute(0, 3, 1, 2).double(), (1, 1, 1, 1), mode="reflect")
    d = xp.unfold(2, 4, 2).unfold(3, 4, 2)                                  # [B, C, TH, TW, 4, 4]
    V = torch.einsum("ik,bcxykl,jl->bijxyc", BT, d, BT)                     # [B, 4, 4, TH, TW, C]
    return V.reshape(V.shape[0], 16, *V.shape[3:])


def test_weight_transform(emul):
    torch.manual_seed(0)
    w = torch.randn(8, 12, 3, 3) * 0.05
    u = torch.zeros(16, 8, 12)
    emul.wino_emul_weight(_p(w), 8, 12, _p(u))
    ref = torch.einsum("ik,ockl,jl->ijoc", G, w.double(), G).reshape(16, 8, 12)
    assert float((u.double() - ref).abs().max()) < 1e-8
    assert torch.equal(u, ref.float())   # fp64 arithmetic, one rounding


@pytest.mark.parametrize("H,W,Cc,nthreads", [(32, 32, 16, 64), (8, 16, 8, 3), (4, 4, 4, 1), (16, 48, 8, 256)])
def test_input_pass(emul, H, W, Cc, nthreads):
    torch.manual_seed(1)
    B = 2
    x = torch.randn(B, H, W, Cc) * 3
    res = torch.randn(B, H, W, Cc)
    mr = torch.stack([x.mean((1, 2)), 1.0 / torch.sqrt(x.var((1, 2), unbiased=False) + 1e-5)], -1).contiguous()
    # plain: just pad + transform + split
    hi, lo, _ = run_input(emul, x, nthreads=nthreads)
    ref = ref_input_transform(x) * 16.0
    got = _recon(hi, lo)
    assert float((got[..., :Cc] - ref).abs().max() / ref.abs().max()) < 2e-6
    assert int(hi[..., Cc:].abs().max()) == 0                      # channels padded up to the 64-wide K block stay zero
    # InstanceNorm + ReLU + residual + act_out (with a channel offset in a wider act_out and operand buffer)
    hi, lo, act = run_input(emul, x, mean_rstd=mr, relu=True, residual=res, want_act=True, nthreads=nthreads,
                            Cp_total=128, c_off=8, act_extra=4)
    v = torch.relu((x - mr[..., 0].view(B, 1, 1, Cc)) * mr[..., 1].view(B, 1, 1, Cc)) + res
    assert torch.equal(act[..., 4:], v) and float(act[..., :4].abs().max()) == 0.0
    ref = ref_input_transform(v) * 16.0
    got = _recon(hi, lo)
    assert float((got[..., 8:8 + Cc] - ref).abs().max() / ref.abs().max()) < 2e-6
    assert int(hi[..., :8].abs().max()) == 0 and int(lo[..., :8].abs().max()) == 0
    assert int(hi[..., 8 + Cc:].abs().max()) == 0


def test_input_pass_bf16_and_saturation(emul):
    torch.manual_seed(2)
    x = torch.randn(1, 8, 8, 4)
    hi, lo, _ = run_input(emul, x, scale=1.0, fmt=1)
    ref = ref_input_transform(x)
    assert float((_recon(hi, lo, 1)[..., :4] - ref).abs().max() / ref.abs().max()) < 1e-4
    x[0, 3, 3, 0] = 3.0e4                                   # 16 * 3e4 overflows fp16: must saturate, not turn into NaN
    hi, lo, _ = run_input(emul, x)
    assert torch.isfinite(_recon(hi, lo)).all()


@pytest.mark.parametrize("H,W,Cc,nthreads", [(32, 32, 32, 64), (4, 16, 64, 1), (8, 32, 32, 5)])
def test_output_pass(emul, H, W, Cc, nthreads):
    torch.manual_seed(3)
    B = 3
    T = (H // 2) * (W // 2)
    m = torch.randn(16, B * T, Cc)
    bias = torch.randn(Cc)
    addend = torch.randn(H * W, Cc)                          # one sample's worth of rows: broadcast over the batch
    y, stats = run_output(emul, m, B, H, W, Cc, bias=bias, addend=addend, nthreads=nthreads)
    M = m.double().view(4, 4, B, H // 2, W // 2, Cc)
    Y = torch.einsum("ai,ijbxyc,ej->bxaye c".replace(" ", ""), AT, M, AT)   # [B, TH, 2, TW, 2, C]
    ref = Y.reshape(B, H, W, Cc) + bias.double() + addend.double().view(1, H, W, Cc)
    assert float((y.double() - ref).abs().max()) < 2e-5
    # statistics partials -> mean / biased variance of y exactly as tsnet_instnorm_reduce merges them
    s = stats.double().view(B, H * W // 32, Cc, 2)
    mean = s[..., 0].sum(1) / (H * W)
    q = (s[..., 1] + s[..., 0] ** 2 / 32.0).sum(1)
    var = q / (H * W) - mean ** 2
    yd = y.double().view(B, -1, Cc)
    assert float((mean - yd.mean(1)).abs().max()) < 1e-5
    assert float((var - yd.var(1, unbiased=False)).abs().max() / yd.var(1, unbiased=False).max()) < 1e-5
    y2, _ = run_output(emul, m, B, H, W, Cc, want_stats=False, nthreads=nthreads)
    assert float((y2.double() - Y.reshape(B, H, W, Cc)).abs().max()) < 2e-5


def test_whole_chain_equals_reflect_pad_conv(emul):
    """T -> 16 plane GEMMs over the hi/lo operands (layouts of tsnet_wino_gemm_fwd: A = V[b, p] rows = tiles, B rows
    p * Cout + o, M = [16, B * tiles, Cout]) -> I  ==  Conv2d(ReflectionPad2d(1)(x)) in fp64."""
    torch.manual_seed(4)
    B, H, W, Cin, Cout = 2, 32, 32, 16, 32
    x = torch.randn(B, H, W, Cin) * 2
    w = torch.randn(Cout, Cin, 3, 3) * 0.05
    bias = torch.randn(Cout)
    u = torch.zeros(16, Cout, Cin)
    emul.wino_emul_weight(_p(w), Cout, Cin, _p(u))
    wscale = 2.0 ** np.floor(np.log2(8192.0 / float(u.abs().max())))
    us = (u * wscale).view(16 * Cout, Cin)
    u_hi = us.half()
    u_lo = (us - u_hi.float()).half()
    hi, lo, _ = run_input(emul, x)
    V = _recon(hi, lo)[..., :Cin].reshape(B, 16, (H // 2) * (W // 2), Cin)  # [B, 16, tiles, C]
    U = (u_hi.double() + u_lo.double()).view(16, Cout, Cin)
    # hi*hi + hi*lo + lo*hi (the dropped lo*lo term is 2^-22 relative) ~ full product of the reconstructed operands
    M = torch.einsum("bptc,poc->pbto", V, U).reshape(16, B * (H // 2) * (W // 2), Cout) / (wscale * 16.0)
    y, _ = run_output(emul, M.float().contiguous(), B, H, W, Cout, bias=bias)
    ref = F.conv2d(F.pad(x.permute(0, 3, 1, 2).double(), (1, 1, 1, 1), mode="reflect"), w.double(), bias.double())
    ref = ref.permute(0, 2, 3, 1)
    assert float((y.double() - ref).abs().max() / ref.abs().max()) < 2e-6


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("H,W,Cc,relu,with_res,nthreads", [(32, 32, 64, True, False, 512), (32, 32, 32, False, True, 512),
                                                           (8, 16, 32, True, True, 64), (4, 4, 64, False, False, 32)])
def test_bridge_pass_equals_output_norm_input_chain(emul, H, W, Cc, relu, with_res, nthreads, variant):
    """The fused bridge (csrc/wino_passes.cuh: phases A, S, B, C) == output transform + bias + addend -> InstanceNorm ->
    [ReLU | + residual] -> reflect pad + input transform, evaluated in torch fp64."""
    vp = C.c_void_p
    emul.wino_emul_bridge.argtypes = [vp, vp, vp, C.c_longlong, vp, vp, vp, vp, vp] + [C.c_int] * 10 + \
                                     [C.c_float, C.c_float, C.c_int, vp, vp, vp, vp, C.c_float, C.c_int]
    torch.manual_seed(8)
    B = 2
    T = (H // 2) * (W // 2)
    m = torch.randn(16, B * T, Cc)
    bias = torch.randn(Cc)
    addend = torch.randn(H * W, Cc)
    res = torch.randn(B, H, W, Cc) if with_res else None
    Cp = (Cc + 32 + 63) // 64 * 64
    hi = torch.zeros((B, 16, Cp // 64, H // 2, W // 2, 64), dtype=torch.int16)
    lo = torch.zeros_like(hi)
    act = torch.zeros(B, H, W, Cc + 8)
    mr = torch.zeros(B, Cc, 2)
    ms = m_slab(m, B, Cc)
    # correlation operand rows at a (random) rank permutation + per-slab partial sums of squares
    rank = torch.stack([torch.randperm(H * W) for _ in range(B)]).to(torch.int16)
    c_hi = torch.zeros(B * H * W, Cc, dtype=torch.int16)
    c_lo = torch.zeros_like(c_hi)
    CS = 16 if variant else 32
    ssq = torch.zeros(B, Cc // CS, H * W)
    emul.wino_emul_bridge(_p(ms), _p(bias), _p(addend), H * W, _p(res), _p(act), _p(mr), _p(hi), _p(lo), B, H, W, Cc,
                          int(relu), Cp, 32, 0, Cc + 8, 8, 16.0, 1e-5, nthreads, _p(c_hi), _p(c_lo), _p(rank), _p(ssq), 16.0,
                          variant)
    hi, lo = v_logical(hi), v_logical(lo)
    a_out = act[..., 8:]
    idx = (rank.long() & 0xFFFF).view(B, H * W, 1).expand(B, H * W, Cc)
    rows = _recon(c_hi, c_lo).view(B, H * W, Cc).gather(1, idx)                # row at rank[pix] -> position pix
    assert float((rows - a_out.double().view(B, H * W, Cc) * 16.0).abs().max() / (a_out.abs().max() * 16.0)) < 2e-6
    ssq_ref = (a_out.double() ** 2).view(B, H * W, Cc // CS, CS).sum(-1).permute(0, 2, 1)
    assert float((ssq.double() - ssq_ref).abs().max() / ssq_ref.max()) < 1e-6
    if not with_res:
        # without residual / act_out / correlation outputs the in-place phase B is skipped and phase C normalises on read:
        # the operands must be the same bits
        hi2 = torch.zeros((B, 16, Cp // 64, H // 2, W // 2, 64), dtype=torch.int16)
        lo2 = torch.zeros_like(hi2)
        emul.wino_emul_bridge(_p(ms), _p(bias), _p(addend), H * W, None, None, None, _p(hi2), _p(lo2), B, H, W, Cc,
                              int(relu), Cp, 32, 0, Cc, 0, 16.0, 1e-5, nthreads, None, None, None, None, 16.0, variant)
        assert torch.equal(v_logical(hi2), hi) and torch.equal(v_logical(lo2), lo)
    M = m.double().view(4, 4, B, H // 2, W // 2, Cc)
    y = torch.einsum("ai,ijbxyc,ej->bxayec", AT, M, AT).reshape(B, H, W, Cc) + bias.double() + \
        addend.double().view(1, H, W, Cc)
    mean = y.mean((1, 2), keepdim=True)
    var = y.var((1, 2), unbiased=False, keepdim=True)
    v = (y - mean) / torch.sqrt(var + 1e-5)
    if relu:
        v = torch.relu(v)
    if with_res:
        v = v + res.double()
    assert float((mr[..., 0].double() - mean.view(B, Cc)).abs().max()) < 1e-6
    assert float((mr[..., 1].double() * torch.sqrt(var + 1e-5).view(B, Cc) - 1).abs().max()) < 1e-6
    assert float((act[..., 8:].double() - v).abs().max()) < 1e-5 and float(act[..., :8].abs().max()) == 0.0
    # the operands are the transform of the fp32 activations the pass itself produced (act_out)
    ref = ref_input_transform(act[..., 8:]) * 16.0
    got = _recon(hi, lo)
    assert float((got[..., 32:32 + Cc] - ref).abs().max() / ref.abs().max()) < 2e-6
    assert int(hi[..., :32].abs().max()) == 0 and (32 + Cc == Cp or int(hi[..., 32 + Cc:].abs().max()) == 0)
