"""GPU suite, part 1: every kernel of libtsnet_sm100.so through the C ABI against a plain torch reference of the same
op, plus edge cases of the correlation kernel against the oracle.  Convolutions are checked against an fp64 torch
conv: cuDNN's "fp32" 3x3 algorithms (Winograd class) are themselves ~1e-5 off at C=512, ten times the error of the
kernel under test."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _fp32_reference_math():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from wacv23_tsnet_b200 import lib
    lib.require_device()  # fail loudly if the CUDA extension / device is missing
    yield


def _recon(hi, lo, fmt):
    dt = torch.float16 if fmt == 0 else torch.bfloat16
    return hi.view(dt).float() + lo.view(dt).float()


def _relerr(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# conv tolerance per math mode, relative to max|ref| of an fp64 evaluation, independent of K thanks to the chunked
# accumulation.  fp16x3 / bf16x3 carry 22 / 16 operand bits.
CONV_TOL = {"fp16x3": 2e-6, "bf16x3": 6e-5, "fp16": 3e-3, "bf16": 3e-2}


def _conv_case(B, H, W, Cin, Cout, kind, mode_name, block_n=None, seed=0):
    from wacv23_tsnet_b200 import lib as L, ops
    torch.manual_seed(seed)
    m = ops.MathMode(mode_name)
    x = torch.randn(B, H, W, Cin, device="cuda")
    xn = x.permute(0, 3, 1, 2)
    k = 1 if kind == "1x1" else 3
    w = torch.randn(Cout, Cin, k, k, device="cuda") * 0.05
    b = torch.randn(Cout, device="cuda")
    wd, bd = w.double(), b.double()
    if kind == "1x1":
        tm, Ho, Wo, ref = L.TAPS_SAME, H, W, F.conv2d(xn.double(), wd, bd)
    elif kind == "3x3":
        tm, Ho, Wo = L.TAPS_REFLECT1, H, W
        ref = F.conv2d(F.pad(xn, (1, 1, 1, 1), mode="reflect").double(), wd, bd)
    elif kind == "3x3s2":
        tm, Ho, Wo, ref = L.TAPS_S2ZERO, H // 2, W // 2, F.conv2d(xn.double(), wd, bd, stride=2, padding=1)
    else:  # up3x3: the up-sampling itself is fp32 in the reference too
        up = F.interpolate(xn, scale_factor=2, mode="bilinear", align_corners=False)
        tm, Ho, Wo = L.TAPS_UP2REFLECT1, 2 * H, 2 * W
        ref = F.conv2d(F.pad(up, (1, 1, 1, 1), mode="reflect").double(), wd, bd)
    ref = ref.permute(0, 2, 3, 1).contiguous().float()
    pc = ops.PackedConv(w, b, m, block_n=block_n)
    hi, lo, g = ops.build_taps(x, m, tm)
    y, stats = ops.conv_gemm(hi, lo, g, pc, "3x3" if kind == "up3x3" else kind, B, Ho, Wo, m, m.act_scale)
    mr = ops.instnorm_reduce(stats, B, Ho * Wo, Cout)
    torch.cuda.synchronize()
    return y, mr, ref


@pytest.mark.parametrize("mode", ["fp16x3", "bf16x3", "fp16"])
@pytest.mark.parametrize("B,H,W,Cin,Cout,kind,bn", [
    (1, 32, 32, 64, 64, "1x1", None),
    (2, 32, 32, 1024, 512, "1x1", None),       # FuseNet.conv / Decoder.map_conv shape
    (2, 32, 32, 128, 128, "3x3", None),
    (2, 64, 64, 64, 128, "3x3s2", None),
    (1, 32, 32, 512, 256, "up3x3", None),       # decoder up-conv 1
])
def test_conv_gemm_vs_torch(B, H, W, Cin, Cout, kind, bn, mode):
    y, mr, ref = _conv_case(B, H, W, Cin, Cout, kind, mode, bn)
    assert _relerr(y, ref) < CONV_TOL[mode]
    if mode.endswith("x3"):
        assert _relerr(mr[..., 0], ref.mean((1, 2))) < 2e-5
        assert _relerr(mr[..., 1], 1.0 / torch.sqrt(ref.var((1, 2), unbiased=False) + 1e-5)) < 5e-5


@pytest.mark.parametrize("B,H,W,Cin,Cout,kind,bn", [
    (3, 32, 32, 512, 512, "3x3", 256),           # the dominant layer: ResnetBlock(512)
    (3, 32, 32, 512, 512, "3x3", 128),
    (1, 32, 32, 1024, 1024, "3x3", 256),         # FuseNet ResnetBlock(1024)
    (2, 256, 256, 64, 128, "3x3s2", None),       # first down-sampling conv (W = 256 > tile width)
    (1, 128, 128, 128, 64, "up3x3", None),       # last up-conv, Cout = 64
    (150, 32, 32, 64, 64, "1x1", None),          # more tiles than SMs: persistent loop + TMEM double buffering
])
def test_conv_gemm_network_shapes(B, H, W, Cin, Cout, kind, bn):
    y, mr, ref = _conv_case(B, H, W, Cin, Cout, kind, "fp16x3", bn)
    assert _relerr(y, ref) < CONV_TOL["fp16x3"]
    assert _relerr(mr[..., 0], ref.mean((1, 2))) < 2e-5


def test_conv_gemm_two_cta_variant_is_bit_exact():
    """conv_gemm2_kernel (tcgen05.mma.cta_group::2, the default for block_n = 256) against the 1-CTA kernel
    (tsnet_conv_desc.flags = TSNET_CONV_ONE_CTA)."""
    from wacv23_tsnet_b200 import ops
    from wacv23_tsnet_b200 import lib as L
    m = ops.MathMode("fp16x3")
    torch.manual_seed(12)
    x = torch.randn(24, 32, 32, 512, device="cuda")
    w = torch.randn(512, 512, 3, 3, device="cuda") * 0.02
    b = torch.randn(512, device="cuda") * 0.1
    pc = ops.PackedConv(w, b, m)
    hi, lo, g = ops.build_taps(x, m, L.TAPS_REFLECT1)
    y2, s2 = ops.conv_gemm(hi, lo, g, pc, "3x3", 24, 32, 32, m, m.act_scale)
    y1, s1 = ops.conv_gemm(hi, lo, g, pc, "3x3", 24, 32, 32, m, m.act_scale, flags=L.CONV_ONE_CTA)
    torch.cuda.synchronize()
    assert torch.equal(y1, y2) and torch.equal(s1, s2)


def test_conv_gemm_tail_wave_split_is_bit_exact():
    """512 -> 512 3x3 over 24 samples = 384 tiles of N = 256 on 148 SMs (2.6 waves): the launcher recomputes the last
    partial wave with N = 128 tiles; per-element accumulation order does not depend on the tile width."""
    from wacv23_tsnet_b200 import ops
    from wacv23_tsnet_b200 import lib as L
    m = ops.MathMode("fp16x3")
    torch.manual_seed(11)
    x = torch.randn(24, 32, 32, 512, device="cuda")
    w = torch.randn(512, 512, 3, 3, device="cuda") * 0.02
    b = torch.randn(512, device="cuda") * 0.1
    pc = ops.PackedConv(w, b, m)
    hi, lo, g = ops.build_taps(x, m, L.TAPS_REFLECT1)
    y1, s1 = ops.conv_gemm(hi, lo, g, pc, "3x3", 24, 32, 32, m, m.act_scale)
    y2, s2 = ops.conv_gemm(hi, lo, g, pc, "3x3", 24, 32, 32, m, m.act_scale, flags=L.CONV_NO_TAIL_SPLIT)
    torch.cuda.synchronize()
    assert torch.equal(y1, y2) and torch.equal(s1, s2)


def test_conv_gemm_is_deterministic_and_batch_invariant():
    """Same sample -> same bits, whatever the batch it rides in (required for shard == single-GPU equality)."""
    from wacv23_tsnet_b200 import lib as L, ops
    m = ops.MathMode("fp16x3")
    torch.manual_seed(1)
    x = torch.randn(5, 32, 32, 256, device="cuda")
    w = torch.randn(256, 256, 3, 3, device="cuda") * 0.05
    pc = ops.PackedConv(w, torch.zeros(256, device="cuda"), m)
    outs = []
    for xb in (x, x[2:3].contiguous(), x):
        hi, lo, g = ops.build_taps(xb, m, L.TAPS_REFLECT1)
        y, st = ops.conv_gemm(hi, lo, g, pc, "3x3", xb.shape[0], 32, 32, m, m.act_scale)
        outs.append((y, ops.instnorm_reduce(st, xb.shape[0], 1024, 256)))
    torch.cuda.synchronize()
    assert torch.equal(outs[0][0], outs[2][0]) and torch.equal(outs[0][1], outs[2][1])
    assert torch.equal(outs[0][0][2:3], outs[1][0]) and torch.equal(outs[0][1][2:3], outs[1][1])


@pytest.mark.parametrize("mode", ["fp16x3", "bf16x3"])
def test_build_taps_modes(mode):
    from wacv23_tsnet_b200 import lib as L, ops
    m = ops.MathMode(mode)
    tol = 2e-6 if m.fmt == 0 else 1e-4
    torch.manual_seed(2)
    x = torch.randn(2, 16, 16, 64, device="cuda") * 2
    xn = x.permute(0, 3, 1, 2)
    hi, lo, _ = ops.build_taps(x, m, L.TAPS_REFLECT1)
    assert _relerr(_recon(hi, lo, m.fmt), F.pad(xn, (1, 1, 1, 1), mode="reflect").permute(0, 2, 3, 1) * m.act_scale) < tol
    hi, lo, _ = ops.build_taps(x, m, L.TAPS_UP2REFLECT1)          # quad kernel: 3x3 neighbourhood -> 2x2 outputs
    up = F.pad(F.interpolate(xn, scale_factor=2, mode="bilinear", align_corners=False), (1, 1, 1, 1), mode="reflect")
    assert _relerr(_recon(hi, lo, m.fmt), up.permute(0, 2, 3, 1) * m.act_scale) < tol
    mr0 = torch.stack([x.mean((1, 2)), 1.0 / torch.sqrt(x.var((1, 2), unbiased=False) + 1e-5)], -1).contiguous()
    hq, lq, _ = ops.build_taps(x, m, L.TAPS_UP2REFLECT1, mean_rstd=mr0, relu=True)
    # one destination pixel per thread (tsnet_taps_desc.flags = TSNET_TAPS_GENERIC_UP2)
    hg, lg, _ = ops.build_taps(x, m, L.TAPS_UP2REFLECT1, mean_rstd=mr0, relu=True, flags=L.TAPS_GENERIC_UP2)
    hg0, lg0, _ = ops.build_taps(x, m, L.TAPS_UP2REFLECT1, flags=L.TAPS_GENERIC_UP2)
    torch.cuda.synchronize()
    assert torch.equal(hq, hg) and torch.equal(lq, lg)             # same expression per output value
    assert _relerr(_recon(hg0, lg0, m.fmt), _recon(hi, lo, m.fmt)) == 0.0
    hi, lo, _ = ops.build_taps(x, m, L.TAPS_S2ZERO)
    xp = F.pad(xn, (1, 1, 1, 1)).permute(0, 2, 3, 1)
    got = _recon(hi, lo, m.fmt).view(2, 4, 9, 9, 64)
    for k, (py, px) in enumerate([(0, 0), (0, 1), (1, 0), (1, 1)]):
        assert _relerr(got[:, k], xp[:, py::2, px::2] * m.act_scale) < tol
    # InstanceNorm + ReLU + residual (ResnetBlock tail) and channel-offset concat
    mr = torch.stack([x.mean((1, 2)), 1.0 / torch.sqrt(x.var((1, 2), unbiased=False) + 1e-5)], -1).contiguous()
    res = torch.randn_like(x)
    act = torch.zeros(2, 16, 16, 128, device="cuda")
    cat_hi = torch.zeros(2, 18, 18, 128, dtype=torch.int16, device="cuda")
    cat_lo = torch.zeros_like(cat_hi)
    ops.build_taps(x, m, L.TAPS_REFLECT1, mean_rstd=mr, relu=True, residual=res, act_out=act, act_c_off=64,
                   taps=(cat_hi, cat_lo), c_off=64)
    ref = F.relu(F.instance_norm(xn, eps=1e-5)).permute(0, 2, 3, 1) + res
    assert _relerr(act[..., 64:], ref) < 1e-6 and float(act[..., :64].abs().max()) == 0.0
    refp = F.pad(ref.permute(0, 3, 1, 2), (1, 1, 1, 1), mode="reflect").permute(0, 2, 3, 1) * m.act_scale
    assert _relerr(_recon(cat_hi, cat_lo, m.fmt)[..., 64:], refp) < tol
    assert int(cat_hi[..., :64].abs().max()) == 0
    # two-part residual: x + torch.cat([r0, r1 broadcast over the batch], C) without materialising the concatenation
    r0 = torch.randn(2, 16, 16, 40, device="cuda")
    r1 = torch.randn(1, 16, 16, 24, device="cuda")
    a2 = torch.empty(2, 16, 16, 64, device="cuda")
    ops.build_taps(x, m, L.TAPS_SAME, mean_rstd=mr, residual=(r0, r1), act_out=a2, want_taps=False)
    ref2 = F.instance_norm(xn, eps=1e-5).permute(0, 2, 3, 1) + torch.cat([r0, r1.expand(2, -1, -1, -1)], -1)
    assert _relerr(a2, ref2) < 1e-6
    # mean over sources (torch.stack(...).mean(1), model/TSNet.py:400)
    x3 = torch.randn(6, 8, 8, 64, device="cuda")
    a3 = torch.empty(2, 8, 8, 64, device="cuda")
    ops.build_taps(x3, m, L.TAPS_SAME, avg_n=3, act_out=a3, want_taps=False)
    assert _relerr(a3, torch.stack([x3[0:2], x3[2:4], x3[4:6]], 1).mean(1)) < 2e-7


@pytest.mark.parametrize("label_nc", [2, 25])
def test_stem_taps_and_stem_conv(label_nc):
    """cat[img/255, lbl] + CoordConv + ReflectionPad2d(3) + 7x7 conv (model/TSNet.py:66, 107-125, 312)."""
    from oracle import tsnet_oracle as O
    from wacv23_tsnet_b200 import ops
    m = ops.MathMode("fp16x3")
    torch.manual_seed(3)
    img = torch.rand(2, 3, 128, 128, device="cuda") * 255 - 100
    lbl = (torch.rand(2, label_nc, 128, 128, device="cuda") > 0.7).float()
    w = torch.randn(64, 3 + label_nc + 3, 7, 7, device="cuda") * 0.02
    b = torch.randn(64, device="cuda") * 0.1
    pc = ops.PackedConv(w, b, m, fold_kw=True)
    hi, lo, g = ops.stem_taps(img, 255.0, lbl, pc.Cp, m)
    full = O.coord_channels(torch.cat([img.cpu() / 255.0, lbl.cpu()], 1)).cuda()
    cin = full.shape[1]
    fp = F.pad(full, (3, 3, 3, 3), mode="reflect")
    got = _recon(hi, lo, m.fmt)
    for s in range(7):
        assert _relerr(got[..., s * cin:(s + 1) * cin], fp[:, :, :, s:s + 128].permute(0, 2, 3, 1) * m.act_scale) < 2e-6
    assert float(got[..., 7 * cin:].abs().max()) == 0.0
    y, st = ops.conv_gemm(hi, lo, g, pc, "7x1", 2, 128, 128, m, m.act_scale)   # vertical-reuse stem kernel
    ref = F.conv2d(fp.double(), w.double(), b.double()).permute(0, 2, 3, 1).float()
    assert _relerr(y, ref) < CONV_TOL["fp16x3"]
    from wacv23_tsnet_b200 import lib as L
    y2, _ = ops.conv_gemm(hi, lo, g, pc, "7x1", 2, 128, 128, m, m.act_scale, flags=L.CONV_NO_VR)  # plain implicit GEMM
    assert torch.equal(y, y2)   # same accumulation order per element -> bit-identical
    # InstanceNorm statistics from its partials (different 32-pixel grouping than the plain kernel: 2 rows x 16 px)
    mr = ops.instnorm_reduce(st, 2, 128 * 128, 64)
    y64 = y.double().view(2, -1, 64)
    assert float((mr[..., 0] - y64.mean(1).float()).abs().max()) < 1e-5
    assert _relerr(mr[..., 1], (1.0 / torch.sqrt(y64.var(1, unbiased=False) + 1e-5)).float()) < 1e-5


def test_l2norm_and_head():
    from wacv23_tsnet_b200 import ops
    m = ops.MathMode("fp16x3")
    torch.manual_seed(4)
    f = torch.randn(2, 64, 512, device="cuda")
    f[0, 3] = 0  # all-zero vector: F.normalize gives 0 (eps clamp)
    hi, lo = ops.l2norm_split(f, m)
    assert _relerr(_recon(hi, lo, m.fmt).view(2, 64, 512), F.normalize(f, dim=2) * m.corr_scale) < 1e-6
    perm = torch.stack([torch.randperm(64) for _ in range(2)]).to(torch.int16).cuda()   # row p of sample b -> rank
    hi2, lo2 = ops.l2norm_split(f, m, rank=perm.data_ptr())
    torch.cuda.synchronize()
    idx = perm.long().view(2, 64, 1).expand(2, 64, 512)
    assert torch.equal(hi2.view(2, 64, 512).gather(1, idx), hi.view(2, 64, 512))
    assert torch.equal(lo2.view(2, 64, 512).gather(1, idx), lo.view(2, 64, 512))
    a = torch.randn(2, 64, 64, 64, device="cuda")
    wh = torch.randn(3, 64, 7, 7, device="cuda") * 0.02
    bh = torch.randn(3, device="cuda") * 0.1
    ref = torch.tanh(F.conv2d(F.pad(a.permute(0, 3, 1, 2), (3, 3, 3, 3), mode="reflect"), wh, bh))
    assert float((ops.head_conv_tanh(a, wh, bh) - ref).abs().max()) < 1e-5
    a = torch.randn(1, 256, 256, 64, device="cuda")
    fill = (-0.4, -0.44, -0.438)
    out = ops.head_conv_tanh(a, wh, bh, fore=(64, 192), fill=fill)
    ref = torch.tanh(F.conv2d(F.pad(a.permute(0, 3, 1, 2), (3, 3, 3, 3), mode="reflect"), wh, bh))
    assert float((out[..., 64:192] - ref[..., 64:192]).abs().max()) < 1e-5
    assert torch.equal(out[..., :64], torch.tensor(fill, device="cuda").view(1, 3, 1, 1).expand(1, 3, 256, 64))


def _corr_inputs(B, n, kind, seed):
    g = torch.Generator().manual_seed(seed)
    tar = torch.relu(torch.randn(B, 512, 32, 32, generator=g))
    tar[0, :, 3, 5] = 0  # all-zero target vector -> zero logits -> uniform softmax row
    srcs = [torch.randn(B, 512, 32, 32, generator=g) * 3 for _ in range(n)]
    if kind == "rect_u8":
        tb = torch.zeros(B, 1, 256, 256, dtype=torch.uint8)
        tb[:, :, 40:200, 30:220] = 1
        sbs = []
        for i in range(n):
            sb = torch.zeros(B, 1, 256, 256, dtype=torch.uint8)
            sb[:, :, 20 + 10 * i:180, 50:230 - 10 * i] = 1
            sbs.append(sb)
    elif kind == "mixed_u8":  # sample 0: rectangles, sample 1: all ones / all zeros, sample 2+: random bits
        g2 = torch.Generator().manual_seed(seed + 100)
        tb = torch.randint(0, 2, (B, 1, 256, 256), generator=g2).to(torch.uint8)
        sbs = [torch.randint(0, 2, (B, 1, 256, 256), generator=g2).to(torch.uint8) for _ in range(n)]
        tb[0] = 0
        tb[0, :, 64:200, 16:140] = 1
        for i, sb in enumerate(sbs):
            sb[0] = 0
            sb[0, :, 30 + 20 * i:220, 40:250 - 30 * i] = 1
        if B > 1:
            tb[1] = 1
            for i, sb in enumerate(sbs):
                sb[1] = i % 2
    elif kind == "random_f32":
        tb = torch.randint(0, 2, (B, 1, 256, 256), generator=g).float()
        sbs = [torch.randint(0, 2, (B, 1, 256, 256), generator=g).float() for _ in range(n)]
    elif kind == "all_one":
        tb = torch.ones(B, 1, 256, 256)
        sbs = [torch.ones(B, 1, 256, 256) for _ in range(n)]
    elif kind == "all_zero_tar":  # nothing matches: every logit 0 -> warp grid = mean coordinate = (0, 0)
        tb = torch.zeros(B, 1, 256, 256)
        sbs = [torch.ones(B, 1, 256, 256) for _ in range(n)]
    else:  # soft (non-binary) float masks: general weight mt*ms + (1-mt)(1-ms)
        tb = torch.rand(B, 1, 256, 256, generator=g)
        sbs = [torch.rand(B, 1, 256, 256, generator=g) for _ in range(n)]
    return tar, srcs, tb, sbs


def _run_corr(tar, srcs, tb, sbs, m, sort=True, want_mean=True, one_cta=False, normalized=False):
    from wacv23_tsnet_b200 import ops
    B, n = tar.shape[0], len(srcs)
    tar_d = tar.permute(0, 2, 3, 1).contiguous().cuda().view(B, 1024, 512)
    src_d = torch.stack([s.permute(0, 2, 3, 1).contiguous() for s in srcs]).cuda().view(n, B, 1024, 512)
    coord = torch.cat([torch.linspace(-1, 1, 32), torch.linspace(-1, 1, 32)]).cuda()
    out, grids = ops.corr_chain(tar_d, src_d, tb.squeeze(1).contiguous().cuda(),
                                [s.squeeze(1).contiguous().cuda() for s in sbs], coord, m, want_grids=True,
                                want_mean=want_mean, sort=sort, one_cta=one_cta, normalized=normalized)
    torch.cuda.synchronize()
    return out, grids


@pytest.mark.parametrize("B,n,kind", [(1, 1, "random_f32"), (2, 3, "rect_u8"), (1, 5, "random_f32"), (1, 8, "rect_u8"),
                                      (1, 2, "all_one"), (1, 2, "all_zero_tar"), (1, 2, "soft"), (3, 2, "mixed_u8")])
def test_corr_warp_vs_oracle(B, n, kind):
    from oracle import tsnet_oracle as O
    from wacv23_tsnet_b200 import ops
    m = ops.MathMode("fp16x3")
    tar, srcs, tb, sbs = _corr_inputs(B, n, kind, seed=10 + n)
    ref_mean, ref_grids = O.corr_warp(tar, srcs, tb, sbs)                        # the reference's fp32 arithmetic
    tru_mean, tru_grids = O.corr_warp(tar.double(), [s.double() for s in srcs], tb, sbs)  # fp64 "truth"
    out, grids = _run_corr(tar, srcs, tb, sbs, m)
    gerr = max(float((grids[i].cpu() - ref_grids[i]).abs().max()) for i in range(n))
    assert gerr < 5e-5, gerr          # vs the fp32 reference, [-1, 1] units; 5e-5 = 8e-4 feature pixels
    assert _relerr(out.view(B, 32, 32, 512).permute(0, 3, 1, 2).cpu(), ref_mean) < 1e-3
    # against fp64 truth: the reference's own fp32 evaluation is 3.5e-6 .. 6e-6 off (softmax(100 x) amplifies every
    # rounding); the kernel carries in addition the systematic truncation of tcgen05's accumulate
    # and must stay within 4x the reference's error, i.e. inside the end-to-end fp32 noise floor of 1.4e-5.
    k_err = max(float((grids[i].cpu().double() - tru_grids[i]).abs().max()) for i in range(n))
    r_err = max(float((ref_grids[i].double() - tru_grids[i]).abs().max()) for i in range(n))
    assert k_err < max(4.0 * r_err, 1.5e-5), (k_err, r_err)
    if kind == "all_zero_tar":
        assert float(torch.stack([g for g in grids]).abs().max()) < 1e-5
    # the class-sorted order with tile skipping and the raster order (every tile computed) are the same function
    out_r, grids_r = _run_corr(tar, srcs, tb, sbs, m, sort=False)
    assert float((grids - grids_r).abs().max()) < 5e-6   # summation order only
    assert _relerr(out, out_r) < 2e-4   # 3e-6 grid units x the feature gradient
    # uint8 and float masks with the same {0,1} content must give identical bits (integer-exact mask path)
    if kind in ("rect_u8", "mixed_u8"):
        out2, grids2 = _run_corr(tar, srcs, tb.float(), [s.float() for s in sbs], m)
        assert torch.equal(out, out2) and torch.equal(grids, grids2)


def test_corr_two_cta_and_one_cta_tile_kernels_agree():
    """The default tile kernel is the 2-CTA one (tcgen05.mma.cta_group::2: a CTA pair computes 256 target rows, each CTA
    loads half of the source chunk); tsnet_corr_desc.one_cta selects the 1-CTA kernel.  Same math per element, different
    work list granularity (tile pairs) -> identical up to the summation order of the state merge."""
    from wacv23_tsnet_b200 import ops
    m = ops.MathMode("fp16x3")
    for kind, B, n in (("rect_u8", 2, 3), ("mixed_u8", 3, 2), ("soft", 1, 2)):
        tar, srcs, tb, sbs = _corr_inputs(B, n, kind, seed=31)
        out2, grids2 = _run_corr(tar, srcs, tb, sbs, m)
        out1, grids1 = _run_corr(tar, srcs, tb, sbs, m, one_cta=True)
        # operands normalised up front (tsnet_l2norm_split) vs the default: un-normalised operands + reciprocal norms
        # applied to the similarity in the softmax FMA -- the same function up to fp32 rounding
        outn, gridsn = _run_corr(tar, srcs, tb, sbs, m, normalized=True)
        assert float((grids2 - gridsn).abs().max()) < 1e-5, kind
        assert _relerr(out2, outn) < 3e-4, kind
        assert float((grids2 - grids1).abs().max()) < 5e-6, kind
        assert _relerr(out2, out1) < 2e-4, kind


def test_corr_strongly_correlated_features():
    """Peaked softmax rows (cos ~ 0.95 at the true match, logits near 100) -- the regime of a trained network; the
    random-feature cases above have row maxima near 20."""
    from oracle import tsnet_oracle as O
    from wacv23_tsnet_b200 import ops
    m = ops.MathMode("fp16x3")
    g = torch.Generator().manual_seed(5)
    B, n = 2, 3
    tar = torch.randn(B, 512, 32, 32, generator=g)
    srcs = [torch.roll(tar, shifts=(2 * i + 1, -(i + 2)), dims=(2, 3)) + 0.3 * torch.randn(B, 512, 32, 32, generator=g)
            for i in range(n)]
    _, _, tb, sbs = _corr_inputs(B, n, "rect_u8", seed=3)
    _, ref = O.corr_warp(tar, srcs, tb, sbs)
    _, tru = O.corr_warp(tar.double(), [s.double() for s in srcs], tb, sbs)
    _, grids = _run_corr(tar, srcs, tb, sbs, m, want_mean=False)
    k_err = max(float((grids[i].cpu().double() - tru[i]).abs().max()) for i in range(n))
    r_err = max(float((ref[i].double() - tru[i]).abs().max()) for i in range(n))
    assert k_err < max(4.0 * r_err, 1.5e-5), (k_err, r_err)


def test_corr_chain_other_geometry():
    """16 x 16 feature maps with 256 channels (hw = 256: one column chunk, two row tiles; C = 256: four K blocks):
    the correlation chain is not tied to the 32 x 32 x 512 geometry of the reference network."""
    from oracle import tsnet_oracle as O
    from wacv23_tsnet_b200 import ops
    m = ops.MathMode("fp16x3")
    g = torch.Generator().manual_seed(21)
    B, n, Cc, hh = 3, 2, 256, 16
    tar = torch.relu(torch.randn(B, Cc, hh, hh, generator=g))
    srcs = [torch.randn(B, Cc, hh, hh, generator=g) * 2 for _ in range(n)]
    tb = torch.zeros(B, 1, 128, 128, dtype=torch.uint8)
    tb[:, :, 20:100, 10:90] = 1
    tb[2] = 1
    sbs = [torch.randint(0, 2, (B, 1, 128, 128), generator=g).to(torch.uint8) for _ in range(n)]
    sbs[1][2] = 0                                   # sample 2, source 1: every pair mismatched -> closed form only
    ref_mean, ref_grids = O.corr_warp(tar, srcs, tb, sbs)
    tar_d = tar.permute(0, 2, 3, 1).contiguous().cuda().view(B, hh * hh, Cc)
    src_d = torch.stack([s.permute(0, 2, 3, 1).contiguous() for s in srcs]).cuda().view(n, B, hh * hh, Cc)
    coord = torch.cat([torch.linspace(-1, 1, hh), torch.linspace(-1, 1, hh)]).cuda()
    out, grids = ops.corr_chain(tar_d, src_d, tb.squeeze(1).contiguous().cuda(),
                                [s.squeeze(1).contiguous().cuda() for s in sbs], coord, m, want_grids=True, want_mean=True)
    torch.cuda.synchronize()
    gerr = max(float((grids[i].cpu() - ref_grids[i]).abs().max()) for i in range(n))
    assert gerr < 5e-5, gerr
    assert _relerr(out.view(B, hh, hh, Cc).permute(0, 3, 1, 2).cpu(), ref_mean) < 1e-3
    assert float(grids[1, 2].abs().max()) < 1e-5    # all logits 0 -> uniform softmax -> mean coordinate (0, 0)


def test_corr_prepare_plan_is_integer_exact():
    """tsnet_corr_prepare: nearest down-sampling, stable class sort (ones, soft, zeros), rank tables and tile classes
    against a numpy restatement (bit-exact integer work)."""
    import numpy as np
    from oracle import tsnet_oracle as O
    from wacv23_tsnet_b200 import ops
    m = ops.MathMode("fp16x3")
    B, n = 3, 2
    _, _, tb, sbs = _corr_inputs(B, n, "mixed_u8", seed=1)
    coord = torch.cat([torch.linspace(-1, 1, 32), torch.linspace(-1, 1, 32)]).cuda()
    plan = ops.corr_prepare(tb.squeeze(1).contiguous().cuda(), [s.squeeze(1).contiguous().cuda() for s in sbs], coord,
                            B, 512, 32, 32, m)
    torch.cuda.synchronize()
    NM = (n + 1) * B
    rank = plan.ws[:NM * 1024 * 2].view(torch.int16).view(NM, 1024).cpu().numpy().astype(np.int64)
    masks = [O.nearest_downsample_mask(tb.numpy(), 32, 32)] + [O.nearest_downsample_mask(s.numpy(), 32, 32) for s in sbs]
    for q in range(n + 1):
        for b in range(B):
            mv = masks[q][b].reshape(-1)
            key = np.where(mv == 1, 0, np.where(mv == 0, 2, 1))
            order = np.argsort(key, kind="stable")
            want = np.empty(1024, np.int64)
            want[order] = np.arange(1024)
            assert (rank[(0 if q == 0 else B + (q - 1) * B) + b] == want).all()


def test_argument_errors_are_reported_not_crashed():
    from wacv23_tsnet_b200 import lib as L, ops
    m = ops.MathMode("fp16x3")
    x = torch.randn(1, 10, 10, 64, device="cuda")  # 100 pixels: not a multiple of the 128-pixel tile
    w = torch.randn(64, 64, 1, 1, device="cuda")
    pc = ops.PackedConv(w, torch.zeros(64, device="cuda"), m)
    hi, lo, g = ops.build_taps(x, m, L.TAPS_SAME)
    with pytest.raises(L.TSNetLibraryError, match="multiple of 128"):
        ops.conv_gemm(hi, lo, g, pc, "1x1", 1, 10, 10, m, m.act_scale)


def test_stem_taps_classmap_equals_onehot():
    """uint8 class-index labels expanded inside the loader (vl2ch, utils/misc.py:50-67) == feeding the one-hot planes."""
    from wacv23_tsnet_b200 import ops
    m = ops.MathMode("fp16x3")
    torch.manual_seed(6)
    for L_ in (2, 25):
        cls = torch.randint(0, L_, (2, 128, 128), device="cuda", dtype=torch.uint8)
        onehot = torch.stack([(cls == c).float() for c in range(L_)], 1).contiguous()
        img = torch.rand(2, 3, 128, 128, device="cuda") * 255 - 100
        Cp = (7 * (3 + L_ + 3) + 63) // 64 * 64
        a = ops.stem_taps(img, 255.0, onehot, Cp, m)
        b = ops.stem_taps(img, 255.0, cls, Cp, m, label_nc=L_)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
        a = ops.stem_taps(None, 1.0, onehot, (7 * (L_ + 3) + 63) // 64 * 64, m)
        b = ops.stem_taps(None, 1.0, cls, (7 * (L_ + 3) + 63) // 64 * 64, m, label_nc=L_)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_postprocess_u8_matches_demo_arithmetic():
    """demo/demo_face.py:194-199 + sample_img (:96-105) restated with torch CPU fp32 ops (each individually rounded, as
    numpy / torch evaluate the reference's expressions), fed the SAME per-frame statistics: the bytes must be equal.
    The statistics themselves (tsnet_plane_stats, fp64 fixed order) are checked against tensor.mean / tensor.std."""
    from wacv23_tsnet_b200 import ops
    torch.manual_seed(7)
    rec = torch.tanh(torch.randn(3, 3, 256, 256, device="cuda"))
    ref_mean = torch.tensor([0.05, -0.02, 0.01], device="cuda")
    ref_std = torch.tensor([0.21, 0.19, 0.2], device="cuda")
    import numpy as np
    img_mean = (np.array((101.84807705937696, 112.10832843463207, 111.65973036298041), dtype=np.float32) / 255)
    stats = ops.plane_stats(rec, 9, 256 * 256)
    got = ops.postprocess_u8(rec, ref_mean, ref_std, img_mean, gen_stats=stats)
    torch.cuda.synchronize()
    rc = rec.cpu()
    assert _relerr(stats[:, 0].cpu(), rc.view(9, -1).double().mean(1).float()) < 1e-5
    assert _relerr(stats[:, 1].cpu(), rc.view(9, -1).double().std(1).float()) < 1e-6
    gm = stats[:, 0].cpu().view(3, 3, 1, 1)
    gs = stats[:, 1].cpu().view(3, 3, 1, 1)
    y = (rc - gm) / gs * ref_std.cpu().view(1, 3, 1, 1) + ref_mean.cpu().view(1, 3, 1, 1)   # demo_face.py:196-198
    y = y.permute(0, 2, 3, 1).numpy() + img_mean                                              # sample_img :99-100
    y[y < 0] = 0
    y[y > 1] = 1
    y *= 255
    ref = torch.from_numpy(np.ascontiguousarray(y[..., ::-1]).astype("uint8"))                # BGR -> RGB, astype
    assert torch.equal(got.cpu(), ref)


def test_conv_gemm_addend_broadcast_over_sources():
    """conv over cat[a_i, t] == conv_a(a_i) + conv_t(t): the source-independent half is computed once and added (fp32)
    in the GEMM epilogue with row index modulo (FuseNet, model/TSNet.py:196-198)."""
    from wacv23_tsnet_b200 import lib as L, ops
    m = ops.MathMode("fp16x3")
    torch.manual_seed(8)
    n, B, C = 3, 2, 128
    a = torch.randn(n * B, 32, 32, C, device="cuda")
    t = torch.randn(B, 32, 32, C, device="cuda")
    w = torch.randn(256, 2 * C, 3, 3, device="cuda") * 0.05
    b = torch.randn(256, device="cuda")
    cat = torch.cat([a, t.repeat(n, 1, 1, 1)], -1).permute(0, 3, 1, 2)
    ref = F.conv2d(F.pad(cat, (1, 1, 1, 1), mode="reflect").double(), w.double(), b.double()).permute(0, 2, 3, 1).float()
    pc_t = ops.PackedConv(w, None, m, cin_range=(C, 2 * C))
    pc_a = ops.PackedConv(w, b, m, cin_range=(0, C))
    th, tl, tg = ops.build_taps(t, m, L.TAPS_REFLECT1)
    y_t, _ = ops.conv_gemm(th, tl, tg, pc_t, "3x3", B, 32, 32, m, m.act_scale, want_stats=False)
    ah, al, ag = ops.build_taps(a, m, L.TAPS_REFLECT1)
    y, st = ops.conv_gemm(ah, al, ag, pc_a, "3x3", n * B, 32, 32, m, m.act_scale, addend=y_t)
    mr = ops.instnorm_reduce(st, n * B, 1024, 256)
    torch.cuda.synchronize()
    assert _relerr(y, ref) < CONV_TOL["fp16x3"]
    assert _relerr(mr[..., 0], ref.mean((1, 2))) < 2e-5   # statistics include the addend


@pytest.mark.parametrize("Cin,Cout,kind,tmode,relu,res", [
    (128, 256, "3x3", "reflect", True, False),     # ResnetBlock conv1: conv -> IN -> ReLU -> pad
    (256, 256, "3x3", "reflect", False, True),     # ResnetBlock conv2: conv -> IN -> + x -> pad (+ fp32 output)
    (128, 512, "3x3s2", "reflect", True, False),   # third down-sampling conv of the encoders
    (256, 1024, "3x3", "same", False, True),       # FuseNet conv2 -> operand of the 1x1 conv
])
def test_conv_gemm_fused_instance_norm_epilogue(Cin, Cout, kind, tmode, relu, res):
    """tsnet_conv_desc.fuse_in: 8-CTA clusters exchange the InstanceNorm statistics over DSMEM and write the next
    layer's operands directly; must equal conv -> F.instance_norm -> ReLU / + residual -> ReflectionPad2d."""
    from wacv23_tsnet_b200 import lib as L, ops
    m = ops.MathMode("fp16x3")
    torch.manual_seed(9)
    B = 5
    Hin = 64 if kind == "3x3s2" else 32
    x = torch.randn(B, Hin, Hin, Cin, device="cuda")
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05
    b = torch.randn(Cout, device="cuda")
    xn = x.permute(0, 3, 1, 2).double()
    if kind == "3x3s2":
        conv = F.conv2d(xn, w.double(), b.double(), stride=2, padding=1)
        tin = ops.build_taps(x, m, L.TAPS_S2ZERO)
    else:
        conv = F.conv2d(F.pad(xn, (1, 1, 1, 1), mode="reflect"), w.double(), b.double())
        tin = ops.build_taps(x, m, L.TAPS_REFLECT1)
    ref = F.instance_norm(conv, eps=1e-5)
    if relu:
        ref = F.relu(ref)
    resid = torch.randn(B, 32, 32, Cout, device="cuda") if res else None
    if res:
        ref = ref + resid.permute(0, 3, 1, 2).double()
    act = torch.zeros(B, 32, 32, Cout + 64, device="cuda")
    tm = L.TAPS_REFLECT1 if tmode == "reflect" else L.TAPS_SAME
    _, Hd, Wd = ops.taps_geometry(tm, 32, 32)
    th = torch.zeros(B, Hd, Wd, Cout + 64, dtype=torch.int16, device="cuda")
    tl = torch.zeros_like(th)
    pc = ops.PackedConv(w, b, m)
    outs = []
    for _ in range(2):
        ops.conv_gemm(tin[0], tin[1], tin[2], pc, kind, B, 32, 32, m, m.act_scale,
                      fuse=dict(relu=relu, tmode=tm, residual=resid, act_out=act, act_c_off=64, taps=(th, tl), c_off=64))
        torch.cuda.synchronize()
        outs.append((act.clone(), th.clone(), tl.clone()))
    assert all(torch.equal(a, b_) for a, b_ in zip(outs[0], outs[1]))            # deterministic
    ref_nhwc = ref.permute(0, 2, 3, 1).float()
    assert _relerr(act[..., 64:], ref_nhwc) < 5e-6 and float(act[..., :64].abs().max()) == 0.0
    refp = ref if tmode == "same" else F.pad(ref, (1, 1, 1, 1), mode="reflect")
    got = _recon(th, tl, m.fmt)
    assert _relerr(got[..., 64:], refp.permute(0, 2, 3, 1).float() * m.act_scale) < 5e-6
    assert int(th[..., :64].abs().max()) == 0


@pytest.mark.parametrize("B,n", [(2, 3), (1, 1), (1, 8)])
def test_warp_mean_taps_vs_grid_sample(B, n):
    """K2: F.grid_sample (bilinear, zeros, align_corners=False) of every source at given grids + mean over sources,
    as fp32 and as the hi/lo operand window of the consumer (model/TSNet.py:366, :392, :163)."""
    from wacv23_tsnet_b200 import ops
    m = ops.MathMode("fp16x3")
    g = torch.Generator().manual_seed(40 + n)
    srcs = [torch.randn(B, 512, 32, 32, generator=g) * 3 for _ in range(n)]
    grids = [torch.rand(B, 32, 32, 2, generator=g) * 2.2 - 1.1 for _ in range(n)]   # includes out-of-range samples
    ref = torch.stack([F.grid_sample(s, gr, align_corners=False) for s, gr in zip(srcs, grids)], 1).mean(1)
    src_d = [s.permute(0, 2, 3, 1).contiguous().view(B, 1024, 512).cuda() for s in srcs]
    grid_d = torch.stack(grids).cuda().contiguous()
    hi = torch.zeros(B, 32, 32, 1024, dtype=torch.int16, device="cuda")
    lo = torch.zeros_like(hi)
    out = ops.warp_mean_taps(src_d, grid_d, B, 32, 32, 512, m, taps=(hi, lo), c_off=512, want_mean=True)
    torch.cuda.synchronize()
    ref_nhwc = ref.permute(0, 2, 3, 1).cuda()
    assert _relerr(out.view(B, 32, 32, 512), ref_nhwc) < 2e-6
    assert _relerr(_recon(hi, lo, m.fmt)[..., 512:], ref_nhwc * m.act_scale) < 2e-6
    assert int(hi[..., :512].abs().max()) == 0


# ---------------------------------------------------------------------------------------------------------------------
# Winograd F(2x2, 3x3) path of the ResnetBlock convolutions (model/TSNet.py:10-49)
# ---------------------------------------------------------------------------------------------------------------------
# tolerance of the whole Winograd convolution against an fp64 reflect-pad conv, relative to max|ref|: the input /
# output transforms are fp32 additions (3e-7 class), operands carry 22 bits, accumulation is promoted every 2 K blocks.
WINO_TOL = 3e-6


def _wino_case(B, Cin, Cout, seed=0, flags=0, with_addend=False):
    from wacv23_tsnet_b200 import lib as L, ops
    torch.manual_seed(seed)
    m = ops.MathMode("fp16x3")
    x = torch.randn(B, 32, 32, Cin, device="cuda")
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05
    b = torch.randn(Cout, device="cuda")
    ref = F.conv2d(F.pad(x.permute(0, 3, 1, 2), (1, 1, 1, 1), mode="reflect").double(), w.double(), b.double())
    ref = ref.permute(0, 2, 3, 1).contiguous()
    addend = None
    if with_addend:
        addend = torch.randn(1, 32, 32, Cout, device="cuda")
        ref = ref + addend.double()
    pw = ops.PackedWino(w, b, m)
    taps = ops.build_taps(x, m, L.TAPS_WINO)
    y, stats = ops.wino_conv(taps, pw, B, 32, 32, m, m.act_scale, addend=addend, flags=flags)
    mr = ops.instnorm_reduce(stats, B, 1024, Cout)
    torch.cuda.synchronize()
    return y, mr, ref.float(), (x, w, b, taps, pw)


@pytest.mark.parametrize("B,Cin,Cout", [
    (3, 512, 512),       # the dominant layer: ResnetBlock(512) of img_enc / the decoder
    (1, 1024, 1024),     # FuseNet ResnetBlock(1024), second conv
    (2, 512, 1024),      # FuseNet first conv, one half of the channel concatenation
    (1, 64, 256),        # one K block per plane
    (37, 128, 256),      # 74 pixel tiles per plane: more items than CTA pairs, persistent loop across planes
])
def test_wino_conv_vs_fp64(B, Cin, Cout):
    y, mr, ref, _ = _wino_case(B, Cin, Cout)
    assert _relerr(y, ref) < WINO_TOL
    assert _relerr(mr[..., 0], ref.mean((1, 2))) < 2e-5
    assert _relerr(mr[..., 1], 1.0 / torch.sqrt(ref.var((1, 2), unbiased=False) + 1e-5)) < 5e-5


def test_wino_conv_addend_one_cta_and_batch_invariance():
    """(a) fp32 addend per output pixel (FuseNet's target half) enters before the statistics; (b) the 1-CTA kernel
    (TSNET_CONV_ONE_CTA) and the default 2-CTA pair kernel are bit-identical; (c) a sample gives the same bits
    whatever batch it rides in (what makes batch sharding exact)."""
    from wacv23_tsnet_b200 import lib as L, ops
    y, mr, ref, (x, w, b, taps, pw) = _wino_case(3, 256, 256, seed=5, with_addend=True)
    assert _relerr(y, ref) < WINO_TOL
    assert _relerr(mr[..., 0], ref.mean((1, 2))) < 2e-5
    m = ops.MathMode("fp16x3")
    y2, s2 = ops.wino_conv(taps, pw, 3, 32, 32, m, m.act_scale)
    y1, s1 = ops.wino_conv(taps, pw, 3, 32, 32, m, m.act_scale, flags=L.CONV_ONE_CTA)
    t1 = ops.build_taps(x[1:2].contiguous(), m, L.TAPS_WINO)
    ya, sa = ops.wino_conv(t1, pw, 1, 32, 32, m, m.act_scale)
    torch.cuda.synchronize()
    assert torch.equal(y1, y2) and torch.equal(s1, s2)
    assert torch.equal(ya, y2[1:2]) and torch.equal(sa, s2[32:64])


def test_wino_passes_equal_host_emulation():
    """The CUDA kernels of the transform passes and the host emulation the CPU suite runs are the same source
    (csrc/wino_passes.cuh): operands and outputs must agree bit for bit (the statistics up to FMA contraction)."""
    import ctypes as C
    import subprocess
    from wacv23_tsnet_b200 import lib as L, ops
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "oracle")], check=True, capture_output=True)
    emul = C.CDLL(os.path.join(root, "oracle", "_build", "libwino_emul.so"))
    vp = C.c_void_p
    emul.wino_emul_input.argtypes = [vp, vp, vp, vp, vp, vp] + [C.c_int] * 10 + [C.c_float, C.c_int]
    emul.wino_emul_output.argtypes = [vp, vp, vp, C.c_longlong, vp, vp] + [C.c_int] * 5
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    m = ops.MathMode("fp16x3")
    torch.manual_seed(6)
    B, Cc = 2, 64
    x = torch.randn(B, 32, 32, Cc) * 3
    res = torch.randn(B, 32, 32, Cc)
    mr = torch.stack([x.mean((1, 2)), 1.0 / torch.sqrt(x.var((1, 2), unbiased=False) + 1e-5)], -1).contiguous()
    hi_c = torch.zeros(B, 16, 16, 16, Cc, dtype=torch.int16)
    lo_c = torch.zeros_like(hi_c)
    act_c = torch.zeros(B, 32, 32, Cc)
    emul.wino_emul_input(p(x), p(mr), p(res), p(act_c), p(hi_c), p(lo_c), B, 32, 32, Cc, 1, Cc, 0, 0, Cc, 0,
                         m.act_scale, 64)
    act_g = torch.zeros(B, 32, 32, Cc, device="cuda")
    x_g, mr_g, res_g = x.cuda(), mr.cuda(), res.cuda()
    hi_g, lo_g, _ = ops.build_taps(x_g, m, L.TAPS_WINO, mean_rstd=mr_g, relu=True, residual=res_g, act_out=act_g)
    torch.cuda.synchronize()
    assert torch.equal(hi_g.cpu().view_as(hi_c), hi_c) and torch.equal(lo_g.cpu().view_as(lo_c), lo_c)
    assert torch.equal(act_g.cpu(), act_c)
    mm = torch.randn(16, B * 256, Cc)
    bias = torch.randn(Cc)
    y_c = torch.zeros(B, 32, 32, Cc)
    st_c = torch.zeros(B * 32, Cc, 2)
    emul.wino_emul_output(p(mm), p(bias), None, 0, p(y_c), p(st_c), B, 32, 32, Cc, 64)
    y_g = torch.empty(B, 32, 32, Cc, device="cuda")
    st_g = torch.empty(B * 32, Cc, 2, device="cuda")
    mm_g, bias_g = mm.cuda(), bias.cuda()
    L.check(L.load().tsnet_wino_output(p(mm_g), B, 32, 32, Cc, p(bias_g), None, 0, p(y_g), p(st_g),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert torch.equal(y_g.cpu(), y_c)
    assert _relerr(st_g.cpu(), st_c) < 1e-6


@pytest.mark.parametrize("variant", [0, 1])   # 32-channel slabs (1 CTA / SM) and 16-channel slabs (2 CTAs / SM)
@pytest.mark.parametrize("relu,with_res,with_addend", [(True, False, False), (False, True, False), (True, False, True)])
def test_wino_bridge_vs_separate_passes(relu, with_res, with_addend, variant):
    """tsnet_wino_bridge (output transform + InstanceNorm + ReLU / residual + input transform in one pass, statistics
    CTA-local in fp64) against the separate passes tsnet_wino_output -> tsnet_instnorm_reduce -> tsnet_build_taps(WINO)."""
    from wacv23_tsnet_b200 import lib as L, ops
    m = ops.MathMode("fp16x3")
    torch.manual_seed(15)
    B, Cin, Cout = 3, 128, 256
    x = torch.randn(B, 32, 32, Cin, device="cuda")
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05
    b = torch.randn(Cout, device="cuda")
    res = torch.randn(B, 32, 32, Cout, device="cuda") if with_res else None
    addend = torch.randn(1, 32, 32, Cout, device="cuda") if with_addend else None
    pw = ops.PackedWino(w, b, m)
    taps = ops.build_taps(x, m, L.TAPS_WINO)
    mbuf = ops.wino_gemm(taps, pw, B, 32, 32, m, m.act_scale)
    # separate passes
    y, stats = ops.wino_output(mbuf, pw, B, 32, 32, addend=addend)
    mr = ops.instnorm_reduce(stats, B, 1024, Cout)
    act_s = torch.zeros(B, 32, 32, Cout, device="cuda")
    hs, ls, _ = ops.build_taps(y, m, L.TAPS_WINO, mean_rstd=mr, relu=relu, residual=res, act_out=act_s)
    # fused bridge
    act_f = torch.zeros(B, 32, 32, Cout + 64, device="cuda")
    mr_f = torch.zeros(B, Cout, 2, device="cuda")
    hf = torch.zeros(B * 16, 16, 16, Cout + 64, dtype=torch.int16, device="cuda")
    lf = torch.zeros_like(hf)
    ops.wino_bridge(mbuf, pw, B, 32, 32, m, relu=relu, addend=addend, residual=res, act_out=act_f, act_c_off=64,
                    taps=(hf, lf), c_off=64, mean_rstd_out=mr_f, variant=variant)
    torch.cuda.synchronize()
    assert _relerr(mr_f[..., 0], mr[..., 0]) < 1e-5 and _relerr(mr_f[..., 1], mr[..., 1]) < 1e-5
    assert _relerr(act_f[..., 64:], act_s) < 2e-6 and float(act_f[..., :64].abs().max()) == 0.0
    vl = ops.wino_v_logical   # operand planes are stored K-block-major
    assert _relerr(_recon(vl(hf), vl(lf), m.fmt)[..., 64:], _recon(vl(hs), vl(ls), m.fmt)) < 4e-6
    assert int(vl(hf)[..., :64].abs().max()) == 0
    # correlation operand emission (the last img_enc block): rows at a rank permutation + partial sums of squares must
    # equal what the stand-alone operand pass makes of the same activations
    rank = torch.stack([torch.randperm(1024) for _ in range(B)]).to(torch.int16).cuda()
    corr = dict(hi=torch.zeros(B * 1024, Cout, dtype=torch.int16, device="cuda"), rank=rank.data_ptr(),
                ssq=torch.zeros(B, Cout // (16 if variant else 32), 1024, device="cuda"))
    corr["lo"] = torch.zeros_like(corr["hi"])
    ops.wino_bridge(mbuf, pw, B, 32, 32, m, relu=relu, addend=addend, residual=res, corr=corr, variant=variant)
    rn_b = ops.corr_norms(corr["ssq"], B, 1024, corr["ssq"].shape[1], rank.data_ptr())
    oh, ol, rn_o = ops.corr_operands(act_f[..., 64:].contiguous().view(B, 1024, Cout), m, rank=rank.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(corr["hi"], oh) and torch.equal(corr["lo"], ol)
    assert _relerr(rn_b, rn_o) < 1e-6
    # deterministic, and a sample does not depend on the batch it rides in
    h2, l2, _ = ops.wino_bridge(mbuf, pw, B, 32, 32, m, relu=relu, addend=addend, residual=res, variant=variant)
    t1 = ops.build_taps(x[1:2].contiguous(), m, L.TAPS_WINO)
    m1 = ops.wino_gemm(t1, pw, 1, 32, 32, m, m.act_scale)
    h1, l1, _ = ops.wino_bridge(m1, pw, 1, 32, 32, m, relu=relu, addend=addend,
                                residual=None if res is None else res[1:2].contiguous(), variant=variant)
    torch.cuda.synchronize()
    assert torch.equal(h2[16:32], h1) and torch.equal(l2[16:32], l1)
    assert torch.equal(vl(h2), vl(hf)[..., 64:]) and torch.equal(vl(l2), vl(lf)[..., 64:])


def test_stem_conv_direct_input():
    """tsnet_stem_conv_fwd: the kw-folded operand tile is generated in shared memory from the raw NCHW inputs (no tap
    source in HBM).  img_enc (3 + 2 + 3 = 8 channels: the layout equals the materialised one) must be bit-identical to
    tsnet_stem_taps + the vertical-reuse kernel; lbl_enc (2 + 3 channels padded to 8) is checked against an fp64 conv;
    uint8 frames + class-index labels give the same bits as their fp32 expansions."""
    from oracle import tsnet_oracle as O
    from wacv23_tsnet_b200 import ops
    m = ops.MathMode("fp16x3")
    torch.manual_seed(21)
    B, H, W, Lc = 3, 64, 128, 2
    u8 = torch.randint(0, 256, (B, 3, H, W), device="cuda", dtype=torch.uint8)
    mean = (101.848, 112.108, 111.66)
    img = u8.float() - torch.tensor(mean, device="cuda").view(1, 3, 1, 1)
    cls = torch.randint(0, Lc, (B, H, W), device="cuda", dtype=torch.uint8)
    lbl = torch.stack([(cls == c).float() for c in range(Lc)], 1).contiguous()
    # ---- img_enc stem
    w = torch.randn(64, 3 + Lc + 3, 7, 7, device="cuda") * 0.02
    b = torch.randn(64, device="cuda") * 0.1
    pc_old = ops.PackedConv(w, b, m, fold_kw=True)
    hi, lo, g = ops.stem_taps(img, 255.0, lbl, pc_old.Cp, m)
    y_old, st_old = ops.conv_gemm(hi, lo, g, pc_old, "7x1", B, H, W, m, m.act_scale)
    pc = ops.PackedConv(w, b, m, fold_kw=True, fold_cin=ops.STEM_FOLD)
    y, st = ops.stem_conv(img, 255.0, lbl, pc, m)
    y8, st8 = ops.stem_conv(u8, 255.0, cls, pc, m, label_nc=Lc, img_mean=mean)
    torch.cuda.synchronize()
    assert torch.equal(y, y_old) and torch.equal(st, st_old)
    assert torch.equal(y8, y) and torch.equal(st8, st)
    full = O.coord_channels(torch.cat([img.cpu() / 255.0, lbl.cpu()], 1)).cuda()
    ref = F.conv2d(F.pad(full, (3, 3, 3, 3), mode="reflect").double(), w.double(), b.double()).permute(0, 2, 3, 1).float()
    assert _relerr(y, ref) < CONV_TOL["fp16x3"]
    # ---- lbl_enc stem (no image; 5 channels padded to the 8-channel folded tap)
    w2 = torch.randn(64, Lc + 3, 7, 7, device="cuda") * 0.02
    pc2 = ops.PackedConv(w2, b, m, fold_kw=True, fold_cin=ops.STEM_FOLD)
    y2, st2 = ops.stem_conv(None, 1.0, lbl, pc2, m)
    y2c, _ = ops.stem_conv(None, 1.0, cls, pc2, m, label_nc=Lc)
    torch.cuda.synchronize()
    full2 = O.coord_channels(lbl.cpu()).cuda()
    ref2 = F.conv2d(F.pad(full2, (3, 3, 3, 3), mode="reflect").double(), w2.double(), b.double()).permute(0, 2, 3, 1).float()
    assert _relerr(y2, ref2) < CONV_TOL["fp16x3"] and torch.equal(y2c, y2)
    mr = ops.instnorm_reduce(st2, B, H * W, 64)
    assert _relerr(mr[..., 0], ref2.mean((1, 2))) < 2e-5
