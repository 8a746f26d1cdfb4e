"""Thin tensor-level wrappers over the C ABI (include/tsnet_b200.h).

torch is used only for device memory and the current stream; every function enqueues the kernels of one C-ABI call
of libtsnet_sm100.so.  `LAUNCHES` counts the calls; `kernel_launches()` is the library's own count of CUDA kernel
launches (bench.py reports it as gpu_launches).
"""
import ctypes as C
import math

import torch

from . import lib as L

LAUNCHES = 0
PROFILE = None  # bench.py sets this to a dict: key -> [(start_event, end_event), ...] around every launch


class _Prof:
    """CUDA events on the launching stream around one kernel launch (only while PROFILE is a dict)."""

    def __init__(self, key):
        self.key = key

    def __enter__(self):
        self.e0 = None
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.e0 is not None and PROFILE is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.setdefault(self.key, []).append((self.e0, e1))
        return False


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def kernel_launches():
    return int(L.load().tsnet_launch_count())


def _count():
    global LAUNCHES
    LAUNCHES += 1


def _f32(t):
    assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous(), (t.dtype, t.device, t.is_contiguous())
    return t


class MathMode:
    """Operand format of the tensor-core path.

    fp16x3 (default): x = hi + lo in fp16 (22 significant bits while lo is normal), three MMAs per K step,
                      fp32 accumulate -- fp32-faithful (DESIGN.md "precision").
    bf16x3:           same with bf16 (16 significant bits), no range management needed.
    fp16 / bf16:      hi term only -- fast, NOT parity grade (reported separately, never the headline).
    Scales are powers of two so they are exact; they keep lo terms out of the fp16 subnormal range.
    """

    def __init__(self, name="fp16x3"):
        assert name in ("fp16x3", "bf16x3", "fp16", "bf16"), name
        self.name = name
        self.fmt = L.FMT_FP16 if name.startswith("fp16") else L.FMT_BF16
        self.split = 1 if name.endswith("x3") else 0
        self.act_scale = 16.0 if self.fmt == L.FMT_FP16 else 1.0
        self.corr_scale = 4096.0 if self.fmt == L.FMT_FP16 else 1.0

    def weight_scale(self, w):
        if self.fmt != L.FMT_FP16:
            return 1.0
        m = float(w.abs().max())
        if not math.isfinite(m) or m == 0.0:
            return 1.0
        return 2.0 ** math.floor(math.log2(8192.0 / m))  # max |w * s| in (4096, 8192]


class PackedConv:
    """A conv weight packed for tsnet_conv_gemm_fwd (K-major hi/lo, zero padded)."""

    def __init__(self, weight, bias, mode, fold_kw=False, Cp=None, block_n=None, cin_range=None, fold_cin=None):
        L.require_device()
        w = weight.detach()
        if cin_range is not None:  # a slice of the input channels (conv over a channel-concatenation, split by operand)
            w = w[:, cin_range[0]:cin_range[1]].contiguous()
        w = _f32(w)
        Cout, Cin, KH, KW = w.shape
        self.Cout, self.Cin, self.KH, self.KW, self.fold_kw = Cout, Cin, KH, KW, bool(fold_kw)
        # fold_cin: channels of every folded horizontal tap padded to fold_cin >= Cin (tsnet_stem_conv_fwd: 8)
        self.fold_cin = fold_cin if (fold_kw and fold_cin) else None
        assert self.fold_cin is None or self.fold_cin >= Cin
        need = KW * (self.fold_cin or Cin) if fold_kw else Cin
        self.Cp = Cp if Cp is not None else (need + 63) // 64 * 64
        assert self.Cp >= need and self.Cp % 64 == 0
        self.num_taps = KH if fold_kw else KH * KW
        # block_n = None: chosen per launch from the tile count (see conv_gemm); rows are padded so any width fits
        self.block_n = block_n
        pad_to = block_n if block_n is not None else (64 if Cout <= 64 else (128 if Cout <= 128 else 256))
        self.Cout_pad = (Cout + pad_to - 1) // pad_to * pad_to
        self.scale = mode.weight_scale(w)
        K = self.num_taps * self.Cp
        self.w_hi = torch.empty((self.Cout_pad, K), dtype=torch.int16, device=w.device)
        self.w_lo = torch.empty_like(self.w_hi)
        self.bias = None if bias is None else _f32(bias.detach())
        L.check(L.load().tsnet_pack_conv_weight(_ptr(w), Cout, Cin, KH, KW,
                                                (self.fold_cin or 1) if self.fold_kw else 0, self.Cp, self.Cout_pad,
                                                C.c_float(self.scale), mode.fmt, _ptr(self.w_hi), _ptr(self.w_lo),
                                                _stream()))
        _count()


class PackedWino:
    """A 3x3 conv weight in the Winograd F(2x2, 3x3) domain, packed for tsnet_wino_gemm_fwd: U = G g G^T
    (tsnet_wino_weight_transform, fp64 arithmetic) as 16 K-major hi/lo matrices [16 * Cout, Cp]."""

    def __init__(self, weight, bias, mode, cin_range=None):
        L.require_device()
        w = weight.detach()
        if cin_range is not None:
            w = w[:, cin_range[0]:cin_range[1]]
        w = _f32(w.contiguous())
        Cout, Cin, KH, KW = w.shape
        assert (KH, KW) == (3, 3) and Cout % 256 == 0 and Cin % 64 == 0, (tuple(w.shape),)
        self.Cout, self.Cin, self.Cp = Cout, Cin, Cin
        u = torch.empty((16, Cout, Cin), dtype=torch.float32, device=w.device)
        L.check(L.load().tsnet_wino_weight_transform(_ptr(w), Cout, Cin, _ptr(u), _stream()))
        _count()
        self.scale = mode.weight_scale(u)
        self.u_hi = torch.empty((16 * Cout, Cin), dtype=torch.int16, device=w.device)
        self.u_lo = torch.empty_like(self.u_hi)
        self.bias = None if bias is None else _f32(bias.detach())
        # U viewed as a 1x1 conv weight [16 * Cout, Cin, 1, 1]
        L.check(L.load().tsnet_pack_conv_weight(_ptr(u), 16 * Cout, Cin, 1, 1, 0, Cin, 16 * Cout,
                                                C.c_float(self.scale), mode.fmt, _ptr(self.u_hi), _ptr(self.u_lo),
                                                _stream()))
        _count()


def wino_v_logical(t):
    """Operand planes as stored by build_taps(TAPS_WINO) / wino_bridge -- K-block-major, [B*16][Cp/64][TH][TW][64], kept
    in a tensor of bookkeeping shape [B*16, TH, TW, Cp] -- re-ordered to the logical [B*16, TH, TW, Cp] (tests)."""
    B16, TH, TW, Cp = t.shape
    return t.reshape(B16, Cp // 64, TH, TW, 64).permute(0, 2, 3, 1, 4).reshape(B16, TH, TW, Cp)


def wino_ok(H, W, Cin, Cout):
    """Shapes the Winograd path covers: every ResnetBlock convolution of the network (32 x 32, 512 / 1024 channels)."""
    return (H % 2 == 0 and W % 2 == 0 and ((H // 2) * (W // 2)) % 128 == 0 and (W // 2) % 8 == 0 and
            (W // 2 <= 128 and 128 % (W // 2) == 0) and Cin % 64 == 0 and Cout % 256 == 0)


def wino_gemm(taps, pw, B, H, W, mode, act_scale, m_buf=None, flags=0, chunk_kb=0):
    """The 16 plane GEMMs of a Winograd 3x3 conv as one batched launch: taps = (hi, lo, geom) from
    build_taps(mode TAPS_WINO) or wino_bridge.  Returns M, fp32 [16, B * H/2 * W/2, Cout] (flat buffer)."""
    hi, lo, geom = taps
    TH, TW = H // 2, W // 2
    assert geom == (16, TH, TW) and hi.shape == (B * 16, TH, TW, pw.Cp), (geom, tuple(hi.shape))
    n_m = 16 * B * TH * TW * pw.Cout
    if m_buf is None or m_buf.numel() < n_m:
        m_buf = torch.empty(n_m, dtype=torch.float32, device=hi.device)
    d = L.WinoGemmDesc()
    d.B, d.TH, d.TW, d.C, d.Cout = B, TH, TW, pw.Cp, pw.Cout
    d.split, d.fmt, d.out_scale, d.chunk_kb, d.flags = mode.split, mode.fmt, 1.0 / (pw.scale * act_scale), chunk_kb, flags
    with _Prof(("wino_gemm", "3x3", B, H, W, pw.Cin, pw.Cout, 9)):
        L.check(L.load().tsnet_wino_gemm_fwd(C.byref(d), _ptr(hi), _ptr(lo), _ptr(pw.u_hi), _ptr(pw.u_lo), _ptr(m_buf),
                                             _stream()))
    _count()
    return m_buf


def wino_output(m_buf, pw, B, H, W, want_stats=True, addend=None):
    """Output transform + bias (+ addend) + InstanceNorm partial statistics: (y_raw [B, H, W, Cout], stats or None)."""
    dev = m_buf.device
    y = torch.empty((B, H, W, pw.Cout), dtype=torch.float32, device=dev)
    stats = torch.empty((B * H * W // 32, pw.Cout, 2), dtype=torch.float32, device=dev) if want_stats else None
    rows = 0
    if addend is not None:
        assert addend.shape[-1] == pw.Cout and addend.is_contiguous() and addend.dtype == torch.float32
        rows = addend.numel() // pw.Cout
    with _Prof(("wino_output",)):
        L.check(L.load().tsnet_wino_output(_ptr(m_buf), B, H, W, pw.Cout, _ptr(pw.bias), _ptr(addend), rows, _ptr(y),
                                           _ptr(stats), _stream()))
    _count()
    return y, stats


def wino_conv(taps, pw, B, H, W, mode, act_scale, want_stats=True, addend=None, m_buf=None, flags=0, chunk_kb=0):
    """3x3 reflect-pad conv via Winograd (GEMM + output pass).  Returns (y_raw [B, H, W, Cout], stats_partial or None)
    -- the outputs of conv_gemm."""
    m_buf = wino_gemm(taps, pw, B, H, W, mode, act_scale, m_buf=m_buf, flags=flags, chunk_kb=chunk_kb)
    return wino_output(m_buf, pw, B, H, W, want_stats=want_stats, addend=addend)


def wino_bridge(m_buf, pw, B, H, W, mode, relu=False, addend=None, residual=None, act_out=None, act_c_off=0, taps=None,
                c_off=0, mean_rstd_out=None, act_scale=None, corr=None, variant=0):
    """Fused layer boundary between two Winograd convolutions (tsnet_wino_bridge): output transform + bias (+ addend) ->
    InstanceNorm -> [ReLU] -> [+ residual] -> [fp32 act_out] -> input transform.  Returns (hi, lo, (16, H/2, W/2)).
    corr = dict(hi, lo [B*H*W, C] int16, rank = device pointer of the [B, H*W] rank table, ssq fp32 [B, C/32, H*W]): the
    activations are also written as the un-normalised operand rows of the correlation (corr_norms finishes the norms)."""
    Cc = pw.Cout
    dev = m_buf.device
    if taps is None:
        hi = torch.empty((B * 16, H // 2, W // 2, Cc), dtype=torch.int16, device=dev)
        lo = torch.empty_like(hi)
    else:
        hi, lo = taps
    assert hi.shape[:3] == (B * 16, H // 2, W // 2)
    d = L.WinoBridgeDesc()
    d.B, d.H, d.W, d.C, d.relu = B, H, W, Cc, int(relu)
    d.Cp_total, d.c_off, d.fmt = hi.shape[3], c_off, mode.fmt
    d.scale, d.eps = (mode.act_scale if act_scale is None else act_scale), 1e-5
    rows = 0
    if addend is not None:
        assert addend.shape[-1] == Cc and addend.is_contiguous() and addend.dtype == torch.float32
        rows = addend.numel() // Cc
    d.addend_rows = rows
    d.variant = variant   # 0: 32-channel slabs, 1 CTA / SM; 1: 16-channel slabs, 2 CTAs / SM
    if act_out is not None:
        assert act_out.dtype == torch.float32 and act_out.is_contiguous() and act_out.shape[:3] == (B, H, W)
        d.act_C_total, d.act_c_off = act_out.shape[-1], act_c_off
    if residual is not None:
        assert _f32(residual).shape == (B, H, W, Cc)
    if corr is not None:
        assert corr["hi"].shape == (B * H * W, Cc) and corr["ssq"].shape == (B, Cc // (16 if variant else 32), H * W)
        d.corr_hi, d.corr_lo = corr["hi"].data_ptr(), corr["lo"].data_ptr()
        d.corr_rank, d.corr_ssq, d.corr_scale = corr["rank"], corr["ssq"].data_ptr(), mode.act_scale
        corr["done"] = True
    with _Prof(("wino_bridge",)):
        L.check(L.load().tsnet_wino_bridge(C.byref(d), _ptr(m_buf), _ptr(pw.bias), _ptr(addend), _ptr(residual),
                                           _ptr(act_out), _ptr(mean_rstd_out), _ptr(hi), _ptr(lo), _stream()))
    _count()
    return hi, lo, (16, H // 2, W // 2)


def taps_geometry(mode, H, W):
    """(planes, Hd, Wd) of the tap source built from an H x W activation."""
    if mode == L.TAPS_WINO:
        return 16, H // 2, W // 2
    if mode == L.TAPS_SAME:
        return 1, H, W
    if mode == L.TAPS_REFLECT1:
        return 1, H + 2, W + 2
    if mode == L.TAPS_S2ZERO:
        return 4, H // 2 + 1, W // 2 + 1
    if mode == L.TAPS_UP2REFLECT1:
        return 1, 2 * H + 2, 2 * W + 2
    raise ValueError(mode)


def conv_taps(kind):
    """tap tables (dy, dx, plane) for the conv kinds of the network."""
    if kind == "1x1":
        return [(0, 0, 0)]
    if kind == "3x3":  # stride 1 on a pad-1 tap source
        return [(r, s, 0) for r in range(3) for s in range(3)]
    if kind == "3x3s2":  # stride 2 on the parity-split zero-padded tap source
        return [(r >> 1, s >> 1, (r & 1) * 2 + (s & 1)) for r in range(3) for s in range(3)]
    if kind == "7x1":  # kw-folded 7x7 stem
        return [(r, 0, 0) for r in range(7)]
    raise ValueError(kind)


_NUM_SMS = None


def _pick_block_n(pc, m_tiles):
    """Widest N tile (best operand reuse) that still gives every SM a tile; small batches (the demos' bs=1) fall back
    to narrower tiles so that e.g. a 512->512 conv over 3 samples is 192 tiles instead of 48."""
    global _NUM_SMS
    if pc.block_n is not None:
        return pc.block_n
    if _NUM_SMS is None:
        _NUM_SMS = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    widest = 64 if pc.Cout <= 64 else (128 if pc.Cout <= 128 else 256)
    bn = widest
    while bn > 64 and m_tiles * (pc.Cout_pad // bn) < _NUM_SMS:
        bn //= 2
    return bn


def conv_gemm(taps_hi, taps_lo, geom, pc, kind, B, H, W, mode, act_scale, want_stats=True, y=None, stats=None,
              addend=None, fuse=None, flags=0):
    """taps_* : int16 [B*planes, Hp, Wp, Cp]; geom = (planes, Hp, Wp). Returns (y_raw [B,H,W,Cout], stats)."""
    planes, Hp, Wp = geom
    d = L.ConvDesc()
    d.B, d.H, d.W, d.Cout, d.Cout_pad, d.Cp = B, H, W, pc.Cout, pc.Cout_pad, pc.Cp
    d.Hp, d.Wp, d.planes = Hp, Wp, planes
    taps = conv_taps(kind)
    assert len(taps) == pc.num_taps, (kind, pc.num_taps)
    d.num_taps = len(taps)
    for t, (dy, dx, pl) in enumerate(taps):
        d.tap_dy[t], d.tap_dx[t], d.tap_plane[t] = dy, dx, pl
    d.block_n, d.split, d.fmt = _pick_block_n(pc, B * H * W // 128), mode.split, mode.fmt
    d.out_scale = 1.0 / (pc.scale * act_scale)
    d.flags = flags
    if addend is not None:  # fp32 [rows, Cout], broadcast over the leading batch dimension by row index modulo
        assert addend.shape[-1] == pc.Cout and addend.is_contiguous() and addend.dtype == torch.float32
        d.addend, d.addend_rows = addend.data_ptr(), addend.numel() // pc.Cout
    assert taps_hi.shape == (B * planes, Hp, Wp, pc.Cp), (tuple(taps_hi.shape), (B * planes, Hp, Wp, pc.Cp))
    if fuse is not None:
        # fused InstanceNorm epilogue: fuse = dict(relu, tmode, residual, act_out, act_c_off, taps=(hi, lo), c_off)
        d.fuse_in, d.fuse_relu, d.fuse_mode = 1, int(fuse.get("relu", False)), fuse["tmode"]
        d.fuse_act_scale, d.fuse_eps = act_scale, 1e-5
        keep = []
        if fuse.get("residual") is not None:
            r = _f32(fuse["residual"])
            assert r.shape == (B, H, W, pc.Cout)
            d.fuse_residual = r.data_ptr()
        if fuse.get("act_out") is not None:
            a = fuse["act_out"]
            assert a.dtype == torch.float32 and a.is_contiguous() and a.shape[:3] == (B, H, W)
            d.fuse_act_out, d.fuse_act_C_total, d.fuse_act_c_off = a.data_ptr(), a.shape[3], fuse.get("act_c_off", 0)
        if fuse.get("taps") is not None:
            th, tl = fuse["taps"]
            _, Hd, Wd = taps_geometry(fuse["tmode"], H, W)
            assert th.shape[:3] == (B, Hd, Wd) and th.is_contiguous() and tl.shape == th.shape
            d.fuse_taps_hi, d.fuse_taps_lo = th.data_ptr(), tl.data_ptr()
            d.fuse_taps_Cp, d.fuse_taps_c_off = th.shape[3], fuse.get("c_off", 0)
        with _Prof(("conv_gemm_fused_in", kind, B, H, W, pc.Cin * (pc.KW if pc.fold_kw else 1), pc.Cout, len(taps))):
            L.check(L.load().tsnet_conv_gemm_fwd(C.byref(d), _ptr(taps_hi), _ptr(taps_lo), _ptr(pc.w_hi),
                                                 _ptr(pc.w_lo), _ptr(pc.bias), None, None, _stream()))
        _count()
        return None, None
    if y is None:
        y = torch.empty((B, H, W, pc.Cout), dtype=torch.float32, device=taps_hi.device)
    if want_stats and stats is None:
        stats = torch.empty((B * H * W // 32, pc.Cout, 2), dtype=torch.float32, device=taps_hi.device)
    cin_eff = pc.Cin * (pc.KW if pc.fold_kw else 1)
    with _Prof(("conv_gemm", kind, B, H, W, cin_eff, pc.Cout, len(taps))):
        L.check(L.load().tsnet_conv_gemm_fwd(C.byref(d), _ptr(taps_hi), _ptr(taps_lo), _ptr(pc.w_hi), _ptr(pc.w_lo),
                                             _ptr(pc.bias), _ptr(y), _ptr(stats) if want_stats else None, _stream()))
    _count()
    return y, (stats if want_stats else None)


def instnorm_reduce(stats, B, HW, Cch, eps=1e-5, out=None):
    if out is None:
        out = torch.empty((B, Cch, 2), dtype=torch.float32, device=stats.device)
    with _Prof(("instnorm_reduce",)):
        L.check(L.load().tsnet_instnorm_reduce(_ptr(stats), B, HW, Cch, C.c_float(eps), _ptr(out), _stream()))
    _count()
    return out


def build_taps(raw, mode, tmode, mean_rstd=None, relu=False, residual=None, act_out=None, act_c_off=0,
               taps=None, c_off=0, want_taps=True, avg_n=1, act_scale=None, flags=0):
    """raw: fp32 [avg_n*B, H, W, C].  Returns (taps_hi, taps_lo, geom) (None, None, geom when want_taps=False).
    `taps` = (hi, lo) preallocated destination (for channel-concatenation), else allocated here."""
    Bt, H, W, Cch = raw.shape
    B = Bt // avg_n
    planes, Hd, Wd = taps_geometry(tmode, H, W)
    d = L.TapsDesc()
    d.B, d.H, d.W, d.C, d.mode, d.relu = B, H, W, Cch, tmode, int(relu)
    d.fmt = mode.fmt
    d.scale = mode.act_scale if act_scale is None else act_scale
    d.avg_n = avg_n
    d.flags = flags
    keep = None
    if isinstance(residual, (tuple, list)):  # torch.cat([r0, r1], -1) with r1's batch broadcast, never materialised
        r0, r1 = _f32(residual[0]), _f32(residual[1])
        assert r0.shape[:3] == (B, H, W) and r1.shape[1:3] == (H, W) and r0.shape[3] + r1.shape[3] == Cch
        d.residual2, d.res_split, d.res2_batch = r1.data_ptr(), r0.shape[3], r1.shape[0]
        keep, residual = (r0, r1), r0
    hi = lo = None
    if want_taps:
        if taps is None:
            Cp = (Cch + 63) // 64 * 64
            hi = torch.empty((B * planes, Hd, Wd, Cp), dtype=torch.int16, device=raw.device)
            lo = torch.empty_like(hi)
            if Cp != Cch:
                hi.zero_()
                lo.zero_()
        else:
            hi, lo = taps
        assert hi.shape[:3] == (B * planes, Hd, Wd), (tuple(hi.shape), (B * planes, Hd, Wd))
        d.Cp_total, d.c_off = hi.shape[3], c_off
    if act_out is not None:
        d.act_C_total, d.act_c_off = act_out.shape[-1], act_c_off
    with _Prof(("build_taps", tmode)):
        L.check(L.load().tsnet_build_taps(C.byref(d), _ptr(_f32(raw)), _ptr(mean_rstd), _ptr(residual), _ptr(act_out),
                                          _ptr(hi), _ptr(lo), _stream()))
    _count()
    return hi, lo, (planes, Hd, Wd)


def stem_taps(img, img_div, lbl, Cp, mode, label_nc=None, img_mean=None):
    """img [B,3,H,W] NCHW or None: fp32 (divided by img_div in the kernel) or uint8 BGR together with img_mean (3 floats:
    (u8 - mean) / img_div is evaluated in the loader); lbl [B,L,H,W] fp32 one-hot planes, or a uint8 class-index map
    [B,H,W] together with label_nc.  Returns (hi, lo, geom) of the kw-folded tap source."""
    if lbl.dtype == torch.uint8:
        assert lbl.dim() == 3 and label_nc is not None and lbl.is_contiguous()
        (B, H, W), Clbl, kind = lbl.shape, label_nc, 1
    else:
        (B, Clbl, H, W), kind = _f32(lbl).shape, 0
    Cimg, img_kind, mean3 = 0, 0, None
    if img is not None:
        Cimg = img.shape[1]
        assert img.is_cuda and img.is_contiguous()
        if img.dtype == torch.uint8:
            assert img_mean is not None and Cimg == 3, "uint8 images need the 3 channel means"
            img_kind, mean3 = 1, (C.c_float * 3)(*[float(v) for v in img_mean])
        else:
            _f32(img)
    hi = torch.empty((B, H + 6, W, Cp), dtype=torch.int16, device=lbl.device)
    lo = torch.empty_like(hi)
    with _Prof(("stem_taps",)):
        L.check(L.load().tsnet_stem_taps(_ptr(img), Cimg, img_kind, mean3, C.c_float(img_div), _ptr(lbl), Clbl, kind,
                                         B, H, W, Cp, mode.fmt, C.c_float(mode.act_scale), _ptr(hi), _ptr(lo),
                                         _stream()))
    _count()
    return hi, lo, (1, H + 6, W)


STEM_FOLD = 8   # channels per folded tap of tsnet_stem_conv_fwd


def stem_conv_ok(Cimg, Clbl, H, W, Cout):
    """Shapes the direct-input stem kernel covers (the face configuration of both encoders)."""
    return Cimg + Clbl + 3 <= STEM_FOLD and Cout == 64 and H % 8 == 0 and W % 16 == 0


def stem_conv(img, img_div, lbl, pc, mode, label_nc=None, img_mean=None, want_stats=True):
    """ReflectionPad2d(3) + 7x7 conv of an encoder stem straight from the raw NCHW inputs (tsnet_stem_conv_fwd): no
    tap source is materialised.  Arguments as stem_taps; pc = PackedConv(..., fold_kw=True, fold_cin=STEM_FOLD).
    Returns (y_raw [B, H, W, 64], stats_partial or None)."""
    assert pc.fold_kw and pc.fold_cin == STEM_FOLD and pc.Cp == 64 and pc.Cout_pad == 64
    d = L.StemConvDesc()
    if lbl.dtype == torch.uint8:
        assert lbl.dim() == 3 and label_nc is not None and lbl.is_contiguous()
        (B, H, W), d.Clbl, d.lbl_kind = lbl.shape, label_nc, 1
    else:
        (B, d.Clbl, H, W), d.lbl_kind = _f32(lbl).shape, 0
    d.B, d.H, d.W = B, H, W
    d.Cimg, d.img_kind = 0, 0
    if img is not None:
        d.Cimg = img.shape[1]
        assert img.is_cuda and img.is_contiguous() and img.shape == (B, d.Cimg, H, W)
        if img.dtype == torch.uint8:
            assert img_mean is not None and d.Cimg == 3, "uint8 images need the 3 channel means"
            d.img_kind = 1
            for k in range(3):
                d.img_mean[k] = float(img_mean[k])
        else:
            _f32(img)
    d.img_div = float(img_div)
    d.Cout, d.split, d.fmt = pc.Cout, mode.split, mode.fmt
    d.act_scale, d.out_scale = mode.act_scale, 1.0 / (pc.scale * mode.act_scale)
    y = torch.empty((B, H, W, pc.Cout), dtype=torch.float32, device=lbl.device)
    stats = torch.empty((B * H * W // 32, pc.Cout, 2), dtype=torch.float32, device=lbl.device) if want_stats else None
    with _Prof(("conv_gemm", "7x1", B, H, W, (d.Cimg + d.Clbl + 3) * 7, pc.Cout, 7)):
        L.check(L.load().tsnet_stem_conv_fwd(C.byref(d), _ptr(img), _ptr(lbl), _ptr(pc.w_hi), _ptr(pc.w_lo),
                                             _ptr(pc.bias), _ptr(y), _ptr(stats), _stream()))
    _count()
    return y, stats


class CorrPlan:
    """Prepared correlation workspace of one forward (tsnet_corr_prepare): class-sorted order of every mask,
    tile classes, work list.  `rank_t` / `rank_s` are the position -> sorted-rank tables l2norm_split needs."""

    def __init__(self, desc, ws, rank_t, rank_s, B, n, Cch, h, w):
        self.desc, self.ws, self.rank_t, self.rank_s = desc, ws, rank_t, rank_s
        self.B, self.n, self.C, self.h, self.w = B, n, Cch, h, w


def corr_prepare(tar_bbox, src_bbox_list, coord_table, B, Cch, h, w, mode, temperature=100.0, sort=True,
                 one_cta=False, chunk_kb=0, normalized=False):
    """model/TSNet.py:322-323, :347-348 (nearest down-sampling of the bbox masks) + the class-sorted work plan of the
    correlation kernel.  bboxes [B, Hb, Wb] uint8 or fp32 (full resolution)."""
    n = len(src_bbox_list)
    assert tar_bbox.dtype in (torch.uint8, torch.float32)
    assert all(bb.dtype == tar_bbox.dtype and bb.is_contiguous() for bb in src_bbox_list) and tar_bbox.is_contiguous()
    d = L.CorrDesc()
    d.B, d.n_src, d.C, d.h, d.w = B, n, Cch, h, w
    d.bbox_h, d.bbox_w = tar_bbox.shape[-2], tar_bbox.shape[-1]
    d.bbox_dtype = 0 if tar_bbox.dtype == torch.uint8 else 1
    d.temperature, d.split, d.fmt = temperature, mode.split, mode.fmt
    # operands: un-normalised features at the activation scale + reciprocal norms (default, corr_operands / the bridge
    # pass), or L2-normalised unit vectors at corr_scale (l2norm_split)
    d.operand_scale = mode.corr_scale * mode.corr_scale if normalized else mode.act_scale * mode.act_scale
    d.sort = 1 if sort else 0
    d.one_cta, d.chunk_kb = int(one_cta), chunk_kb
    lib = L.load()
    nbytes = lib.tsnet_corr_workspace_bytes(C.byref(d))
    if nbytes == 0:
        raise L.TSNetLibraryError(f"corr_prepare: {lib.tsnet_last_error().decode()}")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=tar_bbox.device)
    bb_ptrs = (C.c_void_p * n)(*[bb.data_ptr() for bb in src_bbox_list])
    with _Prof(("corr_prepare",)):
        L.check(lib.tsnet_corr_prepare(C.byref(d), _ptr(tar_bbox), bb_ptrs, _ptr(_f32(coord_table)), _ptr(ws), nbytes,
                                       _stream()))
    _count()
    rank_t = lib.tsnet_corr_rank_table(C.byref(d), _ptr(ws), 0)
    rank_s = lib.tsnet_corr_rank_table(C.byref(d), _ptr(ws), 1)
    plan = CorrPlan(d, ws, rank_t, rank_s, B, n, Cch, h, w)
    plan.normalized = normalized
    plan._keep = (tar_bbox, src_bbox_list, coord_table)
    return plan


def l2norm_split(fea, mode, out=None, rank=None):
    """fea fp32 [B, hw, C] (NHWC flattened) -> (hi, lo) int16 [B*hw, C] of F.normalize(dim=C) * corr_scale.
    rank: device pointer (int) of a [B, hw] uint16 position -> row table (CorrPlan.rank_t / rank_s) or None."""
    B, HW, Cch = fea.shape
    if out is None:
        hi = torch.empty((B * HW, Cch), dtype=torch.int16, device=fea.device)
        lo = torch.empty_like(hi)
    else:
        hi, lo = out
    with _Prof(("l2norm_split",)):
        L.check(L.load().tsnet_l2norm_split(_ptr(_f32(fea)), B, HW, Cch, mode.fmt, C.c_float(mode.corr_scale),
                                            C.c_void_p(rank) if rank else None, _ptr(hi), _ptr(lo), _stream()))
    _count()
    return hi, lo


def corr_operands(fea, mode, rank=None, out=None):
    """fea fp32 [B, hw, C] -> (hi, lo, rnorm): un-normalised operands x * act_scale as int16 hi/lo [B*hw, C] and
    rnorm [B*hw] = 1 / max(||x||, 1e-12), rows at their sorted rank (tsnet_corr_operands)."""
    B, HW, Cch = fea.shape
    if out is None:
        hi = torch.empty((B * HW, Cch), dtype=torch.int16, device=fea.device)
        lo = torch.empty_like(hi)
        rn = torch.empty((B * HW,), dtype=torch.float32, device=fea.device)
    else:
        hi, lo, rn = out
    with _Prof(("corr_operands",)):
        L.check(L.load().tsnet_corr_operands(_ptr(_f32(fea)), B, HW, Cch, mode.fmt, C.c_float(mode.act_scale),
                                             C.c_void_p(rank) if rank else None, _ptr(hi), _ptr(lo), _ptr(rn), _stream()))
    _count()
    return hi, lo, rn


def corr_norms(ssq_part, B, HW, slabs, rank):
    """[B, slabs, HW] partial sums of squares (tsnet_wino_bridge) -> rnorm [B*HW] at the sorted rank."""
    rn = torch.empty((B * HW,), dtype=torch.float32, device=ssq_part.device)
    with _Prof(("corr_norms",)):
        L.check(L.load().tsnet_corr_norms(_ptr(_f32(ssq_part)), B, HW, slabs, C.c_void_p(rank) if rank else None,
                                          _ptr(rn), _stream()))
    _count()
    return rn


def corr_warp(plan, tar_ops, src_ops, src_fea_list, mode, want_grids=False, want_mean=True, taps=None, c_off=0):
    """plan = corr_prepare(...); tar_ops = (hi, lo[, rnorm]) [B*hw, C] and src_ops = (hi, lo[, rnorm]) [n*B*hw, C] from
    corr_operands (or l2norm_split for a `normalized` plan) with the plan's rank tables; src_fea_list n x fp32
    [B, hw, C].  Returns (out_mean fp32 [B, hw, C] or None,
    grids [n, B, h, w, 2] or None); `taps` = (hi, lo) [B, h, w, Cp]: the mean is also written as hi/lo operands into
    the channel window [c_off, c_off + C)."""
    n, B, h, w, Cch = plan.n, plan.B, plan.h, plan.w, plan.C
    dev = tar_ops[0].device
    out = torch.empty((B, h * w, Cch), dtype=torch.float32, device=dev) if want_mean else None
    grids = torch.empty((n, B, h, w, 2), dtype=torch.float32, device=dev) if want_grids else None
    assert out is not None or grids is not None or taps is not None
    hi, lo = taps if taps is not None else (None, None)
    need_fea = want_mean or taps is not None
    fea_ptrs = (C.c_void_p * n)(*[_f32(f).data_ptr() for f in src_fea_list]) if need_fea else None
    lib = L.load()
    rn_t = tar_ops[2] if len(tar_ops) > 2 else None
    rn_s = src_ops[2] if len(src_ops) > 2 else None
    assert (rn_t is None) == (rn_s is None) == bool(getattr(plan, "normalized", False)), "operand kind / plan mismatch"
    with _Prof(("corr_tiles",)):
        L.check(lib.tsnet_corr_tiles(C.byref(plan.desc), _ptr(tar_ops[0]), _ptr(tar_ops[1]), _ptr(src_ops[0]),
                                     _ptr(src_ops[1]), _ptr(rn_t), _ptr(rn_s), _ptr(plan.ws), plan.ws.numel(),
                                     _stream()))
    _count()
    with _Prof(("corr_finish",)):
        L.check(lib.tsnet_corr_finish(C.byref(plan.desc), fea_ptrs, _ptr(out), _ptr(grids), _ptr(hi), _ptr(lo),
                                      0 if hi is None else hi.shape[-1], c_off, C.c_float(mode.act_scale),
                                      _ptr(plan.ws), plan.ws.numel(), _stream()))
    _count()
    return out, grids


def corr_chain(tar_fea, src_fea, tar_bbox, src_bbox_list, coord_table, mode, temperature=100.0, want_grids=True,
               want_mean=False, taps=None, c_off=0, sort=True, one_cta=False, normalized=False):
    """The whole transformation branch on raw features (model/TSNet.py:319-366, :392): tar_fea fp32 [B, hw, C],
    src_fea fp32 [n, B, hw, C] -> (mean of the warped sources [B, hw, C] or None, warp grids [n, B, h, w, 2] or None)."""
    n, B, hw, Cch = src_fea.shape
    h = w = int(round(hw ** 0.5))
    plan = corr_prepare(tar_bbox, src_bbox_list, coord_table, B, Cch, h, w, mode, temperature=temperature, sort=sort,
                        one_cta=one_cta, normalized=normalized)
    prep = l2norm_split if normalized else corr_operands
    tar_ops = prep(tar_fea, mode, rank=plan.rank_t)
    src_ops = prep(src_fea.view(n * B, hw, Cch), mode, rank=plan.rank_s)
    return corr_warp(plan, tar_ops, src_ops, [src_fea[i] for i in range(n)], mode, want_grids=want_grids,
                     want_mean=want_mean, taps=taps, c_off=c_off)


def corr_grids(tar_fea, src_fea, tar_bbox, src_bbox_list, coord_table, mode, temperature=100.0, sort=True):
    return corr_chain(tar_fea, src_fea, tar_bbox, src_bbox_list, coord_table, mode, temperature=temperature,
                      sort=sort)[1]


def warp_mean_taps(src_fea_list, grids, B, h, w, Cch, mode, taps=None, c_off=0, want_mean=False):
    """grids [n,B,h,w,2] + n x fp32 [B,hw,C] source features -> mean_i grid_sample(src_i, G_i): fp32 [B,hw,C] (optional)
    and / or written as hi/lo operands into the channel window [c_off, c_off+C) of `taps` = (hi, lo) [B,h,w,Cp]."""
    n = len(src_fea_list)
    out = torch.empty((B, h * w, Cch), dtype=torch.float32, device=grids.device) if want_mean else None
    hi, lo = taps if taps is not None else (None, None)
    fea_ptrs = (C.c_void_p * n)(*[_f32(f).data_ptr() for f in src_fea_list])
    with _Prof(("warp_mean_taps",)):
        L.check(L.load().tsnet_warp_mean_taps(fea_ptrs, n, _ptr(_f32(grids)), B, h, w, Cch, _ptr(out), _ptr(hi), _ptr(lo),
                                              0 if hi is None else hi.shape[-1], c_off, mode.fmt,
                                              C.c_float(mode.act_scale), _stream()))
    _count()
    return out


def train_extras(src_imgs, src_divs, tar_img_raw, tar_div, grids, pg_mean, sg_mean, fore=None, fill=None):
    """Train-mode branches of forward() (model/TSNet.py:327-331, 372-390, 402-405; TSNet_pose.py:395-396).
    src_imgs: n raw NCHW [B,3,H,W]; grids [n,B,h,w,2]; pg_mean / sg_mean fp32 [B,hw,C] or None (pose).
    Returns (warp [n,B,3,H,W], losses float[2] = (loss_warp, loss_align))."""
    n = len(src_imgs)
    B, _, H, W = src_imgs[0].shape
    h, w = grids.shape[2], grids.shape[3]
    dev = grids.device
    warp = torch.empty((n, B, 3, H, W), dtype=torch.float32, device=dev)
    losses = torch.empty(2, dtype=torch.float32, device=dev)
    lib = L.load()
    nbytes = lib.tsnet_train_extras_workspace_bytes(B, n, h, w)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    img_ptrs = (C.c_void_p * n)(*[_f32(x).data_ptr() for x in src_imgs])
    divs = (C.c_float * n)(*[float(d) for d in src_divs])
    fx0, fx1 = fore if fore is not None else (0, 0)
    fill3 = (C.c_float * 3)(*fill) if fill is not None else None
    Cch = pg_mean.shape[-1] if pg_mean is not None else 0
    with _Prof(("train_extras",)):
        L.check(lib.tsnet_train_extras_fwd(img_ptrs, divs, n, _ptr(_f32(tar_img_raw)), C.c_float(tar_div),
                                           _ptr(_f32(grids)), B, H, W, h, w,
                                           _ptr(None if pg_mean is None else _f32(pg_mean)),
                                           _ptr(None if sg_mean is None else _f32(sg_mean)), Cch, fx0, fx1, fill3,
                                           _ptr(warp), _ptr(losses), _ptr(ws), nbytes, _stream()))
    for _ in range(6):
        _count()
    return warp, losses


def head_conv_tanh(act, weight, bias, fore=None, fill=None, mean_rstd=None, relu=False):
    """act fp32 NHWC [B,H,W,Cin] -> NCHW [B,3,H,W] = tanh(conv7x7(reflectpad3(act))) (+ pose compositing).
    mean_rstd [B,Cin,2] / relu: `act` is a raw conv output, InstanceNorm (+ ReLU) is applied inside the loader."""
    B, H, W, Cin = act.shape
    out = torch.empty((B, 3, H, W), dtype=torch.float32, device=act.device)
    x0, x1 = fore if fore is not None else (0, 0)
    fill_arr = (C.c_float * 3)(*(fill if fill is not None else (0.0, 0.0, 0.0)))
    with _Prof(("head_conv_tanh",)):
        L.check(L.load().tsnet_head_conv_tanh(_ptr(_f32(act)), _ptr(None if mean_rstd is None else _f32(mean_rstd)),
                                              int(relu), B, H, W, Cin, _ptr(_f32(weight.detach())),
                                              _ptr(_f32(bias.detach())), x0, x1, fill_arr, _ptr(out), _stream()))
    _count()
    return out


def plane_stats(x, planes, n, div=1.0):
    """(mean, unbiased std) of x / div over each of `planes` contiguous runs of n floats -> fp32 [planes, 2]
    (tensor.mean / tensor.std of the reference, fp64 accumulation in a fixed order)."""
    out = torch.empty((planes, 2), dtype=torch.float32, device=x.device)
    with _Prof(("plane_stats",)):
        L.check(L.load().tsnet_plane_stats(_ptr(_f32(x)), planes, n, C.c_float(div), _ptr(out), _stream()))
    _count()
    return out


def postprocess_u8(rec, ref_mean, ref_std, img_mean, gen_stats=None):
    """rec fp32 NCHW [B,3,H,W] -> uint8 RGB [B,H,W,3]: the demos' colour re-normalisation + sample_img
    (demo/demo_face.py:96-105, 194-199).  ref_mean / ref_std: CUDA tensors with 3 floats; img_mean: 3 python floats
    (IMG_MEAN / 255); gen_stats: [B*3, 2] statistics of the generated planes (default: tsnet_plane_stats of rec)."""
    B, _, H, W = rec.shape
    if gen_stats is None:
        gen_stats = plane_stats(rec, B * 3, H * W)
    out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=rec.device)
    im = (C.c_float * 3)(*[float(v) for v in img_mean])
    with _Prof(("postprocess_u8",)):
        L.check(L.load().tsnet_postprocess_u8(_ptr(_f32(rec)), B, H, W, _ptr(_f32(gen_stats)),
                                              _ptr(_f32(ref_mean.reshape(3))), _ptr(_f32(ref_std.reshape(3))), im,
                                              _ptr(out), _stream()))
    _count()
    return out


def direct_conv_fp32(x_nhwc, weight, bias, stride=1, pad=0, reflect=False):
    """Validation-only fp32 direct convolution (NHWC in / out)."""
    B, H, W, Cin = x_nhwc.shape
    Cout, _, K, _ = weight.shape
    Ho, Wo = (H + 2 * pad - K) // stride + 1, (W + 2 * pad - K) // stride + 1
    y = torch.empty((B, Ho, Wo, Cout), dtype=torch.float32, device=x_nhwc.device)
    L.check(L.load().tsnet_direct_conv_fp32(_ptr(_f32(x_nhwc)), B, H, W, Cin, _ptr(_f32(weight)), _ptr(bias), Cout, K,
                                            stride, pad, int(reflect), _ptr(y), _stream()))
    _count()
    return y
