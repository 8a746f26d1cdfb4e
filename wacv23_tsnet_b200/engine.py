"""Forward engine: sequences the sm_100a kernels for TSNet.forward().

Reference path being replaced: model/TSNet.py:309-407 (face) and model/TSNet_pose.py:325-417 (pose): the
is_train=False forward and, on request (`train=`), the forward-only is_train=True branches.  All activations stay on
the device in NHWC; the only torch ops used are allocations.
The n sources are run through img_enc / fuse_net as ONE batch of n*B samples (shared weights), the target
through lbl_enc / dec as a batch of B.
"""
import torch

from . import lib as L
from . import ops
from .ops import MathMode, PackedConv


class ForwardEngine:
    def __init__(self, img_enc, lbl_enc, fuse_net, dec, label_nc, n_blocks_dec, n_downsampling=3, ngf=64,
                 math_mode="fp16x3", winograd=True):
        if n_downsampling != 3 or ngf != 64:
            # FuseNet(ngf=1024) is hard-coded in the reference (model/TSNet.py:227): only 64 * 2^3 = 512 fits.
            raise ValueError("TS-Net geometry requires ngf=64, n_downsampling=3 (FuseNet is fixed at 1024 channels)")
        self.nets = dict(img_enc=img_enc, lbl_enc=lbl_enc, fuse_net=fuse_net, dec=dec)
        self.label_nc = label_nc
        self.n_blocks_dec = n_blocks_dec
        self.mode = MathMode(math_mode)
        # TSNET_FUSED_IN=1: InstanceNorm + ReLU + residual + tap building fused into the GEMM epilogue of the 32x32
        # layers (8-CTA clusters exchanging statistics over DSMEM, tsnet_conv_desc.fuse_in).  Correct and tested, but
        # measured SLOWER on B200 (54.4 vs 49.8 ms/step): only 16 such clusters (128 of 148 SMs) can be co-resident
        # and the longer epilogue serialises with the MMA pipe -- see DESIGN.md section 4.  Default: unfused.
        import os
        self.fused_in = os.environ.get("TSNET_FUSED_IN", "0") == "1"
        # Winograd F(2x2, 3x3) for every ResnetBlock convolution (img_enc x 18, FuseNet x 2 + its target half, decoder
        # x 2 n_blocks): 2.25 x fewer tensor-core MACs for the layers that take ~60 % of the step.
        # winograd=False keeps the direct implicit GEMM for them (A/B comparisons, tests).
        # winograd="unfused" keeps the separate output-transform / instnorm_reduce / input-transform passes between two
        # Winograd layers instead of the fused bridge pass (tsnet_wino_bridge).
        self.winograd = bool(winograd)
        self.bridge = self.winograd and winograd != "unfused"
        # K blocks (of 64 channels) accumulated in TMEM before the partial sum is promoted to fp32 registers in the
        # Winograd plane GEMMs (K = Cin per plane).  Measured on the B200 (tools/wino_bench.py, 512 -> 512 x 96 samples):
        # 2 -> 0.572 ms / 6.4e-7 of max|ref| vs an fp64 conv; 4 -> 0.503 ms / 1.1e-6; 8 (no promotion) -> 0.474 ms / 2.1e-6
        # Whole-forward parity (tools/parity_report.py, worst of the 7 goldens, tolerance 1e-3 / 5e-5): chunk 2 -> image
        # 8.7e-4, grids 2.4e-5 (the direct path: 9.3e-4 / 2.3e-5); chunk 4 -> 1.07e-3 / 2.6e-5; chunk 8 -> 1.3e-3 / 3.4e-5.
        # The error is set by the 18-convolution img_enc chain: chunk 2 there and chunk 4 in FuseNet / decoder gives exactly
        # the all-2 worst case (8.749e-4 / 2.44e-5), while chunk 4 in img_enc alone already gives 1.07e-3.  Default: that
        # mix.  The value is an int, or a dict {net name: chunk} with key "default".
        self.wino_chunk_kb = {"img_enc": 2, "default": 4}
        self.wino_flags = 0      # tsnet_wino_gemm_desc.flags (experiments: L.CONV_SMALL_FIRST)
        self.bridge_variant = 0  # tsnet_wino_bridge_desc.variant: 0 = 32-channel slabs (1 CTA / SM), 1 = 16-channel (2 / SM)
        self._packs = {}
        self._coord = {}
        self._src_cache = None   # opt-in source-feature cache (see forward(src_key=...))
        # encoder stems straight from the raw network inputs (tsnet_stem_conv_fwd) wherever the channels fit one
        # 8-channel folded tap (face configuration) instead of the materialised tap source (tsnet_stem_taps).  Measured on
        # the B200 (bs=32, n=3): layer DRAM traffic 4.96 GB -> 1.78 GB, but the three producer warps are latency-bound on
        # their scattered plane loads: 2.42 + 0.71 ms against 1.43 + 0.45 + 0.70 ms (stem_taps) for the materialised
        # path.  Opt-in until the producers are software-pipelined across tiles.
        self.direct_stem = False

    def invalidate(self):
        """Drop every packed weight.  Needed after parameter writes that do not bump the tensor version
        (`p.data.copy_()`, `init.normal_(p.data)`, `dist.broadcast(p.data)`): the pack cache is keyed on
        (data_ptr, _version) and would otherwise keep serving the old operands."""
        self._packs.clear()
        self._src_cache = None

    # ------------------------------------------------------------------ weights
    def _pack(self, net, wkey, fold_kw=False, block_n=None, cin_range=None, with_bias=True, fold_cin=None):
        """Packed weight for parameter `wkey` of sub-net `net`; re-packed when the parameter changed
        (load_state_dict / optimizer step bump the tensor version)."""
        sd = self.nets[net].state_dict(keep_vars=True)
        w, b = sd[wkey + ".weight"], sd[wkey + ".bias"]
        sig = (w.data_ptr(), w._version, b.data_ptr(), b._version, self.mode.name)
        key = (net, wkey, cin_range, with_bias, fold_cin)
        hit = self._packs.get(key)
        if hit is None or hit[0] != sig:
            hit = (sig, PackedConv(w, b if with_bias else None, self.mode, fold_kw=fold_kw, block_n=block_n,
                                   cin_range=cin_range, fold_cin=fold_cin))
            self._packs[key] = hit
        return hit[1]

    def _pack_wino(self, net, wkey, cin_range=None, with_bias=True):
        """Winograd-domain packed weight (U = G g G^T, 16 K-major hi/lo matrices); same cache policy as _pack."""
        sd = self.nets[net].state_dict(keep_vars=True)
        w, b = sd[wkey + ".weight"], sd[wkey + ".bias"]
        sig = (w.data_ptr(), w._version, b.data_ptr(), b._version, self.mode.name)
        key = (net, wkey, cin_range, with_bias, "wino")
        hit = self._packs.get(key)
        if hit is None or hit[0] != sig:
            hit = (sig, ops.PackedWino(w, b if with_bias else None, self.mode, cin_range=cin_range))
            self._packs[key] = hit
        return hit[1]

    def _chunk(self, net):
        ck = self.wino_chunk_kb
        return ck.get(net, ck.get("default", 2)) if isinstance(ck, dict) else ck

    def _tmode3(self, H, W, Cin, Cout):
        """Operand format a 3x3 reflect-pad convolution consumes: Winograd planes where the path applies."""
        return L.TAPS_WINO if (self.winograd and ops.wino_ok(H, W, Cin, Cout)) else L.TAPS_REFLECT1

    def _coord_table(self, h, w, device):
        key = (h, w, str(device))
        if key not in self._coord:
            self._coord[key] = torch.cat([torch.linspace(-1, 1, h), torch.linspace(-1, 1, w)]).float().to(device)
        return self._coord[key]

    # ------------------------------------------------------------------ building blocks
    def _conv(self, taps, pc, kind, B, H, W, norm=True, addend=None):
        hi, lo, geom = taps
        y, stats = ops.conv_gemm(hi, lo, geom, pc, kind, B, H, W, self.mode, self.mode.act_scale, want_stats=norm,
                                 addend=addend)
        mr = ops.instnorm_reduce(stats, B, H * W, pc.Cout) if norm else None
        return y, mr

    def _conv3(self, taps, net, wkey, B, H, W, norm=True, addend=None, cin_range=None, with_bias=True):
        """3x3 stride-1 reflect-pad convolution (ResnetBlock, model/TSNet.py:27,42) on a REFLECT1 tap source (direct
        implicit GEMM) or on Winograd planes (taps geometry (16, H/2, W/2)).  Returns (y_raw, mean_rstd or None)."""
        if taps[2][0] == 16:
            pw = self._pack_wino(net, wkey, cin_range=cin_range, with_bias=with_bias)
            y, stats = ops.wino_conv(taps, pw, B, H, W, self.mode, self.mode.act_scale, want_stats=norm, addend=addend,
                                     chunk_kb=self._chunk(net), flags=self.wino_flags)
            mr = ops.instnorm_reduce(stats, B, H * W, pw.Cout) if norm else None
            return y, mr
        pc = self._pack(net, wkey, cin_range=cin_range, with_bias=with_bias)
        return self._conv(taps, pc, "3x3", B, H, W, norm=norm, addend=addend)

    def _conv_in(self, taps, pc, kind, B, H, W, tmode, relu=False, residual=None, need_act=False, dest=None, c_off=0,
                 want_taps=True, addend=None, act_out=None, act_c_off=0, corr_out=None):
        """conv -> InstanceNorm -> [ReLU] -> [+ residual] -> [fp32 act_out] -> tap source of the next layer (tmode).
        `pc` is a PackedConv, or (net, wkey[, cin_range]) for a ResnetBlock 3x3 convolution (direct or Winograd,
        decided by the operand format of `taps`).
        Returns ((hi, lo, geom) or None, act_out or None)."""
        m = self.mode
        is3 = isinstance(pc, tuple)
        Cout = (self.nets[pc[0]].state_dict(keep_vars=True)[pc[1] + ".weight"].shape[0]) if is3 else pc.Cout
        if act_out is None and need_act:
            act_out = torch.empty((B, H, W, Cout), dtype=torch.float32, device=taps[0].device)
        if (self.fused_in and H * W == 1024 and tmode in (L.TAPS_SAME, L.TAPS_REFLECT1) and taps[2][0] != 16 and
                not isinstance(residual, tuple)):
            if is3:
                pc = self._pack(pc[0], pc[1], cin_range=pc[2] if len(pc) > 2 else None)
            planes, Hd, Wd = ops.taps_geometry(tmode, H, W)
            if want_taps and dest is None:
                hi = torch.empty((B, Hd, Wd, pc.Cout), dtype=torch.int16, device=taps[0].device)
                dest = (hi, torch.empty_like(hi))
            ops.conv_gemm(taps[0], taps[1], taps[2], pc, kind, B, H, W, m, m.act_scale, addend=addend,
                          fuse=dict(relu=relu, tmode=tmode, residual=residual, act_out=act_out, act_c_off=act_c_off,
                                    taps=dest if want_taps else None, c_off=c_off))
            return ((dest[0], dest[1], (planes, Hd, Wd)) if want_taps else None), act_out
        if is3 and self.bridge and taps[2][0] == 16 and tmode == L.TAPS_WINO and want_taps:
            # Winograd layer feeding a Winograd layer: GEMM + ONE bridge pass (no y_raw / statistics round trip)
            pw = self._pack_wino(pc[0], pc[1], cin_range=pc[2] if len(pc) > 2 else None)
            mbuf = ops.wino_gemm(taps, pw, B, H, W, m, m.act_scale, chunk_kb=self._chunk(pc[0]), flags=self.wino_flags)
            t = ops.wino_bridge(mbuf, pw, B, H, W, m, relu=relu, addend=addend, residual=residual, act_out=act_out,
                                act_c_off=act_c_off, taps=dest, c_off=c_off, corr=corr_out,
                                variant=self.bridge_variant)
            return t, act_out
        if is3:
            y, mr = self._conv3(taps, pc[0], pc[1], B, H, W, addend=addend, cin_range=pc[2] if len(pc) > 2 else None)
        else:
            y, mr = self._conv(taps, pc, kind, B, H, W, addend=addend)
        t = ops.build_taps(y, m, tmode, mean_rstd=mr, relu=relu, residual=residual, act_out=act_out,
                           act_c_off=act_c_off, taps=dest, c_off=c_off, want_taps=want_taps)
        return (t if want_taps else None), act_out

    def _resblock(self, net, prefix, taps, x_res, B, H, W, dim, need_act=True, want_taps=True,
                  tmode_out=L.TAPS_REFLECT1, corr_out=None):
        """ResnetBlock (model/TSNet.py:10-49). taps = operand of conv_block.1 built from x (REFLECT1 or Winograd
        planes), x_res = fp32 x (residual).  Returns (taps of the output in tmode_out, fp32 output or None)."""
        t1, _ = self._conv_in(taps, (net, prefix + "conv_block.1"), "3x3", B, H, W, self._tmode3(H, W, dim, dim),
                              relu=True)
        return self._conv_in(t1, (net, prefix + "conv_block.5"), "3x3", B, H, W, tmode_out, residual=x_res,
                             need_act=need_act, want_taps=want_taps, corr_out=corr_out)

    def _encoder(self, net, img, img_div, lbl, n_blocks, final_tmode=None, img_mean=None, corr_out=None):
        """Encoder.forward (model/TSNet.py:52-125). Returns (fp32 NHWC feature [X,32,32,512], taps of it in
        final_tmode or None).  For n_blocks = 0 (lbl_enc) the feature is relu(IN(conv)), for img_enc it is the
        residual stream."""
        m = self.mode
        X, H, W = lbl.shape[0], lbl.shape[-2], lbl.shape[-1]
        Cimg = 0 if img is None else img.shape[1]
        Clbl = self.label_nc if lbl.dtype == torch.uint8 else lbl.shape[1]
        if self.direct_stem and ops.stem_conv_ok(Cimg, Clbl, H, W, 64):
            pc = self._pack(net, "model.1", fold_kw=True, fold_cin=ops.STEM_FOLD)
            y, stats = ops.stem_conv(img, img_div, lbl, pc, m, label_nc=self.label_nc, img_mean=img_mean)
            mr = ops.instnorm_reduce(stats, X, H * W, pc.Cout)
        else:
            pc = self._pack(net, "model.1", fold_kw=True)
            t = ops.stem_taps(img, img_div, lbl, pc.Cp, m, label_nc=self.label_nc, img_mean=img_mean)
            y, mr = self._conv(t, pc, "7x1", X, H, W)
        for k, idx in enumerate((4, 7)):
            t = ops.build_taps(y, m, L.TAPS_S2ZERO, mean_rstd=mr, relu=True)
            pc = self._pack(net, f"model.{idx}")
            H, W = H // 2, W // 2
            y, mr = self._conv(t, pc, "3x3s2", X, H, W)
        t = ops.build_taps(y, m, L.TAPS_S2ZERO, mean_rstd=mr, relu=True)
        pc = self._pack(net, "model.10")
        H, W = H // 2, W // 2
        if n_blocks == 0:   # lbl_enc: the feature is relu(IN(conv))
            _, fea = self._conv_in(t, pc, "3x3s2", X, H, W, L.TAPS_SAME, relu=True, need_act=True, want_taps=False)
            return fea, None
        dim = pc.Cout
        t, x = self._conv_in(t, pc, "3x3s2", X, H, W, self._tmode3(H, W, dim, dim), relu=True, need_act=True)
        for blk in range(n_blocks):
            last = blk == n_blocks - 1
            t, x = self._resblock(net, f"model.{13 + blk}.", t, x, X, H, W, dim,
                                  tmode_out=final_tmode if last else self._tmode3(H, W, dim, dim),
                                  corr_out=corr_out if last else None)
        return x, t

    # ------------------------------------------------------------------ whole forward
    @torch.no_grad()
    def forward(self, src_imgs, img_divs, src_lbls, src_bboxes, tar_lbl, tar_bbox, return_flow=False,
                pose_fill=None, collect=None, train=None, img_mean=None, src_key=None):
        """src_imgs / src_lbls: lists of n fp32 NCHW CUDA tensors [B,3,256,256] / [B,L,256,256] (labels may instead be
        uint8 class-index maps [B,256,256]: vl2ch is then evaluated inside the stem loader; images NOT yet /255:
        img_divs[i] is the divisor set_*_input would have applied: 255, or 1 for use_prev sources); src_bboxes / tar_bbox: [B,256,256] uint8|fp32.
        Returns (rec_tar_img NCHW fp32, list of warp grids [B,h,w,2] or None).
        `collect`: optional dict receiving intermediates (tests).
        `src_key`: optional hashable identifying the CONTENT of the source images + labels (opt-in source-feature cache
        for the demo loops, which re-feed the same source frames for every driving frame, demo/demo_face.py:170-192): when
        it equals the key of the previous forward (and the img_enc weights are unchanged) img_enc is not re-run.
        `train`: optional dict {"tar_img": raw NCHW target image, "align": bool}: the is_train=True branches of the
        reference forward (image-space warp, warp / alignment losses) are evaluated and returned in it as
        "warp" [n,B,3,H,W] and "losses" (device float[2])."""
        L.require_device()
        m = self.mode
        n = len(src_imgs)
        B, H0, W0 = tar_lbl.shape[0], tar_lbl.shape[-2], tar_lbl.shape[-1]
        dev = tar_lbl.device
        h, w, Cf = H0 // 8, W0 // 8, 512
        hw = h * w

        # ---- encoders.  The last img_enc block writes the operand of FuseNet's first conv (source half) directly.
        # ---- argument validation (a mismatched source would otherwise be read out of bounds on the device)
        if not (len(img_divs) == len(src_lbls) == len(src_bboxes) == n and n >= 1):
            raise ValueError(f"forward: {n} source images but {len(src_lbls)} labels / {len(src_bboxes)} bboxes")
        for i in range(n):
            if src_imgs[i].shape != (B, 3, H0, W0):
                raise ValueError(f"forward: source image {i} has shape {tuple(src_imgs[i].shape)}, expected {(B, 3, H0, W0)}")
            if (src_lbls[i].shape[0] != B or src_lbls[i].shape[-2:] != (H0, W0) or
                    src_lbls[i].dtype != src_lbls[0].dtype or src_lbls[i].shape != src_lbls[0].shape):
                raise ValueError(f"forward: source label {i} {tuple(src_lbls[i].shape)} / {src_lbls[i].dtype} does not "
                                 f"match batch {B}, size {(H0, W0)} or the format of source label 0")
            if src_bboxes[i].shape != tar_bbox.shape or src_bboxes[i].dtype != tar_bbox.dtype:
                raise ValueError(f"forward: source bbox {i} {tuple(src_bboxes[i].shape)} / {src_bboxes[i].dtype} does not "
                                 f"match the target bbox {tuple(tar_bbox.shape)} / {tar_bbox.dtype}")
        if tar_bbox.shape[0] != B:
            raise ValueError("forward: target bbox batch size differs from the target label's")
        if src_imgs[0].dtype == torch.uint8 and (img_mean is None or len(set(img_divs)) > 1):
            if img_mean is None:
                raise ValueError("forward: uint8 source images need img_mean")
            mean_t = torch.tensor(img_mean, dtype=torch.float32, device=dev).view(1, 3, 1, 1)
            src_imgs, img_mean = [im.float() - mean_t for im in src_imgs], None
        if src_imgs[0].dtype != torch.uint8:
            img_mean = None
        if n > 1 and len(set(img_divs)) > 1:
            # use_prev mixes /255 and raw sources (model/TSNet.py:270-276): divide before the batched kernel
            src_imgs = [im if dv == 1.0 else im / dv for im, dv in zip(src_imgs, img_divs)]
            img_divs = [1.0] * n
        # ---- transformation branch, step 1 (model/TSNet.py:322-323, :347-348): the masks -> class-sorted order + work
        # list.  Needs only the bboxes, and the rank tables must exist before the last img_enc block writes the sources'
        # correlation operands.
        plan = ops.corr_prepare(tar_bbox.contiguous(), [bb.contiguous() for bb in src_bboxes],
                                self._coord_table(h, w, dev), B, Cf, h, w, m)
        corr_out = None
        if self.bridge and self._tmode3(h, w, Cf, 2 * Cf) == L.TAPS_WINO:
            s_hi = torch.empty((n * B * hw, Cf), dtype=torch.int16, device=dev)
            self._ssq_slabs = Cf // (16 if self.bridge_variant else 32)
            corr_out = dict(hi=s_hi, lo=torch.empty_like(s_hi), rank=plan.rank_s,
                            ssq=torch.empty((n * B, self._ssq_slabs, hw), dtype=torch.float32, device=dev), done=False)
        cache_sig = None
        if src_key is not None:
            cache_sig = (src_key, n, B, tuple(img_divs), m.name, self.winograd, self.bridge, str(self.wino_chunk_kb),
                         self.direct_stem,
                         tuple((p.data_ptr(), p._version) for p in self.nets["img_enc"].parameters()))
        if cache_sig is not None and self._src_cache is not None and self._src_cache[0] == cache_sig:
            src_fea, fuse_taps, cached_ssq = self._src_cache[1]
            corr_out = None   # cached features: their operand rows are re-ranked for this frame's masks below
        else:
            cached_ssq = None
            img_cat = torch.cat(src_imgs, 0) if n > 1 else src_imgs[0]
            lbl_cat = torch.cat(src_lbls, 0) if n > 1 else src_lbls[0]
            src_fea, fuse_taps = self._encoder("img_enc", img_cat.contiguous(), float(img_divs[0]),
                                               lbl_cat.contiguous(), 9, final_tmode=self._tmode3(h, w, Cf, 2 * Cf),
                                               img_mean=img_mean, corr_out=corr_out)       # [n*B, h, w, 512]
            if cache_sig is not None:
                # the per-pixel partial sums of squares are position-indexed (independent of this frame's mask ranks): a
                # later cache hit rebuilds bit-identical reciprocal norms from them
                ssq = corr_out["ssq"] if (corr_out is not None and corr_out["done"]) else None
                self._src_cache = (cache_sig, (src_fea, fuse_taps, ssq))
            else:
                self._src_cache = None
        tar_fea, _ = self._encoder("lbl_enc", None, 1.0, tar_lbl.contiguous(), 0)        # [B, h, w, 512]

        # ---- transformation branch (model/TSNet.py:319-366, 392)
        # prepare: masks -> class-sorted order + work list; operands are written in that order; the tensor-core tiles
        # emit partial softmax states; the finish kernel merges them, gathers the 4 bilinear taps, averages the sources
        # and writes the decoder's map_conv operand directly (torch.cat([pg, sg]) channels [0, 512), model/TSNet.py:163)
        tar_ops = ops.corr_operands(tar_fea.view(B, hw, Cf), m, rank=plan.rank_t)
        if corr_out is not None and corr_out["done"]:
            # the last img_enc bridge pass wrote the sources' operand rows + partial sums of squares
            src_ops = (corr_out["hi"], corr_out["lo"],
                       ops.corr_norms(corr_out["ssq"], n * B, hw, corr_out["ssq"].shape[1], plan.rank_s))
        else:
            src_ops = ops.corr_operands(src_fea.view(n * B, hw, Cf), m, rank=plan.rank_s)
            if cached_ssq is not None:   # cache hit after a bridge-emitting miss: same norms as the miss computed
                src_ops = (src_ops[0], src_ops[1],
                           ops.corr_norms(cached_ssq, n * B, hw, cached_ssq.shape[1], plan.rank_s))
        src_fea_v = src_fea.view(n, B, hw, Cf)
        dec_hi = torch.empty((B, h, w, 2 * Cf), dtype=torch.int16, device=dev)
        dec_lo = torch.empty_like(dec_hi)
        want_align = train is not None and train.get("align", True)
        pg_mean, grids = ops.corr_warp(plan, tar_ops, src_ops, [src_fea_v[i] for i in range(n)], m,
                                       want_grids=return_flow or train is not None,
                                       want_mean=collect is not None or want_align, taps=(dec_hi, dec_lo), c_off=0)

        # ---- synthesis branch: FuseNet on all sources at once (model/TSNet.py:177-200, 396-400).
        # conv1(reflpad(cat[s_i, t])) = W[:, :512] * reflpad(s_i) + W[:, 512:] * reflpad(t): pad and conv are linear, so
        # the target half is evaluated ONCE per frame and added (fp32) in the epilogue of the per-source GEMM.
        # x of "x + conv_block(x)" is cat[src_fea_i, tar_fea]: read through two pointers by the pass that adds it, the
        # 1024-channel concatenation is never written
        cat_act = (src_fea, tar_fea)
        t_taps = ops.build_taps(tar_fea, m, self._tmode3(h, w, Cf, 2 * Cf))
        y_t, _ = self._conv3(t_taps, "fuse_net", "model.0.conv_block.1", B, h, w, norm=False,
                             cin_range=(Cf, 2 * Cf), with_bias=False)                      # [B, h, w, 1024]
        t1, _ = self._conv_in(fuse_taps, ("fuse_net", "model.0.conv_block.1", (0, Cf)), "3x3", n * B, h, w,
                              self._tmode3(h, w, 2 * Cf, 2 * Cf), relu=True, addend=y_t)
        tfo, _ = self._conv_in(t1, ("fuse_net", "model.0.conv_block.5"), "3x3", n * B, h, w, L.TAPS_SAME,
                               residual=cat_act)
        pcf = self._pack("fuse_net", "conv")
        sg, _ = self._conv(tfo, pcf, "1x1", n * B, h, w, norm=False)                     # [n*B, h, w, 512]

        # ---- decoder (model/TSNet.py:128-174): map_conv on cat[pg_mean, mean_i sg_i]
        sg_mean = (torch.empty((B, h, w, Cf), dtype=torch.float32, device=dev)
                   if (collect is not None or want_align) else None)
        ops.build_taps(sg, m, L.TAPS_SAME, taps=(dec_hi, dec_lo), c_off=Cf, avg_n=n, act_out=sg_mean)
        pcm = self._pack("dec", "map_conv")
        x, _ = self._conv((dec_hi, dec_lo, (1, h, w)), pcm, "1x1", B, h, w, norm=False)
        nb = self.n_blocks_dec
        Hc, Wc = h, w
        if nb > 0:
            t = ops.build_taps(x, m, self._tmode3(Hc, Wc, Cf, Cf))
            for blk in range(nb):
                last = blk == nb - 1
                t, x = self._resblock("dec", f"model{blk}.0.", t, x, B, Hc, Wc, Cf,
                                      tmode_out=L.TAPS_UP2REFLECT1 if last else self._tmode3(Hc, Wc, Cf, Cf),
                                      need_act=not last)
        else:
            t = ops.build_taps(x, m, L.TAPS_UP2REFLECT1)
        for i in range(3):
            pc = self._pack("dec", f"model{nb + i}.2")
            Hc, Wc = Hc * 2, Wc * 2
            y, mr = self._conv(t, pc, "3x3", B, Hc, Wc)
            if i < 2:
                t = ops.build_taps(y, m, L.TAPS_UP2REFLECT1, mean_rstd=mr, relu=True)
        # last IN + ReLU (model/TSNet.py:149-150) is applied inside the head kernel's tile loader
        sd = self.nets["dec"].state_dict(keep_vars=True)
        hw_key = f"model{nb + 3}.1"
        fore, fill = (None, None) if pose_fill is None else ((64, 192), pose_fill)
        rec = ops.head_conv_tanh(y, sd[hw_key + ".weight"], sd[hw_key + ".bias"], fore=fore, fill=fill, mean_rstd=mr,
                                 relu=True)

        if train is not None:
            # train-only branches of the reference forward (model/TSNet.py:327-331, 372-390, 402-405)
            src_raw = train.get("src_img_raw", src_imgs)
            src_div = train.get("src_img_div", img_divs)
            fore, fill = (None, None) if pose_fill is None else ((64, 192), pose_fill)
            warp, losses = ops.train_extras(src_raw, src_div, train["tar_img"], 255.0, grids,
                                            pg_mean if want_align else None,
                                            sg_mean.view(B, hw, Cf) if want_align else None, fore=fore, fill=fill)
            train.update(warp=warp, losses=losses)
        if collect is not None:
            collect.update(src_fea=src_fea_v, tar_fea=tar_fea, pg_mean=pg_mean.view(B, h, w, Cf), sg_mean=sg_mean)
        grid_list = [grids[i] for i in range(n)] if return_flow else None
        return rec, grid_list
