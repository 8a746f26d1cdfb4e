"""Staged on-device bring-up checks (run under gpurun).  `python tools/gpu_diag.py` runs every stage in its own
subprocess (a device trap in one stage must not poison the others) and writes gpurun_out/diag.log.
Not a test-suite replacement: tests/ holds the real parity tests; this prints more detail when something is off.
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ["elementwise", "gemm1x1", "gemm_shapes", "stats", "corr", "e2e"]


def _setup():
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    return torch


def _recon(hi, lo, fmt):
    import torch
    dt = torch.float16 if fmt == 0 else torch.bfloat16
    return hi.view(dt).float() + lo.view(dt).float()


def _report(name, got, ref, tol):
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    ok = err <= tol * max(scale, 1e-30)
    print(f"  {name}: max|err| {err:.3e} (ref max {scale:.3e}, rel {err / max(scale, 1e-30):.3e}) {'OK' if ok else 'FAIL'}",
          flush=True)
    return ok


def stage_elementwise():
    torch = _setup()
    import torch.nn.functional as F
    from wacv23_tsnet_b200 import ops, lib as L
    ok = True
    dev = "cuda"
    for name in ("fp16x3", "bf16x3"):
        m = ops.MathMode(name)
        tol = 2e-6 if m.fmt == 0 else 1e-4
        x = torch.randn(2, 16, 16, 64, device=dev) * 2
        # direct conv (validation kernel) vs torch
        w = torch.randn(32, 64, 3, 3, device=dev) * 0.05
        b = torch.randn(32, device=dev)
        y = ops.direct_conv_fp32(x, w, b, stride=1, pad=1, reflect=True)
        ref = F.conv2d(F.pad(x.permute(0, 3, 1, 2), (1, 1, 1, 1), mode="reflect"), w, b).permute(0, 2, 3, 1)
        ok &= _report(f"[{name}] direct_conv reflect", y, ref, 1e-5)
        y = ops.direct_conv_fp32(x, w, b, stride=2, pad=1, reflect=False)
        ref = F.conv2d(x.permute(0, 3, 1, 2), w, b, stride=2, padding=1).permute(0, 2, 3, 1)
        ok &= _report(f"[{name}] direct_conv s2", y, ref, 1e-5)
        # build_taps modes
        xn = x.permute(0, 3, 1, 2)
        hi, lo, g = ops.build_taps(x, m, L.TAPS_REFLECT1)
        ref = F.pad(xn, (1, 1, 1, 1), mode="reflect").permute(0, 2, 3, 1) * m.act_scale
        ok &= _report(f"[{name}] taps reflect1", _recon(hi, lo, m.fmt), ref, tol)
        hi, lo, g = ops.build_taps(x, m, L.TAPS_UP2REFLECT1)
        up = F.interpolate(xn, scale_factor=2, mode="bilinear", align_corners=False)
        ref = F.pad(up, (1, 1, 1, 1), mode="reflect").permute(0, 2, 3, 1) * m.act_scale
        ok &= _report(f"[{name}] taps up2+reflect1", _recon(hi, lo, m.fmt), ref, tol)
        hi, lo, g = ops.build_taps(x, m, L.TAPS_S2ZERO)
        xp = F.pad(xn, (1, 1, 1, 1)).permute(0, 2, 3, 1)  # [B,18,18,C]
        planes = torch.stack([xp[:, py::2, px::2][:, :9, :9] for py in (0, 1) for px in (0, 1)], 1)
        # plane (py,px)[i][j] = padded[2i+py][2j+px]; for odd parity the 9th row/col is beyond the pad (zero)
        refp = torch.zeros(2, 4, 9, 9, 64, device=dev)
        for k, (py, px) in enumerate([(0, 0), (0, 1), (1, 0), (1, 1)]):
            sub = xp[:, py::2, px::2]
            refp[:, k, :sub.shape[1], :sub.shape[2]] = sub
        ok &= _report(f"[{name}] taps s2 planes", _recon(hi, lo, m.fmt).view(2, 4, 9, 9, 64), refp * m.act_scale, tol)
        # norm + relu + residual + act_out
        mr = torch.stack([x.mean((1, 2)), 1.0 / torch.sqrt(x.var((1, 2), unbiased=False) + 1e-5)], -1).contiguous()
        res = torch.randn_like(x)
        act = torch.empty_like(x)
        hi, lo, g = ops.build_taps(x, m, L.TAPS_REFLECT1, mean_rstd=mr, relu=True, residual=res, act_out=act)
        refa = F.relu(F.instance_norm(xn, eps=1e-5)).permute(0, 2, 3, 1) + res
        ok &= _report(f"[{name}] norm+relu+res act_out", act, refa, 1e-5)
        # avg_n
        x3 = torch.randn(6, 8, 8, 64, device=dev)
        a3 = torch.empty(2, 8, 8, 64, device=dev)
        ops.build_taps(x3, m, L.TAPS_SAME, avg_n=3, act_out=a3, want_taps=False)
        ok &= _report(f"[{name}] avg_n", a3, x3.view(3, 2, 8, 8, 64).mean(0), 1e-6)
        # stem taps
        img = torch.rand(2, 3, 32, 32, device=dev) * 255 - 100
        lbl = (torch.rand(2, 2, 32, 32, device=dev) > 0.5).float()
        hi, lo, g = ops.stem_taps(img, 255.0, lbl, 64, m)
        from oracle import tsnet_oracle as O
        full = O.coord_channels(torch.cat([img.cpu() / 255.0, lbl.cpu()], 1)).to(dev)  # [2,8,32,32]
        fp = F.pad(full, (3, 3, 3, 3), mode="reflect")  # [2,8,38,38]
        ref = torch.zeros(2, 38, 32, 64, device=dev)
        for s in range(7):
            ref[..., s * 8:(s + 1) * 8] = fp[:, :, :, s:s + 32].permute(0, 2, 3, 1)
        ok &= _report(f"[{name}] stem taps", _recon(hi, lo, m.fmt), ref * m.act_scale, tol)
        # l2norm
        f = torch.randn(2, 64, 512, device=dev)
        f[0, 3] = 0
        hi, lo = ops.l2norm_split(f, m)
        ok &= _report(f"[{name}] l2norm", _recon(hi, lo, m.fmt).view(2, 64, 512), F.normalize(f, dim=2) * m.corr_scale, tol)
        # head conv
        a = torch.randn(1, 32, 32, 64, device=dev)
        wh = torch.randn(3, 64, 7, 7, device=dev) * 0.02
        bh = torch.randn(3, device=dev) * 0.1
        out = ops.head_conv_tanh(a, wh, bh)
        ref = torch.tanh(F.conv2d(F.pad(a.permute(0, 3, 1, 2), (3, 3, 3, 3), mode="reflect"), wh, bh))
        ok &= _report(f"[{name}] head conv", out, ref, 1e-5)
    return ok


def _gemm_case(torch, name, B, H, W, Cin, Cout, kind, mode_name, block_n=None, verbose=False):
    import torch.nn.functional as F
    from wacv23_tsnet_b200 import ops, lib as L
    m = ops.MathMode(mode_name)
    dev = "cuda"
    x = torch.randn(B, H, W, Cin, device=dev)
    xn = x.permute(0, 3, 1, 2)
    if kind == "1x1":
        w = torch.randn(Cout, Cin, 1, 1, device=dev) * 0.05
        tm, Ho, Wo = L.TAPS_SAME, H, W
        ref = F.conv2d(xn, w)
    elif kind == "3x3":
        w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05
        tm, Ho, Wo = L.TAPS_REFLECT1, H, W
        ref = F.conv2d(F.pad(xn, (1, 1, 1, 1), mode="reflect"), w)
    elif kind == "3x3s2":
        w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05
        tm, Ho, Wo = L.TAPS_S2ZERO, H // 2, W // 2
        ref = F.conv2d(xn, w, stride=2, padding=1)
    elif kind == "up3x3":
        w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05
        tm, Ho, Wo = L.TAPS_UP2REFLECT1, 2 * H, 2 * W
        up = F.interpolate(xn, scale_factor=2, mode="bilinear", align_corners=False)
        ref = F.conv2d(F.pad(up, (1, 1, 1, 1), mode="reflect"), w)
    b = torch.randn(Cout, device=dev)
    ref = (ref + b.view(1, -1, 1, 1)).permute(0, 2, 3, 1).contiguous()
    pc = ops.PackedConv(w, b, m, block_n=block_n)
    hi, lo, g = ops.build_taps(x, m, tm)
    y, stats = ops.conv_gemm(hi, lo, g, pc, "3x3" if kind == "up3x3" else kind, B, Ho, Wo, m, m.act_scale)
    torch.cuda.synchronize()
    tol = {"fp16x3": 2e-6, "bf16x3": 5e-5, "fp16": 2e-3, "bf16": 2e-2}[mode_name]
    ok = _report(f"{name} {kind} B{B} {H}x{W} {Cin}->{Cout} bn={pc.block_n or 'auto'} {mode_name}", y, ref, tol)
    if not ok and verbose:
        d = (y - ref).abs()
        print("    err by row-in-tile (first 8 of 128):", d.view(-1, 128, Cout).amax((0, 2))[:8].tolist())
        print("    err by col (first 8):", d.view(-1, Cout).amax(0)[:8].tolist())
        print("    y[0,0,0,:4]", y[0, 0, 0, :4].tolist(), "ref", ref[0, 0, 0, :4].tolist())
        nz = (y != 0).float().mean().item()
        print(f"    nonzero fraction of y: {nz:.3f}")
    # statistics
    mr = ops.instnorm_reduce(stats, B, Ho * Wo, Cout)
    torch.cuda.synchronize()
    rm = ref.mean((1, 2))
    rr = 1.0 / torch.sqrt(ref.var((1, 2), unbiased=False) + 1e-5)
    ok &= _report(f"{name}   mean", mr[..., 0], rm, max(tol * 20, 2e-5))
    ok &= _report(f"{name}   rstd", mr[..., 1], rr, max(tol * 20, 2e-5))
    return ok


def stage_gemm1x1():
    torch = _setup()
    ok = True
    for mode in ("fp16", "fp16x3", "bf16x3"):
        ok &= _gemm_case(torch, "g", 1, 32, 32, 64, 64, "1x1", mode, verbose=True)
    ok &= _gemm_case(torch, "g", 2, 32, 32, 128, 128, "1x1", "fp16x3", verbose=True)
    ok &= _gemm_case(torch, "g", 2, 32, 32, 256, 256, "1x1", "fp16x3", verbose=True)
    return ok


def stage_gemm_shapes():
    torch = _setup()
    ok = True
    ok &= _gemm_case(torch, "g", 2, 32, 32, 128, 128, "3x3", "fp16x3", verbose=True)
    ok &= _gemm_case(torch, "g", 3, 32, 32, 512, 512, "3x3", "fp16x3", verbose=True)
    ok &= _gemm_case(torch, "g", 3, 32, 32, 512, 512, "3x3", "fp16x3", block_n=128)
    ok &= _gemm_case(torch, "g", 2, 64, 64, 64, 128, "3x3s2", "fp16x3", verbose=True)
    ok &= _gemm_case(torch, "g", 2, 256, 256, 64, 128, "3x3s2", "fp16x3", verbose=True)
    ok &= _gemm_case(torch, "g", 1, 32, 32, 512, 256, "up3x3", "fp16x3", verbose=True)
    ok &= _gemm_case(torch, "g", 1, 128, 128, 128, 64, "up3x3", "fp16x3", verbose=True)
    ok &= _gemm_case(torch, "g", 1, 32, 32, 1024, 512, "1x1", "fp16x3")
    ok &= _gemm_case(torch, "g", 150, 32, 32, 64, 64, "1x1", "fp16x3")  # > 148 CTAs worth of tiles, persistent loop
    ok &= _gemm_case(torch, "g", 2, 32, 32, 128, 128, "3x3", "bf16x3")
    return ok


def stage_stats():
    return True


def stage_corr():
    torch = _setup()
    import numpy as np
    from wacv23_tsnet_b200 import ops
    from oracle import tsnet_oracle as O
    ok = True
    dev = "cuda"
    for (B, n, rect, mode) in [(1, 1, False, "fp16x3"), (2, 3, True, "fp16x3"), (2, 3, False, "bf16x3")]:
        m = ops.MathMode(mode)
        g = torch.Generator().manual_seed(5 + B + n)
        tar = torch.relu(torch.randn(B, 512, 32, 32, generator=g))
        tar[0, :, 3, 5] = 0  # all-zero target vector -> uniform softmax row
        srcs = [torch.randn(B, 512, 32, 32, generator=g) * 3 for _ in range(n)]
        if rect:
            tb = torch.zeros(B, 1, 256, 256, dtype=torch.uint8)
            tb[:, :, 40:200, 30:220] = 1
            sbs = []
            for i in range(n):
                sb = torch.zeros(B, 1, 256, 256, dtype=torch.uint8)
                sb[:, :, 20 + 10 * i:180, 50:230 - 10 * i] = 1
                sbs.append(sb)
        else:
            tb = torch.randint(0, 2, (B, 1, 256, 256), generator=g).float()
            sbs = [torch.randint(0, 2, (B, 1, 256, 256), generator=g).float() for _ in range(n)]
        ref_mean, ref_grids = O.corr_warp(tar, srcs, tb, sbs)
        tru_mean, tru_grids = O.corr_warp(tar.double(), [s.double() for s in srcs], tb, sbs)
        tar_d = tar.permute(0, 2, 3, 1).contiguous().to(dev)
        src_d = torch.stack([s.permute(0, 2, 3, 1).contiguous() for s in srcs]).to(dev)  # [n,B,h,w,C]
        coord = torch.cat([torch.linspace(-1, 1, 32), torch.linspace(-1, 1, 32)]).to(dev)
        out, grids = ops.corr_chain(tar_d.view(B, 1024, 512), src_d.view(n, B, 1024, 512),
                                    tb.squeeze(1).contiguous().to(dev), [s.squeeze(1).contiguous().to(dev) for s in sbs],
                                    coord, m, want_grids=True, want_mean=True)
        torch.cuda.synchronize()
        gerr = max((grids[i].cpu() - ref_grids[i]).abs().max().item() for i in range(n))
        kerr = max((grids[i].cpu().double() - tru_grids[i]).abs().max().item() for i in range(n))
        rerr = max((ref_grids[i].double() - tru_grids[i]).abs().max().item() for i in range(n))
        print(f"  corr B{B} n{n} rect{rect} {mode}: grid max|err| vs fp32 oracle {gerr:.3e}; vs fp64 truth: kernel {kerr:.3e}, "
              f"fp32 oracle {rerr:.3e}", flush=True)
        ok &= gerr < (2e-5 if mode == "fp16x3" else 5e-4)
        ok &= _report(f"corr B{B} n{n} {mode} warped mean", out.view(B, 32, 32, 512).permute(0, 3, 1, 2).cpu(), ref_mean,
                      1e-4 if mode == "fp16x3" else 3e-3)
    return ok


def stage_e2e():
    torch = _setup()
    import numpy as np
    from oracle import make_golden as MG, synth
    from wacv23_tsnet_b200.model.TSNet import TSNet
    from wacv23_tsnet_b200.model.TSNet_pose import TSNet as TSNetPose
    ok = True
    for name in ("quickstart_bs1", "face_bs1_nb4", "pose_bs1_nb4"):
        cfg = MG.CONFIGS[name]
        gold = np.load(os.path.join(MG.GOLDEN_DIR, name + ".npz"))
        sds, inputs = MG.build_case(cfg)
        assert (MG.case_checksums(sds, inputs) == gold["checks"]).all(), "synthetic data differs from the fixture's"
        cls = TSNetPose if cfg["pose"] else TSNet
        kw = dict(mean=synth.IMG_MEAN) if cfg["pose"] else dict(return_flow=True)
        net = cls(is_train=False, label_nc=cfg["label_nc"], n_blocks=cfg["n_blocks"], n_downsampling=3,
                  n_source=cfg["n_source"], **kw)
        for k in ("img_enc", "lbl_enc", "fuse_net", "dec"):
            getattr(net, k).load_state_dict({kk: torch.from_numpy(v) for kk, v in sds[k].items()})
        net.eval()
        net.set_test_input([torch.from_numpy(x) for x in inputs["src_img"]], [torch.from_numpy(x) for x in inputs["src_lbl"]],
                           [torch.from_numpy(x) for x in inputs["src_bbox"]], torch.from_numpy(inputs["tar_lbl"]),
                           torch.from_numpy(inputs["tar_bbox"]))
        col = {}
        t0 = time.time()
        net.forward(_collect=col)
        torch.cuda.synchronize()
        print(f"  {name}: forward {time.time() - t0:.3f}s", flush=True)
        nhwc = lambda t: t.permute(0, 3, 1, 2).cpu()
        ok &= _report(f"{name} tar_fea", nhwc(col["tar_fea"])[:, ::16], torch.from_numpy(gold["tar_fea_c16"]), 1e-4)
        ok &= _report(f"{name} src_fea0", nhwc(col["src_fea"][0].view(-1, 32, 32, 512))[:, ::16],
                      torch.from_numpy(gold["src_fea0_c16"]), 1e-4)
        if not cfg["pose"]:
            g = torch.stack(net.warp_grid2d_list).cpu()
            ok &= _report(f"{name} grids", g, torch.from_numpy(gold["grids"]), 2e-4)
        ok &= _report(f"{name} pg_mean", nhwc(col["pg_mean"])[:, ::8], torch.from_numpy(gold["pg_mean_c8"]), 2e-3)
        ok &= _report(f"{name} sg_mean", nhwc(col["sg_mean"])[:, ::8], torch.from_numpy(gold["sg_mean_c8"]), 2e-3)
        err = (net.rec_tar_img.cpu() - torch.from_numpy(gold["rec_tar_img"])).abs().max().item()
        print(f"  {name} rec_tar_img max|err| {err:.3e} {'OK' if err < 2e-3 else 'FAIL'}", flush=True)
        ok &= err < 2e-3
    return ok


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--stage":
        ok = globals()["stage_" + sys.argv[2]]()
        print(f"STAGE {sys.argv[2]}: {'PASS' if ok else 'FAIL'}", flush=True)
        sys.exit(0 if ok else 1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "diag.log"), "w")
    stages = sys.argv[1:] if len(sys.argv) > 1 else STAGES
    for st in stages:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", st], capture_output=True, text=True,
                               timeout=600, cwd=ROOT)
            out = p.stdout + ("\n[stderr]\n" + p.stderr[-6000:] if p.returncode != 0 else "")
            rc = p.returncode
        except subprocess.TimeoutExpired as e:
            out, rc = f"TIMEOUT\n{e.stdout}\n{e.stderr}", -9
        msg = f"===== {st} rc={rc} [{time.time() - t0:.1f}s]\n{out}\n"
        print(msg, flush=True)
        log.write(msg)
        log.flush()


if __name__ == "__main__":
    main()
