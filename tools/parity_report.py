"""Errors of the whole forward against every reference golden for a given kernel configuration (run under gpurun):
   python tools/parity_report.py [--winograd bridge|unfused|off] [--chunk-kb N]
Test infrastructure (reads tests/golden, uses the oracle's synthetic data generator)."""
import argparse, contextlib, io, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import make_golden as MG, synth
from wacv23_tsnet_b200.model.TSNet import TSNet
from wacv23_tsnet_b200.model.TSNet_pose import TSNet as TSNetPose

ap = argparse.ArgumentParser()
ap.add_argument("--winograd", default="bridge")
ap.add_argument("--small-first", dest="small_first", action="store_true")
ap.add_argument("--direct-stem", dest="direct_stem", action="store_true")
ap.add_argument("--chunk-kb", dest="chunk_kb", nargs="*", default=["2"],
                help="ints, or per-net specs like img_enc=2,default=4")
args = ap.parse_args()
wino = {"bridge": True, "unfused": "unfused", "off": False}[args.winograd]
for ck in args.chunk_kb:
    worst_i, worst_g = 0.0, 0.0
    for name, cfg in MG.CONFIGS.items():
        gold = np.load(os.path.join(MG.GOLDEN_DIR, name + ".npz"))
        sds, inputs = MG.build_case(cfg)
        cls = TSNetPose if cfg["pose"] else TSNet
        kw = dict(mean=synth.IMG_MEAN) if cfg["pose"] else dict(return_flow=True)
        with contextlib.redirect_stdout(io.StringIO()):
            net = cls(is_train=False, label_nc=cfg["label_nc"], n_blocks=cfg["n_blocks"], n_downsampling=3,
                      n_source=cfg["n_source"], winograd=wino, **kw)
        net._engine.wino_flags = 8 if args.small_first else 0
        net._engine.direct_stem = args.direct_stem
        net._engine.wino_chunk_kb = (int(ck) if ck.isdigit() else
                                     {kv.split("=")[0]: int(kv.split("=")[1]) for kv in ck.split(",")})
        for k in ("img_enc", "lbl_enc", "fuse_net", "dec"):
            getattr(net, k).load_state_dict({kk: torch.from_numpy(v) for kk, v in sds[k].items()})
        t = torch.from_numpy
        with torch.no_grad():
            net.set_test_input([t(x) for x in inputs["src_img"]], [t(x) for x in inputs["src_lbl"]],
                               [t(x) for x in inputs["src_bbox"]], t(inputs["tar_lbl"]), t(inputs["tar_bbox"]))
            net.forward()
        diff = (net.rec_tar_img.cpu() - t(gold["rec_tar_img"])).abs()
        ie = float(diff.max())
        imean, ip = float(diff.mean()), float(torch.quantile(diff.flatten()[::7].double(), 0.9999))
        ge = 0.0 if cfg["pose"] else float((torch.stack(net.warp_grid2d_list).cpu() - t(gold["grids"])).abs().max())
        worst_i, worst_g = max(worst_i, ie), max(worst_g, ge)
        print(f"winograd={args.winograd} small_first={args.small_first} chunk_kb={ck} {name}: img {ie:.3e} (mean {imean:.2e}, p99.99 {ip:.2e}) grid {ge:.3e}", flush=True)
    print(f"winograd={args.winograd} small_first={args.small_first} chunk_kb={ck} WORST: img {worst_i:.3e} (tol 1e-3) grid {worst_g:.3e} (tol 5e-5)", flush=True)
