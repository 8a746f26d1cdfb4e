"""Pin the oracle against the real reference and write tests/golden/*.npz.

Run in the BUILD CONTAINER only (needs /root/reference):   python -m oracle.make_golden

For every config below it
  1. builds deterministic weights / inputs with oracle/synth.py,
  2. runs the UNMODIFIED reference forward on CPU (oracle/ref_harness.py),
  3. runs the restatement oracle/tsnet_oracle.py on the same data and requires BIT-EXACT agreement
     of rec_tar_img and of every warp grid with the reference (torch.equal),
  4. stores the reference outputs + subsampled oracle intermediates as the fixture.
Fixtures carry checksums of the synthetic weights / inputs so a host that generates different
synthetic data is detected loudly instead of failing parity mysteriously.
"""
import os
import sys
import time

import numpy as np
import torch

from . import ref_harness, synth, tsnet_oracle

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> config.  (config 1 = quick_start parity gate; 2 = FaceForensics; 3 = pose;
# 5 = n_source sweep members.)
CONFIGS = {
    "quickstart_bs1": dict(kind="qs", bs=1, label_nc=2, n_blocks=0, n_source=3, pose=False, bias_std=0.0),
    "face_bs1_nb4": dict(kind="ds", bs=1, label_nc=2, n_blocks=4, n_source=3, pose=False, bias_std=0.05),
    "pose_bs1_nb4": dict(kind="ds", bs=1, label_nc=25, n_blocks=4, n_source=3, pose=True, bias_std=0.05),
    "face_bs2_n1": dict(kind="ds", bs=2, label_nc=2, n_blocks=0, n_source=1, pose=False, bias_std=0.0),
    "face_bs1_n5": dict(kind="qs", bs=1, label_nc=2, n_blocks=0, n_source=5, pose=False, bias_std=0.0),
    # round 2: the top point of the n_source sweep (BASELINE.json config 5) and a batched FaceForensics forward
    "face_bs1_n8": dict(kind="ds", bs=1, label_nc=2, n_blocks=0, n_source=8, pose=False, bias_std=0.05),
    "face_bs2_nb4_n3": dict(kind="ds", bs=2, label_nc=2, n_blocks=4, n_source=3, pose=False, bias_std=0.05),
}


def build_case(cfg, seed=1234):
    sds = synth.make_state_dicts(cfg["label_nc"], cfg["n_blocks"], seed=seed, bias_std=cfg["bias_std"])
    if cfg["kind"] == "qs":
        inputs = synth.quick_start_inputs(cfg["bs"], cfg["label_nc"], cfg["n_source"], seed=seed)
    else:
        inputs = synth.dataset_like_inputs(cfg["bs"], cfg["label_nc"], cfg["n_source"], seed=seed, pose=cfg["pose"])
    return sds, inputs


def case_checksums(sds, inputs):
    ws = [synth.checksum(sds[n][k]) for n in sorted(sds) for k in sorted(sds[n])]
    xs = [synth.checksum(a) for key in ("src_img", "src_lbl", "src_bbox") for a in inputs[key]]
    xs += [synth.checksum(inputs[k]) for k in ("tar_lbl", "tar_bbox")]
    fold = lambda v: synth.checksum(np.array(v, dtype=np.uint64)) & 0x7FFFFFFFFFFFFFFF
    return np.array([fold(ws), fold(xs)], np.int64)


# train-mode forward branches: name -> (base config, use_prev)
TRAIN_CONFIGS = {
    "train_quickstart_bs1": ("quickstart_bs1", None),
    "train_pose_bs1_nb4": ("pose_bs1_nb4", None),
    "train_face_bs2_n1_useprev": ("face_bs2_n1", [True]),
}


def make_train_goldens():
    """Reference forward with the is_train branches active (see ref_harness.reference_forward) vs the restatement:
    bit-exact warp_src_img_list / loss_warp / loss_align required; fixtures store the warped images at half
    resolution plus the losses."""
    for name, (base, use_prev) in TRAIN_CONFIGS.items():
        t0 = time.time()
        cfg = CONFIGS[base]
        sds, inputs = build_case(cfg)
        mean = synth.IMG_MEAN if cfg["pose"] else None
        ref = ref_harness.reference_forward(sds, inputs, cfg["label_nc"], cfg["n_blocks"], pose=cfg["pose"],
                                            pose_mean=mean, n_source=cfg["n_source"], train=True, use_prev=use_prev)
        ora = tsnet_oracle.tsnet_forward(sds, inputs, cfg["n_blocks"], pose_mean=mean, train=True, use_prev=use_prev)
        ok = torch.equal(ref["rec_tar_img"], ora["rec_tar_img"])
        ok &= all(torch.equal(a, b) for a, b in zip(ref["warp_src_img_list"], ora["warp_src_img_list"]))
        ok &= torch.equal(ref["loss_warp"], ora["loss_warp"])
        if not cfg["pose"]:
            ok &= torch.equal(ref["loss_align"], ora["loss_align"])
        print(f"{name}: oracle==reference (image, warped sources, losses) {ok}  [{time.time()-t0:.1f}s]")
        if not ok:
            sys.exit(f"oracle restatement of the train-mode branches is NOT bit-exact with the reference on {name}")
        np.savez_compressed(
            os.path.join(GOLDEN_DIR, name + ".npz"),
            rec_tar_img_s2=ref["rec_tar_img"][..., ::2, ::2].numpy(),
            warp_s2=torch.stack(ref["warp_src_img_list"])[..., ::2, ::2].numpy(),
            loss_warp=ref["loss_warp"].numpy(),
            loss_align=(ref["loss_align"].numpy() if not cfg["pose"] else np.float32(0)),
            checks=case_checksums(sds, inputs),
        )


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "train":
        if not ref_harness.available():
            sys.exit("reference not found: goldens can only be (re)generated in the build container")
        torch.set_num_threads(os.cpu_count())
        return make_train_goldens()
    if not ref_harness.available():
        sys.exit("reference not found: goldens can only be (re)generated in the build container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    only = set(sys.argv[1:])   # optional: names of the configs to (re)generate
    for name, cfg in CONFIGS.items():
        if only and name not in only:
            continue
        t0 = time.time()
        sds, inputs = build_case(cfg)
        mean = synth.IMG_MEAN if cfg["pose"] else None
        ref = ref_harness.reference_forward(sds, inputs, cfg["label_nc"], cfg["n_blocks"], pose=cfg["pose"],
                                            pose_mean=mean, n_source=cfg["n_source"])
        ora = tsnet_oracle.tsnet_forward(sds, inputs, cfg["n_blocks"], pose_mean=mean)
        ok_img = torch.equal(ref["rec_tar_img"], ora["rec_tar_img"])
        ok_grid = all(torch.equal(a, b) for a, b in zip(ref.get("grids", []), ora["grids"]))
        d = (ref["rec_tar_img"] - ora["rec_tar_img"]).abs().max().item()
        print(f"{name}: oracle==reference image {ok_img} (max|d|={d:.3g}) grids {ok_grid}  [{time.time()-t0:.1f}s]")
        if not (ok_img and ok_grid):
            sys.exit(f"oracle restatement is NOT bit-exact with the reference on {name}")
        np.savez_compressed(
            os.path.join(GOLDEN_DIR, name + ".npz"),
            rec_tar_img=ref["rec_tar_img"].numpy(),
            grids=torch.stack(ora["grids"]).numpy(),
            pg_mean_c8=ora["pg_mean"][:, ::8].numpy(),
            sg_mean_c8=ora["sg_mean"][:, ::8].numpy(),
            tar_fea_c16=ora["tar_fea"][:, ::16].numpy(),
            src_fea0_c16=ora["src_fea"][0][:, ::16].numpy(),
            checks=case_checksums(sds, inputs),
        )


if __name__ == "__main__":
    main()
