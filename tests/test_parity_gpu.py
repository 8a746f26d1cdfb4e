"""GPU suite, part 2: whole-forward parity of the drop-in TSNet classes (through the public class surface and the
C ABI underneath) against the reference outputs stored in tests/golden/, plus size-independent properties at the
benchmark's full batch size."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# Stated tolerances (DESIGN.md "precision"): the reference CPU fp32 forward itself sits 3e-4..5e-4 (image) and
# 1.5e-5 (grids) away from an fp64 evaluation of the same network -- random-init weights + softmax(100 x) amplify
# rounding noise by ~1e3.  We require agreement with the reference within IMG_TOL / GRID_TOL below.
# Measured on B200 (fp16x3): image 3.7e-4 .. 5.8e-4, grids 1.8e-5, features 2.9e-6 rel, pg_mean 2.4e-4 rel.
# IMG_TOL: the reference's own fp32 forward is 3.2e-4 .. 5.1e-4 (max-abs, per golden) away from an fp64 evaluation of
# itself, and the max over 4e5 pixels of the difference between two fp32-level evaluations is a noisy statistic: every
# bit-different but equally valid kernel configuration measured on the B200 lands between 8.6e-4 and 1.0e-3 on the worst
# golden (direct GEMM 9.3e-4, Winograd chunk 2 8.7e-4, + un-normalised correlation operands 1.0e-3, chunk 3 9.3e-4;
# tools/parity_report.py).  2.5x the reference's noise floor keeps a real regression (chunk 8: 1.3e-3, TF32: 0.3) out.
IMG_TOL = 1.25e-3   # max-abs on rec_tar_img, tanh range (-1, 1)
IMG_MEAN_TOL = 1.5e-4   # mean-abs on rec_tar_img: the robust companion of the max (measured 2.3e-5 .. 6.5e-5 per golden)
GRID_TOL = 5e-5     # max-abs on warp grids, [-1, 1] units (= 8e-4 feature pixels); reference fp32-vs-fp64: 1.4e-5
FEA_TOL = 2e-5      # encoder features, relative to max|ref|
MIX_TOL = 1e-3      # pg_mean / sg_mean, relative to max|ref|


def _build(name, math_mode="fp16x3"):
    from oracle import make_golden as MG, synth
    from wacv23_tsnet_b200.model.TSNet import TSNet
    from wacv23_tsnet_b200.model.TSNet_pose import TSNet as TSNetPose
    cfg = MG.CONFIGS[name]
    gold = np.load(os.path.join(MG.GOLDEN_DIR, name + ".npz"))
    sds, inputs = MG.build_case(cfg)
    assert (MG.case_checksums(sds, inputs) == gold["checks"]).all(), "synthetic data differs from the fixture's"
    cls = TSNetPose if cfg["pose"] else TSNet
    kw = dict(mean=synth.IMG_MEAN) if cfg["pose"] else dict(return_flow=True)
    net = cls(is_train=False, label_nc=cfg["label_nc"], n_blocks=cfg["n_blocks"], n_downsampling=3,
              n_source=cfg["n_source"], math_mode=math_mode, **kw)
    for k in ("img_enc", "lbl_enc", "fuse_net", "dec"):
        getattr(net, k).load_state_dict({kk: torch.from_numpy(v) for kk, v in sds[k].items()})
    net.eval()
    return cfg, gold, inputs, net


def _feed(net, inputs, sl=slice(None)):
    net.set_test_input([torch.from_numpy(x[sl]) for x in inputs["src_img"]],
                       [torch.from_numpy(x[sl]) for x in inputs["src_lbl"]],
                       [torch.from_numpy(x[sl]) for x in inputs["src_bbox"]],
                       torch.from_numpy(inputs["tar_lbl"][sl]), torch.from_numpy(inputs["tar_bbox"][sl]))


@pytest.mark.parametrize("name", ["quickstart_bs1", "face_bs1_nb4", "pose_bs1_nb4", "face_bs2_n1", "face_bs1_n5",
                                  "face_bs1_n8", "face_bs2_nb4_n3"])
def test_forward_matches_reference_golden(name):
    cfg, gold, inputs, net = _build(name)
    _feed(net, inputs)
    col = {}
    with torch.no_grad():
        net.forward(_collect=col)
    torch.cuda.synchronize()
    nchw = lambda t: t.permute(0, 3, 1, 2).cpu()
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    assert rel(nchw(col["tar_fea"])[:, ::16], torch.from_numpy(gold["tar_fea_c16"])) < FEA_TOL
    assert rel(nchw(col["src_fea"][0].reshape(-1, 32, 32, 512))[:, ::16], torch.from_numpy(gold["src_fea0_c16"])) < FEA_TOL
    assert rel(nchw(col["pg_mean"])[:, ::8], torch.from_numpy(gold["pg_mean_c8"])) < MIX_TOL
    assert rel(nchw(col["sg_mean"])[:, ::8], torch.from_numpy(gold["sg_mean_c8"])) < MIX_TOL
    if not cfg["pose"]:
        g = torch.stack(net.warp_grid2d_list).cpu()
        assert g.shape == gold["grids"].shape
        assert float((g - torch.from_numpy(gold["grids"])).abs().max()) < GRID_TOL
    out = net.rec_tar_img
    assert out.is_cuda and out.dtype == torch.float32 and tuple(out.shape) == (cfg["bs"], 3, 256, 256)
    assert float((out.cpu() - torch.from_numpy(gold["rec_tar_img"])).abs().max()) < IMG_TOL
    assert float((out.cpu() - torch.from_numpy(gold["rec_tar_img"])).abs().mean()) < IMG_MEAN_TOL
    if cfg["pose"]:  # compositing is exact outside the foreground columns
        assert torch.equal(out.cpu()[..., :64], torch.from_numpy(gold["rec_tar_img"])[..., :64])


@pytest.mark.parametrize("name", ["train_quickstart_bs1", "train_pose_bs1_nb4", "train_face_bs2_n1_useprev"])
def test_train_mode_forward_branches_match_reference_golden(name):
    """is_train=True forward (set_train_input + forward): rec_tar_img, warp_src_img_list
    (unfold -> grid_sample -> fold image warp + re-normalisation [+ pose compositing]), loss_warp, loss_align against
    the reference fixture (tests/golden/train_*.npz, written by `python -m oracle.make_golden train`)."""
    from oracle import make_golden as MG, synth
    from wacv23_tsnet_b200.model.TSNet import TSNet
    from wacv23_tsnet_b200.model.TSNet_pose import TSNet as TSNetPose
    base, use_prev = MG.TRAIN_CONFIGS[name]
    cfg = MG.CONFIGS[base]
    gold = np.load(os.path.join(MG.GOLDEN_DIR, name + ".npz"))
    sds, inputs = MG.build_case(cfg)
    assert (MG.case_checksums(sds, inputs) == gold["checks"]).all()
    cls = TSNetPose if cfg["pose"] else TSNet
    kw = dict(mean=synth.IMG_MEAN) if cfg["pose"] else {}
    net = cls(is_train=True, label_nc=cfg["label_nc"], n_blocks=cfg["n_blocks"], n_downsampling=3,
              n_source=cfg["n_source"], **kw)
    for k in ("img_enc", "lbl_enc", "fuse_net", "dec"):
        getattr(net, k).load_state_dict({kk: torch.from_numpy(v) for kk, v in sds[k].items()})
    net.set_train_input([torch.from_numpy(x) for x in inputs["src_img"]],
                        [torch.from_numpy(x) for x in inputs["src_lbl"]],
                        [torch.from_numpy(x) for x in inputs["src_bbox"]], torch.from_numpy(inputs["tar_img"]),
                        torch.from_numpy(inputs["tar_lbl"]), torch.from_numpy(inputs["tar_bbox"]), use_prev=use_prev)
    with torch.no_grad():
        net.forward()
    torch.cuda.synchronize()
    assert float((net.rec_tar_img.cpu()[..., ::2, ::2] - torch.from_numpy(gold["rec_tar_img_s2"])).abs().max()) < IMG_TOL
    ref = torch.from_numpy(gold["warp_s2"])
    got = torch.stack(net.warp_src_img_list).cpu()[..., ::2, ::2]
    assert got.shape == ref.shape
    # the warped image moves with the warp grid: 2e-5 grid units = 3e-4 cells of an image whose neighbouring 8x8 patches
    # are unrelated (random inputs) -> a few 1e-4 of the value range
    assert float((got - ref).abs().max()) < 2e-3 * float(ref.abs().max())
    assert abs(float(net.loss_warp) - float(gold["loss_warp"])) < 1e-3 * abs(float(gold["loss_warp"]))
    if cfg["pose"]:  # compositing is exact outside the foreground columns
        assert torch.equal(got[..., :32], ref[..., :32])
    else:
        assert abs(float(net.loss_align) - float(gold["loss_align"])) < 1e-4


def test_demo_call_sequence_5d_lists_uint8_bbox_and_source_count():
    """demo/demo_face.py:170-194 style: 5-D tensors used as lists, uint8 bboxes, set_source_num, .data.cpu()."""
    cfg, gold, inputs, net = _build("face_bs1_nb4")
    imgs = torch.from_numpy(np.stack(inputs["src_img"]))        # [n, 1, 3, 256, 256]
    lbls = torch.from_numpy(np.stack(inputs["src_lbl"]))
    bbs = torch.from_numpy(np.stack(inputs["src_bbox"]))        # uint8
    assert bbs.dtype == torch.uint8
    with torch.no_grad():
        net.set_test_input(imgs, lbls, bbs, torch.from_numpy(inputs["tar_lbl"]), torch.from_numpy(inputs["tar_bbox"]))
        net.forward()
        a = net.rec_tar_img.data.cpu()
        assert float((a - torch.from_numpy(gold["rec_tar_img"])).abs().max()) < IMG_TOL
        # float masks with the same content: identical bits
        net.set_test_input(imgs, lbls, bbs.float(), torch.from_numpy(inputs["tar_lbl"]),
                           torch.from_numpy(inputs["tar_bbox"]).float())
        net.forward()
        assert torch.equal(net.rec_tar_img.data.cpu(), a)
        # fewer sources through set_source_num (model/TSNet.py:296): uses the first n of the staged lists
        net.set_source_num(1)
        net.forward()
        assert tuple(net.rec_tar_img.shape) == (1, 3, 256, 256) and len(net.warp_grid2d_list) == 1
        assert not torch.equal(net.rec_tar_img.data.cpu(), a)


def test_source_feature_cache_is_exact_and_detects_changes():
    """Opt-in cache for the demo loop (demo/demo_face.py:170-192 re-feeds the same source frames for every driving
    frame): a hit skips img_enc and gives the same bits; an in-place edit of the sources (version bump) or new weights
    are detected."""
    from wacv23_tsnet_b200 import ops
    cfg, gold, inputs, net = _build("face_bs1_nb4")
    imgs = torch.from_numpy(np.stack(inputs["src_img"]))
    lbls = torch.from_numpy(np.stack(inputs["src_lbl"]))
    bbs = torch.from_numpy(np.stack(inputs["src_bbox"]))
    tl, tb = torch.from_numpy(inputs["tar_lbl"]), torch.from_numpy(inputs["tar_bbox"])
    with torch.no_grad():
        net.set_test_input(imgs, lbls, bbs, tl, tb)
        l0 = ops.kernel_launches()
        net.forward()
        miss = ops.kernel_launches() - l0
        a = net.rec_tar_img.clone()
        net.enable_source_cache(True)
        for k in range(3):
            net.set_test_input(imgs, lbls, bbs, tl, tb)
            l0 = ops.kernel_launches()
            net.forward()
            hit = ops.kernel_launches() - l0
            assert torch.equal(net.rec_tar_img, a)
        assert hit < miss - 40, (hit, miss)          # the 9 ResnetBlocks + stem of img_enc were skipped
        net.set_test_input(imgs, lbls, bbs, tl, tb.flip(-1).contiguous())    # a new driving frame, same sources
        net.forward()
        assert not torch.equal(net.rec_tar_img, a)
        imgs.mul_(0.5)                                # in-place edit: the version counter changes
        net.set_test_input(imgs, lbls, bbs, tl, tb)
        net.forward()
        b = net.rec_tar_img.clone()
        assert not torch.equal(b, a)
        net.enable_source_cache(False)
        net.set_test_input(imgs, lbls, bbs, tl, tb)
        net.forward()
        assert torch.equal(net.rec_tar_img, b)


def test_load_state_dict_repacks_weights():
    from oracle import make_golden as MG
    cfg, gold, inputs, net = _build("quickstart_bs1")
    _feed(net, inputs)
    with torch.no_grad():
        net.forward()
        a = net.rec_tar_img.clone()
        sds2, _ = MG.build_case(dict(cfg), seed=77)
        net.dec.load_state_dict({k: torch.from_numpy(v) for k, v in sds2["dec"].items()})
        net.forward()
        b = net.rec_tar_img.clone()
        assert not torch.equal(a, b)  # stale packed weights would reproduce `a`
        sds, _ = MG.build_case(cfg)
        net.dec.load_state_dict({k: torch.from_numpy(v) for k, v in sds["dec"].items()})
        net.forward()
        assert torch.equal(net.rec_tar_img, a)  # and the forward is bit-reproducible


def test_full_batch_properties():
    """bs=32, n_blocks=4, n_source=3 (BASELINE.json config 2 size): no cross-sample dependence -- every row of the
    batched forward is bit-identical to the same sample run alone or inside another batch composition, which is what
    makes the 8-GPU batch split exact."""
    from oracle import synth
    from wacv23_tsnet_b200.model.TSNet import TSNet
    torch.manual_seed(1234)
    net = TSNet(is_train=False, label_nc=2, n_blocks=4, n_downsampling=3, n_source=3, return_flow=True)
    net.eval()
    inputs = synth.dataset_like_inputs(32, 2, 3, seed=5)
    with torch.no_grad():
        _feed(net, inputs)
        net.forward()
        full = net.rec_tar_img.clone()
        grids = torch.stack(net.warp_grid2d_list).clone()
        assert torch.isfinite(full).all() and float(full.abs().max()) <= 1.0
        assert float(grids.abs().max()) <= 1.0 + 1e-6      # expected coordinates are convex combinations
        for sl in (slice(0, 1), slice(13, 14), slice(8, 16), slice(16, 32)):
            _feed(net, inputs, sl)
            net.forward()
            assert torch.equal(net.rec_tar_img, full[sl]), f"rows {sl} differ from the batched forward"
            assert torch.equal(torch.stack(net.warp_grid2d_list), grids[:, sl])
        # permuting the sources permutes the grids and leaves the (mean-fused) image unchanged up to fp32 re-association
        _feed(net, {k: (v[::-1] if isinstance(v, list) else v) for k, v in inputs.items()}, slice(0, 4))
        net.forward()
        assert torch.equal(torch.stack(net.warp_grid2d_list), grids[:, 0:4].flip(0))
        assert float((net.rec_tar_img - full[0:4]).abs().max()) < 2e-3


def test_bs32_rows_match_the_oracle():
    """The benchmark's batch size checked directly: bs=32, n_blocks=4, n_source=3 (BASELINE.json config 2) on the GPU,
    rows 0, 17 and 31 against the oracle (bit-exact restatement of the reference forward) run on the host for just
    those samples."""
    from oracle import synth, tsnet_oracle as O
    from wacv23_tsnet_b200.model.TSNet import TSNet
    sds = synth.make_state_dicts(2, 4, seed=1234, bias_std=0.05)
    inputs = synth.dataset_like_inputs(32, 2, 3, seed=7)
    net = TSNet(is_train=False, label_nc=2, n_blocks=4, n_downsampling=3, n_source=3, return_flow=True)
    for k in ("img_enc", "lbl_enc", "fuse_net", "dec"):
        getattr(net, k).load_state_dict({kk: torch.from_numpy(v) for kk, v in sds[k].items()})
    net.eval()
    with torch.no_grad():
        _feed(net, inputs)
        net.forward()
    out = net.rec_tar_img.cpu()
    grids = torch.stack(net.warp_grid2d_list).cpu()
    tsd = O.to_torch_sd(sds)
    for r in (0, 17, 31):
        one = {k: ([a[r:r + 1] for a in v] if isinstance(v, list) else v[r:r + 1]) for k, v in inputs.items()}
        ref = O.tsnet_forward(tsd, one, 4)
        assert float((out[r:r + 1] - ref["rec_tar_img"]).abs().max()) < IMG_TOL, r
        assert float((grids[:, r:r + 1] - torch.stack(ref["grids"])).abs().max()) < GRID_TOL, r


@pytest.mark.parametrize("variant", [False, "unfused"])
def test_direct_and_winograd_paths_agree(variant):
    """winograd=False keeps the direct implicit GEMM for the ResnetBlock convolutions, winograd="unfused" the separate
    transform passes instead of the fused bridge: all paths compute the same function (each within the parity tolerance
    of the reference; against each other within the same noise floor)."""
    cfg, gold, inputs, net = _build("face_bs1_nb4")
    from wacv23_tsnet_b200.model.TSNet import TSNet
    ref_net = TSNet(is_train=False, label_nc=cfg["label_nc"], n_blocks=cfg["n_blocks"], n_downsampling=3,
                    n_source=cfg["n_source"], return_flow=True, winograd=variant)
    for k in ("img_enc", "lbl_enc", "fuse_net", "dec"):
        getattr(ref_net, k).load_state_dict(getattr(net, k).state_dict())
    ref_net.eval()
    with torch.no_grad():
        _feed(net, inputs)
        net.forward()
        _feed(ref_net, inputs)
        ref_net.forward()
    g = torch.from_numpy(gold["rec_tar_img"])
    assert float((ref_net.rec_tar_img.cpu() - g).abs().max()) < IMG_TOL
    assert float((net.rec_tar_img.cpu() - g).abs().max()) < IMG_TOL
    assert float((net.rec_tar_img - ref_net.rec_tar_img).abs().max()) < IMG_TOL
    assert not torch.equal(net.rec_tar_img, ref_net.rec_tar_img)   # the paths really are different kernels


def test_data_writes_need_invalidate_weights():
    """Writes through `.data` do not bump the tensor version that keys the packed-weight cache: after
    invalidate_weights() the forward must see the new values."""
    cfg, gold, inputs, net = _build("quickstart_bs1")
    _feed(net, inputs)
    with torch.no_grad():
        net.forward()
        a = net.rec_tar_img.clone()
        p = net.dec.state_dict(keep_vars=True)["map_conv.weight"]
        p.data.copy_(p.data * 1.5)
        net.invalidate_weights()
        net.forward()
        assert not torch.equal(net.rec_tar_img, a)


def test_fast_modes_run_and_are_flagged_non_parity():
    """Single-pass modes exist as speed points only: they must run, but they are NOT within the parity tolerance
    (documents why the headline uses fp16x3)."""
    cfg, gold, inputs, net = _build("quickstart_bs1", math_mode="bf16")
    _feed(net, inputs)
    with torch.no_grad():
        net.forward()
    err = float((net.rec_tar_img.cpu() - torch.from_numpy(gold["rec_tar_img"])).abs().max())
    assert np.isfinite(err) and err > IMG_TOL


def test_cuda_graph_replay_is_bit_identical_to_eager():
    """the whole forward captured as one CUDA graph (the demos' one-frame-per-forward loop)."""
    cfg, gold, inputs, net = _build("face_bs1_nb4")
    with torch.no_grad():
        _feed(net, inputs)
        net.forward()
        eager = net.rec_tar_img.clone()
        eager_grids = torch.stack(net.warp_grid2d_list).clone()
        net.enable_cuda_graph(True)
        for _ in range(3):  # capture, then two replays with freshly staged inputs
            _feed(net, inputs)
            net.forward()
            assert torch.equal(net.rec_tar_img, eager)
            assert torch.equal(torch.stack(net.warp_grid2d_list), eager_grids)
        # different input content through the same graph
        inputs2 = {k: ([a[::-1].copy() if a.ndim == 3 else a[..., ::-1].copy() for a in v] if isinstance(v, list)
                       else v[..., ::-1].copy()) for k, v in inputs.items()}
        _feed(net, inputs2)
        net.forward()
        g_out = net.rec_tar_img.clone()
        net.enable_cuda_graph(False)
        _feed(net, inputs2)
        net.forward()
        assert torch.equal(net.rec_tar_img, g_out)


_SHARD_WORKER = r"""
import os, sys, torch
sys.path.insert(0, {root!r})
from oracle import synth
from wacv23_tsnet_b200 import dist as D
from wacv23_tsnet_b200.model.TSNet import TSNet
rank, local, world = D.init_from_env("nccl")
torch.cuda.set_device(local)
torch.manual_seed(100 + rank)                      # different initial weights per rank: the broadcast must fix that
net = TSNet(is_train=False, label_nc=2, n_blocks=1, n_downsampling=3, n_source=2)
D.broadcast_generator(net, src=0)
net.eval()
inp = synth.dataset_like_inputs(4 * world, 2, 2, seed=11)   # the GLOBAL batch, identical on every rank
t = lambda a: torch.from_numpy(a)
full = dict(src_img=[t(a) for a in inp["src_img"]], src_lbl=[t(a) for a in inp["src_lbl"]],
            src_bbox=[t(a) for a in inp["src_bbox"]], tar_lbl=t(inp["tar_lbl"]), tar_bbox=t(inp["tar_bbox"]))
mine = D.shard_inputs(full, rank, world)
with torch.no_grad():
    net.set_test_input(mine["src_img"], mine["src_lbl"], mine["src_bbox"], mine["tar_lbl"], mine["tar_bbox"])
    net.forward()
    gathered = D.all_gather_frames(net.rec_tar_img)          # [4*world, 3, 256, 256], rank-major
    net.set_test_input(full["src_img"], full["src_lbl"], full["src_bbox"], full["tar_lbl"], full["tar_bbox"])
    net.forward()                                             # the whole batch on ONE GPU
assert torch.equal(gathered, net.rec_tar_img), "sharded forward differs from the single-GPU forward"
sys.stdout.write("rank%dok\n" % rank); sys.stdout.flush()
torch.distributed.destroy_process_group()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpu_batch_shards_equal_single_gpu(tmp_path):
    """rank r owns rows [r*B/R, (r+1)*B/R); NCCL only broadcasts the weights and gathers the frames.
    The gathered sharded output must be BIT-identical to the same global batch run on one GPU."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "shard_worker.py"
    script.write_text(_SHARD_WORKER.format(root=root))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29733", str(script)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "rank0ok" in r.stdout and "rank1ok" in r.stdout


def test_frame_pipeline_matches_synchronous_forward():
    """wacv23_tsnet_b200.pipeline.FramePipeline (overlapped H2D / forward / D2H) returns, in order, exactly what
    set_test_input + forward + .cpu() returns for each batch."""
    from oracle import synth
    from wacv23_tsnet_b200.pipeline import FramePipeline
    cfg, gold, inputs, net = _build("quickstart_bs1")
    batches = []
    for k in range(4):
        inp = synth.quick_start_inputs(2, 2, 3, seed=50 + k)
        batches.append({key: ([torch.from_numpy(a).pin_memory() for a in v] if isinstance(v, list)
                              else torch.from_numpy(v).pin_memory()) for key, v in inp.items() if key != "tar_img"})
    expect = []
    with torch.no_grad():
        for bt in batches:
            net.set_test_input(bt["src_img"], bt["src_lbl"], bt["src_bbox"], bt["tar_lbl"], bt["tar_bbox"])
            net.forward()
            expect.append(net.rec_tar_img.cpu())
    got = list(FramePipeline(net).run(batches))
    assert len(got) == len(expect)
    for g, e in zip(got, expect):
        assert torch.equal(g, e)


def test_classmap_labels_give_identical_forward():
    """uint8 class-index labels ([B,H,W]) staged on the device == the one-hot float planes."""
    from oracle import synth
    from wacv23_tsnet_b200.model.TSNet import TSNet
    torch.manual_seed(3)
    net = TSNet(is_train=False, label_nc=2, n_blocks=0, n_downsampling=3, n_source=2)
    net.eval()
    inp = synth.dataset_like_inputs(2, 2, 2, seed=9)                 # one-hot labels
    cls = lambda oh: torch.from_numpy(oh.argmax(1).astype(np.uint8))
    t = torch.from_numpy
    with torch.no_grad():
        net.set_test_input([t(a) for a in inp["src_img"]], [t(a) for a in inp["src_lbl"]],
                           [t(a) for a in inp["src_bbox"]], t(inp["tar_lbl"]), t(inp["tar_bbox"]))
        net.forward()
        a = net.rec_tar_img.clone()
        net.set_test_input([t(a_) for a_ in inp["src_img"]], [cls(a_) for a_ in inp["src_lbl"]],
                           [t(a_) for a_ in inp["src_bbox"]], cls(inp["tar_lbl"]), t(inp["tar_bbox"]))
        net.forward()
        assert torch.equal(net.rec_tar_img, a)


def test_uint8_images_give_identical_forward_and_bad_shapes_raise():
    """SURVEY section 8f row 2 (remainder): uint8 BGR source images + the dataset mean staged on the device; the
    loader evaluates (u8 - mean) / 255 (dataset/dataset_video_face.py:329 + model/TSNet.py:286) -> same bits as feeding
    the mean-subtracted fp32 images.  Mismatched source tensors raise instead of being read out of bounds."""
    from oracle import synth
    from wacv23_tsnet_b200.model.TSNet import TSNet
    torch.manual_seed(4)
    net = TSNet(is_train=False, label_nc=2, n_blocks=0, n_downsampling=3, n_source=2, img_mean=synth.IMG_MEAN)
    net.eval()
    inp = synth.dataset_like_inputs(2, 2, 2, seed=13)
    t = torch.from_numpy
    g = torch.Generator().manual_seed(0)
    u8 = [torch.randint(0, 256, (2, 3, 256, 256), generator=g).to(torch.uint8) for _ in range(2)]
    mean = torch.tensor(np.asarray(synth.IMG_MEAN, dtype=np.float32)).view(1, 3, 1, 1)
    f32 = [x.float() - mean for x in u8]                                # what the dataset would have produced
    with torch.no_grad():
        net.set_test_input(f32, [t(a) for a in inp["src_lbl"]], [t(a) for a in inp["src_bbox"]], t(inp["tar_lbl"]),
                           t(inp["tar_bbox"]))
        net.forward()
        a = net.rec_tar_img.clone()
        net.set_test_input(u8, [t(a_) for a_ in inp["src_lbl"]], [t(a_) for a_ in inp["src_bbox"]], t(inp["tar_lbl"]),
                           t(inp["tar_bbox"]))
        assert net._src_img_raw[0].dtype == torch.uint8
        net.forward()
        assert torch.equal(net.rec_tar_img, a)
        assert torch.equal(net.src_img_list[1], f32[1].cuda() / 255.0)
        with pytest.raises(ValueError):
            net.set_test_input(u8, [t(a_) for a_ in inp["src_lbl"]], [t(a_)[:1] for a_ in inp["src_bbox"]],
                               t(inp["tar_lbl"]), t(inp["tar_bbox"]))
            net.forward()
        with pytest.raises(ValueError):
            net.set_test_input([x[:, :, :128] for x in u8], [t(a_) for a_ in inp["src_lbl"]],
                               [t(a_) for a_ in inp["src_bbox"]], t(inp["tar_lbl"]), t(inp["tar_bbox"]))
            net.forward()


def test_use_prev_sources_skip_the_255_division():
    """set_train_input(use_prev=...) (model/TSNet.py:266-276): flagged sources are used as they are, the others /255."""
    from oracle import synth
    from wacv23_tsnet_b200.model.TSNet import TSNet
    torch.manual_seed(5)
    net = TSNet(is_train=False, label_nc=2, n_blocks=0, n_downsampling=3, n_source=3)
    net.eval()
    inp = synth.dataset_like_inputs(1, 2, 3, seed=21)
    t = torch.from_numpy
    imgs = [t(a) for a in inp["src_img"]]
    pre = [imgs[0], imgs[1] / 255.0, imgs[2]]                     # source 1 arrives already scaled ("previous output")
    with torch.no_grad():
        net.set_train_input(pre, [t(a) for a in inp["src_lbl"]], [t(a) for a in inp["src_bbox"]], t(inp["tar_img"]),
                            t(inp["tar_lbl"]), t(inp["tar_bbox"]), use_prev=[False, True, False])
        net.forward()
        a = net.rec_tar_img.clone()
        net.set_test_input(imgs, [t(x) for x in inp["src_lbl"]], [t(x) for x in inp["src_bbox"]], t(inp["tar_lbl"]),
                           t(inp["tar_bbox"]))
        net.forward()
    # identical up to the rounding of x/255 done on the host (true division) vs in the mixed-divisor staging path
    assert float((net.rec_tar_img - a).abs().max()) < 2e-3
