"""CPU suite, part 2: the C-ABI library (loads, exports everything the header declares), the host-side class
surface (state_dict layout, input staging, loud failure without a device) and the multi-rank plumbing (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from wacv23_tsnet_b200 import lib
    return lib


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "tsnet_b200.h")).read()
    declared = set(re.findall(r"\b(tsnet_[a-z0-9_]+)\s*\(", hdr))
    declared = {d for d in declared if not d.endswith("_desc")}
    assert len(declared) >= 12
    h = ctypes.CDLL(built.LIB_PATH)
    for name in declared:
        assert hasattr(h, name), f"{name} declared in include/tsnet_b200.h but not exported"
    assert declared == set(built.EXPORTED_SYMBOLS), declared ^ set(built.EXPORTED_SYMBOLS)
    assert built.load().tsnet_abi_version() == 3


def test_struct_layouts_match_header(built):
    # sizes the C compiler gives the descriptor structs (catches ctypes / header drift without a GPU)
    src = '#include "tsnet_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(tsnet_conv_desc),' \
          ' sizeof(tsnet_taps_desc), sizeof(tsnet_corr_desc), sizeof(tsnet_wino_gemm_desc),' \
          ' sizeof(tsnet_wino_bridge_desc), sizeof(tsnet_stem_conv_desc));return 0;}'
    exe = "/tmp/tsnet_sizeof"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src.encode(), check=True)
    sizes = [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(built.ConvDesc), ctypes.sizeof(built.TapsDesc), ctypes.sizeof(built.CorrDesc),
                     ctypes.sizeof(built.WinoGemmDesc), ctypes.sizeof(built.WinoBridgeDesc),
                     ctypes.sizeof(built.StemConvDesc)]


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-box behaviour")
def test_no_device_fails_loudly(built, cuda_off_shim):
    from wacv23_tsnet_b200.model.TSNet import TSNet
    net = TSNet(is_train=False, label_nc=2, n_blocks=0, n_downsampling=3, n_source=1)
    x = torch.zeros(1, 3, 256, 256)
    net.set_test_input([x], [torch.zeros(1, 2, 256, 256)], [torch.zeros(1, 256, 256)], torch.zeros(1, 2, 256, 256),
                       torch.zeros(1, 256, 256))
    with pytest.raises(built.TSNetLibraryError):
        net.forward()  # no CPU / PyTorch fallback exists


def test_state_dict_layout_and_loading(cuda_off_shim):
    from oracle import synth
    from wacv23_tsnet_b200.model.TSNet import TSNet
    from wacv23_tsnet_b200.model.TSNet_pose import TSNet as TSNetPose
    for cls, L, nb in ((TSNet, 2, 4), (TSNetPose, 25, 4), (TSNet, 2, 0)):
        net = cls(is_train=False, label_nc=L, n_blocks=nb, n_downsampling=3, n_source=3)
        expect = {"img_enc": synth.encoder_shapes(3 + L, 9), "lbl_enc": synth.encoder_shapes(L, 0),
                  "fuse_net": synth.fuse_shapes(), "dec": synth.decoder_shapes(nb)}
        for name, shapes in expect.items():
            sd = getattr(net, name).state_dict()
            assert {k: tuple(v.shape) for k, v in sd.items()} == shapes
            assert list(sd.keys()) == list(shapes.keys())  # same order as the reference's state_dict
        # reference init statistics: weights ~ N(0, 0.02^2), biases 0 (model/networks.py:67-103)
        w = net.img_enc.state_dict()["model.13.conv_block.1.weight"]
        assert abs(float(w.std()) - 0.02) < 5e-4 and abs(float(w.mean())) < 5e-4
        assert float(net.dec.state_dict()["map_conv.bias"].abs().max()) == 0.0
    sds = synth.make_state_dicts(2, 0, seed=5)
    net = TSNet(is_train=False, label_nc=2, n_blocks=0, n_downsampling=3)
    for k in ("img_enc", "lbl_enc", "fuse_net", "dec"):
        getattr(net, k).load_state_dict({kk: torch.from_numpy(v) for kk, v in sds[k].items()})  # strict


def test_rejects_impossible_geometry_and_training(cuda_off_shim):
    from wacv23_tsnet_b200.model.TSNet import TSNet
    with pytest.raises(ValueError):
        TSNet(is_train=False, label_nc=2, n_downsampling=4)  # FuseNet(1024) only fits ngf*2^n = 512
    # is_train=True is forward-only here: the generator is built (no VGG19 download, no discriminators) and everything
    # that needs a backward pass raises
    net = TSNet(is_train=True, label_nc=2, n_downsampling=3)
    assert net.is_train and not hasattr(net, "netD")
    for fn in (net.optimize_parameters, net.setup, net.print_learning_rate, net.get_current_losses):
        with pytest.raises(NotImplementedError):
            fn()


def test_input_staging_matches_reference_semantics(cuda_off_shim):
    """set_*_input: /255 (except use_prev), bbox unsqueeze(1), 5-D 'list' tensors, uint8 masks (model/TSNet.py:266-294)."""
    from wacv23_tsnet_b200.model.TSNet import TSNet
    net = TSNet(is_train=False, label_nc=2, n_blocks=0, n_downsampling=3, n_source=3)
    g = torch.Generator().manual_seed(0)
    imgs = torch.rand(3, 1, 3, 256, 256, generator=g) * 255  # demo style: [n, 1, 3, H, W] iterated on dim 0
    lbls = torch.zeros(3, 1, 2, 256, 256)
    bbs = torch.randint(0, 2, (3, 1, 256, 256), generator=g).to(torch.uint8)
    net.set_test_input(imgs, lbls, bbs, torch.zeros(1, 2, 256, 256), bbs[0])
    assert len(net.src_img_list) == 3 and net.src_img_list[0].shape == (1, 3, 256, 256)
    assert torch.equal(net.src_img_list[1], imgs[1] / 255.0)
    assert net.src_bbox_list[2].shape == (1, 1, 256, 256) and net.src_bbox_list[2].dtype == torch.uint8
    assert net.tar_bbox.shape == (1, 1, 256, 256)
    net.set_train_input(list(imgs), list(lbls), list(bbs.float()), imgs[0], lbls[0], bbs[0].float(),
                        use_prev=[False, True, False])
    assert torch.equal(net.src_img_list[1], imgs[1]) and torch.equal(net.src_img_list[0], imgs[0] / 255.0)
    assert torch.equal(net.tar_img, imgs[0] / 255.0)
    net.set_source_num(2)
    assert net.n_source == 2


@pytest.mark.skipif(not os.path.isdir("/root/reference/model"), reason="reference checkout only in the build container")
def test_same_seed_gives_reference_weights(cuda_off_shim):
    """Same construction order + same RNG consumption as the reference => identical parameters from one seed."""
    import contextlib
    import io
    from oracle import ref_harness
    from wacv23_tsnet_b200.model.TSNet import TSNet
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(io.StringIO()):
        mine = TSNet(is_train=False, label_nc=2, n_blocks=1, n_downsampling=3)
    with ref_harness._reference_on_path(), contextlib.redirect_stdout(io.StringIO()):
        from model.TSNet import TSNet as RefTSNet
        torch.manual_seed(1234)
        ref = RefTSNet(is_train=False, label_nc=2, n_blocks=1, n_downsampling=3)
    for name in ("img_enc", "lbl_enc", "fuse_net", "dec"):
        a, b = getattr(mine, name).state_dict(), getattr(ref, name).state_dict()
        assert list(a.keys()) == list(b.keys())
        assert all(torch.equal(a[k], b[k]) for k in a)


def test_shard_range():
    from wacv23_tsnet_b200 import dist as D
    assert [D.shard_range(256, r, 8) for r in (0, 1, 7)] == [(0, 32), (32, 64), (224, 256)]
    with pytest.raises(ValueError):
        D.shard_range(30, 0, 4)


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from wacv23_tsnet_b200 import dist as D
torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self
rank, local, world = D.init_from_env("gloo")
from wacv23_tsnet_b200.model.TSNet import TSNet
torch.manual_seed(100 + rank)                       # deliberately different initial weights per rank
net = TSNet(is_train=False, label_nc=2, n_blocks=0, n_downsampling=3)
D.broadcast_generator(net, src=0)
sig = torch.stack([p.double().sum() for n in D.GENERATOR_NETS for p in getattr(net, n).parameters()]).sum()
sigs = [torch.zeros_like(sig) for _ in range(world)]
dist.all_gather(sigs, sig)
assert all(torch.equal(s, sigs[0]) for s in sigs), "replicas differ after broadcast"
full = {{"tar_lbl": torch.arange(8.).view(8, 1), "src_img": [torch.arange(8.).view(8, 1) + 10 * i for i in range(3)]}}
mine = D.shard_inputs(full, rank, world)
assert mine["tar_lbl"].flatten().tolist() == [4. * rank + k for k in range(4)]
frames = D.all_gather_frames(mine["src_img"][1])
assert torch.equal(frames, full["src_img"][1])      # shards tile the batch in rank order
assert D.max_over_ranks(float(rank), "cpu") == world - 1
sys.stdout.write("rank%dok\n" % rank); sys.stdout.flush()
"""


def test_two_rank_gloo_plumbing(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "rank0ok" in r.stdout and "rank1ok" in r.stdout


def test_corr_host_side_validation_and_workspace(built):
    """Host-side parts of the correlation ABI that need no device: workspace sizing and argument checks."""
    import ctypes as C
    lib = built.load()
    d = built.CorrDesc()
    d.B, d.n_src, d.C, d.h, d.w = 32, 3, 512, 32, 32
    d.bbox_h = d.bbox_w = 256
    d.temperature, d.split, d.fmt, d.operand_scale, d.sort = 100.0, 1, 0, 4096.0 ** 2, 1
    n32 = lib.tsnet_corr_workspace_bytes(C.byref(d))
    # rank u16 + mask f32 per map, coordinates per source map, 8 float4 states per (source, row)
    assert n32 >= 128 * 1024 * 6 + 96 * 1024 * 8 + 3 * 32 * 1024 * 8 * 16
    d.B = 64
    assert lib.tsnet_corr_workspace_bytes(C.byref(d)) > n32
    for field, bad in (("n_src", 13), ("h", 30), ("C", 520), ("bbox_dtype", 2), ("B", 0)):
        e = built.CorrDesc()
        for f, _ in built.CorrDesc._fields_:
            setattr(e, f, getattr(d, f))
        setattr(e, field, bad)
        assert lib.tsnet_corr_workspace_bytes(C.byref(e)) == 0
        assert lib.tsnet_corr_prepare(C.byref(e), None, None, None, None, 0, None) < 0
        assert len(lib.tsnet_last_error()) > 0
    assert lib.tsnet_launch_count() >= 0
