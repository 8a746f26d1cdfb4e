"""Import the UNMODIFIED reference (/root/reference) and run its forward on CPU.

TEST INFRASTRUCTURE ONLY.  Works only in the build container (the GPU box has no
/root/reference); used by oracle/make_golden.py to pin oracle/tsnet_oracle.py and to write
tests/golden/*.npz.  The single shim: `.cuda()` is made an identity because
model/networks.py:116 calls it unconditionally and this container has no GPU.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("TSNET_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "TSNet.py"))


@contextlib.contextmanager
def _reference_on_path():
    """Temporarily make `import model.TSNet` resolve to the reference (our repo root also has a
    `model/` drop-in shim, so we must isolate sys.modules)."""
    saved = {k: v for k, v in sys.modules.items() if k == "model" or k.startswith("model.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        yield
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            del sys.modules[k]
        sys.modules.update(saved)


@contextlib.contextmanager
def _cuda_is_identity():
    if torch.cuda.is_available():
        yield
        return
    t_cuda, m_cuda = torch.Tensor.cuda, torch.nn.Module.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = t_cuda, m_cuda


@torch.no_grad()
def reference_forward(sds, inputs, label_nc, n_blocks, pose=False, pose_mean=None, n_source=3, train=False,
                      use_prev=None):
    """Build reference TSNet(is_train=False, ...), load `sds` (numpy state dicts), run set_test_input +
    forward() on CPU.  Returns dict with rec_tar_img (+ warp grids for the face variant).
    train=True: the is_train=True branches of forward() are exercised.  Constructing the reference with is_train=True
    needs a VGG19 download (model/TSNet.py:545), so the generator-only model is built and its `is_train` attribute is
    flipped before set_train_input + forward(): forward() itself reads nothing else that is_train=True would create."""
    with _reference_on_path(), _cuda_is_identity(), contextlib.redirect_stdout(io.StringIO()):
        if pose:
            from model.TSNet_pose import TSNet
            kw = dict(use_mask=True, mean=np.asarray(pose_mean, np.float32))
        else:
            from model.TSNet import TSNet
            kw = dict(return_flow=True)
        net = TSNet(is_train=False, label_nc=label_nc, n_blocks=n_blocks, n_downsampling=3,
                    n_source=n_source, **kw)
        for name in ("img_enc", "lbl_enc", "fuse_net", "dec"):
            getattr(net, name).load_state_dict({k: torch.from_numpy(v) for k, v in sds[name].items()})
        net.eval()
        if train:
            net.is_train = True
            net.set_train_input([torch.from_numpy(x) for x in inputs["src_img"]],
                                [torch.from_numpy(x) for x in inputs["src_lbl"]],
                                [torch.from_numpy(x) for x in inputs["src_bbox"]],
                                torch.from_numpy(inputs["tar_img"]), torch.from_numpy(inputs["tar_lbl"]),
                                torch.from_numpy(inputs["tar_bbox"]), use_prev=use_prev)
        else:
            net.set_test_input([torch.from_numpy(x) for x in inputs["src_img"]],
                               [torch.from_numpy(x) for x in inputs["src_lbl"]],
                               [torch.from_numpy(x) for x in inputs["src_bbox"]],
                               torch.from_numpy(inputs["tar_lbl"]), torch.from_numpy(inputs["tar_bbox"]))
        net.forward()
        out = {"rec_tar_img": net.rec_tar_img.clone()}
        if not pose:
            out["grids"] = [g.clone() for g in net.warp_grid2d_list]
        if train:
            out["warp_src_img_list"] = [x.clone() for x in net.warp_src_img_list]
            out["loss_warp"] = net.loss_warp.clone()
            if not pose:
                out["loss_align"] = net.loss_align.clone()
        return out
