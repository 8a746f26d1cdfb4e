"""CPU suite, part 4: the drop-in boundary proven on the reference's REAL callers.

The north star asks that demo/demo_face.py and demo/demo_pose.py "run unmodified".  Here the two scripts are executed
AS THEY ARE (runpy, `__main__`) from /root/reference with this repository's `model/` package first on sys.path, so their
`from model.TSNet import TSNet` / `from model.TSNet_pose import TSNet` bind to the B200 implementation.  Everything the
scripts need besides the model is stubbed: the hard-coded checkpoint / output paths (os.path.exists, os.makedirs,
torch.load returning a checkpoint dict in the layout train_face.py:350-358 writes), the dataset classes (one synthetic
video pair with the shapes and dtypes dataset/dataset_video_face.py:248-530 / dataset_video_pose.py:275-607 produce),
utils.misc (Logger, vl2ch with the semantics of utils/misc.py:50-67), cv2, imageio, tqdm.

On this CPU box the run must get through model construction, the four load_state_dict calls, eval(), the data loop,
vl2ch, set_test_input and INTO model.forward(), and stop there with TSNetLibraryError (no device: there is no CPU
fallback).  The GPU box has no /root/reference (the -m gpu tests may not read it), so the finite-frame half of this
check is tests/test_parity_gpu.py::test_demo_call_sequence_5d_lists_uint8_bbox_and_source_count, which replays the
same call sequence with the same tensor shapes / dtypes on the device.
"""
import os
import runpy
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "demo")) or torch.cuda.is_available(),
                                reason="needs the reference checkout (build container) and no GPU")


def _vl2ch(label_nc):
    def vl2ch(lbl, type_):   # class-index map [T, H, W] -> one-hot float planes [T, label_nc, H, W]
        lbl = lbl.long()
        return torch.stack([(lbl == c).float() for c in range(label_nc)], dim=1)
    return vl2ch


class _VideoPair(torch.utils.data.Dataset):
    """One (subject video, driving video) item: images mean-subtracted BGR float32 [T,3,256,256], class-index labels
    [T,256,256], uint8 rectangular bboxes [T,256,256], frame names."""

    def __init__(self, label_nc, frames=4, **kw):
        self.label_nc, self.T = label_nc, frames
        self.kw = kw

    def __len__(self):
        return 1

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(3)
        T = self.T

        def video(tag):
            img = torch.rand(T, 3, 256, 256, generator=g) * 255 - 110
            lbl = torch.randint(0, self.label_nc, (T, 256, 256), generator=g).to(torch.uint8)
            bb = torch.zeros(T, 256, 256, dtype=torch.uint8)
            bb[:, 40:220, 50:200] = 1
            names = ["%s_%04d.png" % (tag, k) for k in range(T)]
            return img, lbl, bb, names
        s, d = video("val024"), video("test114")
        return (*s, *d)


def _run_demo(script, model_mod, label_nc, dataset_mod, dataset_cls, monkeypatch, tmp_path):
    from wacv23_tsnet_b200 import lib
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        mod = __import__("wacv23_tsnet_b200.model." + model_mod, fromlist=["TSNet"])
        donor = mod.TSNet(is_train=False, label_nc=label_nc, n_blocks=4, n_downsampling=3, n_source=3)
    ckpt = {"example": 0, "img_enc": donor.img_enc.state_dict(), "lbl_enc": donor.lbl_enc.state_dict(),
            "dec": donor.dec.state_dict(), "fuse_net": donor.fuse_net.state_dict(), "netD": {}}
    reached = {}

    # ---- stub modules the demo imports besides the model
    misc = types.ModuleType("utils.misc")

    class Logger:
        def __init__(self, path, stream):
            self.stream = stream

        def write(self, m):
            self.stream.write(m)

        def flush(self):
            self.stream.flush()
    misc.Logger, misc.vl2ch, misc.vl2im = Logger, _vl2ch(label_nc), (lambda lbl, t: np.zeros(lbl.shape + (3,), np.uint8))
    utils = types.ModuleType("utils")
    utils.misc = misc
    ds = types.ModuleType("dataset." + dataset_mod)
    setattr(ds, dataset_cls, lambda **kw: _VideoPair(label_nc, **kw))
    dataset = types.ModuleType("dataset")
    setattr(dataset, dataset_mod, ds)
    cv2 = types.ModuleType("cv2")
    cv2.COLOR_BGR2RGB = 4
    cv2.cvtColor = lambda img, code: img[..., ::-1]
    imageio = types.ModuleType("imageio")
    imageio.mimsave = lambda *a, **k: None
    tqdm = types.ModuleType("tqdm")
    tqdm.tqdm = lambda it, *a, **k: it
    for name, m in (("utils", utils), ("utils.misc", misc), ("dataset", dataset), ("dataset." + dataset_mod, ds),
                    ("cv2", cv2), ("imageio", imageio), ("tqdm", tqdm)):
        monkeypatch.setitem(sys.modules, name, m)
    for name in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
        monkeypatch.delitem(sys.modules, name)
    monkeypatch.syspath_prepend(ROOT)            # this repository's `model/` shim wins over the reference's package

    # ---- the hard-coded paths of the demo
    real_exists, real_isfile, real_makedirs = os.path.exists, os.path.isfile, os.makedirs
    fake = lambda p: isinstance(p, str) and p.startswith(("/data/hfn5052", "/home/hfn5052"))
    monkeypatch.setattr(os.path, "exists", lambda p: True if fake(p) else real_exists(p))
    monkeypatch.setattr(os.path, "isfile", lambda p: True if fake(p) else real_isfile(p))
    monkeypatch.setattr(os, "makedirs", lambda p, *a, **k: None if fake(p) else real_makedirs(p, *a, **k))
    monkeypatch.setattr(torch, "load", lambda *a, **k: ckpt)
    monkeypatch.setattr(sys, "argv", [script])   # defaults: one DataLoader worker, batch size 1

    # ---- record that the forward is entered through the drop-in class
    real_forward = mod.TSNet.forward

    def spy(self, *a, **k):
        reached["forward"] = True
        reached["inputs"] = (len(self.src_lbl_list), tuple(self._src_img_raw[0].shape), self.src_bbox_list[0].dtype,
                             tuple(self.tar_lbl.shape))
        return real_forward(self, *a, **k)
    base = __import__("wacv23_tsnet_b200.model.TSNet", fromlist=["TSNet"]).TSNet
    monkeypatch.setattr(base, "forward", spy)

    stdout = sys.stdout
    try:
        with pytest.raises(lib.TSNetLibraryError):
            runpy.run_path(os.path.join(REF, "demo", script), run_name="__main__")
    finally:
        sys.stdout = stdout      # the demo replaces sys.stdout with its Logger
    assert reached.get("forward"), "the demo never reached TSNet.forward()"
    return reached


def test_demo_face_runs_unmodified_up_to_the_device(cuda_off_shim, monkeypatch, tmp_path):
    r = _run_demo("demo_face.py", "TSNet", 2, "dataset_video_face", "FaceDatasetTest", monkeypatch, tmp_path)
    n, img_shape, bb_dtype, tar_lbl_shape = r["inputs"]
    assert n == 3 and img_shape == (1, 3, 256, 256) and bb_dtype == torch.uint8 and tar_lbl_shape == (1, 2, 256, 256)


def test_demo_pose_runs_unmodified_up_to_the_device(cuda_off_shim, monkeypatch, tmp_path):
    r = _run_demo("demo_pose.py", "TSNet_pose", 25, "dataset_video_pose", "PoseDatasetTestVideo", monkeypatch, tmp_path)
    n, img_shape, bb_dtype, tar_lbl_shape = r["inputs"]
    assert n == 3 and img_shape == (1, 3, 256, 256) and bb_dtype == torch.uint8 and tar_lbl_shape == (1, 25, 256, 256)
