"""The demos' operating point (demo/demo_face.py:185-192: ONE driving frame per forward, the same 3 source frames every
time): ms per frame for eager launches, CUDA-graph replay, and graph + the opt-in source-feature cache.  Host tensors
are staged per frame as the demo does (set_test_input on CPU tensors).  Run under gpurun:  python tools/demo_point.py"""
import contextlib, io, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wacv23_tsnet_b200.model.TSNet import TSNet
from oracle import synth  # data generator only

L, nb, n = 2, 4, 3
torch.manual_seed(1234)
with contextlib.redirect_stdout(io.StringIO()):
    net = TSNet(is_train=False, label_nc=L, n_blocks=nb, n_downsampling=3, n_source=n)
net.eval()
inp = synth.dataset_like_inputs(8, L, n, seed=1234)
t = torch.from_numpy
src_img = torch.stack([t(a[:1]) for a in inp["src_img"]])     # [n, 1, 3, 256, 256]: the demo's 5-D "lists"
src_lbl = torch.stack([t(a[:1]) for a in inp["src_lbl"]])
src_bb = torch.stack([t(a[:1]) for a in inp["src_bbox"]])
tar_lbl, tar_bb = t(inp["tar_lbl"]), t(inp["tar_bbox"])         # 8 driving frames


def frame(k):
    net.set_test_input(src_img, src_lbl, src_bb, tar_lbl[k % 8:k % 8 + 1], tar_bb[k % 8:k % 8 + 1])
    net.forward()
    return net.rec_tar_img


out = {}
with torch.no_grad():
    for mode in ("eager", "graph", "graph+source_cache", "eager+source_cache"):
        net.enable_cuda_graph(mode.startswith("graph"))
        net.enable_source_cache(mode.endswith("source_cache"))
        for k in range(6):
            frame(k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for k in range(40):
            frame(k)
        e1.record()
        torch.cuda.synchronize()
        out[mode] = {"ms_per_frame": e0.elapsed_time(e1) / 40}
        print(f"{mode}: {out[mode]['ms_per_frame']:.3f} ms/frame", flush=True)
print(json.dumps({"demo_operating_point_bs1_n3_nb4": out}))
