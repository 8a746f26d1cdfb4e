"""bs=32 forward: eager launches vs CUDA-graph replay (TSNet(cuda_graph=True)); inputs resident."""
import contextlib, io, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wacv23_tsnet_b200.model.TSNet import TSNet
from oracle import synth  # data generator only
L, nb, n, bs = 2, 4, 3, int(os.environ.get("B", 32))
torch.manual_seed(1234)
with contextlib.redirect_stdout(io.StringIO()):
    net = TSNet(is_train=False, label_nc=L, n_blocks=nb, n_downsampling=3, n_source=n)
net.eval()
inp = synth.dataset_like_inputs(bs, L, n, seed=1234)
dev = {k: ([torch.from_numpy(a).cuda() for a in v] if isinstance(v, list) else torch.from_numpy(v).cuda())
       for k, v in inp.items() if k != "tar_img"}
def step():
    net.set_test_input(dev["src_img"], dev["src_lbl"], dev["src_bbox"], dev["tar_lbl"], dev["tar_bbox"])
    net.forward()
with torch.no_grad():
    for mode in ("eager", "graph", "eager", "graph"):
        net.enable_cuda_graph(mode == "graph")
        for _ in range(4): step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): step()
        e1.record(); torch.cuda.synchronize()
        print(f"{mode}: {e0.elapsed_time(e1) / 20:.3f} ms/step")
