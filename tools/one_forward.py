"""bench.py's workload (FaceForensics config: bs=32, n_source=3, n_blocks=4, fp16x3, inputs resident), N forwards and
nothing else -- the target of the ncu captures in tools/gpu_r2.sh.  Prints `FORWARD_LAUNCHES <n>` = kernel launches of
the last forward (counted by the library wrappers) so that the summariser can cut the launch list."""
import argparse, contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wacv23_tsnet_b200.model.TSNet import TSNet
from oracle import synth  # data generator only

ap = argparse.ArgumentParser()
ap.add_argument("--forwards", type=int, default=3)
ap.add_argument("--batch", type=int, default=32)
args = ap.parse_args()
L, nb, n, bs = 2, 4, 3, args.batch
torch.manual_seed(1234)
with contextlib.redirect_stdout(io.StringIO()):
    net = TSNet(is_train=False, label_nc=L, n_blocks=nb, n_downsampling=3, n_source=n)
net.eval()
inp = synth.dataset_like_inputs(bs, L, n, seed=1234)
dev = {k: ([torch.from_numpy(a).cuda() for a in v] if isinstance(v, list) else torch.from_numpy(v).cuda())
       for k, v in inp.items() if k != "tar_img"}
with torch.no_grad():
    for _ in range(args.forwards):
        net.set_test_input(dev["src_img"], dev["src_lbl"], dev["src_bbox"], dev["tar_lbl"], dev["tar_bbox"])
        net.forward()
    torch.cuda.synchronize()
print("done")
