"""Error-budget study for DESIGN.md section 11 item 5 (CPU only, test infrastructure: uses the oracle).

Question: the 3x3 convolutions of the 32x32 layers are ~28 of the 46 ms of a forward and are power-capped at the
3-term tensor rate; Winograd F(2x2, 3x3) needs 2.25x fewer MACs.  Would its extra rounding error still fit the parity
tolerances (rec_tar_img 2e-3, warp grids 1e-4) when the GEMM operands are 22-bit (fp16 hi + lo) values?

Emulation: input / filter / output transforms in fp32 (as a kernel would do them), the 16 channel contractions with
operands rounded to 22 significant bits and exact accumulation (fp64), result rounded to fp32 -- the same model of the
tensor-core path that gives 2.7e-4 .. 3.4e-4 for the DIRECT convolution (last row).
Usage:  python tools/winograd_study.py [config ...]      (default: quickstart_bs1 face_bs1_nb4)
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import make_golden as MG, synth, tsnet_oracle as O  # noqa: E402

BT = torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=torch.float32)
G = torch.tensor([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], dtype=torch.float32)
AT = torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=torch.float32)


def q22(x):
    """round to 22 significant bits (what an fp16 hi + lo pair holds while lo is normal)"""
    m, e = torch.frexp(x.double())
    return torch.ldexp(torch.round(m * (1 << 22)) / (1 << 22), e)


def conv3x3_direct_q(xp, w, b):
    """direct 3x3 convolution on an already padded input, 22-bit operands, exact accumulation, fp32 result"""
    y = F.conv2d(q22(xp), q22(w)).float()
    return y + b.view(1, -1, 1, 1)


def conv3x3_winograd_q(xp, w, b):
    """Winograd F(2x2, 3x3) on an already padded input [B, C, H+2, W+2] (H, W even)"""
    B_, C, Hp, Wp = xp.shape
    H, W = Hp - 2, Wp - 2
    d = xp.unfold(2, 4, 2).unfold(3, 4, 2)                       # [B, C, H/2, W/2, 4, 4]
    V = torch.einsum("ik,bcxykl,jl->bcxyij", BT, d, BT)            # fp32 input transform
    U = torch.einsum("ik,ockl,jl->ocij", G, w, G)                  # fp32 filter transform
    M = torch.einsum("ocij,bcxyij->boxyij", q22(U), q22(V)).float()  # 16 contractions over channels
    Y = torch.einsum("ik,boxykl,jl->boxyij", AT, M, AT)            # fp32 output transform -> 2 x 2 per tile
    y = Y.permute(0, 1, 2, 4, 3, 5).reshape(B_, w.shape[0], H, W)
    return y + b.view(1, -1, 1, 1)


def make_resblock(conv):
    def resblock(x, sd, prefix):
        y = F.pad(x, (1, 1, 1, 1), mode="reflect")
        y = conv(y, sd[prefix + "conv_block.1.weight"], sd[prefix + "conv_block.1.bias"])
        y = F.relu(O._inorm(y))
        y = F.pad(y, (1, 1, 1, 1), mode="reflect")
        y = conv(y, sd[prefix + "conv_block.5.weight"], sd[prefix + "conv_block.5.bias"])
        return x + O._inorm(y)
    return resblock


def main():
    names = sys.argv[1:] or ["quickstart_bs1", "face_bs1_nb4"]
    torch.set_num_threads(os.cpu_count())
    # sanity: the two emulations agree with an exact convolution on one layer
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 64, 18, 18, generator=g)
    w = torch.randn(32, 64, 3, 3, generator=g) * 0.02
    b = torch.zeros(32)
    ref = F.conv2d(x.double(), w.double())
    rel = lambda y: float((y.double() - ref).abs().max() / ref.abs().max())
    print(f"single layer, relative to max|ref|: direct(22 bit) {rel(conv3x3_direct_q(x, w, b)):.2e}, "
          f"winograd(22 bit) {rel(conv3x3_winograd_q(x, w, b)):.2e}, fp32 direct {rel(F.conv2d(x, w)):.2e}")
    orig = O.resblock
    for name in names:
        cfg = MG.CONFIGS[name]
        gold = np.load(os.path.join(MG.GOLDEN_DIR, name + ".npz"))
        sds, inputs = MG.build_case(cfg)
        mean = synth.IMG_MEAN if cfg["pose"] else None
        for tag, conv in (("direct, 22-bit operands (all resblock convs)", conv3x3_direct_q),
                          ("Winograd F(2x2,3x3), 22-bit operands (all resblock convs)", conv3x3_winograd_q)):
            O.resblock = make_resblock(conv)
            try:
                out = O.tsnet_forward(sds, inputs, cfg["n_blocks"], pose_mean=mean)
            finally:
                O.resblock = orig
            ierr = float((out["rec_tar_img"] - torch.from_numpy(gold["rec_tar_img"])).abs().max())
            gerr = float((torch.stack(out["grids"]) - torch.from_numpy(gold["grids"])).abs().max())
            print(f"{name}: {tag}: rec_tar_img max|err| {ierr:.2e} (tol 2e-3), warp grid max|err| {gerr:.2e} (tol 1e-4)")


if __name__ == "__main__":
    main()
