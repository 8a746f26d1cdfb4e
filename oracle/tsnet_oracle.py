"""CPU restatement of the TS-Net forward hot path  --  TEST INFRASTRUCTURE ONLY.

This file is the *checker* for the sm_100a kernels in wacv23_tsnet_b200/csrc.  Only tests/,
bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke() may import it; the
product path (wacv23_tsnet_b200/) never does and fails loudly without its CUDA library.

It restates, function by function, /root/reference/model/TSNet.py (face) and
/root/reference/model/TSNet_pose.py (pose) in plain fp32 torch-CPU / numpy ops: the whole
`is_train=False` forward plus the forward-only `is_train=True` branches (`train_extras`).
Each function cites the reference lines it follows.

Parity pin: the reference publishes no golden vectors.  The pin is the
reference itself, imported unmodified in the build container by oracle/make_golden.py, which
(1) checks this restatement bit-for-bit against reference `TSNet.forward()` on CPU (five inference
    configs + three train-mode configs: `python -m oracle.make_golden [train]`) and
(2) writes the fixtures in tests/golden/ that travel to the GPU box.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------
# integer / index work (numpy) -- must be bit-exact
# ------------------------------------------------------------------------------------------
def nearest_downsample_mask(bbox, h, w):
    """F.interpolate(bbox, (h, w), mode='nearest') of model/TSNet.py:322, :347.

    ATen nearest: src = min(floor(dst * (in/out)), in-1) computed with a float scale.
    bbox: numpy [B, H, W] (uint8 or float32) -> [B, h, w] same dtype.
    """
    bbox = np.asarray(bbox)
    H, W = bbox.shape[-2:]
    ys = np.minimum(np.floor(np.arange(h, dtype=np.float32) * np.float32(H / h)).astype(np.int64), H - 1)
    xs = np.minimum(np.floor(np.arange(w, dtype=np.float32) * np.float32(W / w)).astype(np.int64), W - 1)
    return bbox[..., ys[:, None], xs[None, :]]


def linspace_table(n):
    """torch.linspace(-1, 1, n) of model/TSNet.py:301-302, restated.

    ATen (CPU) builds it symmetrically with fused multiply-adds: fma(step, k, start) for k < n/2 and
    fma(-step, n-1-k, end) otherwise, step = fl32(2/(n-1)) -- the naive start + k*step is off by one ulp
    (SURVEY.md section 7.3-3b).  The fma is emulated exactly: the product of two fp32 values is exact in fp64.
    """
    step = np.float64(np.float32(2.0) / np.float32(n - 1))
    k = np.arange(n, dtype=np.float64)
    lo = (-1.0 + step * k).astype(np.float32)
    hi = (1.0 - step * (float(n - 1) - k)).astype(np.float32)
    return np.where(np.arange(n) < n // 2, lo, hi).astype(np.float32)


# ------------------------------------------------------------------------------------------
# network blocks (torch CPU fp32)
# ------------------------------------------------------------------------------------------
def coord_channels(x):
    """Encoder.coord_conv, model/TSNet.py:107-125: append x, y in [-1,1] and r = sqrt(x^2+y^2)."""
    bs, _, h, w = x.shape
    xx = torch.arange(w, dtype=x.dtype, device=x.device).view(1, 1, 1, w).expand(bs, 1, h, w)
    yy = torch.arange(h, dtype=x.dtype, device=x.device).view(1, 1, h, 1).expand(bs, 1, h, w)
    xx = 2 * (xx.float() / (w - 1)) - 1
    yy = 2 * (yy.float() / (h - 1)) - 1
    rr = torch.sqrt(torch.pow(xx, 2) + torch.pow(yy, 2))
    return torch.cat((x, xx, yy, rr), dim=1)


def _inorm(x):
    """nn.InstanceNorm2d defaults: affine=False, eps=1e-5, biased variance (model/TSNet.py:53,130)."""
    return F.instance_norm(x, eps=1e-5)


def resblock(x, sd, prefix):
    """ResnetBlock.forward, model/TSNet.py:10-49 (reflect padding, no dropout)."""
    y = F.pad(x, (1, 1, 1, 1), mode="reflect")
    y = F.conv2d(y, sd[prefix + "conv_block.1.weight"], sd[prefix + "conv_block.1.bias"])
    y = F.relu(_inorm(y))
    y = F.pad(y, (1, 1, 1, 1), mode="reflect")
    y = F.conv2d(y, sd[prefix + "conv_block.5.weight"], sd[prefix + "conv_block.5.bias"])
    return x + _inorm(y)


def encoder_forward(x, sd, n_blocks, n_down=3):
    """Encoder.forward (addcoords=True, debug=False, normalization=False), model/TSNet.py:52-105."""
    x = coord_channels(x)
    x = F.pad(x, (3, 3, 3, 3), mode="reflect")
    x = F.relu(_inorm(F.conv2d(x, sd["model.1.weight"], sd["model.1.bias"])))
    idx = 4
    for _ in range(n_down):
        x = F.relu(_inorm(F.conv2d(x, sd[f"model.{idx}.weight"], sd[f"model.{idx}.bias"], stride=2, padding=1)))
        idx += 3
    for _ in range(n_blocks):
        x = resblock(x, sd, f"model.{idx}.")
        idx += 1
    return x


def fuse_forward(src_fea, tar_fea, sd):
    """FuseNet.forward, model/TSNet.py:177-200."""
    x = torch.cat((src_fea, tar_fea), dim=1)
    x = resblock(x, sd, "model.0.")
    return F.conv2d(x, sd["conv.weight"], sd["conv.bias"])


def decoder_forward(pg, sg, sd, n_blocks, n_down=3):
    """Decoder.forward (return_fea=True), model/TSNet.py:128-174. Returns the final image only."""
    x = F.conv2d(torch.cat([pg, sg], dim=1), sd["map_conv.weight"], sd["map_conv.bias"])
    for n in range(n_blocks):
        x = resblock(x, sd, f"model{n}.0.")
    for i in range(n_down):
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
        x = F.pad(x, (1, 1, 1, 1), mode="reflect")
        x = F.conv2d(x, sd[f"model{n_blocks + i}.2.weight"], sd[f"model{n_blocks + i}.2.bias"])
        x = F.relu(_inorm(x))
    x = F.pad(x, (3, 3, 3, 3), mode="reflect")
    k = n_blocks + n_down
    return torch.tanh(F.conv2d(x, sd[f"model{k}.1.weight"], sd[f"model{k}.1.bias"]))


# ------------------------------------------------------------------------------------------
# transformation branch: mask-aware correlation -> softmax -> coordinate expectation -> warp
# ------------------------------------------------------------------------------------------
def get_grid(b, H, W):
    """TSNet.get_grid(normalize=True), model/TSNet.py:299-307: (x, y) channel order after flip(3)."""
    hr = torch.from_numpy(linspace_table(H))
    wr = torch.from_numpy(linspace_table(W))
    gy, gx = torch.meshgrid([hr, wr], indexing="ij")
    return torch.stack([gx, gy], -1).unsqueeze(0).repeat(b, 1, 1, 1).float()


def _down_mask(bbox, h, w):
    """CPU: the numpy restatement above (pinned against the reference).  CUDA tensors (bench.py's PyTorch-CUDA eager
    timing of this port only): the reference's own call, F.interpolate(mode='nearest') (model/TSNet.py:322, :347)."""
    if bbox.is_cuda:
        return F.interpolate(bbox, size=(h, w), mode="nearest")
    return torch.from_numpy(nearest_downsample_mask(bbox.numpy(), h, w))


def corr_warp(tar_fea, src_fea_list, tar_bbox, src_bbox_list, temperature=100.0):
    """model/TSNet.py:319-323, 336-366, 392.

    tar_fea [B,C,h,w], src_fea_list n x [B,C,h,w] fp32; tar_bbox [B,1,H,W] / src_bbox n x [B,1,H,W]
    (uint8 or float).  Returns (mean over sources of warped features [B,C,h,w], list of warp
    grids [B,h,w,2]).  Follows the reference op by op (two masked bmm, add, softmax, matmul).
    """
    b, c, h, w = tar_fea.shape
    dt = tar_fea.dtype  # fp32 = the reference; tests also evaluate this in fp64 as the "truth"
    t = F.normalize(tar_fea, p=2, dim=1).view(b, c, h * w).transpose(1, 2)
    mt = _down_mask(tar_bbox, h, w).view(b, 1, h * w).transpose(1, 2)
    grid2d = get_grid(b, h, w).view(b, h * w, 2).to(dt).to(tar_fea.device)
    if dt == torch.float64:
        mt = mt.to(dt)
    warped, grids = [], []
    for src_fea, src_bbox in zip(src_fea_list, src_bbox_list):
        s = F.normalize(src_fea, p=2, dim=1).view(b, c, h * w)
        ms = _down_mask(src_bbox, h, w).view(b, 1, h * w)
        if dt == torch.float64:
            ms = ms.to(dt)
        a = torch.bmm(t * mt, s * ms) + torch.bmm(t * (1.0 - mt), s * (1.0 - ms))
        p = F.softmax(temperature * a, dim=2)
        g = torch.matmul(p, grid2d).view(b, h, w, 2)
        warped.append(F.grid_sample(src_fea, g, align_corners=False))
        grids.append(g)
    return torch.stack(warped, dim=1).mean(dim=1), grids


def pose_composite(img, mean):
    """model/TSNet_pose.py:276-280, 416-417: keep columns 64:192, fill the rest with -mean/255."""
    fore = torch.zeros((256, 256), dtype=torch.float32)
    fore[:, 64:192] = 1
    fore = fore.view(1, 1, 256, 256)
    mask_img = torch.from_numpy(-np.asarray(mean, np.float32)).view(1, 3, 1, 1).repeat(1, 1, 256, 256) / 255.0
    return img * fore + mask_img * (1 - fore)


# ------------------------------------------------------------------------------------------
# whole forward
# ------------------------------------------------------------------------------------------
def to_torch_sd(sds):
    return {net: {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()} for net, sd in sds.items()}


def stage_inputs(inputs, use_prev=None):
    """set_train_input / set_test_input staging, model/TSNet.py:266-294: images /255 (unless use_prev[i]),
    bbox unsqueeze(1).  `inputs` is the dict produced by oracle/synth.py (numpy)."""
    st = {}
    st["src_img"] = []
    for i, x in enumerate(inputs["src_img"]):
        x = torch.from_numpy(x)
        st["src_img"].append(x if (use_prev is not None and use_prev[i]) else x / 255.0)
    st["src_lbl"] = [torch.from_numpy(x) for x in inputs["src_lbl"]]
    st["src_bbox"] = [torch.from_numpy(x).unsqueeze(1) for x in inputs["src_bbox"]]
    st["tar_lbl"] = torch.from_numpy(inputs["tar_lbl"])
    st["tar_bbox"] = torch.from_numpy(inputs["tar_bbox"]).unsqueeze(1)
    return st


def train_extras(src_img_list, tar_img, grids, pg_mean, sg_mean, pose_mean=None):
    """The train-only branches INSIDE forward(), restated op by op:
    model/TSNet.py:327-331 (reference statistics of the target image), :372-390 (image-space warp through
    unfold -> grid_sample -> fold, per-image re-normalisation, 10 * L1 warp loss), :402-405 (cosine alignment loss of the
    two branch means; face variant only) and model/TSNet_pose.py:395-396 (foreground compositing of the warped image).
    src_img_list / tar_img are the STAGED images (already /255 unless use_prev).  Returns (warp_src_img_list,
    loss_warp, loss_align or None)."""
    b = tar_img.shape[0]
    h, w = grids[0].shape[1:3]
    ref_mean = tar_img.view(b, 3, -1).mean(dim=2).view(b, 3, 1, 1)
    ref_std = tar_img.view(b, 3, -1).std(dim=2).view(b, 3, 1, 1)
    warp_list, losses = [], []
    for src_img, g in zip(src_img_list, grids):
        ori_h = src_img.shape[2]
        down = ori_h // h
        src_down = F.unfold(src_img, down, stride=down).view(b, -1, h, w)
        rec_down = F.grid_sample(src_down, g, align_corners=False).view(b, -1, h * w)
        warp = F.fold(rec_down, 256, down, stride=down)
        gen_mean = warp.view(b, 3, -1).mean(dim=2).view(b, 3, 1, 1)
        gen_std = warp.view(b, 3, -1).std(dim=2).view(b, 3, 1, 1)
        warp = (warp - gen_mean) / gen_std * ref_std + ref_mean
        if pose_mean is not None:
            warp = pose_composite(warp, pose_mean)
        warp_list.append(warp)
        losses.append(10 * F.l1_loss(warp, tar_img))
    loss_warp = sum(losses)
    loss_align = None
    if pose_mean is None:
        loss_align = 1 - (F.cosine_similarity(pg_mean, sg_mean, dim=1)).mean()
    return warp_list, loss_warp, loss_align


@torch.no_grad()
def tsnet_forward(sds, inputs, n_blocks, pose_mean=None, n_source=None, use_prev=None, train=False):
    """TSNet.forward for is_train=False, model/TSNet.py:309-407 (+ TSNet_pose.py:416-417 when pose_mean
    is given).  sds: dict of four state dicts (numpy or torch); inputs: synth dict (numpy).
    Returns dict of intermediates (torch CPU tensors)."""
    if isinstance(next(iter(sds["dec"].values())), np.ndarray):
        sds = to_torch_sd(sds)
    st = stage_inputs(inputs, use_prev)
    n = n_source if n_source is not None else len(st["src_img"])
    src_fea = [encoder_forward(torch.cat([st["src_img"][i], st["src_lbl"][i]], dim=1), sds["img_enc"], 9)
               for i in range(n)]
    tar_fea = encoder_forward(st["tar_lbl"], sds["lbl_enc"], 0)
    pg_mean, grids = corr_warp(tar_fea, src_fea, st["tar_bbox"], st["src_bbox"][:n])
    sg = [fuse_forward(src_fea[i], tar_fea, sds["fuse_net"]) for i in range(n)]
    sg_mean = torch.stack(sg, dim=1).mean(dim=1)
    rec = decoder_forward(pg_mean, sg_mean, sds["dec"], n_blocks)
    if pose_mean is not None:
        rec = pose_composite(rec, pose_mean)
    out = {"src_fea": src_fea, "tar_fea": tar_fea, "grids": grids, "pg_mean": pg_mean,
           "sg_mean": sg_mean, "rec_tar_img": rec}
    if train:  # is_train=True branches of forward(); needs inputs["tar_img"]
        tar_img = torch.from_numpy(inputs["tar_img"]) / 255.0
        warp_list, loss_warp, loss_align = train_extras(st["src_img"][:n], tar_img, grids, pg_mean, sg_mean, pose_mean)
        out.update(warp_src_img_list=warp_list, loss_warp=loss_warp, loss_align=loss_align)
    return out
