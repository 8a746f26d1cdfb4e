"""Deterministic synthetic weights / inputs for TS-Net parity work.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): imported by tests/, bench.py's data
generator and __graft_entry__.smoke().  Nothing here is on the product path.

Everything is produced from a counter-based integer hash (splitmix64) in numpy, with
only exact integer arithmetic before a single float conversion, so the same call yields
bit-identical arrays on every host (torch's CPU normal_()/rand() go through vectorised
libm paths whose last bit may differ between CPU generations; we avoid that on purpose:
the golden fixtures in tests/golden/ were produced in the build container and are
checked on the GPU box).

Distributions follow the reference's own:
  * conv weights ~ N(0, 0.02^2), bias 0           (model/networks.py:67-103 `init_weights`)
    here an Irwin-Hall(12) approximation of the normal -- same mean / variance.
  * quick_start inputs                             (quick_start1.py:18-29)
  * FaceForensics-like inputs                      (train_face.py:29 IMG_MEAN,
    dataset/dataset_video_face.py:179-193 rectangular bbox, utils/misc.py:50-67 one-hot labels)
"""
import zlib

import numpy as np

IMG_MEAN = np.array((101.84807705937696, 112.10832843463207, 111.65973036298041), dtype=np.float32)

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    """x: uint64 array of counters -> uint64 array of well-mixed bits."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _stream(tag, seed):
    """64-bit stream id from a string tag and an integer seed."""
    h = zlib.crc32(tag.encode()) & 0xFFFFFFFF
    return np.uint64(((int(seed) & 0xFFFFFFFF) << 32) | h)


def bits(shape, tag, seed=0, salt=0):
    n = int(np.prod(shape)) if len(shape) else 1
    ctr = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        base = _splitmix64(np.array([_stream(tag, seed) + np.uint64(salt)], dtype=np.uint64))[0]
        return _splitmix64(ctr ^ base).reshape(shape)


def uniform(shape, tag, seed=0, salt=0):
    """U[0,1) with 24 random bits (exactly representable in fp32)."""
    b = bits(shape, tag, seed, salt)
    return ((b >> np.uint64(40)).astype(np.float64) / float(1 << 24)).astype(np.float32)


def randint(shape, high, tag, seed=0):
    b = bits(shape, tag, seed)
    return ((b >> np.uint64(33)) % np.uint64(high)).astype(np.int64)


def normal(shape, std, tag, seed=0):
    """Irwin-Hall(12) ~ N(0,1), scaled by std. Exact integer sum, one fp64 scale, one cast."""
    acc = np.zeros(int(np.prod(shape)), dtype=np.int64)
    for k in range(12):
        acc += (bits((acc.size,), tag, seed, salt=k + 1) >> np.uint64(40)).astype(np.int64)
    x = acc.astype(np.float64) / float(1 << 24) - 6.0
    return (x * float(std)).astype(np.float32).reshape(shape)


# ----------------------------------------------------------------------------------------------
# state_dict layout of the four generator nets (keys measured from the reference:
# model/TSNet.py:52-200).  ngf=64, n_downsampling=3 is the only geometry FuseNet(ngf=1024) allows.
# ----------------------------------------------------------------------------------------------
def encoder_shapes(input_nc, n_blocks, ngf=64, n_down=3, addcoords=True):
    cin = input_nc + (3 if addcoords else 0)
    shapes = {"model.1.weight": (ngf, cin, 7, 7), "model.1.bias": (ngf,)}
    idx = 4
    for i in range(n_down):
        m = 2 ** i
        shapes[f"model.{idx}.weight"] = (ngf * m * 2, ngf * m, 3, 3)
        shapes[f"model.{idx}.bias"] = (ngf * m * 2,)
        idx += 3
    dim = ngf * 2 ** n_down
    for _ in range(n_blocks):
        for j in (1, 5):
            shapes[f"model.{idx}.conv_block.{j}.weight"] = (dim, dim, 3, 3)
            shapes[f"model.{idx}.conv_block.{j}.bias"] = (dim,)
        idx += 1
    return shapes


def fuse_shapes(ngf=1024):
    s = {}
    for j in (1, 5):
        s[f"model.0.conv_block.{j}.weight"] = (ngf, ngf, 3, 3)
        s[f"model.0.conv_block.{j}.bias"] = (ngf,)
    s["conv.weight"] = (ngf // 2, ngf, 1, 1)
    s["conv.bias"] = (ngf // 2,)
    return s


def decoder_shapes(n_blocks, ngf=64, n_down=3, output_nc=3):
    dim = ngf * 2 ** n_down
    s = {"map_conv.weight": (dim, dim * 2, 1, 1), "map_conv.bias": (dim,)}
    for n in range(n_blocks):
        for j in (1, 5):
            s[f"model{n}.0.conv_block.{j}.weight"] = (dim, dim, 3, 3)
            s[f"model{n}.0.conv_block.{j}.bias"] = (dim,)
    for i in range(n_down):
        m = 2 ** (n_down - i)
        s[f"model{n_blocks + i}.2.weight"] = (ngf * m // 2, ngf * m, 3, 3)
        s[f"model{n_blocks + i}.2.bias"] = (ngf * m // 2,)
    s[f"model{n_blocks + n_down}.1.weight"] = (output_nc, ngf, 7, 7)
    s[f"model{n_blocks + n_down}.1.bias"] = (output_nc,)
    return s


def make_state_dicts(label_nc, n_blocks, seed=1234, bias_std=0.0):
    """Four numpy state dicts {'img_enc','lbl_enc','fuse_net','dec'} -> {key: fp32 array}.

    bias_std=0 reproduces the reference init (bias 0); tests also use bias_std>0 so that the
    bias path (present in trained checkpoints) is exercised.
    """
    nets = {
        "img_enc": encoder_shapes(3 + label_nc, 9),
        "lbl_enc": encoder_shapes(label_nc, 0),
        "fuse_net": fuse_shapes(),
        "dec": decoder_shapes(n_blocks),
    }
    out = {}
    for net, shapes in nets.items():
        sd = {}
        for key, shp in shapes.items():
            tag = f"{net}/{key}"
            if key.endswith("weight"):
                sd[key] = normal(shp, 0.02, tag, seed)
            else:
                sd[key] = normal(shp, bias_std, tag, seed) if bias_std > 0 else np.zeros(shp, np.float32)
        out[net] = sd
    return out


# ----------------------------------------------------------------------------------------------
# inputs
# ----------------------------------------------------------------------------------------------
def quick_start_inputs(bs, label_nc=2, n_source=3, seed=1234, size=256):
    """quick_start1.py:18-29 distributions: img U[0,1) (then /255 by set_*_input), labels and
    bbox independent bits per pixel (NOT one-hot, NOT rectangles), all float32."""
    d = {"src_img": [], "src_lbl": [], "src_bbox": []}
    for i in range(n_source):
        d["src_img"].append(uniform((bs, 3, size, size), f"qs/src_img{i}", seed))
        d["src_lbl"].append(randint((bs, label_nc, size, size), 2, f"qs/src_lbl{i}", seed).astype(np.float32))
        d["src_bbox"].append(randint((bs, size, size), 2, f"qs/src_bbox{i}", seed).astype(np.float32))
    d["tar_img"] = uniform((bs, 3, size, size), "qs/tar_img", seed)
    d["tar_lbl"] = randint((bs, label_nc, size, size), 2, "qs/tar_lbl", seed).astype(np.float32)
    d["tar_bbox"] = randint((bs, size, size), 2, "qs/tar_bbox", seed).astype(np.float32)
    return d


def _rect_bbox(bs, size, tag, seed):
    """Axis-aligned rectangle covering roughly 30-70 % of the frame, uint8 {0,1}
    (dataset/dataset_video_face.py:179-193 `get_bbox_image` produces exactly such rectangles)."""
    u = uniform((bs, 4), tag, seed)
    out = np.zeros((bs, size, size), np.uint8)
    for b in range(bs):
        hh = int(size * (0.55 + 0.30 * u[b, 0]))
        ww = int(size * (0.55 + 0.30 * u[b, 1]))
        y0 = int((size - hh) * u[b, 2])
        x0 = int((size - ww) * u[b, 3])
        out[b, y0:y0 + hh, x0:x0 + ww] = 1
    return out


def _onehot_labels(bs, label_nc, size, tag, seed, blocky=8):
    """One-hot float32 labels (utils/misc.py:50-67 `vl2ch`): a blocky random class map with
    class 0 (background) dominant, like a sparse edge / skeleton map."""
    cls = randint((bs, size // blocky, size // blocky), 8 * label_nc, tag, seed)
    cls = np.where(cls < label_nc, cls, 0)
    cls = np.repeat(np.repeat(cls, blocky, axis=1), blocky, axis=2)
    oh = np.zeros((bs, label_nc, size, size), np.float32)
    for c in range(label_nc):
        oh[:, c] = (cls == c)
    return oh


def dataset_like_inputs(bs, label_nc=2, n_source=3, seed=1234, size=256, pose=False):
    """FaceForensics / Youtube-dance shaped inputs: mean-subtracted
    BGR in 0-255 units, one-hot labels, uint8 rectangular bbox."""
    def img(tag):
        x = uniform((bs, 3, size, size), tag, seed) * 255.0 - IMG_MEAN.reshape(1, 3, 1, 1)
        x = x.astype(np.float32)
        if pose:  # person occupies the centre 128 columns (dataset_video_pose.py:162-168)
            x[:, :, :, :size // 4] = -IMG_MEAN.reshape(1, 3, 1, 1)
            x[:, :, :, 3 * size // 4:] = -IMG_MEAN.reshape(1, 3, 1, 1)
        return x
    d = {"src_img": [], "src_lbl": [], "src_bbox": []}
    for i in range(n_source):
        d["src_img"].append(img(f"ds/src_img{i}"))
        d["src_lbl"].append(_onehot_labels(bs, label_nc, size, f"ds/src_lbl{i}", seed))
        d["src_bbox"].append(_rect_bbox(bs, size, f"ds/src_bbox{i}", seed))
    d["tar_img"] = img("ds/tar_img")
    d["tar_lbl"] = _onehot_labels(bs, label_nc, size, "ds/tar_lbl", seed)
    d["tar_bbox"] = _rect_bbox(bs, size, "ds/tar_bbox", seed)
    return d


def checksum(arr):
    """Order-sensitive 64-bit checksum of raw bytes (for 'same weights on this host?' guards)."""
    a = np.ascontiguousarray(arr).view(np.uint8)
    return int(zlib.crc32(a.tobytes())) | (int(zlib.adler32(a.tobytes())) << 32)
