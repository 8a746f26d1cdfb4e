"""TS-Net forward benchmark (BASELINE.json metric: forward frames/sec @256x256, n_src=3; corr+warp HBM GB/s).

  python bench.py --gpus N --steps K --warmup W            # B200 arm (one process per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the reference algorithm on the host cores

A "step" = one TSNet.forward() over one batch of synthetic FaceForensics-shaped input (SURVEY.md section 8d config 2:
bs=32 per GPU, label_nc=2, n_source=3, n_blocks=4, uint8 rectangular bbox).  Weak scaling: every rank owns 32 rows.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "tsnet_forward_frames_per_sec_256x256_nsrc3"
ALGO_BYTES_CORR_PER_FRAME = {1: 6299648, 3: 10502144, 5: 14704640, 8: 21008384}  # SURVEY.md section 8d / BASELINE.md section 3


T0 = time.time()


def log(msg):
    print(f"[bench +{time.time() - T0:6.1f}s] {msg}", file=sys.stderr, flush=True)


def csrc_sha1():
    """Hash of the kernel sources: ties a committed ncu capture to the code it was taken from."""
    import glob
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(ROOT, "wacv23_tsnet_b200", "csrc")
    for f in sorted(glob.glob(os.path.join(d, "*.cu")) + glob.glob(os.path.join(d, "*.cuh")) +
                    glob.glob(os.path.join(d, "*.h"))):
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()


def load_traffic():
    """DRAM bytes per launch from the committed ncu captures (profiles/traffic.json).  The file records the hash of
    the kernel sources the captures were taken from; when csrc/ has changed since, the numbers are stale and `traffic`
    is reported as null instead (with the reason)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.isfile(p):
        return {}
    d = json.load(open(p))
    if d.get("csrc_sha1") != csrc_sha1():
        return {"_stale": f"profiles/traffic.json was captured from csrc sha1 {str(d.get('csrc_sha1'))[:12]}, "
                          f"current sources are {csrc_sha1()[:12]}: traffic not reported"}
    return d


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.rows = gpu_index, None, []
        self.t0, self.t1 = None, None   # only samples taken inside [t0, t1] (the timed region) are reported

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        lo = self.t0 if self.t0 is not None else 0.0
        hi = (self.t1 if self.t1 is not None else time.time()) + 0.25   # a 200 ms sample may land just after the end
        rows = [r for t, r in self.rows if lo <= t <= hi] or [r for _, r in self.rows]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(bs, label_nc, n_source, seed, pose=False):
    from oracle import synth  # data generator only
    return synth.dataset_like_inputs(bs, label_nc, n_source, seed=seed, pose=pose)


def usable_cpus():
    """CPUs this process may actually use: affinity mask, capped by the cgroup CPU quota if one is set."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except (OSError, ValueError):
        pass
    return n


def cpu_forward_fps(sds_np, label_nc, n_blocks, n_source, bs, steps, warmup):
    """Reference algorithm (oracle port of model/TSNet.py:309-407) on the host cores; returns (frames/s, threads).
    The thread count is the best of a short ascending sweep up to all usable CPUs (a bs=1 forward on 128 threads is
    60x SLOWER than on 16 on the GPU box -- oneDNN oversubscription -- and would be an unfairly weak baseline)."""
    from oracle import tsnet_oracle as O
    inp = make_inputs(bs, label_nc, n_source, seed=4321)
    sds = O.to_torch_sd(sds_np)
    ncpu = usable_cpus()
    best_t, best_n = None, None
    for nt in sorted({min(ncpu, c) for c in (8, 16, 32, 64, ncpu)}):
        torch.set_num_threads(nt)
        O.tsnet_forward(sds, inp, n_blocks)
        t0 = time.perf_counter()
        O.tsnet_forward(sds, inp, n_blocks)
        dt = time.perf_counter() - t0
        log(f"cpu baseline: {nt} threads -> {dt:.2f} s / forward (usable cpus {ncpu}, os.cpu_count {os.cpu_count()})")
        if best_t is None or dt < best_t:
            best_t, best_n = dt, nt
        elif dt > 1.5 * best_t:
            break
    torch.set_num_threads(best_n)
    for _ in range(warmup):
        O.tsnet_forward(sds, inp, n_blocks)
    log("cpu baseline: warm-up done")
    t0 = time.perf_counter()
    for _ in range(steps):
        O.tsnet_forward(sds, inp, n_blocks)
    dt = time.perf_counter() - t0
    return bs * steps / dt, torch.get_num_threads(), dt / steps


def torch_cuda_eager_fps(sds_np, label_nc, n_blocks, n_source, bs, steps=5, warmup=2):
    """The reference's stock op sequence (oracle port of model/TSNet.py:309-407: F.conv2d / instance_norm / bmm / softmax
    / grid_sample ...) executed EAGERLY ON THE GPU by PyTorch + cuDNN/cuBLAS -- what running the unmodified reference
    on this B200 costs (north star: 'the reference PyTorch-CUDA forward').  cudnn.benchmark = True as the reference's
    demo scripts set it (demo/demo_face.py:113-114).  Two variants: PyTorch's default (cuDNN convolutions may use
    TF32 -- NOT the fp32 function) and allow_tf32 = False (true fp32)."""
    from oracle import tsnet_oracle as O
    dev = torch.device("cuda")
    inp = make_inputs(bs, label_nc, n_source, seed=4321)
    sds = {net: {k: v.to(dev) for k, v in sd.items()} for net, sd in O.to_torch_sd(sds_np).items()}
    src_img = [torch.from_numpy(x).to(dev) / 255.0 for x in inp["src_img"]]
    src_lbl = [torch.from_numpy(x).to(dev) for x in inp["src_lbl"]]
    src_bbox = [torch.from_numpy(x).to(dev).float().unsqueeze(1) for x in inp["src_bbox"]]
    tar_lbl = torch.from_numpy(inp["tar_lbl"]).to(dev)
    tar_bbox = torch.from_numpy(inp["tar_bbox"]).to(dev).float().unsqueeze(1)

    @torch.no_grad()
    def fwd():
        src_fea = [O.encoder_forward(torch.cat([src_img[i], src_lbl[i]], dim=1), sds["img_enc"], 9)
                   for i in range(n_source)]
        tar_fea = O.encoder_forward(tar_lbl, sds["lbl_enc"], 0)
        pg_mean, _ = O.corr_warp(tar_fea, src_fea, tar_bbox, src_bbox)
        sg = [O.fuse_forward(src_fea[i], tar_fea, sds["fuse_net"]) for i in range(n_source)]
        sg_mean = torch.stack(sg, dim=1).mean(dim=1)
        return O.decoder_forward(pg_mean, sg_mean, sds["dec"], n_blocks)

    out = {}
    old = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.benchmark = True
    try:
        for tag, tf32 in (("default_tf32_convs", True), ("fp32_no_tf32", False)):
            torch.backends.cudnn.allow_tf32 = tf32
            for _ in range(warmup):
                fwd()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                fwd()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[tag] = {"frames_per_s": bs / (ms * 1e-3), "ms_per_step": ms}
            log(f"torch-cuda eager port ({tag}): {ms:.1f} ms / step of bs={bs}")
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    del sds, src_img, src_lbl
    torch.cuda.empty_cache()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import synth
    L, nb, n = 2, 4, 3
    sds = synth.make_state_dicts(L, nb, seed=1234)
    bs = 1
    fps, cores, spf = cpu_forward_fps(sds, L, nb, n, bs, max(1, args.steps), max(1, args.warmup))
    line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": spf * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": "FaceForensics config (bs per step bounded to 1 on CPU), 256x256, label_nc=2, "
                                   "n_source=3, n_blocks=4", "per_step_batch": bs},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} forwards of bs={bs} (oracle/tsnet_oracle.py, torch CPU fp32, "
                                       "bit-exact restatement of the reference forward)"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_b200(args):
    from wacv23_tsnet_b200 import dist as D
    from wacv23_tsnet_b200 import ops
    from wacv23_tsnet_b200.model.TSNet import TSNet
    from wacv23_tsnet_b200.model.TSNet_pose import TSNet as TSNetPose
    import torch.distributed as tdist

    rank, local_rank, world = D.init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L, nb, n, bs = (25 if args.pose else 2), args.n_blocks, args.n_source, args.batch
    peaks = load_peaks()

    import contextlib
    import io
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(io.StringIO()):
        if args.pose:  # BASELINE.json config 3 (Youtube-dance): label_nc=25, foreground compositing
            from oracle.synth import IMG_MEAN
            net = TSNetPose(is_train=False, label_nc=L, n_blocks=nb, n_downsampling=3, n_source=n, use_mask=True,
                            mean=IMG_MEAN, math_mode=args.math, winograd=args.winograd)
        else:
            net = TSNet(is_train=False, label_nc=L, n_blocks=nb, n_downsampling=3, n_source=n, math_mode=args.math,
                        winograd=args.winograd)
    D.broadcast_generator(net, src=0)
    net.eval()
    if args.wino_chunk_kb > 0:
        net._engine.wino_chunk_kb = args.wino_chunk_kb
    net._engine.direct_stem = bool(args.direct_stem)
    if args.bridge_variant >= 0:
        net._engine.bridge_variant = args.bridge_variant
    log("model built")

    from oracle.synth import IMG_MEAN as _MEAN
    mean4 = np.asarray(_MEAN, dtype=np.float32).reshape(1, 3, 1, 1)

    def rank_inputs(r):
        """Synthetic batch of rank r.  Frames are 8-bit pixels (what a video decoder delivers); the fp32 tensors the
        reference's datasets would emit are exactly `u8 - mean`, so the resident run and the end-to-end run process
        the SAME frames."""
        d = make_inputs(bs, L, n, seed=1234 + r, pose=args.pose)
        u8 = [np.clip(np.rint(a + mean4), 0, 255).astype(np.uint8) for a in d["src_img"]]
        d["src_img"] = [a.astype(np.float32) - mean4 for a in u8]
        return d, u8
    inp, src_u8 = rank_inputs(rank)
    host = {k: ([torch.from_numpy(a).pin_memory() for a in v] if isinstance(v, list) else torch.from_numpy(v).pin_memory())
            for k, v in inp.items() if k != "tar_img"}
    devin = {k: ([t.to(dev) for t in v] if isinstance(v, list) else v.to(dev)) for k, v in host.items()}
    # compact host formats for the end-to-end path: uint8 BGR frames (+ the dataset mean applied
    # in the stem loader) and uint8 class-index label maps (vl2ch evaluated in the loader); bit-identical forward
    # (tests/test_parity_gpu.py::test_uint8_images_give_identical_forward..., ::test_classmap_labels_give_identical_forward)
    host_c = dict(src_img=[torch.from_numpy(a).pin_memory() for a in src_u8],
                  src_lbl=[torch.from_numpy(a.argmax(1).astype(np.uint8)).pin_memory() for a in inp["src_lbl"]],
                  src_bbox=host["src_bbox"], tar_lbl=torch.from_numpy(inp["tar_lbl"].argmax(1).astype(np.uint8)).pin_memory(),
                  tar_bbox=host["tar_bbox"])
    net.set_image_mean(_MEAN)
    nbytes = lambda d: sum(t.numel() * t.element_size() for v in d.values() for t in (v if isinstance(v, list) else [v]))
    h2d_fp32 = nbytes(host)
    if args.e2e_inputs == "compact":
        host = host_c
    h2d = nbytes(host)
    d2h = bs * 3 * 256 * 256 * 4
    log(f"inputs staged (e2e host bytes per step: {h2d}; reference fp32 formats would be {h2d_fp32})")

    def step_resident():
        net.set_test_input(devin["src_img"], devin["src_lbl"], devin["src_bbox"], devin["tar_lbl"], devin["tar_bbox"])
        net.forward()

    def step_e2e():
        net.set_test_input([t.cuda(non_blocking=True) for t in host["src_img"]],
                           [t.cuda(non_blocking=True) for t in host["src_lbl"]],
                           [t.cuda(non_blocking=True) for t in host["src_bbox"]],
                           host["tar_lbl"].cuda(non_blocking=True), host["tar_bbox"].cuda(non_blocking=True))
        net.forward()
        return net.rec_tar_img.cpu()

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.kernel_launches()
        if profile:
            ops.PROFILE = {}
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        prof, ops.PROFILE = ops.PROFILE, None
        ms = D.max_over_ranks(ev0.elapsed_time(ev1), dev)
        return ms, ops.kernel_launches() - l0, prof

    import gc
    with torch.no_grad():
        # the sampler process is forked and nvidia-smi initialises BEFORE the warm-up, so that neither overlaps the
        # timed region; only the samples taken inside the timed region are reported
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        for _ in range(max(args.warmup, 3)):
            step_resident()
        torch.cuda.synchronize()
        log("warm-up done")
        gc.collect()
        gc.disable()   # no collector pauses between launches inside the timed regions
        sampler.mark_begin()
        ms, launches, prof = timed(step_resident, args.steps, profile=True)
        sampler.mark_end()
        log(f"timed region done: {ms / args.steps:.2f} ms/step")
        clocks = sampler.stop() if rank == 0 else None
        for _ in range(2):
            step_e2e()
        if args.e2e_mode == "sync":
            ms_e2e, _, _ = timed(step_e2e, args.steps)
        else:
            # public serving API: FramePipeline overlaps H2D of batch i+1 / forward of batch i / D2H of batch i-1;
            # every step still copies its own inputs from pinned host memory and reads its own result back.
            from wacv23_tsnet_b200.pipeline import FramePipeline
            pipe = FramePipeline(net)

            def run_pipe():
                for _ in pipe.run((host for _ in range(args.steps)), copy_out=False):
                    pass
            run_pipe_once = lambda: [None for _ in pipe.run([host, host, host], copy_out=False)]
            run_pipe_once()
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            run_pipe()
            ev1.record()
            barrier()
            ms_e2e = D.max_over_ranks(ev0.elapsed_time(ev1), dev)
        log(f"e2e region done: {ms_e2e / args.steps:.2f} ms/step")
    gc.enable()

    # ---- N > 1: the batch split must be EXACT, not only fast.  Rank 0 regenerates rank 1's inputs, recomputes its
    # first two rows on its own GPU and demands bit-identical frames (outside the timed region).
    shard_check = None
    if world > 1:
        with torch.no_grad():
            step_resident()
            mine = net.rec_tar_img[:2].contiguous()
            if rank == 1:
                tdist.send(mine, dst=0)
            if rank == 0:
                theirs = torch.empty_like(mine)
                tdist.recv(theirs, src=1)
                inp1, _ = rank_inputs(1)
                t = lambda a: torch.from_numpy(a[:2]).to(dev)
                net.set_test_input([t(a) for a in inp1["src_img"]], [t(a) for a in inp1["src_lbl"]],
                                   [t(a) for a in inp1["src_bbox"]], t(inp1["tar_lbl"]), t(inp1["tar_bbox"]))
                net.forward()
                same = bool(torch.equal(net.rec_tar_img, theirs))
                shard_check = {"what": "rows 0-1 of rank 1's shard recomputed on rank 0 (bs=2 instead of inside bs="
                                       f"{bs}) vs the frames rank 1 produced", "bit_identical": same}
                log(f"shard check: bit_identical={same}")
                assert same, "sharded forward differs from the single-GPU forward of the same rows"
        barrier()

    frames = bs * world * args.steps
    value = frames / (ms * 1e-3)
    e2e_value = frames / (ms_e2e * 1e-3)

    # ---- per-kernel rooflines from CUDA events recorded around the launches inside the timed region
    kern = {}
    for key, evs in (prof or {}).items():
        tms = [a.elapsed_time(b) for a, b in evs]
        kern[key] = (sum(tms), len(tms))
    total_ms = sum(v[0] for v in kern.values()) or 1.0
    roof, roof_corr, shares = None, None, {}
    if kern:
        for key, (t, c) in sorted(kern.items(), key=lambda kv: -kv[1][0])[:14]:
            shares[str(key)] = {"ms_per_step": t / args.steps, "launches_per_step": c / args.steps,
                                "share_of_kernel_time": t / total_ms}
        conv_keys = [k for k in kern if k[0] in ("conv_gemm", "wino_gemm")]
        dom = max(conv_keys, key=lambda k: kern[k][0])
        t, c = kern[dom]
        kname, kind, X, Hh, Ww, Cin_eff, Cout, taps = dom
        conv_flops = 2.0 * X * Hh * Ww * Cout * taps * Cin_eff  # fp32-conv flops of the layer one launch belongs to
        # Winograd F(2x2,3x3): the launch executes 16 plane GEMMs over H/2 x W/2 tiles = 16/36 of the conv's MACs;
        # ALGORITHMIC flops of the kernel = what the batched GEMM computes
        flops = conv_flops * (16.0 / 36.0) if kname == "wino_gemm" else conv_flops
        ach = flops / (t / c * 1e-3) / 1e12
        peak = peaks["bf16_sustained"]
        traffic = load_traffic()
        tkey = {"conv_gemm": "conv_gemm_512x512_3x3_x96", "wino_gemm": "wino_gemm_512x512_x96"}[kname]
        conv_traffic = traffic.get(tkey) if (kind, X, Hh, Cin_eff, Cout) == ("3x3", 96, 32, 512, 512) else None
        roof = {"kernel": f"{kname} {kind} {Cin_eff}->{Cout} @{Hh}x{Ww} x{X} samples", "bound": "tensor",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": conv_traffic,
                "traffic_source": traffic.get("_stale") or traffic.get("source"),
                "algorithmic_flops_per_launch": flops,
                "peak_source": peaks["source"] + ", bf16 sustained",
                "note": "fp32-faithful mode issues 3 fp16 MMAs per algorithmic MAC (hi*hi + hi*lo + lo*hi): "
                        "frac <= 0.333 by construction; tensor-pipe utilisation = 3 x frac" if net._engine.mode.split
                else "single-pass 16-bit operands"}
        if kname == "wino_gemm":
            # the whole Winograd convolution layer = input transform pass + batched GEMM + output transform pass
            n_l = c / args.steps
            # transform passes of ALL Winograd layers of the step (fused bridge passes, or separate input / output
            # passes at the edges of a Winograd chain), spread over the Winograd GEMM launches of the step
            t_pass = sum(v[0] for k, v in kern.items()
                         if (k[0] == "build_taps" and k[1] == 4) or k[0] in ("wino_output", "wino_bridge")) / args.steps
            n_gemm = sum(v[1] for k, v in kern.items() if k[0] == "wino_gemm") / args.steps
            layer_ms = t / c + t_pass / n_gemm
            roof["winograd_layer"] = {
                "what": "per 3x3 conv layer of this shape: GEMM launch + the transform passes of the step (bridge = output"
                        " transform + InstanceNorm + input transform in one pass; separate passes at the chain edges) "
                        "averaged over all Winograd GEMM launches of the step",
                "ms": layer_ms, "gemm_ms": t / c, "passes_ms_per_layer": t_pass / n_gemm, "launches_per_step": n_l,
                "conv_equivalent_tflops": conv_flops / (layer_ms * 1e-3) / 1e12,
                "conv_equivalent_frac_of_peak": conv_flops / (layer_ms * 1e-3) / 1e12 / peak}
        # every kernel of the chain that ran: prepare, operand pass(es), [norm finalisation], tiles, finish
        chain = tuple(k for k in ("corr_prepare", "corr_operands", "corr_norms", "l2norm_split", "corr_tiles",
                                  "corr_finish") if (k,) in kern)
        if all(k in chain for k in ("corr_prepare", "corr_tiles", "corr_finish")):
            # the reference's corr+warp (model/TSNet.py:319-366, :392) = ALL four kernels of the chain: mask sort / work
            # list, F.normalize + operand split, tensor-core similarity tiles + softmax, merge + grid_sample + mean
            per = {k: kern[(k,)][0] / args.steps * 1e3 for k in chain}       # us per forward
            t_us = sum(per.values())
            byts = ALGO_BYTES_CORR_PER_FRAME.get(n, 4 * 512 * 1024 * (n + 2) + 4 * 1024 * (n + 1)) * bs
            ach = byts / (t_us * 1e-6) / 1e9
            tiles_ach = byts / (per["corr_tiles"] * 1e-6) / 1e9
            roof_corr = {"kernel": "correlation chain: corr_prepare (mask class sort + work list, one launch) + operand "
                                   "pass(es) (un-normalised hi/lo rows + reciprocal norms; the sources' rows come from the"
                                   " img_enc bridge pass when it is on) + corr_tiles (tcgen05 similarity, F.normalize as a"
                                   " row x column scale, softmax states) + corr_finish (merge + grid_sample + source mean "
                                   "written as map_conv's operand)", "kernels": list(chain), "bound": "hbm", "achieved": ach,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                         "traffic": traffic.get("corr_chain_b32_n3") if (bs, n) == (32, 3) else None,
                         "peak_source": peaks["source"], "algorithmic_bytes_per_launch": byts,
                         "us_per_forward": per,
                         "corr_tiles_alone": {"achieved": tiles_ach, "frac": tiles_ach / peaks["hbm_gbs"]},
                         "note": "fp32-faithful 3-term MMAs make the similarity tensor-bound: the ceiling of corr_tiles "
                                 "alone is ~0.27 of the HBM roofline when every tile is computed (DESIGN.md section 4)"}

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sds_np = {k: {kk: vv.detach().cpu().numpy() for kk, vv in getattr(net, k).state_dict().items()}
                      for k in D.GENERATOR_NETS}
            fps, cores, spf = cpu_forward_fps(sds_np, L, nb, n, 1, 5, 1)  # (compositing is negligible on the CPU side)
            cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                   "sample": "5 forwards of bs=1 of the same config (oracle/tsnet_oracle.py = bit-exact CPU "
                             "restatement of the reference forward), 1 warm-up"}
        eager = None
        fast = None
        if world == 1 and args.fast_point:
            # clearly labelled NON-PARITY speed point: single-pass fp16 operands (fails the parity tolerance; never the
            # headline -- tests/test_parity_gpu.py::test_fast_modes_run_and_are_flagged_non_parity)
            try:
                torch.manual_seed(1234)
                with contextlib.redirect_stdout(io.StringIO()):
                    fnet = TSNet(is_train=False, label_nc=L, n_blocks=nb, n_downsampling=3, n_source=n,
                                 math_mode="fp16", winograd=args.winograd)
                fnet.eval()
                with torch.no_grad():
                    def fstep():
                        fnet.set_test_input(devin["src_img"], devin["src_lbl"], devin["src_bbox"], devin["tar_lbl"],
                                            devin["tar_bbox"])
                        fnet.forward()
                    for _ in range(3):
                        fstep()
                    fms, _, _ = timed(fstep, 5)
                fast = {"math_mode": "fp16 (single pass, NOT parity grade)", "frames_per_s": bs * 5 / (fms * 1e-3),
                        "ms_per_step": fms / 5}
                del fnet
                torch.cuda.empty_cache()
            except Exception as e:
                fast = {"error": repr(e)[:200]}
        if world == 1 and args.torch_cuda_baseline:
            sds_np = {k: {kk: vv.detach().cpu().numpy() for kk, vv in getattr(net, k).state_dict().items()}
                      for k in D.GENERATOR_NETS}
            try:
                eager = torch_cuda_eager_fps(sds_np, L, nb, n, bs)
                eager["what"] = ("oracle port of the reference's torch ops run eagerly on this GPU (cuDNN / cuBLAS), "
                                 "same config and batch; informational -- the reference arm of this tier is the CPU run")
                eager["speedup_of_value_vs_fp32"] = value / world / eager["fp32_no_tf32"]["frames_per_s"]
                eager["speedup_of_value_vs_default_tf32"] = value / world / eager["default_tf32_convs"]["frames_per_s"]
            except Exception as e:  # never let the informational leg break the bench line
                eager = {"error": repr(e)[:200]}
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (fp16 hi/lo 3-term split operands, fp32 accumulate)" if args.math == "fp16x3" else args.math,
                "data": "synthetic",
                "config": {"workload": f"{'Youtube-dance (pose)' if args.pose else 'FaceForensics'} config: bs={bs}/GPU, 256x256, label_nc={L}, n_source={n}, "
                                       f"n_blocks={nb}, uint8 rectangular bbox, random-init weights",
                           "global_batch": bs * world, "parallelism": f"dp{world} (batch rows sharded, no collective)",
                           "math_mode": args.math, "winograd_f2x2_3x3": (args.winograd if isinstance(args.winograd, str) else bool(args.winograd)),
                           "wino_chunk_kb": net._engine.wino_chunk_kb, "direct_stem": net._engine.direct_stem, "bridge_variant": net._engine.bridge_variant,
                           "l2": "per-step working set (>5 GB of activations) far exceeds the 126 MB L2; no flush needed"},
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d * world,
                        "d2h_bytes_per_step": d2h * world, "ms_per_step": ms_e2e / args.steps,
                        "api": "FramePipeline.run (H2D / forward / D2H of consecutive batches overlapped)"
                        if args.e2e_mode == "pipelined" else "set_test_input + forward + rec_tar_img.cpu()",
                        "host_formats": ("uint8 BGR frames + dataset mean, uint8 class-index labels, uint8 bboxes "
                                         "(bit-identical forward; the reference's fp32 formats would be "
                                         f"{h2d_fp32 * world} B per step)") if args.e2e_inputs == "compact"
                        else "fp32 mean-subtracted frames, fp32 one-hot labels, uint8 bboxes (the reference's formats)"},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "roofline_corr_warp": roof_corr,
                "kernel_shares": shares, "cpu_baseline": cpu, "torch_cuda_eager_port": eager,
                "non_parity_speed_point": fast, "shard_check": shard_check}
        emit(line)
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """Rank 0's single JSON line goes to the process's ORIGINAL stdout; everything else that prints to fd 1 (e.g. the
    NCCL version banner under NCCL_DEBUG=VERSION) has been routed to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)   # stdout of this process and of the libraries it loads -> stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="frames per GPU per step")
    ap.add_argument("--n-source", dest="n_source", type=int, default=3)
    ap.add_argument("--n-blocks", dest="n_blocks", type=int, default=4)
    ap.add_argument("--math", default="fp16x3", choices=["fp16x3", "bf16x3", "fp16", "bf16"],
                    help="fp16x3 is the parity-grade default; single-pass modes are non-parity speed points")
    ap.add_argument("--pose", action="store_true", help="TSNet_pose, label_nc=25 (BASELINE.json config 3)")
    ap.add_argument("--no-winograd", dest="winograd", action="store_const", const=False, default=True,
                    help="direct implicit GEMM for the ResnetBlock convolutions (A/B against the Winograd default)")
    ap.add_argument("--wino-chunk-kb", dest="wino_chunk_kb", type=int, default=0,
                    help="experiment: K blocks accumulated in TMEM per promotion in the Winograd GEMMs (default: the "
                         "engine's parity-validated 2; 4 is faster but misses the image tolerance on one golden)")
    ap.add_argument("--direct-stem", dest="direct_stem", action="store_true",
                    help="generate the stem operand inside the stem kernel (tsnet_stem_conv_fwd) instead of materialising "
                         "it with tsnet_stem_taps (less DRAM traffic, measured slower: opt-in)")
    ap.add_argument("--bridge-variant", dest="bridge_variant", type=int, default=-1,
                    help="experiment: tsnet_wino_bridge_desc.variant (0 = 32-channel slabs, 1 = 16-channel slabs, 2 CTAs/SM)")
    ap.add_argument("--winograd-unfused", dest="winograd", action="store_const", const="unfused",
                    help="Winograd with separate transform passes instead of the fused bridge pass (A/B)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-cuda-baseline", dest="torch_cuda_baseline", action="store_false",
                    help="skip timing the reference's torch op sequence eagerly on the GPU (cuDNN / cuBLAS, bs = --batch;"
                         " on by default at N=1: BASELINE.md section 3 step 4)")
    ap.add_argument("--no-fast-point", dest="fast_point", action="store_false",
                    help="skip the labelled non-parity single-pass fp16 speed point")
    ap.add_argument("--e2e-inputs", dest="e2e_inputs", default="compact", choices=["compact", "fp32"],
                    help="host formats of the end-to-end path: compact = uint8 frames + uint8 class maps (default), "
                         "fp32 = the reference callers' formats")
    ap.add_argument("--e2e-mode", dest="e2e_mode", default="pipelined", choices=["pipelined", "sync"],
                    help="pipelined: wacv23_tsnet_b200.pipeline.FramePipeline; sync: set_test_input + forward + .cpu()")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
