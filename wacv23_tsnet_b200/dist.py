"""Multi-GPU plumbing for the TS-Net forward: one process per GPU, batch rows sharded, no collective inside
the forward (InstanceNorm, softmax and the source means are all per-sample: SURVEY.md section 8e).

torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only for
  * broadcasting the generator weights from rank 0 at start-up, and
  * optionally all-gathering the output frames.
"""
import os

import torch
import torch.distributed as dist

GENERATOR_NETS = ("img_enc", "lbl_enc", "fuse_net", "dec")


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (no-op for WORLD_SIZE=1)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_range(global_batch, rank, world):
    """Rows [start, stop) of the global batch owned by `rank`: contiguous equal shards (rank r takes
    rows [r*B/R, (r+1)*B/R))."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def shard_inputs(inputs, rank, world):
    """Slice every tensor (or list of tensors) of an input dict along the batch dimension."""
    out = {}
    for k, v in inputs.items():
        if isinstance(v, (list, tuple)):
            s, e = shard_range(v[0].shape[0], rank, world)
            out[k] = [t[s:e] for t in v]
        else:
            s, e = shard_range(v.shape[0], rank, world)
            out[k] = v[s:e]
    return out


def broadcast_generator(model, src=0):
    """Make every rank hold rank `src`'s generator parameters (bit-identical replicas)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    with torch.no_grad():
        for name in GENERATOR_NETS:
            for p in getattr(model, name).parameters():
                dist.broadcast(p.data, src=src)
    # the broadcast writes through .data (no version bump): drop packed weights / captured graphs of earlier forwards
    if hasattr(model, "invalidate_weights"):
        model.invalidate_weights()


def all_gather_frames(x):
    """[b, ...] per rank -> [world*b, ...] on every rank, rank-major (the global batch order)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    parts = [torch.empty_like(x) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, x.contiguous())
    return torch.cat(parts, 0)


def max_over_ranks(value, device):
    """Max of a python float over ranks (timing is reported as the slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
