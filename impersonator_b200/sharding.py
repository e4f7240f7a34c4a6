"""Multi-GPU plumbing of the hot path: frames shard, weights broadcast once, nothing per step.

The reference runs inference on one GPU at batch 1 (models/imitator.py:166-171, no collectives,
SURVEY.md 2.1).  Frames are independent units (per-frame state is only (cam, verts)), so N GPUs =
N processes (torchrun), rank r takes a contiguous shard of the target-frame list, and the only
communication is ONE broadcast at init of a packed buffer: generator weights (+ optional extras such
as the source image).  Source-side state (encode_src, background) is recomputed per rank (67.7 GFLOP,
< 1 ms) instead of shipping 44 MB of features.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced [start, end) of ``n_items`` for ``rank`` (first n % world ranks get one more)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_state(state_dict, extras=()):
    """-> (flat f32 tensor, layout) with layout = [(key, shape), ...] in sorted-key order, extras appended."""
    keys = sorted(state_dict.keys())
    parts = [state_dict[k].detach().reshape(-1).float() for k in keys] + [e.detach().reshape(-1).float() for e in extras]
    layout = [(k, tuple(state_dict[k].shape)) for k in keys] + [("__extra%d" % i, tuple(e.shape)) for i, e in enumerate(extras)]
    return torch.cat(parts), layout


def unpack_state(flat, layout):
    out, off = {}, 0
    for key, shape in layout:
        n = 1
        for s in shape:
            n *= s
        out[key] = flat[off:off + n].view(shape).clone()
        off += n
    assert off == flat.numel()
    return out


def broadcast_module(net, extras=(), src=0, device=None, stats=None):
    """Make every rank's ``net`` (and ``extras`` tensors) equal to rank ``src``'s with ONE collective.
    Works with NCCL (CUDA tensors) and gloo (CPU tensors).  Returns the received extras.  ``stats`` (dict, optional)
    receives ``bytes`` and, for CUDA tensors, ``ms`` = device time of the collective (CUDA events)."""
    sd = net.state_dict()
    flat, layout = pack_state(sd, extras)
    if device is not None:
        flat = flat.to(device)
    if stats is not None:
        stats["bytes"] = flat.numel() * flat.element_size()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if flat.is_cuda and stats is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            dist.broadcast(flat, src=src)
            e1.record()
            torch.cuda.synchronize()
            stats["ms"] = e0.elapsed_time(e1)
        else:
            dist.broadcast(flat, src=src)
    got = unpack_state(flat, layout)
    new_sd = {k: got[k].to(sd[k].dtype) for k in sd.keys()}
    net.load_state_dict(new_sd)
    return [got["__extra%d" % i] for i in range(len(extras))]
