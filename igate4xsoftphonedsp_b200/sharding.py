"""Multi-GPU layer: bridges shard across ranks, nothing is exchanged per frame.

The reference runs four independent softphone processes per box
(main.cpp:5-13, roip_ed137.cpp:82-100) that share no audio; the unit that must
stay together is the bridge (the legs mixed into conference slot 0,
roip_ed137.cpp:4907-4920).  So rank r owns bridges [r*B/R, (r+1)*B/R) and runs
the fused path on its own GPU with no collective on the hot path.  The only
exchange is the end-of-run gather of the per-channel event summaries
(Functions.cpp:2148-2230 -> 32 B/channel) to rank 0: one
torch.distributed.gather over NCCL (NVLink) -- or gloo in the CPU tests.
"""
import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def bridge_range(rank, world, nbridges):
    """Contiguous bridge slice [b0, b1) of `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (nbridges * rank) // world, (nbridges * (rank + 1)) // world


def channel_range(rank, world, nbridges, legs):
    b0, b1 = bridge_range(rank, world, nbridges)
    return b0 * legs, b1 * legs


def shard_sizes(world, nbridges):
    return [bridge_range(r, world, nbridges)[1] - bridge_range(r, world, nbridges)[0] for r in range(world)]


def gather_records(local, nbridges, legs, group=None, dst=0):
    """Gathers per-channel records (any fixed-size record dtype) to rank `dst`.

    local : this rank's records for its channel slice -- numpy structured array
            (gloo / host) or a torch tensor [n_local, words] on the rank's GPU (nccl).
    Returns the concatenated [nbridges*legs] records on `dst`, None elsewhere.
    Ranks may own different numbers of bridges; shards are padded to the largest.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [s * legs for s in shard_sizes(world, nbridges)]
    nmax = max(sizes)
    is_np = isinstance(local, np.ndarray)
    if is_np:
        itemsize = local.dtype.itemsize
        assert local.shape[0] == sizes[rank]
        buf = torch.zeros(nmax * itemsize, dtype=torch.uint8)
        buf[: local.nbytes] = torch.from_numpy(np.ascontiguousarray(local).view(np.uint8).reshape(-1))
    else:
        assert local.shape[0] == sizes[rank]
        flat = local.contiguous().view(torch.uint8).reshape(-1)
        # bytes per record from the record shape, also on a rank that owns no bridge (world > nbridges)
        itemsize = local.element_size()
        for d in local.shape[1:]:
            itemsize *= int(d)
        buf = torch.zeros(nmax * itemsize, dtype=torch.uint8, device=local.device)
        buf[: flat.numel()] = flat
    outs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, outs, dst=dst, group=group)
    if rank != dst:
        return None
    parts = [o[: sizes[r] * itemsize] for r, o in enumerate(outs)]
    cat = torch.cat(parts)
    if is_np:
        return cat.numpy().view(local.dtype).reshape(-1)
    return cat.view(local.dtype).reshape(-1, *local.shape[1:])
