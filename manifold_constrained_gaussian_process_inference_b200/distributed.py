"""Multi-GPU plumbing: chains shard across ranks with NO collective on the hot path (every rank rebuilds the identical
GP tables locally); one all-gather of the retained scalar draws (θ, σ, lp) at the end of a run feeds R-hat / ESS
(SURVEY.md section 8(e)).  On GPUs both exchange steps (the warm-up's pooled window statistics and the final all-gather) run
inside libmagi_b200 with NCCL on the sampler's stream (``init_device_comm`` / ``allgather_draws_device``; csrc/comm.cu);
``torch.distributed`` only carries the 128-byte NCCL unique id between the ranks.  ``allgather_draws`` is the
transport-agnostic host path (gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_chains(n_chains_total: int, rank: int, world: int):
    """Contiguous block partition: returns (first_chain, n_local).  The first chain id doubles as the RNG stream offset,
    so the union of all ranks' chains is independent of the world size."""
    base, rem = divmod(n_chains_total, world)
    n_local = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n_local


def allgather_draws(local_draws, group=None):
    """local_draws: torch tensor (n_iter, n_local_chains, n_cols) on this rank's device (CUDA for NCCL, CPU for gloo).
    Returns (n_iter, n_chains_total, n_cols) with chains in global order (ranks may hold different chain counts)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_draws
    world = dist.get_world_size(group)
    n_iter, n_local, n_cols = local_draws.shape
    counts = [None] * world
    dist.all_gather_object(counts, int(n_local), group=group)          # host-side exchange of the shard sizes: no device sync
    nmax = max(counts)
    pad = local_draws
    if n_local < nmax:
        pad = torch.cat([local_draws, local_draws.new_zeros((n_iter, nmax - n_local, n_cols))], dim=1)
    pad = pad.contiguous()
    out = local_draws.new_empty((world, n_iter, nmax, n_cols))
    dist.all_gather_into_tensor(out, pad, group=group) if local_draws.is_cuda else dist.all_gather(list(out.unbind(0)), pad, group=group)
    return torch.cat([out[r, :, :counts[r], :] for r in range(world)], dim=1)


def device_draws_as_tensor(target):
    """Zero-copy torch view of the on-device draw store of ``target`` (n_stored, n_chains, n_cols)."""
    import torch
    from .samplers import hmc_draws_device_view
    ptr, ns, nc, ncol = hmc_draws_device_view(target)
    if ns == 0:
        return torch.empty((0, nc, ncol), dtype=torch.float64, device="cuda:%d" % target.device)

    class _Holder:
        pass
    holder = _Holder()
    holder.__cuda_array_interface__ = {"shape": (ns, nc, ncol), "typestr": "<f8", "data": (ptr, False), "version": 2}
    return torch.as_tensor(holder, device="cuda:%d" % target.device)


def make_window_allreduce(target, group=None):
    """The ``magi_allreduce_fn`` callback (include/magi_b200.h) for ``run_hmc_sampler(window_allreduce=...)``: an in-place
    sum over ranks of a device buffer, enqueued on the sampler's stream through ``torch.distributed`` (NCCL).  Returns a
    ctypes function pointer; keep a reference to it for as long as the sampler may call it."""
    import torch
    import torch.distributed as dist
    from . import _lib
    dev = "cuda:%d" % target.device

    def _cb(ptr, n, stream, user):
        try:
            class _Holder:
                pass
            holder = _Holder()
            holder.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}
            t = torch.as_tensor(holder, device=dev)
            ext = torch.cuda.ExternalStream(int(stream) if stream else 0, device=dev) if stream else torch.cuda.default_stream(dev)
            with torch.cuda.stream(ext):
                if dist.is_initialized() and dist.get_world_size(group) > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return 0
        except Exception:                                   # nothing may propagate through the C frame
            import traceback
            traceback.print_exc()
            return 1
    return _lib.ALLREDUCE_FN(_cb)


def init_device_comm(target, group=None):
    """Creates the NCCL communicator libmagi_b200 uses for ``target``'s sampler (collective over ``group``): rank 0 draws
    the unique id, torch.distributed broadcasts its 128 bytes, every rank calls ``magi_comm_init``; one small all-reduce /
    all-gather then warms the communicator so that the first real collective is not charged its set-up."""
    import ctypes
    import torch.distributed as dist
    from . import _lib
    L = _lib.lib()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        _lib.check(L.magi_nccl_unique_id(buf))
    box = [buf.raw]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    _lib.check(L.magi_comm_init(target._h, box[0], rank, world))
    _lib.check(L.magi_comm_warmup(target._h, None))


def allgather_draws_device(target, stream: int = 0, out=None):
    """All-gather of the on-device draw store over the library's communicator (``init_device_comm``): returns a CUDA tensor
    (n_iter, n_chains_total, n_cols) with chains in global order.  Every rank must hold equally many chains.  ``out``: optional
    preallocated (world, n_iter, n_local, n_cols) float64 CUDA tensor to receive the ranks' blocks."""
    import ctypes
    import torch
    from . import _lib
    from .samplers import hmc_draws_device_view
    L = _lib.lib()
    _, ns, nc, ncol = hmc_draws_device_view(target)
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    if out is None:
        out = torch.empty((world, ns, nc, ncol), dtype=torch.float64, device="cuda:%d" % target.device)
    assert tuple(out.shape) == (world, ns, nc, ncol) and out.is_contiguous()
    _lib.check(L.magi_hmc_allgather_draws(target._h, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(stream)))
    return out.permute(1, 0, 2, 3).reshape(ns, world * nc, ncol)
