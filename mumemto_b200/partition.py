"""Partition-merge across ranks: one partition per GPU, one all-gather of the partition tables, anchor merge on the device.

Mirrors what `mumemto merge p1.mums p2.mums ...` does through files (reference mumemto/merge_mums.py:171-192 -> anchor_merge,
src/merge_candidates.cpp:170-255): every partition was run with -M -n and shares its first sequence (the anchor); partitions are folded
in rank order, which fixes the column order of the merged table.  The exchange is a `torch.distributed.all_gather` (NCCL between GPUs;
gloo in the CPU tests) of {lengths, offsets, strands, thresholds[0..L0]}, padded to the largest partition.
"""
import numpy as np
import torch


class _DevArray:
    """Zero-copy view of engine-owned device memory for torch (CUDA array interface)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def local_partition(engine, device):
    """(lengths i32[n], offsets i64[n,N], strands u8[n,N], thresholds i16[L0+1]) of the engine's last find() as tensors on `device`."""
    n, N = engine.num_matches, engine.num_docs
    if device.type == "cuda":
        pl, po, ps, pt = engine.mum_table_device()
        L = engine.lib.mmb_threshold_count(engine.h) // 2
        mk = lambda ptr, shape, ts: torch.as_tensor(_DevArray(ptr, shape, ts), device=device) if int(np.prod(shape)) else torch.empty(shape, device=device)
        return (mk(pl, (n,), "<i4").to(torch.int32), mk(po, (n, N), "<i8").to(torch.int64), mk(ps, (n, N), "|u1").to(torch.uint8),
                mk(pt, (L,), "<i2").to(torch.int16))
    lengths, offsets, strands = engine.mums()
    thr = engine.thresholds()
    return (torch.from_numpy(lengths.view(np.int32).copy()), torch.from_numpy(offsets.copy()), torch.from_numpy(strands.copy()),
            torch.from_numpy(thr[: len(thr) // 2].view(np.int16).copy()))


def gather_partitions(part, dist, world):
    """all_gather of one partition per rank -> list of (lengths, offsets, strands, thresholds, n_mums, ndocs) in rank order."""
    lengths, offsets, strands, thr = part
    dev = lengths.device
    meta = torch.tensor([lengths.shape[0], offsets.shape[1] if offsets.dim() == 2 else 0, thr.shape[0]], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    metas = [m.tolist() for m in metas]
    if len({m[2] for m in metas}) != 1:
        raise ValueError("partitions do not share their first sequence (threshold lengths differ)")
    nmax, dmax = max(m[0] for m in metas), max(m[1] for m in metas)

    def padded(t, shape):
        buf = torch.zeros(shape, dtype=t.dtype, device=dev)
        if t.numel():
            if len(shape) == 1:
                buf[: t.shape[0]] = t
            else:
                buf[: t.shape[0], : t.shape[1]] = t
        return buf

    out = []
    for t, shape in ((lengths, (max(nmax, 1),)), (offsets, (max(nmax, 1), max(dmax, 1))), (strands, (max(nmax, 1), max(dmax, 1))), (thr, (thr.shape[0],))):
        mine = padded(t, shape)
        raw = mine.view(torch.uint8).reshape(-1)                 # NCCL and gloo have no 16-bit integer type: exchange bytes
        bufs = [torch.empty_like(raw) for _ in range(world)]
        dist.all_gather(bufs, raw)
        out.append([b.view(mine.dtype).reshape(shape) for b in bufs])
    parts = []
    for r, (n, nd, _) in enumerate(metas):
        # un-pad the column dimension so that rows are contiguous [n][nd]
        parts.append((out[0][r][:n].contiguous(), out[1][r][:n, :nd].contiguous(), out[2][r][:n, :nd].contiguous(), out[3][r].contiguous(), n, nd))
    return parts


def merge_gathered(engine, parts):
    """Fold the gathered partitions with the engine's merge kernels; returns (lengths, offsets, strands, thresholds) as numpy."""
    anchor_len = parts[0][3].shape[0]
    on_device = parts[0][0].is_cuda
    if on_device:
        args = [(p[0].data_ptr(), p[1].data_ptr(), p[2].data_ptr(), p[3].data_ptr(), p[4], p[5]) for p in parts]
    else:
        host = [(p[0].numpy().view(np.uint32), p[1].numpy(), p[2].numpy(), p[3].numpy().view(np.uint16), p[4], p[5]) for p in parts]
        args = [(np.ascontiguousarray(a), np.ascontiguousarray(b), np.ascontiguousarray(c), np.ascontiguousarray(d), n, nd) for a, b, c, d, n, nd in host]
    engine.anchor_merge(args, anchor_len, on_device=on_device)
    return engine.merge_result()


def partition_merge(engine, dist, world, device):
    """Everything after the per-rank find(): exchange + merge.  Every rank ends up with the merged result."""
    parts = gather_partitions(local_partition(engine, device), dist, world)
    return merge_gathered(engine, parts)
