"""Partition-merge across ranks: one partition per GPU, one all-gather of the partition tables, anchor merge on the device.

Mirrors what `mumemto merge p1.mums p2.mums ...` does through files (reference mumemto/merge_mums.py:171-192 -> anchor_merge,
src/merge_candidates.cpp:170-255): every partition was run with -M -n and shares its first sequence (the anchor); partitions are folded
in rank order, which fixes the column order of the merged table.  The exchange is a `torch.distributed.all_gather` (NCCL between GPUs;
gloo in the CPU tests) of {lengths, offsets, strands, thresholds[0..L0]} packed into one byte payload per rank.
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


def _layout(n, nd, L):
    """Byte offsets of (lengths u32[n], offsets i64[n][nd], strands u8[n][nd], thresholds u16[L]) inside one rank's payload."""
    a8 = lambda x: (x + 7) & ~7
    o_off = a8(4 * n)
    o_str = o_off + 8 * n * nd
    o_thr = a8(o_str + n * nd)
    return 0, o_off, o_str, o_thr, (o_thr + 2 * L + 15) & ~15


def gather_partitions(part, dist, world):
    """ONE payload all-gather (plus a 3-integer size exchange that shapes it): every rank packs its partition table into a byte buffer
    (NCCL and gloo have no 16-bit integer type), padded to the largest partition; the gathered buffer is then viewed in place.
    Returns a list of (lengths, offsets, strands, thresholds, n_mums, ndocs) in rank order."""
    lengths, offsets, strands, thr = part
    dev = lengths.device
    n, nd, L = int(lengths.shape[0]), int(offsets.shape[1]) if offsets.dim() == 2 else 0, int(thr.shape[0])
    meta = torch.tensor([n, nd, L], dtype=torch.int64, device=dev)
    metas = torch.zeros(world * 3, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(metas, meta)
    metas = metas.view(world, 3).tolist()
    if len({m[2] for m in metas}) != 1:
        raise ValueError("partitions do not share their first sequence (threshold lengths differ)")
    stride = max(_layout(*m)[4] for m in metas)
    mine = torch.zeros(stride, dtype=torch.uint8, device=dev)
    o_len, o_off, o_str, o_thr, _ = _layout(n, nd, L)
    if n:
        mine[o_len:o_len + 4 * n].view(torch.int32).copy_(lengths)
        mine[o_off:o_off + 8 * n * nd].view(torch.int64).copy_(offsets.reshape(-1))
        mine[o_str:o_str + n * nd].copy_(strands.reshape(-1))
    if L:
        mine[o_thr:o_thr + 2 * L].view(torch.int16).copy_(thr)
    everything = torch.empty(world * stride, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(everything, mine)
    parts = []
    for r, (rn, rnd, rL) in enumerate(metas):
        base = r * stride
        q_len, q_off, q_str, q_thr, _ = _layout(rn, rnd, rL)
        parts.append((everything[base + q_len: base + q_len + 4 * rn].view(torch.int32),
                      everything[base + q_off: base + q_off + 8 * rn * rnd].view(torch.int64).view(rn, rnd),
                      everything[base + q_str: base + q_str + rn * rnd].view(rn, rnd),
                      everything[base + q_thr: base + q_thr + 2 * rL].view(torch.int16), rn, rnd))
    return parts


def merge_gathered(engine, parts, fetch=True):
    """Fold the gathered partitions with the engine's merge kernels; returns (lengths, offsets, strands, thresholds) as numpy, or
    nothing with fetch=False (the merged table and its .mums text stay on the device)."""
    anchor_len = parts[0][3].shape[0]
    on_device = parts[0][0].is_cuda
    if on_device:
        args = [(p[0].data_ptr(), p[1].data_ptr(), p[2].data_ptr(), p[3].data_ptr(), p[4], p[5]) for p in parts]
    else:
        host = [(p[0].numpy().view(np.uint32), p[1].numpy(), p[2].numpy(), p[3].numpy().view(np.uint16), p[4], p[5]) for p in parts]
        args = [(np.ascontiguousarray(a), np.ascontiguousarray(b), np.ascontiguousarray(c), np.ascontiguousarray(d), n, nd) for a, b, c, d, n, nd in host]
    engine.anchor_merge(args, anchor_len, on_device=on_device)
    return engine.merge_result() if fetch else None


def partition_merge(engine, dist, world, device, fetch=True):
    """Everything after the per-rank find(): exchange + merge.  Every rank ends up with the merged result."""
    parts = gather_partitions(local_partition(engine, device), dist, world)
    return merge_gathered(engine, parts, fetch=fetch)


def gather_slices(engine, dist, world, device):
    """SURVEY 8e(2), one collection sharded over ranks by slice of the suffix order: the output bytes of every rank's slice, in rank order
    (= the bytes of the unsharded run).  One size exchange and ONE all-gather of the padded text slices; returns (bytes tensor on
    `device` holding all slices back to back, total length)."""
    n = int(engine.lib.mmb_text_size(engine.h))
    sizes = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(sizes, torch.tensor([n], dtype=torch.int64, device=device))
    sizes = sizes.tolist()
    stride = (max(sizes) + 15) & ~15
    mine = torch.zeros(max(stride, 16), dtype=torch.uint8, device=device)
    if n:
        if device.type == "cuda":
            mine[:n].copy_(torch.as_tensor(_DevArray(engine.lib.mmb_text_device(engine.h), (n,), "|u1"), device=device))
        else:
            mine[:n].copy_(torch.from_numpy(np.frombuffer(engine.text(), dtype=np.uint8).copy()))
    everything = torch.empty(world * mine.numel(), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(everything, mine)
    out = torch.cat([everything[r * mine.numel(): r * mine.numel() + sizes[r]] for r in range(world)]) if sum(sizes) else everything[:0]
    return out, int(sum(sizes))
