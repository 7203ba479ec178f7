"""Multi-GPU plumbing for the sqeazy hot path: one process per GPU, torch.distributed for the rendezvous.

The path shards without data-path collectives:
  * batches / time-lapses: stack v goes to rank v mod G (`stacks_for_rank`);
  * one large stack: contiguous z-slabs (`zslab_for_rank`), each rank encodes its slab into its own blob.
Two small reductions make the sharded result identical to the single-GPU one:
  * quantiser: the 65536-bin histogram is summed over ranks (NCCL all-reduce, 256 KiB) so every rank derives
    the same LUT (`allreduce_histogram`, then encode with `global_hist=`);
  * rmestbkrd: the four sampled-face histograms are summed over the ranks that own the faces/rows, then every rank
    evaluates the same 99 % support (`global_background_threshold`) and encodes with remove_background(threshold=T).
Everything here works with the gloo backend on CPU tensors (tests) and with nccl on CUDA tensors (bench).
"""
from __future__ import annotations

import numpy as np


def stacks_for_rank(n_stacks: int, rank: int, world: int):
    """volume v -> GPU v mod G"""
    return list(range(rank, n_stacks, world))


def zslab_for_rank(Z: int, rank: int, world: int):
    """contiguous z-range [z0, z1) of rank; slabs differ by at most one frame"""
    base, rem = divmod(Z, world)
    z0 = rank * base + min(rank, rem)
    return z0, z0 + base + (1 if rank < rem else 0)


def allreduce_histogram(hist, group=None):
    """in-place sum over ranks of a 65536-bin int32 histogram tensor (wraps mod 2^32 like the reference's uint32 bins)"""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


def background_sample_ranges(shape, l2_bytes: int, frame_portion: int):
    """element ranges (hist index, start voxel, count) of the whole volume that rmestbkrd samples
    (encoders/background_scheme_utils.hpp:35-105): z = 0 / z = Z-1 face portions, rows y = 0 / y = Y-1 at z in {1, Z/2, Z-2}"""
    Z, Y, X = shape
    frame = Y * X
    out = [(0, 0, frame_portion), (1, (Z - 1) * frame, frame_portion)]
    for i, y in enumerate((0, Y - 1)):
        for z in (1, Z // 2, Z - 2):
            if 0 <= z < Z:
                out.append((2 + i, z * frame + y * X, X))
    return out


def global_background_threshold(slab, z0: int, shape, l2_bytes: int = -1, histogram_fn=None, support_fn=None, group=None):
    """rmestbkrd threshold of the WHOLE volume from z-slabs spread over ranks.
    slab: this rank's voxels [z0, z0 + slab.shape[0]) as a flat-indexable tensor; histogram_fn(sub_tensor, hist_row) accumulates
    a 65536-bin histogram (sqeazy_b200.histogram_device on CUDA); support_fn(hist_row_numpy) -> float."""
    import torch

    import sqeazy_b200 as sq

    Z, Y, X = shape
    frame = Y * X
    portion = sq.rmest_frame_portion(frame, l2_bytes)
    lo, hi = z0 * frame, (z0 + slab.shape[0]) * frame
    hists = torch.zeros((4, 65536), dtype=torch.int32, device=slab.device)
    flat = slab.reshape(-1)
    for h, start, count in background_sample_ranges(shape, l2_bytes, portion):
        a, b = max(start, lo), min(start + count, hi)
        if a < b:
            histogram_fn(flat[a - lo : b - lo], hists[h])
    if slab.is_cuda:
        torch.cuda.synchronize()
    allreduce_histogram(hists, group)
    hn = hists.cpu().numpy().view(np.uint32)
    supports = np.array([support_fn(hn[i]) for i in range(4)], dtype=np.float32)
    return int(np.uint16(supports.min())), supports
