"""Band sharding across the GPUs of one box (one process per GPU, torch.distributed).

The canvas is cut into `world` horizontal bands.  A rank fuses its band (plus a 5-row halo each side, so the 11x11
SSIM window never needs a neighbour's pixels) from the tiles that reach it -- per-tile pyramids have tile-local
borders, so a rank recomputes the pyramid rows its band depends on (libsrfuse derives those row ranges itself from
`row_begin/row_end`) and the data path needs NO collective.  The only exchange is the all-reduce of the additive
PSNR / SSIM partial sums (srf_iqa_sums): three numbers per rank, over NCCL on the GPUs (gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

from . import _native
from ._arrays import torch

IQA_HALO = 5  # radius of the 11-tap SSIM window


def band(rank: int, world: int, rows: int) -> Tuple[int, int]:
    """canvas rows [r0, r1) owned by `rank`: ceil(rows / world) rows each, the last band takes what is left"""
    per = -(-rows // world)
    return min(rank * per, rows), min((rank + 1) * per, rows)


def balanced_bands(world: int, rows: int, cols: int, positions_yx, sizes_hw) -> List[Tuple[int, int]]:
    """Cut the canvas rows into `world` contiguous bands of equal COST rather than equal height.  A canvas row costs
    one unit of per-pixel work (assessment + canvas write) plus one unit per covering tile (gather + that tile's
    pyramid rows); with 20-30 % tile overlap the doubly covered rows are ~60 % dearer, and equal-height bands leave
    the ranks that own an overlap zone ~20 % slower than the rest.  0.65 : 1 is the measured ratio of the two terms
    on B200 (profiles/README.md)."""
    import numpy as np
    cost = np.full(rows, 0.65)
    for (y, x), (h, w) in zip(positions_yx, sizes_hw):
        wfrac = max(0, min(cols, x + w) - max(0, x)) / cols
        cost[max(0, y):min(rows, y + h)] += wfrac
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    cuts = [0]
    for k in range(1, world):
        cuts.append(max(int(np.searchsorted(cum, cum[-1] * k / world)), cuts[-1]))
    cuts.append(rows)
    return [(min(a, rows), min(b, rows)) for a, b in zip(cuts[:-1], cuts[1:])]


def widen(band_rows: Tuple[int, int], rows: int) -> Tuple[int, int]:
    """rows a rank must FUSE for a band: the band plus the SSIM halo, clipped to the canvas"""
    r0, r1 = band_rows
    if r0 >= r1:
        return r0, r0
    return max(r0 - IQA_HALO, 0), min(r1 + IQA_HALO, rows)


def fuse_rows(rank: int, world: int, rows: int) -> Tuple[int, int]:
    """rows a rank must FUSE: its (equal-height) band widened by the SSIM halo, clipped to the canvas"""
    return widen(band(rank, world, rows), rows)


def tiles_for_rows(positions_yx: Sequence[Tuple[int, int]], heights: Sequence[int], r0: int, r1: int) -> List[int]:
    """indices (list order kept -- it is the accumulation order) of the tiles that intersect canvas rows [r0, r1)"""
    return [i for i, ((y, _), h) in enumerate(zip(positions_yx, heights)) if y < r1 and y + h > r0]


def allreduce_sums(sums, total_elems: int, group=None, device=None):
    """sum srf_iqa_sums over the ranks of `group` (in place) and return it.  sq_err and ssim_count are integers below
    2^53, exact in float64; bits_a / bits_b travel as their <= 1 / > 1 class."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sums
    # one collective: the two OR-of-bytes fields only matter through "is any byte > 1", which sums fine
    t = torch.tensor([float(sums.sq_err), sums.ssim_sum, float(sums.ssim_count), float(sums.bits_a > 1), float(sums.bits_b > 1)],
                     dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    v = t.tolist()
    sums.sq_err, sums.ssim_sum, sums.ssim_count = int(v[0]), v[1], int(v[2])
    sums.n_elems = int(total_elems)
    sums.bits_a = 255 if v[3] > 0 else min(int(sums.bits_a), 1)
    sums.bits_b = 255 if v[4] > 0 else min(int(sums.bits_b), 1)
    return sums


def allreduce_device_sums(d_sums, total_elems: int, group=None):
    """`d_sums`: the 5-double CUDA tensor QualityAssessmentModule.psnr_ssim_sums_device filled.  One NCCL all-reduce
    (skipped at world size 1), ONE device->host read; returns an srf_iqa_sums with the job-wide totals."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(d_sums, op=dist.ReduceOp.SUM, group=group)
    v = d_sums.tolist()
    s = _native.srf_iqa_sums()
    s.sq_err, s.ssim_sum, s.ssim_count, s.n_elems = int(v[0]), v[1], int(v[2]), int(total_elems)
    s.bits_a, s.bits_b = (255 if v[3] > 0 else 1), (255 if v[4] > 0 else 1)
    return s
