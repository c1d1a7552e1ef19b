"""torchrun worker of tests/test_gpu_configs.py::test_multi_rank_bands_equal_single_rank_canvas (one rank per GPU).

Every rank builds the same seeded workload (a 3 x 3 grid of 1536^2 tiles with 300-pixel overlaps, 6 levels: every band
cuts through tiles, overlaps and 4-way corners), fuses ONLY its cost-balanced band (+ the SSIM halo) from the tiles that
reach it, scores its band, and the sums travel through distributed.SumsExchange (libsrfuse's NVLink peer exchange; NCCL
only if peer mapping is unavailable).  Rank 0 also computes the whole canvas alone and compares:
    reassembled bands == single-rank canvas, byte for byte; sq_err totals equal exactly; SSIM totals to 1e-9.
Prints one JSON line (rank 0)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    from helpers import synth
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    if rank == 0:
        ge.build()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()
    import super_resolution_system_b200 as srs
    D = srs.distributed

    rng = np.random.default_rng(2026)
    T, step, levels = 1536, 1236, 6
    tiles = [synth(rng, T, T) for _ in range(9)]
    pos = [(j * step, i * step) for j in range(3) for i in range(3)]
    H, W = 2 * step + T - 37, 2 * step + T - 101  # the last row / column overhangs the canvas
    ref = synth(rng, H, W)

    bands = D.balanced_bands(world, H, W, pos, [t.shape[:2] for t in tiles])
    r0, r1 = bands[rank]
    e0, e1 = D.widen(bands[rank], H)
    mine = D.tiles_for_rows(pos, [t.shape[0] for t in tiles], e0, e1)
    bm = srs.BlendingModule(num_levels=levels, device=local)
    qa = srs.QualityAssessmentModule(device=f"cuda:{local}")
    infos = [srs.TileInfo(torch.from_numpy(tiles[i]).to(dev), pos[i][1], pos[i][0], 0, 0) for i in mine]
    d_ref = torch.from_numpy(ref).to(dev)
    canvas = torch.zeros((H, W, 3), dtype=torch.uint8, device=dev)
    ex = D.SumsExchange(dev)
    results = []
    for _ in range(3):  # repeated steps: graph replay + consecutive exchanges (both parity slots)
        bm.laplacian_fusion(infos, output_shape=(H, W), row_range=(e0, e1), out=canvas)
        qa.psnr_ssim_sums_device(d_ref, canvas, ex.local, row_range=(r0, r1))
        results.append(ex.reduce(H * W * 3))
    tot = results[-1]
    assert all((r.sq_err, r.ssim_count, r.ssim_sum) == (tot.sq_err, tot.ssim_count, tot.ssim_sum) for r in results)

    # the host-memory variant: own rows uploaded once, halo rows shipped GPU -> GPU, same bytes and sums
    import torch as _t
    pinned = [_t.from_numpy(t).pin_memory().numpy() for t in tiles]
    h_infos = [srs.TileInfo(pinned[i], pos[i][1], pos[i][0], 0, 0) for i in range(len(tiles))]
    h_out = _t.zeros((H, W, 3), dtype=_t.uint8).pin_memory().numpy()
    fuser = D.BandedHostFuser(bm, qa, h_infos, _t.from_numpy(ref).pin_memory().numpy(), (H, W), bands, rank, world, dev, exchange=ex)
    for _ in range(2):
        ftot = fuser.step(h_out)
    fuser_ok = (ftot.sq_err, ftot.ssim_count, ftot.ssim_sum) == (tot.sq_err, tot.ssim_count, tot.ssim_sum) and \
        bool(np.array_equal(h_out[r0:r1], canvas[r0:r1].cpu().numpy()))
    ok_t = torch.tensor([1 if fuser_ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
    total_h2d = torch.tensor([float(fuser.h2d_bytes), float(fuser.p2p_bytes)], dtype=torch.float64, device=dev)
    dist.all_reduce(total_h2d)

    # every rank must hold the same totals: gather them (plumbing) and compare on rank 0
    mine_t = torch.tensor([float(tot.sq_err), tot.ssim_sum, float(tot.ssim_count)], dtype=torch.float64, device=dev)
    every = [torch.empty_like(mine_t) for _ in range(world)]
    dist.all_gather(every, mine_t)
    # reassemble: each rank contributes its own band's rows
    own = torch.zeros((H, W, 3), dtype=torch.uint8, device=dev)
    own[r0:r1] = canvas[r0:r1]
    dist.reduce(own, dst=0, op=dist.ReduceOp.MAX)  # bands are disjoint and the rest is zero: MAX == concatenation
    if rank == 0:
        all_infos = [srs.TileInfo(torch.from_numpy(t).to(dev), x, y, 0, 0) for t, (y, x) in zip(tiles, pos)]
        full = bm.laplacian_fusion(all_infos, output_shape=(H, W))
        s1 = qa.psnr_ssim_sums(d_ref, full)
        rep = {"world": world, "bands": bands, "exchange": ex.kind,
               "canvas_equal": bool(torch.equal(full, own)),
               "sq_err_equal": int(tot.sq_err) == int(s1.sq_err) and int(tot.ssim_count) == int(s1.ssim_count),
               "ssim_rel_err": abs(tot.ssim_sum - s1.ssim_sum) / abs(s1.ssim_sum),
               "all_ranks_agree": all(bool(torch.equal(e, every[0])) for e in every),
               "host_fuser_equal": bool(ok_t.item() == 1), "host_fuser_h2d_bytes": total_h2d[0].item(), "host_fuser_p2p_bytes": total_h2d[1].item(),
               "tile_bytes_plus_reference": float(sum(t.nbytes for t in tiles) + ref.nbytes)}
        print(json.dumps(rep), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
