#!/usr/bin/env python
"""Per-rank device time of an N-band job, measured on ONE GPU: for every rank of `bench.py --gpus N` this runs that rank's
band (same tiles, same band cut, same kernels, CUDA-graph replay) without the all-reduce, so the load balance and the
parts of a step that do not shrink with the band can be read without an N-GPU box.

    python profiles/band_probe.py --gpus 8 [--workload cfg4_200mp] [--steps 10]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--workload", default="cfg4_200mp")
    args = ap.parse_args()
    import torch
    import workloads
    import __graft_entry__ as ge
    ge.build()
    import super_resolution_system_b200 as pkg
    wl = workloads.build(args.workload)
    dev = torch.device("cuda", 0)
    native, D = pkg._native, pkg.distributed
    ctx = native.context(0)
    H, W = wl.canvas_hw
    bands = D.balanced_bands(args.gpus, H, W, wl.positions_yx, [t.shape[:2] for t in wl.tiles])
    d_all = [torch.from_numpy(t).to(dev) for t in wl.tiles]
    d_ref = torch.from_numpy(wl.reference).to(dev)
    d_canvas = torch.zeros((H, W, 3), dtype=torch.uint8, device=dev)
    d_sums = torch.zeros(5, dtype=torch.float64, device=dev)
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    out = []
    for rank, (r0, r1) in enumerate(bands):
        e0, e1 = D.widen((r0, r1), H)
        mine = D.tiles_for_rows(wl.positions_yx, [t.shape[0] for t in wl.tiles], e0, e1)
        bm = pkg.BlendingModule(num_levels=wl.levels, device=0)
        qa = pkg.QualityAssessmentModule(device="cuda:0")
        infos = [pkg.TileInfo(d_all[i], wl.positions_yx[i][1], wl.positions_yx[i][0], 0, 0) for i in mine]
        plan = bm.plan_fusion(infos, output_shape=(H, W))

        def step():
            bm.laplacian_fusion(plan, row_range=(e0, e1), out=d_canvas)
            qa.psnr_ssim_sums_device(d_ref, d_canvas, d_sums, row_range=(r0, r1))

        for _ in range(4):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.steps
        # the same steps with the bench's per-step read-back of the sums (a host round trip every step): the difference is
        # the launch latency a step exposes once the host waits for its result before enqueueing the next one
        a.record()
        for _ in range(args.steps):
            step()
            d_sums.tolist()
        b.record()
        torch.cuda.synchronize()
        ms_synced = a.elapsed_time(b) / args.steps
        # the two-deep pipeline of bench.py's default loop (assess of step k overlaps fuse of step k + 1, two contexts)
        pipe = D.PipelinedFuseAssess(pkg, infos, (H, W), wl.levels, 0, d_ref, (r0, r1), (e0, e1), D.SumsExchange(dev))
        for _ in range(8):
            pipe.submit()
        pipe.join()
        torch.cuda.synchronize()
        a.record(pipe.s_fuse)
        for _ in range(args.steps):
            pipe.submit()
        pipe.join()
        b.record()
        torch.cuda.synchronize()
        ms_pipe = a.elapsed_time(b) / args.steps
        del pipe
        native.profile_enable(ctx, True)
        native.profile_read(ctx)
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
        agg = {}
        for name, t, _ in native.profile_read(ctx):
            agg[name] = agg.get(name, 0.0) + t / args.steps
        native.profile_enable(ctx, False)
        rec = {"rank": rank, "band": [r0, r1], "rows": r1 - r0, "tiles": len(mine), "ms": round(ms, 4), "ms_with_per_step_readback": round(ms_synced, 4), "ms_pipelined": round(ms_pipe, 4),
               "kernels": {k: round(v, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])}}
        out.append(rec)
        print(json.dumps(rec), flush=True)
    worst = max(r["ms"] for r in out)
    print(json.dumps({"gpus": args.gpus, "max_ms": worst, "max_ms_pipelined": max(r["ms_pipelined"] for r in out),
                      "max_ms_with_per_step_readback": max(r["ms_with_per_step_readback"] for r in out), "mean_ms": round(sum(r["ms"] for r in out) / len(out), 4),
                      "sum_ms": round(sum(r["ms"] for r in out), 4)}))


if __name__ == "__main__":
    main()
