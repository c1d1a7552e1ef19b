#!/usr/bin/env python
"""bench.py -- fused canvas megapixels per second of the tile -> Laplacian-fuse -> PSNR/SSIM hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4_200mp] [--impl b200|reference]

One "step" = fuse every upscaled tile of the workload into the uint8 canvas AND score that canvas (PSNR + 11x11
Gaussian SSIM) against the comparison canvas.  Metric: canvas MP / s (BASELINE.json).

  value    whole-job MP/s with tiles + comparison canvas already resident in HBM (CUDA events, max over ranks)
  e2e      the same step through the public host API (SuperResolutionPipeline.fuse_and_assess): pinned HOST tiles and
           comparison canvas in, HOST canvas + scores out, copies inside the timed region
  roofline the dominant kernel of the step: algorithmic bytes / its CUDA-event time (per-kernel events from the
           library's profiling hooks, over K more steps) against MEASURED_PEAKS.json's HBM figure
  cpu_baseline  the reference's CPU path (oracle/reference_port.py, same arithmetic and cost structure) on a bounded
           sample of the same workload, on this box's host cores (rank 0, N=1 only)

N > 1 (torchrun, one rank per GPU): the canvas is split into N horizontal bands (strong scaling on the same canvas);
every rank fuses its band (+5 halo rows each side so SSIM needs no neighbour exchange) from the tiles that reach it
and the PSNR/SSIM partial sums are all-reduced over NCCL.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
METRIC = "fused_canvas_megapixels_per_second"
UNIT = "MP/s"
IQA_HALO = 5


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE JSON line: park the real stdout and point fd 1 at stderr so that library chatter
# (NCCL's version banner, for one) cannot get in front of it
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception as e:  # pragma: no cover
            log("clock sampling unavailable:", e)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for t, line in self.rows:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 6 or not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            try:
                sm.append(float(p[0]))
                mx = max(mx, float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------- reference arm
def cpu_sample(wl, threads: int):
    """bounded sample of the workload for the CPU path: the first two tiles of the grid (they overlap) fused into
    their bounding canvas, and PSNR + SSIM on a 2048-row strip of that canvas"""
    tiles, pos = wl.tiles[:2], wl.positions_yx[:2]
    H = min(wl.canvas_hw[0], max(y + t.shape[0] for t, (y, x) in zip(tiles, pos)))
    W = min(wl.canvas_hw[1], max(x + t.shape[1] for t, (y, x) in zip(tiles, pos)))
    return tiles, pos, (H, W), wl.reference[:H, :W]


def cpu_step(port, sample, levels, threads):
    tiles, pos, hw, ref = sample
    t0 = time.perf_counter()
    canvas = port.laplacian_fuse(tiles, pos, hw, levels, "cosine", threads)
    t1 = time.perf_counter()
    rows = min(2048, hw[0])
    port.psnr(ref[:rows], canvas[:rows])
    port.ssim(ref[:rows], canvas[:rows])
    t2 = time.perf_counter()
    mp_fuse, mp_iqa = hw[0] * hw[1] / 1e6, rows * hw[1] / 1e6
    # seconds per canvas megapixel of the whole step = fuse s/MP + assess s/MP
    return 1.0 / ((t1 - t0) / mp_fuse + (t2 - t1) / mp_iqa), (t1 - t0), (t2 - t1), mp_fuse, mp_iqa


def run_reference(args, wl):
    import oracle
    from oracle import reference_port as port
    oracle.pyr.build()
    threads = os.cpu_count() or 1
    try:
        import cv2
        cv2.setNumThreads(threads)
    except Exception:
        pass
    sample = cpu_sample(wl, threads)
    for _ in range(args.warmup):
        cpu_step(port, sample, wl.levels, threads)
    vals, t_all = [], time.perf_counter()
    for _ in range(args.steps):
        vals.append(cpu_step(port, sample, wl.levels, threads))
    wall = time.perf_counter() - t_all
    v = float(np.mean([x[0] for x in vals]))
    desc = (f"tiles 0-1 of {wl.name} fused into their {sample[2][1]}x{sample[2][0]} canvas ({vals[0][3]:.1f} MP, "
            f"{np.mean([x[1] for x in vals]):.2f} s) + PSNR/SSIM on a 2048-row strip ({vals[0][4]:.1f} MP, "
            f"{np.mean([x[2] for x in vals]):.2f} s); value = 1/(fuse s/MP + assess s/MP)")
    return {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, "tiles": wl.grid[0] * wl.grid[1], "grid": list(wl.grid), "canvas": [wl.canvas_hw[1], wl.canvas_hw[0]],
                   "levels": wl.levels, "sample": desc},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


# ---------------------------------------------------------------------------------------------- B200 arm
def run_b200(args, wl):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        ge.build()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import super_resolution_system_b200 as pkg
    native = pkg._native
    ctx = native.context(local)
    lib = native.load()

    H, W = wl.canvas_hw
    D = pkg.distributed
    bands = D.balanced_bands(world, H, W, wl.positions_yx, [t.shape[:2] for t in wl.tiles])  # equal cost, not equal height
    r0, r1 = bands[rank]
    e0, e1 = D.widen(bands[rank], H)  # rows this rank fuses (own band + SSIM halo)
    mine = D.tiles_for_rows(wl.positions_yx, [t.shape[0] for t in wl.tiles], e0, e1)
    bm = pkg.BlendingModule(num_levels=wl.levels, device=local)
    qa = pkg.QualityAssessmentModule(device=f"cuda:{local}")

    # ---- device-resident inputs
    d_tiles = [pkg.TileInfo(torch.from_numpy(wl.tiles[i]).to(dev), wl.positions_yx[i][1], wl.positions_yx[i][0], 0, 0) for i in mine]
    d_ref = torch.from_numpy(wl.reference).to(dev)
    d_canvas = torch.zeros((H, W, 3), dtype=torch.uint8, device=dev)

    d_sums = torch.zeros(5, dtype=torch.float64, device=dev)
    plan = bm.plan_fusion(d_tiles, output_shape=(H, W))  # marshal the tile list once, like any serving loop would

    def step_device():
        bm.laplacian_fusion(plan, row_range=(e0, e1), out=d_canvas)
        qa.psnr_ssim_sums_device(d_ref, d_canvas, d_sums, row_range=(r0, r1))  # enqueue only
        s = D.allreduce_device_sums(d_sums, H * W * 3)  # NCCL all-reduce of 5 doubles (N > 1), one read-back
        return qa.finish(s)

    def timed(fn, steps, sampler=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        a.record()
        out = None
        for _ in range(steps):
            out = fn()
        b.record()
        torch.cuda.synchronize()
        t1 = time.time()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out, (t0, t1)

    side = torch.cuda.Stream(device=dev)  # a real stream: repeated fuse calls are replayed as a CUDA graph
    torch.cuda.set_stream(side)
    for _ in range(args.warmup):
        scores = step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    l0 = lib.srf_launch_count(ctx)
    ms_total, scores, (t0, t1) = timed(step_device, args.steps)
    launches = int(lib.srf_launch_count(ctx) - l0)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms_step = ms_total / args.steps
    value = (H * W / 1e6) / (ms_step / 1e3)

    # ---- per-kernel breakdown (same steps again with the library's event hooks on)
    native.profile_enable(ctx, True)
    native.profile_read(ctx)
    ms_prof, _, _ = timed(step_device, args.steps)
    prof = native.profile_read(ctx)
    native.profile_enable(ctx, False)
    agg = {}
    for name, ms, nbytes in prof:
        a = agg.setdefault(name, [0.0, 0.0, 0])
        a[0] += ms
        a[1] += nbytes
        a[2] += 1
    peak, peak_src = measured_peak()
    kernels = []
    for name, (ms, nbytes, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        gbs = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        kernels.append({"kernel": name, "launches_per_step": n / args.steps, "ms_per_step": ms / args.steps,
                        "algorithmic_GB_per_step": nbytes / args.steps / 1e9, "GBps": gbs, "frac_of_peak": gbs / peak})
    top = kernels[0]
    traffic = None  # DRAM bytes of one launch of the dominant kernel, from the committed ncu --set full capture
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
        if wl.name == "cfg4_200mp" and world == 1 and top["launches_per_step"] == 1.0:
            traffic = tj["kernels"].get(top["kernel"], {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": top["kernel"], "achieved": top["GBps"], "peak": peak, "unit": "GB/s",
                "frac": top["frac_of_peak"], "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": top["algorithmic_GB_per_step"] * 1e9 / max(top["launches_per_step"], 1),
                "note": "final_gather and psnr_ssim are instruction / on-chip bound, not HBM bound (ncu, profiles/"
                        "r01_ncu_full_final2_summary.txt: issue-active 59 % / 77 %, smem data pipe 55 % / FMA pipe 61 %, "
                        "DRAM 20 % / 11 %): their DRAM traffic equals the algorithmic bytes, the gap to the HBM roofline is "
                        "arithmetic, not re-reads",
                "share_of_step": top["ms_per_step"] / (ms_prof / args.steps),
                "whole_step": {"algorithmic_GB": (wl.tile_bytes + 3 * H * W + 6 * H * W) / 1e9,
                               "GBps": (wl.tile_bytes + 9 * H * W) / 1e9 / (ms_step / 1e3),
                               "frac": (wl.tile_bytes + 9 * H * W) / 1e9 / (ms_step / 1e3) / peak}}

    # ---- end to end through the public host API (pinned host buffers, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        pipe = pkg.SuperResolutionPipeline(pkg.PipelineConfig(num_pyramid_levels=wl.levels, enable_qa=True))
        pipe.blending_module.device = local
        h_tiles = []
        for i in (range(len(wl.tiles)) if world == 1 else mine):
            pt = torch.from_numpy(wl.tiles[i]).pin_memory()
            h_tiles.append(pkg.TileInfo(pt.numpy(), wl.positions_yx[i][1], wl.positions_yx[i][0], 0, 0))
        h_ref = torch.from_numpy(wl.reference).pin_memory().numpy()
        h_out = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory().numpy()
        moved = [0, 0]

        if world == 1:
            def step_host():  # stage 3 + 4 in one library call (band-pipelined over three streams)
                rep = pipe.fuse_and_assess(h_tiles, h_ref, (H, W), out=h_out)[1]
                moved[0], moved[1] = native.last_transfer_bytes(ctx)
                return rep["psnr"], rep["ms_ssim"]
        else:
            hplan = bm.plan_fusion(h_tiles, output_shape=(H, W))

            def step_host():  # every rank: the band-pipelined host call on ITS band, then one all-reduce of the sums
                _, s = bm.fuse_and_assess(hplan, h_ref, out=h_out, row_range=(r0, r1))
                moved[0], moved[1] = native.last_transfer_bytes(ctx)
                s = D.allreduce_sums(s, H * W * 3, device=dev)
                return qa.finish(s)

        for _ in range(max(1, args.warmup // 2)):
            rep = step_host()
        n_e2e = max(2, args.steps // 2)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ta = time.perf_counter()
        for _ in range(n_e2e):
            rep = step_host()
        torch.cuda.synchronize()
        tb = time.perf_counter()
        if world > 1:
            dist.barrier()
        tt = torch.tensor([tb - ta, float(moved[0]), float(moved[1])], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = tt.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            tt[0] = tmax[0]  # time: slowest rank; bytes: summed over ranks
        secs, h2d_b, d2h_b = tt.tolist()
        e2e = {"value": (H * W / 1e6) / (secs / n_e2e), "unit": UNIT, "h2d_bytes_per_step": int(h2d_b), "d2h_bytes_per_step": int(d2h_b),
               "h2d_bytes_untrimmed": int(wl.tile_bytes + wl.reference.nbytes), "ms_per_step": 1e3 * secs / n_e2e, "steps": n_e2e,
               "api": "SuperResolutionPipeline.fuse_and_assess (srf_fuse_assess_host)" if world == 1 else
                      "per rank: BlendingModule.fuse_and_assess(row_range=band) (srf_fuse_assess_host) + all-reduce of the sums",
               "psnr": rep[0], "ms_ssim": rep[1]}
        del h_tiles, h_ref, h_out

    # ---- the reference's CPU path beside it (bounded sample), rank 0 at N = 1 only
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import oracle
        from oracle import reference_port as port
        oracle.pyr.build()
        threads = os.cpu_count() or 1
        sample = cpu_sample(wl, threads)
        v, tf, tq, mpf, mpq = cpu_step(port, sample, wl.levels, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"tiles 0-1 fused into {sample[2][1]}x{sample[2][0]} ({mpf:.1f} MP, {tf:.2f} s) + PSNR/SSIM on a "
                         f"2048-row strip ({mpq:.1f} MP, {tq:.2f} s); oracle/reference_port.py, cv2+NumPy+SciPy, 1 run"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": wl.name, "tiles": len(wl.tiles), "grid": list(wl.grid), "canvas": [W, H], "megapixels": H * W / 1e6,
                       "levels": wl.levels, "tile": 4096, "l2": "inputs (1.2 GB of tiles + 0.6 GB comparison canvas) exceed the 126 MB L2",
                       "parallelism": f"band{world}" if world > 1 else "single"},
            "psnr": scores[0], "ms_ssim": scores[1],
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks, "ms_per_step_profiled": ms_prof / args.steps,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4_200mp")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0
    import workloads
    t = time.time()
    wl = workloads.build(args.workload, max_tiles=2 if args.impl == "reference" else 0)
    log(f"[bench] workload {wl.name}: {len(wl.tiles)} tiles, canvas {wl.canvas_hw[1]}x{wl.canvas_hw[0]} built in {time.time() - t:.1f}s")
    if args.impl == "reference":
        emit(run_reference(args, wl))
    else:
        run_b200(args, wl)
    return 0


if __name__ == "__main__":
    sys.exit(main())
