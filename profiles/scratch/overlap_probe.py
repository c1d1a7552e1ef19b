"""fuse (stream A) and psnr_ssim (stream B, independent canvas) alone vs together: how much do the two kernels' stalls interleave?"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import workloads
import __graft_entry__ as ge
ge.build()
import super_resolution_system_b200 as pkg
dev = torch.device("cuda", 0)
wl = workloads.build("cfg4_200mp")
H, W = wl.canvas_hw
bm = pkg.BlendingModule(num_levels=wl.levels, device=0)
qa = pkg.QualityAssessmentModule(device="cuda:0")
tiles = [pkg.TileInfo(torch.from_numpy(wl.tiles[i]).to(dev), wl.positions_yx[i][1], wl.positions_yx[i][0], 0, 0) for i in range(len(wl.tiles))]
d_ref = torch.from_numpy(wl.reference).to(dev)
canvas = torch.zeros((H, W, 3), dtype=torch.uint8, device=dev)
A, B = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
plan = bm.plan_fusion(tiles, output_shape=(H, W))
sums = torch.zeros(5, dtype=torch.float64, device=dev)
with torch.cuda.stream(A):
    for _ in range(3):
        bm.laplacian_fusion(plan, out=canvas)
torch.cuda.synchronize()
canvas2 = canvas.clone()

def fuse():
    with torch.cuda.stream(A):
        bm.laplacian_fusion(plan, out=canvas)

def ssim(rows=None):
    with torch.cuda.stream(B):
        qa.psnr_ssim_sums_device(d_ref, canvas2, sums, row_range=rows or (0, H))

def timed(fns, reps=8):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); ea = torch.cuda.Event(enable_timing=True); eb = torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        A.wait_event(e0); B.wait_event(e0)
        for f in fns:
            f()
        ea.record(A); eb.record(B)
        torch.cuda.synchronize()
        best = min(best, max(e0.elapsed_time(ea), e0.elapsed_time(eb)))
    return best

out = {"fuse_alone_ms": timed([fuse]), "ssim_alone_ms": timed([ssim]), "both_ms": timed([fuse, ssim]), "both_ssim_first_ms": timed([ssim, fuse])}
out["serial_sum_ms"] = out["fuse_alone_ms"] + out["ssim_alone_ms"]
print(json.dumps(out))
