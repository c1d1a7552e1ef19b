import cProfile, pstats, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, workloads
import __graft_entry__ as ge
ge.build()
import super_resolution_system_b200 as pkg
wl = workloads.build("cfg4_200mp")
dev = torch.device("cuda", 0)
D = pkg.distributed
H, W = wl.canvas_hw
bands = D.balanced_bands(8, H, W, wl.positions_yx, [t.shape[:2] for t in wl.tiles])
r0, r1 = bands[2]; e0, e1 = D.widen(bands[2], H)
mine = D.tiles_for_rows(wl.positions_yx, [t.shape[0] for t in wl.tiles], e0, e1)
d_all = {i: torch.from_numpy(wl.tiles[i]).to(dev) for i in mine}
d_ref = torch.from_numpy(wl.reference).to(dev)
d_canvas = torch.zeros((H, W, 3), dtype=torch.uint8, device=dev)
side = torch.cuda.Stream(device=dev); torch.cuda.set_stream(side)
bm = pkg.BlendingModule(num_levels=wl.levels, device=0); qa = pkg.QualityAssessmentModule(device="cuda:0")
infos = [pkg.TileInfo(d_all[i], wl.positions_yx[i][1], wl.positions_yx[i][0], 0, 0) for i in mine]
plan = bm.plan_fusion(infos, output_shape=(H, W))
red = D.SumsExchange(dev)
def step():
    bm.laplacian_fusion(plan, row_range=(e0, e1), out=d_canvas)
    qa.psnr_ssim_sums_device(d_ref, d_canvas, red.local, row_range=(r0, r1))
    return qa.finish(red.reduce(H * W * 3))
for _ in range(20): step()
torch.cuda.synchronize()
# host-only cost of the enqueue calls: GPU kept far behind? no -- measure the calls with the GPU idle-synced each step
t = {"fuse": 0.0, "ssim": 0.0, "reduce": 0.0, "finish": 0.0}
n = 500
for _ in range(n):
    a = time.perf_counter(); bm.laplacian_fusion(plan, row_range=(e0, e1), out=d_canvas)
    b = time.perf_counter(); qa.psnr_ssim_sums_device(d_ref, d_canvas, red.local, row_range=(r0, r1))
    c = time.perf_counter(); s = red.reduce(H * W * 3)
    d = time.perf_counter(); qa.finish(s)
    e = time.perf_counter()
    t["fuse"] += b - a; t["ssim"] += c - b; t["reduce"] += d - c; t["finish"] += e - d
print({k: round(v / n * 1e6, 1) for k, v in t.items()}, "us per call (reduce includes waiting for the GPU)")
pr = cProfile.Profile(); pr.enable()
for _ in range(500): step()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:3500])
