import os, sys, time, json
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
d_ref = torch.from_numpy(wl.reference).to(dev)
d_canvas = torch.zeros((H, W, 3), dtype=torch.uint8, device=dev)
d_all = [torch.from_numpy(t).to(dev) for t in wl.tiles]
side = torch.cuda.Stream(device=dev); torch.cuda.set_stream(side)
qa = pkg.QualityAssessmentModule(device="cuda:0")
red = D.SumsExchange(dev)
for world, rank in ((8, 2), (1, 0)):
    bands = D.balanced_bands(world, H, W, wl.positions_yx, [t.shape[:2] for t in wl.tiles])
    r0, r1 = bands[rank]; e0, e1 = D.widen(bands[rank], H)
    mine = D.tiles_for_rows(wl.positions_yx, [t.shape[0] for t in wl.tiles], e0, e1)
    infos = [pkg.TileInfo(d_all[i], wl.positions_yx[i][1], wl.positions_yx[i][0], 0, 0) for i in mine]
    for tag, env in (("fork", {}), ("serial", {"SRF_GATHER_SERIAL": "1"})):
        os.environ.update(env)
        bm = pkg.BlendingModule(num_levels=wl.levels, device=0)
        plan = bm.plan_fusion(infos, output_shape=(H, W))
        def step(sync):
            bm.laplacian_fusion(plan, row_range=(e0, e1), out=d_canvas)
            qa.psnr_ssim_sums_device(d_ref, d_canvas, red.local, row_range=(r0, r1))
            if sync: return red.reduce(H * W * 3)
        for _ in range(10): step(True)
        res = {}
        for sync in (False, True):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(50): step(sync)
            b.record(); torch.cuda.synchronize()
            res["synced" if sync else "free"] = round(a.elapsed_time(b) / 50, 4)
        # host time of the fuse call with the GPU idle
        tt = 0.0
        for _ in range(200):
            torch.cuda.synchronize()
            t0 = time.perf_counter(); bm.laplacian_fusion(plan, row_range=(e0, e1), out=d_canvas); tt += time.perf_counter() - t0
        res["fuse_call_host_us"] = round(tt / 200 * 1e6, 1)
        print(json.dumps({"world": world, "rank": rank, "mode": tag, **res}), flush=True)
        for k in env: del os.environ[k]
