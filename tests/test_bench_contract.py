"""CPU suite, part 4: the bench.py contract -- the reference arm runs without a GPU and prints exactly ONE JSON line with
the keys the driver reads; BASELINE.json's configs map to workloads the bench knows."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    # --ref-sample: the bounded 2-tile sample (the default arm runs the WHOLE config, ~1.5 min per step -- the driver's job)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-sample"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fused_canvas_megapixels_per_second" and d["unit"] == "MP/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    # "reference" = the reference's own modules from oracle/_ref (present wherever build() saw /root/reference), else the port
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert cb["kind"] == ("reference" if os.path.isdir("/root/reference") or os.path.exists(os.path.join(ROOT, "oracle", "_ref", "blending_module.py")) else "port")
    assert d["e2e"] == {"value": d["value"], "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "cfg4_200mp" and d["config"]["tiles"] == 24 and d["config"]["levels"] == 7


def test_workloads_cover_baseline_configs():
    sys.path.insert(0, ROOT)
    import workloads
    names = set(workloads.CONFIGS) if hasattr(workloads, "CONFIGS") else set()
    src = open(os.path.join(ROOT, "workloads.py")).read()
    for name in ("cfg1_100mp_720p", "cfg2_100mp", "cfg3_150mp", "cfg4_200mp"):
        assert name in names or name in src, name


def test_traffic_figures_are_tied_to_the_sources_of_their_own_kernel(tmp_path, monkeypatch):
    """bench.traffic_for attaches a committed ncu DRAM figure only while the digest of THAT kernel's translation units is the
    one recorded with the figure: a change to srf_iqa.cu drops the SSIM figure and keeps the fuse kernels', a change to the
    shared header drops all of them; launch counts that differ from the capture's drop the figure too"""
    sys.path.insert(0, ROOT)
    import shutil
    import bench
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
        tj = json.load(f)
    assert set(tj["kernels"]) >= {"psnr_ssim", "final_gather", "up_level", "down_u8", "down_f32"}
    for name, k in tj["kernels"].items():  # the committed table matches the committed sources
        assert k["src_digest"] == bench.csrc_digest(name), name
        got, dig = bench.traffic_for(name, "cfg4_200mp", 1, k["launches"])
        assert got == k["dram_bytes_per_step"] and dig == k["src_digest"]
        assert bench.traffic_for(name, "cfg4_200mp", 1, k["launches"] + 1)[0] is None
        assert bench.traffic_for(name, "cfg4_200mp", 2, k["launches"])[0] is None  # per-rank traffic of a band is another number
    # a scratch copy of the tree: touch one translation unit at a time
    root2 = tmp_path / "repo"
    shutil.copytree(os.path.join(ROOT, "super-resolution-system_b200", "csrc"), root2 / "super-resolution-system_b200" / "csrc")
    shutil.copytree(os.path.join(ROOT, "profiles"), root2 / "profiles", ignore=shutil.ignore_patterns("*.csv", "*.txt", "scratch"))
    monkeypatch.setattr(bench, "ROOT", str(root2))
    with open(root2 / "super-resolution-system_b200" / "csrc" / "srf_iqa.cu", "a") as f:
        f.write("// touched\n")
    assert bench.traffic_for("psnr_ssim", "cfg4_200mp", 1, 1)[0] is None
    assert bench.traffic_for("final_gather", "cfg4_200mp", 1, tj["kernels"]["final_gather"]["launches"])[0] is not None
    with open(root2 / "super-resolution-system_b200" / "csrc" / "srf_common.cuh", "a") as f:
        f.write("// touched\n")
    assert bench.traffic_for("final_gather", "cfg4_200mp", 1, tj["kernels"]["final_gather"]["launches"])[0] is None


def test_cpu_sample_is_a_two_by_two_block_scaled_per_tile():
    """the bounded CPU sample: tiles 0, 1 and the two below them (every overlap kind), and the value scales the fusion per
    tile and the assessment per megapixel to the whole workload"""
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench

    class WL:
        grid = (6, 4)
        canvas_hw = (11546, 17320)
        tiles = [np.zeros((8, 8, 3), np.uint8) for _ in range(24)]
        positions_yx = [(3482 * (i // 6), 3278 * (i % 6)) for i in range(24)]
        reference = np.zeros((11546, 17320, 1), np.uint8)

    tiles, pos, hw, ref = bench.cpu_sample(WL)
    assert pos == [(0, 0), (0, 3278), (3482, 0), (3482, 3278)] and hw == (3490, 3286) and ref.shape[:2] == hw
    WL.tiles, WL.positions_yx = WL.tiles[:2], WL.positions_yx[:2]
    assert len(bench.cpu_sample(WL)[0]) == 2  # not enough tiles for a block: the first two
    calls = {"fuse": 0}

    def fuse(t, p, hw_, levels):
        calls["fuse"] += 1
        return np.zeros(hw_ + (3,), np.uint8)

    backend = ("port", fuse, lambda a, b: 30.0, lambda a, b: 0.9, "stub")
    a = np.zeros((64, 64, 3), np.uint8)
    v, tf, tq, mpf, mpq, p, s = bench.cpu_step(backend, [a, a, a, a], [(0, 0)] * 4, (64, 64), np.zeros((64, 64, 3), np.uint8), 3,
                                               iqa_rows=32, whole=(24, 200.0))
    assert calls["fuse"] == 1 and (p, s) == (30.0, 0.9) and mpq == 32 * 64 / 1e6
    assert v == 200.0 / (tf / 4 * 24 + tq / mpq * 200.0)
