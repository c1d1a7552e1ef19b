import os, sys, collections
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
ge.build()
import super_resolution_system_b200 as srs
from helpers import synth
def run(infos, out_hw, levels, kind, env):
    os.environ.update(env)
    r = srs.BlendingModule(num_levels=levels).laplacian_fusion(infos, output_shape=out_hw, weight_type=srs.WeightType(kind))
    for k in env: del os.environ[k]
    return r
for name, shapes, pos, out_hw, env in [
    ("two_vertical", [(640, 704), (640, 704)], [(0, 0), (500, 0)], (1140, 704), {}),
    ("two_vertical_serial", [(640, 704), (640, 704)], [(0, 0), (500, 0)], (1140, 704), {"SRF_GATHER_SERIAL": "1"}),
    ("two_vertical_x6", [(640, 704), (640, 704)], [(0, 6), (500, 6)], (1140, 720), {}),
    ("two_vertical_x20", [(640, 704), (640, 704)], [(0, 20), (500, 20)], (1140, 740), {}),
    ("two_vertical_swapped_order", [(640, 704), (640, 704)], [(500, 0), (0, 0)], (1140, 704), {}),
    ("two_vertical_L3", [(640, 704), (640, 704)], [(0, 0), (500, 0)], (1140, 704), {"LV": "3"}),
]:
    hist = collections.Counter()
    lv = int(env.pop("LV", "5"))
    for seed in range(8):
        rng = np.random.default_rng(seed)
        tiles = [synth(rng, h, w, 3) for h, w in shapes]
        infos = [srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(tiles, pos)]
        a = run(infos, out_hw, lv, "sigmoid", dict(env))
        b = run(infos, out_hw, lv, "sigmoid", {"SRF_GATHER_ONE_CLASS": "1"})
        d = a.astype(int) - b.astype(int)
        ys, xs, cs = np.nonzero(d)
        for y, x, c in zip(ys.tolist(), xs.tolist(), cs.tolist()):
            hist[(y, x, int(d[y, x, c]))] += 1
    print(name, "diffs (row, col, delta):", sorted(hist.items())[:30])
