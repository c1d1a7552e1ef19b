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
for name, shapes, pos, out_hw in [
    ("single_in_canvas", [(640, 704)], [(64, 128)], (800, 1000)),
    ("single_at_origin", [(640, 704)], [(0, 0)], (640, 704)),
    ("two_vertical", [(640, 704), (640, 704)], [(0, 0), (500, 0)], (1140, 704)),
]:
    hist = collections.Counter()
    for seed in range(6):
        rng = np.random.default_rng(seed)
        tiles = [synth(rng, h, w, 3) for h, w in shapes]
        infos = [srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(tiles, pos)]
        a = run(infos, out_hw, 5, "sigmoid", {})
        b = run(infos, out_hw, 5, "sigmoid", {"SRF_GATHER_ONE_CLASS": "1"})
        ys, xs, cs = np.nonzero(a.astype(int) - b.astype(int))
        for y, x in zip(ys.tolist(), xs.tolist()):
            hist[(y, x % 128)] += 1
    print(name, "diffs by (row, col mod 128):", sorted(hist.items())[:40])
