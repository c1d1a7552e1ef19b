import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
ge.build()
import super_resolution_system_b200 as srs
import oracle
from helpers import synth
rng = np.random.default_rng(43)
tiles, pos = [], []
for j in range(2):
    for i in range(2):
        tiles.append(synth(rng, 640, 704, 3)); pos.append((j * 500, i * 520))
out_hw = (1100, 1190)
infos = [srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(tiles, pos)]
for levels, kind in ((3, "cosine"), (5, "sigmoid"), (5, "cosine"), (5, "linear")):
    want = oracle.fusion_oracle.laplacian_fuse(tiles, pos, out_hw, levels, kind, "c")
    res = {}
    for tag, env in (("classified", {}), ("no_edge", {"SRF_GATHER_NO_EDGE": "1"}), ("generic", {"SRF_GATHER_ONE_CLASS": "1"}), ("legacy", {"SRF_LEGACY_GATHER": "1"})):
        os.environ.update(env)
        res[tag] = srs.BlendingModule(num_levels=levels).laplacian_fusion(infos, output_shape=out_hw, weight_type=srs.WeightType(kind))
        for k in env: del os.environ[k]
        d = res[tag].astype(int) - want.astype(int)
        print(levels, kind, tag, "vs oracle: mismatches", int((d != 0).sum()), "max", int(np.abs(d).max()))
    for tag in ("no_edge", "generic", "legacy"):
        d = res[tag].astype(int) - res["classified"].astype(int)
        ys, xs, cs = np.nonzero(d)
        print("   classified vs", tag, ":", len(ys), "diffs", list(zip(ys[:12].tolist(), xs[:12].tolist(), cs[:12].tolist(), d[ys[:12], xs[:12], cs[:12]].tolist())))
