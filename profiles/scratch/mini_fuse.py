import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import __graft_entry__ as ge; ge.build()
import super_resolution_system_b200 as srs, oracle
from helpers import synth
rng = np.random.default_rng(1)
T = 256
tiles = [synth(rng, T, T) for _ in range(2)]
pos = [(0, 0), (0, 200)]
out_hw = (T, 456)
bm = srs.BlendingModule(num_levels=4)
got = bm.laplacian_fusion([srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(tiles, pos)], output_shape=out_hw)
want = oracle.fusion_oracle.laplacian_fuse(tiles, pos, out_hw, 4, "cosine", "c")
print("equal", np.array_equal(got, want), np.abs(got.astype(int) - want).max())
