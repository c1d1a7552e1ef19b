import numpy as np


def synth(rng, h, w, c=3):
    """textured, non-saturated, non-flat uint8 image; NumPy-only (box-blurred noise, contrast-stretched)"""
    f = rng.random((h + 8, w + 8, c), dtype=np.float32) * 255
    for ax in (0, 1):  # 9-tap box blur, twice ~ triangular
        for _ in range(2):
            cs = np.cumsum(f, axis=ax, dtype=np.float64)
            n = f.shape[ax]
            hi = np.take(cs, np.minimum(np.arange(n) + 4, n - 1), axis=ax)
            lo = np.take(cs, np.maximum(np.arange(n) - 5, 0), axis=ax)
            f = ((hi - lo) / 9).astype(np.float32)
    f = f[4:4 + h, 4:4 + w]
    f = (f - f.mean()) * 6 + 128 + rng.normal(0, 4, f.shape)
    out = np.clip(f, 0, 255).astype(np.uint8)
    return out if c > 1 else out[:, :, 0]


def fuse_case(golden, name):
    n, levels, oh, ow = (int(v) for v in golden[f"fuse_{name}_n"])
    tiles = [golden[f"fuse_{name}_t{i}"] for i in range(n)]
    pos = [tuple(int(v) for v in p) for p in golden[f"fuse_{name}_pos"]]
    return tiles, pos, (oh, ow), levels, str(golden[f"fuse_{name}_kind"]), golden[f"fuse_{name}_out"]


FUSE_CASES = ["grid", "ragged", "linear6", "gray", "single1", "deep"]


def canvas_agreement(a, b):
    """(max abs diff, fraction of exactly equal elements)"""
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return int(d.max()), float((d == 0).mean())


def merge_case(golden):
    """the reference's merge_tiles case: [(data, gx, gy, out_w, out_h, (top, bottom, left, right))], canvas (w, h)"""
    meta, data = golden["merge_meta"], golden["merge_tiles"]
    tiles = [(data[i], int(m[0]), int(m[1]), int(m[4]), int(m[5]), tuple(int(v) for v in m[6:10])) for i, m in enumerate(meta)]
    return tiles, (300, 200)
