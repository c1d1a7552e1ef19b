"""Band sharding across the GPUs of one box (one process per GPU, torch.distributed).

The canvas is cut into `world` horizontal bands.  A rank fuses its band (plus a 5-row halo each side, so the 11x11
SSIM window never needs a neighbour's pixels) from the tiles that reach it -- per-tile pyramids have tile-local
borders, so a rank recomputes the pyramid rows its band depends on (libsrfuse derives those row ranges itself from
`row_begin/row_end`).  The only exchange is the sum of the additive PSNR / SSIM partial sums (srf_iqa_sums): on the
GPUs it is libsrfuse's own peer-memory exchange (SumsExchange -> srf_allreduce_metrics: one small kernel per rank that
writes into every peer's inbox over NVLink, no NCCL call in the step); torch.distributed only carries the 64-byte IPC
handles at start-up, and is the fallback collective where peer mapping is unavailable (and gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

from . import _native
from ._arrays import torch

IQA_HALO = 5  # radius of the 11-tap SSIM window


def _cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if part:
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_device_node(device_index: int, sysfs: str = "/sys/bus/pci/devices"):
    """Pin the calling process to the host cores next to GPU `device_index` (its PCIe root's NUMA node), so that the
    pinned buffers it allocates afterwards are first-touched on that node and host<->device copies do not cross the
    socket interconnect.  Call before allocating pinned memory.  Returns {"node", "cpus"} or None when the box gives
    no locality (single node, no sysfs entry, or the allowed cpu set does not reach that node) -- then nothing changes."""
    import os
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"{sysfs}/{bdf}/numa_node") as f:
            node = int(f.read())
        with open(f"{sysfs}/{bdf}/local_cpulist") as f:
            local = _cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        use = local & allowed
        if node < 0 or not use or use == allowed:
            return None
        os.sched_setaffinity(0, use)
        return {"node": node, "cpus": len(use)}
    except (OSError, ValueError, AttributeError):
        return None


def band(rank: int, world: int, rows: int) -> Tuple[int, int]:
    """canvas rows [r0, r1) owned by `rank`: ceil(rows / world) rows each, the last band takes what is left"""
    per = -(-rows // world)
    return min(rank * per, rows), min((rank + 1) * per, rows)


def balanced_bands(world: int, rows: int, cols: int, positions_yx, sizes_hw, row_cost: float = 1.13,
                   tile_cost: float = 100.0) -> List[Tuple[int, int]]:
    """Cut the canvas rows into `world` contiguous bands of equal COST rather than equal height.  Cost of a band, in
    units of "one tile covering one full canvas row":
        row_cost per canvas row        the per-pixel work: assessment, the canvas write and the gather's per-block part
      + 1 per covering tile and row    the gather's per-tile work (scaled by the covered fraction of the row)
      + tile_cost per tile touched     what a band pays for every tile it reaches at all, however few rows: that tile's
                                       coarse pyramid levels and halo rows (scaled by tile width / 4096)
    The constants are a least-squares fit to the per-rank kernel times of the 8-band 200 MP job measured with
    profiles/band_probe.py on the round-2 kernels (0.177 us per canvas row, 0.157 us per covered row, 15.8 us per
    touched 4096-wide tile; the fit reproduces every rank's time to 2 %): at that band height the pyramid kernels are
    launch-latency and halo bound, i.e. a per-tile constant, not a per-row cost.  The cuts minimise the most expensive
    band (binary search on the bound, greedy fill)."""
    import numpy as np
    if world <= 1 or rows <= 0:
        return [(0, max(rows, 0))] + [(max(rows, 0), max(rows, 0))] * (max(world, 1) - 1)
    per_row = np.full(rows, float(row_cost))
    tiles = []
    for (y, x), (h, w) in zip(positions_yx, sizes_hw):
        wc = max(0, min(cols, x + w) - max(0, x))
        if wc == 0 or y >= rows or y + h <= 0:
            continue
        per_row[max(0, y):min(rows, y + h)] += wc / cols
        tiles.append((y, y + h, tile_cost * wc / 4096.0))
    cum = np.concatenate([[0.0], np.cumsum(per_row)])

    def cost(r0, r1):
        e0, e1 = r0 - IQA_HALO, r1 + IQA_HALO
        return cum[r1] - cum[r0] + sum(f for (y0, y1, f) in tiles if y0 < e1 and y1 > e0)

    def fill(bound):
        """greedy: every band as tall as `bound` allows; None if more than `world` bands would be needed"""
        cuts, s = [0], 0
        for _ in range(world):
            lo, hi = s, rows  # largest e with cost(s, e) <= bound (cost is monotone in e)
            if cost(s, min(s + 1, rows)) > bound:
                return None
            while lo < hi:
                mid = (lo + hi + 1) // 2
                if cost(s, mid) <= bound:
                    lo = mid
                else:
                    hi = mid - 1
            s = lo
            cuts.append(s)
            if s >= rows:
                break
        return cuts if cuts[-1] >= rows else None

    lo_b, hi_b = cost(0, rows) / world, cost(0, rows)
    for _ in range(40):
        mid = 0.5 * (lo_b + hi_b)
        if fill(mid) is None:
            lo_b = mid
        else:
            hi_b = mid
    cuts = fill(hi_b)
    cuts += [rows] * (world + 1 - len(cuts))
    return [(a, b) for a, b in zip(cuts[:-1], cuts[1:])]


def widen(band_rows: Tuple[int, int], rows: int) -> Tuple[int, int]:
    """rows a rank must FUSE for a band: the band plus the SSIM halo, clipped to the canvas"""
    r0, r1 = band_rows
    if r0 >= r1:
        return r0, r0
    # the start is rounded down to an even row: tile origins are even (2 x the low-resolution origin), and the gather
    # kernel's predicate-free interior path needs canvas blocks of even parity relative to the tile -- an odd start
    # puts every block of the band on the generic path (+60 % on that kernel)
    return max(r0 - IQA_HALO, 0) & ~1, min(r1 + IQA_HALO, rows)


def fuse_rows(rank: int, world: int, rows: int) -> Tuple[int, int]:
    """rows a rank must FUSE: its (equal-height) band widened by the SSIM halo, clipped to the canvas"""
    return widen(band(rank, world, rows), rows)


def tiles_for_rows(positions_yx: Sequence[Tuple[int, int]], heights: Sequence[int], r0: int, r1: int) -> List[int]:
    """indices (list order kept -- it is the accumulation order) of the tiles that intersect canvas rows [r0, r1)"""
    return [i for i, ((y, _), h) in enumerate(zip(positions_yx, heights)) if y < r1 and y + h > r0]


def allreduce_sums(sums, total_elems: int, group=None, device=None):
    """sum srf_iqa_sums over the ranks of `group` (in place) and return it.  sq_err and ssim_count are integers below
    2^53, exact in float64; bits_a / bits_b travel as their <= 1 / > 1 class."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sums
    # one collective: the two OR-of-bytes fields only matter through "is any byte > 1", which sums fine
    t = torch.tensor([float(sums.sq_err), sums.ssim_sum, float(sums.ssim_count), float(sums.bits_a > 1), float(sums.bits_b > 1)],
                     dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    v = t.tolist()
    sums.sq_err, sums.ssim_sum, sums.ssim_count = int(v[0]), v[1], int(v[2])
    sums.n_elems = int(total_elems)
    sums.bits_a = 255 if v[3] > 0 else min(int(sums.bits_a), 1)
    sums.bits_b = 255 if v[4] > 0 else min(int(sums.bits_b), 1)
    return sums


def allreduce_device_sums(d_sums, total_elems: int, group=None):
    """`d_sums`: the 5-double CUDA tensor QualityAssessmentModule.psnr_ssim_sums_device filled.  One NCCL all-reduce
    (skipped at world size 1), ONE device->host read; returns an srf_iqa_sums with the job-wide totals."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(d_sums, op=dist.ReduceOp.SUM, group=group)
    v = d_sums.tolist()
    s = _native.srf_iqa_sums()
    s.sq_err, s.ssim_sum, s.ssim_count, s.n_elems = int(v[0]), v[1], int(v[2]), int(total_elems)
    s.bits_a, s.bits_b = (255 if v[3] > 0 else 1), (255 if v[4] > 0 else 1)
    return s


class SumsExchange:
    """Job-wide totals of the additive IQA sums, `slots` x 5 doubles {sq_err, ssim_sum, ssim_count, any byte of a > 1, of
    b > 1} (slot 0 = the banded canvas pair; further slots = whole pairs owned by single ranks).

    kind == "peer"  libsrfuse's NVLink exchange (include/srfuse.h: srf_comm_handle / srf_comm_init /
                    srf_allreduce_metrics / srf_comm_result): enqueue-only kernel, totals land in pinned host memory
    kind == "nccl"  torch.distributed all_reduce + one read-back (when the peers' memory cannot be mapped)
    kind == "local" world size 1: only the read-back
    `local` is the device tensor the IQA kernels write a rank's own sums into; reduce() leaves the totals in it."""

    def __init__(self, device, group=None, slots: int = 1, allow_peer: bool = True):
        import ctypes as C
        import torch.distributed as dist
        self.device, self.group, self.slots = device, group, int(slots)
        self.local = torch.zeros(5 * self.slots, dtype=torch.float64, device=device)
        self._scratch = None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.kind = "local"
        self._lib, self._ctx = _native.load(), _native.context(device.index)
        self._host = (C.c_double * (5 * self.slots))()
        if self.world == 1:
            return
        self.kind = "nccl"
        if not allow_peer or 5 * self.slots > 128 or self.world > 16:
            return
        handle = (C.c_uint8 * 64)()
        ok = self._lib.srf_comm_handle(self._ctx, handle) == 0
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
        every = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(every, mine, group=group)
        if ok:
            blob = bytes(torch.stack(every).cpu().numpy().tobytes())
            ok = self._lib.srf_comm_init(self._ctx, self.rank, self.world, blob) == 0
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)  # all ranks or none: a mixed job would deadlock
        if int(flag.item()) == 1:
            self.kind = "peer"

    # -- per-step ----------------------------------------------------------------------------------------------
    def _totals(self):
        """sum self.local over the ranks (in place) and return the 5 * slots totals as Python floats"""
        if self.kind == "peer":
            st = torch.cuda.current_stream(self.device).cuda_stream
            _native.check(self._lib.srf_allreduce_metrics(self._ctx, self.local.data_ptr(), 5 * self.slots, st), self._ctx)
            _native.check(self._lib.srf_comm_result(self._ctx, self._host, 5 * self.slots, st), self._ctx)
            return list(self._host)
        if self.kind == "nccl":
            import torch.distributed as dist
            dist.all_reduce(self.local, op=dist.ReduceOp.SUM, group=self.group)
        return self.local.tolist()

    def enqueue(self):
        """the exchange alone, on the current stream, nothing read: sums self.local over the ranks in place (and, for the
        peer kind, into pinned host memory); fetch() reads the totals of the LAST enqueued exchange"""
        if self.kind == "peer":
            st = torch.cuda.current_stream(self.device).cuda_stream
            _native.check(self._lib.srf_allreduce_metrics(self._ctx, self.local.data_ptr(), 5 * self.slots, st), self._ctx)
        elif self.kind == "nccl":
            import torch.distributed as dist
            dist.all_reduce(self.local, op=dist.ReduceOp.SUM, group=self.group)

    def fetch(self, total_elems: int, stream=None):
        """slot 0 of the last enqueue()d exchange, after synchronising `stream` (default: the current one)"""
        st = torch.cuda.current_stream(self.device) if stream is None else stream
        if self.kind == "peer":
            _native.check(self._lib.srf_comm_result(self._ctx, self._host, 5 * self.slots, st.cuda_stream), self._ctx)
            return self._sums(list(self._host)[:5], total_elems)
        st.synchronize()
        return self._sums(self.local.tolist()[:5], total_elems)

    @staticmethod
    def _sums(v, total_elems):
        s = _native.srf_iqa_sums()
        s.sq_err, s.ssim_sum, s.ssim_count, s.n_elems = int(v[0]), v[1], int(v[2]), int(total_elems)
        s.bits_a, s.bits_b = (255 if v[3] > 0 else 1), (255 if v[4] > 0 else 1)
        return s

    def reduce(self, total_elems: int):
        """slot 0 only: the job-wide srf_iqa_sums of the banded pair"""
        return self._sums(self._totals()[:5], total_elems)

    def reduce_many(self, total_elems):
        v = self._totals()
        return [self._sums(v[5 * k:5 * k + 5], n) for k, n in enumerate(total_elems)]

    def scratch(self, n: int):
        """5 * n doubles for a batched launch's per-pair sums (scatter_scratch moves them to their slots)"""
        if self._scratch is None or self._scratch.numel() < 5 * n:
            self._scratch = torch.zeros(5 * n, dtype=torch.float64, device=self.device)
        return self._scratch

    def scatter_scratch(self, slot_ids):
        key = tuple(slot_ids)
        if getattr(self, "_idx_key", None) != key:
            self._idx = torch.tensor([5 * k + j for k in key for j in range(5)], dtype=torch.long, device=self.device)
            self._idx_key = key
        self.local.index_copy_(0, self._idx, self._scratch[:self._idx.numel()])

    def _load_host(self, slot_sums):
        self.local.zero_()
        if slot_sums:
            vals = torch.zeros(5 * self.slots, dtype=torch.float64)
            for k, s in slot_sums:
                vals[5 * k:5 * k + 5] = torch.tensor([float(s.sq_err), s.ssim_sum, float(s.ssim_count), float(s.bits_a > 1), float(s.bits_b > 1)],
                                                     dtype=torch.float64)
            self.local.copy_(vals)

    def reduce_host(self, sums, total_elems: int):
        """host-side srf_iqa_sums of this rank's band -> job-wide sums (the e2e legs)"""
        if self.world == 1:
            sums.n_elems = int(total_elems)
            return sums
        self._load_host([(0, sums)])
        return self.reduce(total_elems)

    def reduce_many_host(self, slot_sums, total_elems):
        self._load_host(slot_sums)
        return self.reduce_many(total_elems)


class BandedHostFuser:
    """Stage 3 + 4 of the pipeline for a band-sharded job whose tiles live in HOST memory: every tile row crosses PCIe
    exactly ONCE -- to the rank whose band contains it -- and the pyramid halo a band needs beyond its own rows (about
    4 * 2^(levels-1) tile rows per side) travels GPU -> GPU over NVLink as one grouped NCCL send/recv per step, instead
    of every rank uploading its whole dependency cone from the host (srf_fuse_assess_host on a band: 2.95 GB of H2D
    for the 8-rank 200 MP job against 1.5 GB here).  Then the device path runs: srf_laplacian_fuse on the band,
    srf_psnr_ssim_enqueue against this rank's rows of the comparison image, SumsExchange, and the band goes back to
    the host.

    tiles: TileInfo list with HOST uint8 images (pinned memory makes the copies asynchronous); reference: host canvas-
    sized image; out: host canvas the rank's band is written into.  Every rank passes the same tile list and bands."""

    def __init__(self, bm, qa, tiles, reference, canvas_hw, bands, rank: int, world: int, device, exchange=None):
        import ctypes as C
        import numpy as np
        self.bm, self.qa, self.rank, self.world, self.device = bm, qa, rank, world, device
        self.H, self.W = int(canvas_hw[0]), int(canvas_hw[1])
        self.bands = [tuple(b) for b in bands]
        self.r0, self.r1 = self.bands[rank]
        self.e0, self.e1 = widen(self.bands[rank], self.H)
        self.exchange = exchange or SumsExchange(device)
        self.host_tiles = [np.asarray(t.image) for t in tiles]
        self.pos = [(int(t.y), int(t.x)) for t in tiles]
        self.reference = np.asarray(reference)
        lib = _native.load()
        levels = bm.num_levels

        def cone(q, i):  # tile rows / columns rank q's fused rows depend on
            f0, f1 = widen(self.bands[q], self.H)
            if f1 <= f0:
                return 0, 0, 0
            h, w = self.host_tiles[i].shape[:2]
            a, b, c = C.c_int(), C.c_int(), C.c_int()
            _native.check(lib.srf_fuse_input_rows(h, w, self.pos[i][0], self.pos[i][1], levels, self.H, self.W, f0, f1,
                                                  C.byref(a), C.byref(b), C.byref(c)))
            return a.value, b.value, c.value

        def own(q, i):  # tile rows that belong to rank q: those inside its band (the first / last band also take what overhangs the canvas)
            h = self.host_tiles[i].shape[0]
            b0, b1 = self.bands[q]
            if b1 <= b0:
                return 0, 0
            lo = -10 ** 9 if all(self.bands[p][1] <= self.bands[p][0] or p >= q for p in range(world)) else b0
            hi = 10 ** 9 if all(self.bands[p][1] <= self.bands[p][0] or p <= q for p in range(world)) else b1
            y = self.pos[i][0]
            return max(lo - y, 0), max(min(hi - y, h), 0)

        n = len(tiles)
        self.mine = [i for i in range(n) if cone(rank, i)[1] > cone(rank, i)[0]]
        self.cones = {i: cone(rank, i) for i in self.mine}
        # what I upload: my own rows of every tile I work on, restricted to the cone's rows and columns
        self.uploads = []
        for i in self.mine:
            a, b, c = self.cones[i]
            oa, ob = own(rank, i)
            lo, hi = max(a, oa), min(b, ob)
            if hi > lo:
                self.uploads.append((i, lo, hi, c))
        # halo rows: (peer, tile, lo, hi) I receive / send -- every rank derives the same lists from the same geometry
        self.recvs, self.sends = [], []
        for q in range(world):
            for i in range(n):
                a, b, _ = cone(q, i)
                if b <= a:
                    continue
                for p in range(world):
                    if p == q:
                        continue
                    oa, ob = own(p, i)
                    lo, hi = max(a, oa), min(b, ob)
                    if hi > lo:
                        if q == rank:
                            self.recvs.append((p, i, lo, hi))
                        elif p == rank:
                            self.sends.append((q, i, lo, hi, cone(q, i)[2]))
        # a sender must hold (and upload) the rows AND columns it ships even where its own cone would not need them
        held = {i: [lo, hi, c] for i, lo, hi, c in self.uploads}
        for q, i, lo, hi, cq in self.sends:
            if i not in held:
                held[i] = [lo, hi, cq]
            else:
                held[i] = [min(held[i][0], lo), max(held[i][1], hi), max(held[i][2], cq)]
        self.sends = [(q, i, lo, hi) for q, i, lo, hi, _ in self.sends]
        self.uploads = [(i, v[0], v[1], v[2]) for i, v in sorted(held.items())]
        need_dev = sorted(set(self.mine) | set(held))
        self.d_tiles = {i: torch.empty(self.host_tiles[i].shape, dtype=torch.uint8, device=device) for i in need_dev}
        from .blending_module import TileInfo
        self.plan = bm.plan_fusion([TileInfo(self.d_tiles[i], self.pos[i][1], self.pos[i][0], 0, 0) for i in self.mine],
                                   output_shape=(self.H, self.W))
        self.d_canvas = torch.zeros((self.H, self.W) + self.host_tiles[0].shape[2:], dtype=torch.uint8, device=device)
        self.q0, self.q1 = max(self.r0 - IQA_HALO, 0), min(self.r1 + IQA_HALO, self.H)  # comparison rows the SSIM window reads
        self.d_ref = torch.empty((max(self.q1 - self.q0, 0),) + self.reference.shape[1:], dtype=torch.uint8, device=device)
        self.h2d_bytes = self.d2h_bytes = self.p2p_bytes = 0

    def _sub_bands(self):
        """the rank's fused rows cut into a few sub-bands (starts on multiples of 16 rows): uploads of later sub-bands and
        downloads of earlier ones overlap the kernels of the current one"""
        rows = self.e1 - self.e0
        k = max(1, min(4, rows // 1024))
        per = -(-rows // k)
        per = (per + 15) // 16 * 16
        cuts = [min(self.e0 + j * per, self.e1) for j in range(k)] + [self.e1]
        return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]

    def step(self, out):
        """one fuse + assess of this rank's band; returns the job-wide srf_iqa_sums (identical on every rank).
        Three streams: uploads (host -> device), kernels + halo exchange, downloads.  Order of the uploads: first the rows
        the neighbours will ask for (both ends of the band), then the halo exchange, then the rest sub-band by sub-band."""
        import ctypes as C
        import torch.distributed as dist
        lib = _native.load()
        comp = torch.cuda.current_stream(self.device)
        if not hasattr(self, "_up"):
            self._up, self._down = torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)
            self._subs = self._sub_bands()
            # per sub-band and tile: the last tile row its dependency cone needs
            self._cone_hi = []
            for (f0, f1) in self._subs:
                hi = {}
                for i in self.mine:
                    h, w = self.host_tiles[i].shape[:2]
                    a, b, c = C.c_int(), C.c_int(), C.c_int()
                    _native.check(lib.srf_fuse_input_rows(h, w, self.pos[i][0], self.pos[i][1], self.bm.num_levels, self.H, self.W, f0, f1,
                                                          C.byref(a), C.byref(b), C.byref(c)))
                    hi[i] = b.value
                self._cone_hi.append(hi)
        up, down = self._up, self._down
        nch = self.reference.shape[2] if self.reference.ndim == 3 else 1
        h2d = 0
        start = torch.cuda.Event()
        start.record(comp)
        up.wait_event(start)
        down.wait_event(start)
        srcs = {i: torch.from_numpy(self.host_tiles[i]) for i, _, _, _ in self.uploads}
        done = {i: lo for i, lo, hi, c in self.uploads}     # rows [lo, done[i]) of the own range are on the device
        spans = {i: (lo, hi, c) for i, lo, hi, c in self.uploads}

        def upload(i, lo, hi):
            nonlocal h2d
            a, b, c = spans[i]
            lo, hi = max(lo, a), min(hi, b)
            if hi > lo:
                self.d_tiles[i][lo:hi, :c].copy_(srcs[i][lo:hi, :c], non_blocking=True)
                h2d += (hi - lo) * c * nch

        with torch.cuda.stream(up):
            # phase A: rows the neighbours need (they sit at the two ends of my own range); remember them so they are not sent twice
            sent_rows = {}
            for q, i, lo, hi in self.sends:
                upload(i, lo, hi)
                sent_rows.setdefault(i, []).append((lo, hi))
            ev_a = torch.cuda.Event()
            ev_a.record(up)
        # phase B: halo exchange on the compute stream (NCCL orders itself against the current stream)
        comp.wait_event(ev_a)
        ops, p2p = [], 0
        for q, i, lo, hi in self.sends:
            ops.append(dist.P2POp(dist.isend, self.d_tiles[i][lo:hi], q))
        for p_, i, lo, hi in self.recvs:
            ops.append(dist.P2POp(dist.irecv, self.d_tiles[i][lo:hi], p_))
            p2p += self.d_tiles[i][lo:hi].numel()
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        # phase C: sub-bands
        for k, (f0, f1) in enumerate(self._subs):
            with torch.cuda.stream(up):
                for i, (a, b, c) in spans.items():
                    want = min(self._cone_hi[k].get(i, 0), b) if k + 1 < len(self._subs) else b
                    if want > done[i]:
                        lo = done[i]
                        # skip what phase A already brought (ranges at the ends of the own span)
                        pieces = [(lo, want)]
                        for (sa, sb) in sent_rows.get(i, []):
                            nxt = []
                            for (x, y) in pieces:
                                if sb <= x or sa >= y:
                                    nxt.append((x, y))
                                else:
                                    if x < sa:
                                        nxt.append((x, sa))
                                    if sb < y:
                                        nxt.append((sb, y))
                            pieces = nxt
                        for (x, y) in pieces:
                            upload(i, x, y)
                        done[i] = want
                if k + 1 == len(self._subs) and self.q1 > self.q0:  # the comparison rows trail: only the assessment needs them
                    self.d_ref.copy_(torch.from_numpy(self.reference)[self.q0:self.q1], non_blocking=True)
                    h2d += self.d_ref.numel()
                ev_u = torch.cuda.Event()
                ev_u.record(up)
            comp.wait_event(ev_u)
            self.bm.laplacian_fusion(self.plan, row_range=(f0, f1), out=self.d_canvas)
            ev_f = torch.cuda.Event()
            ev_f.record(comp)
            lo, hi = max(f0, self.r0), min(f1, self.r1)
            if hi > lo:
                down.wait_event(ev_f)
                with torch.cuda.stream(down):
                    torch.from_numpy(out)[lo:hi].copy_(self.d_canvas[lo:hi], non_blocking=True)
        if self.r1 > self.r0:
            self.qa.psnr_ssim_sums_device_rows(self.d_ref, self.d_canvas[self.q0:self.q1], self.exchange.local[0:5], image_rows=self.H,
                                               first_row=self.q0, row_range=(self.r0, self.r1))
        else:
            self.exchange.local.zero_()
        total = self.exchange.reduce(self.H * self.W * nch)  # synchronises the compute stream
        down.synchronize()
        self.h2d_bytes, self.p2p_bytes = h2d, p2p
        self.d2h_bytes = max(self.r1 - self.r0, 0) * self.W * nch
        return total


class PipelinedFuseAssess:
    """Steady-state (fuse band -> assess band) loop, two steps deep: the assessment of step k runs on its own stream while
    the fusion of step k + 1 is already building its pyramids -- the two stress different pipes (FP32 filters against
    byte / integer streaming) and, on the short bands of a many-rank job, the one fills the SMs the other's small launches
    leave idle.  Two library contexts (each with its own workspaces and captured CUDA graph) fuse into two canvases in turn;
    a canvas is overwritten only after the assessment that read it has finished (events).  Every step does the whole work
    of the serial loop -- same kernels, same bytes; only the host no longer waits for a step's sums before it enqueues the
    next (they land in `reducer` / pinned host memory; result() returns those of the last step)."""

    def __init__(self, pkg, tiles, canvas_hw, levels: int, device_index: int, reference, band, fuse_rows, reducer, channels: int = 3):
        self.dev = torch.device("cuda", device_index)
        self.H, self.W = canvas_hw
        self.band, self.fuse_rows, self.reducer, self.reference = band, fuse_rows, reducer, reference
        self.bm = [pkg.BlendingModule(num_levels=levels, device=device_index, context_slot=s) for s in (0, 1)]
        self.qa = pkg.QualityAssessmentModule(device=f"cuda:{device_index}")
        self.plan = [b.plan_fusion(tiles, output_shape=canvas_hw) for b in self.bm]
        shape = (self.H, self.W, channels) if channels > 1 else (self.H, self.W)
        self.canvas = [torch.zeros(shape, dtype=torch.uint8, device=self.dev) for _ in (0, 1)]
        self.s_fuse, self.s_iqa = torch.cuda.Stream(device=self.dev), torch.cuda.Stream(device=self.dev)
        self.fused = [torch.cuda.Event() for _ in (0, 1)]
        self.read = [torch.cuda.Event() for _ in (0, 1)]
        self.k = 0
        cur = torch.cuda.current_stream(self.dev)
        self.s_fuse.wait_stream(cur)  # inputs uploaded on the caller's stream
        self.s_iqa.wait_stream(cur)
        for e in self.read:
            e.record(self.s_iqa)

    def submit(self):
        c = self.k & 1
        self.k += 1
        with torch.cuda.stream(self.s_fuse):
            self.s_fuse.wait_event(self.read[c])  # the assessment two steps back has finished with this canvas
            self.bm[c].laplacian_fusion(self.plan[c], row_range=self.fuse_rows, out=self.canvas[c])
            self.fused[c].record(self.s_fuse)
        with torch.cuda.stream(self.s_iqa):
            self.s_iqa.wait_event(self.fused[c])
            self.qa.psnr_ssim_sums_device(self.reference, self.canvas[c], self.reducer.local, row_range=self.band)
            self.read[c].record(self.s_iqa)
            self.reducer.enqueue()

    def result(self):
        """srf_iqa_sums of the last submitted step (synchronises the assessment stream)"""
        return self.reducer.fetch(self.H * self.W * (self.canvas[0].shape[2] if self.canvas[0].dim() == 3 else 1), stream=self.s_iqa)

    def join(self):
        """make the caller's current stream wait for everything submitted"""
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_stream(self.s_fuse)
        cur.wait_stream(self.s_iqa)

    def last_canvas(self):
        return self.canvas[(self.k - 1) & 1]
