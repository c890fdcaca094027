"""Tile-sharded full-resolution inference (BASELINE config 4) and batch sharding (config 3).

Host-side plumbing only (torch + torch.distributed).  The geometry reproduces the reference's
MapReduce wrapper so that a tile-sharded run returns the reference's tiled result:

* tile enumeration  - mapreduce_v2/tile_wrapper.py:101-120 (`TileWrapper._enumerate_tiles`):
  stride = tile - overlap, border tiles shifted back inside the image; the reference emits the
  clamped last row / column more than once and accumulates the duplicates (harmless after
  normalisation, but they do change the blend where tiles overlap).  `unique=False` keeps that
  behaviour bit for bit; `unique=True` runs every distinct tile ONCE and weights it by its
  multiplicity - the same stitched image up to fp32 summation order, with 10 instead of 12 model
  runs for the `middlebury` preset on 1984x2880.  The multi-GPU path shards the distinct tiles.
* per-tile replicate pad to a multiple of 32 and un-pad - tile_wrapper.py:226-247.
* blend weight clamp(sin(pi y) sin(pi x), 1e-4) and weighted accumulation - tile_wrapper.py:36-49,
  328-362; final `stitched / clamp(weight, 1e-4)` - tile_wrapper.py:185.

Multi-GPU: tiles are independent units.  Every rank runs its share of tiles and accumulates
`disp * w` locally; ONE collective - `reduce(sum)` of that `[H,W]` fp32 tensor to rank 0 - stitches the
image.  The weight accumulator `sum w` of the reference depends on the tile geometry
only, so rank 0 forms it locally (`weight_sum`, same accumulation order as the reference) instead of
shipping a second `[H,W]` plane through the collective.  Batch sharding needs one `all_gather` of the
disparities.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tile = Tuple[int, int, int, int]  # y0, y1, x0, x1

#: the reference's presets that matter for the BASELINE configs (mapreduce_v2/tile_presets.py:37-127):
#: name -> (tile_height, tile_width, overlap)
PRESETS = {
    "default": (448, 448, 96),
    "middlebury": (1120, 672, 112),
    "kitti": (448, 1344, 128),
    "sceneflow": (448, 448, 112),
    "booster": (896, 1120, 224),
}


def _axis_starts(extent: int, tile: int, overlap: int) -> List[Tuple[int, int]]:
    step = tile - overlap
    if step <= 0:
        raise ValueError("overlap must be smaller than the tile")
    spans = []
    pos = 0
    while pos < extent:
        hi = min(pos + tile, extent)
        spans.append((max(0, hi - tile), hi))
        pos += step
    return spans


def enumerate_tiles(height: int, width: int, tile_h: int, tile_w: int, overlap: int, unique: bool = False) -> List[Tile]:
    """Row-major list of (y0, y1, x0, x1); `unique=True` drops repeated tiles (first occurrence kept)."""
    tiles = [(y0, y1, x0, x1) for (y0, y1) in _axis_starts(height, tile_h, overlap)
             for (x0, x1) in _axis_starts(width, tile_w, overlap)]
    return list(dict.fromkeys(tiles)) if unique else tiles


def tile_multiplicity(height: int, width: int, tile_h: int, tile_w: int, overlap: int):
    """Distinct tiles in first-occurrence order with the number of times the reference emits them."""
    counts = {}
    for t in enumerate_tiles(height, width, tile_h, tile_w, overlap):
        counts[t] = counts.get(t, 0) + 1
    return list(counts.items())


def blend_weight(h: int, w: int, device=None) -> torch.Tensor:
    """[h, w] cosine window of the reference (tile_wrapper.py:36-49)."""
    if h <= 0 or w <= 0:
        raise ValueError("Tile dimensions must be positive")
    y = torch.linspace(0, 1, h, device=device)
    x = torch.linspace(0, 1, w, device=device)
    wy = torch.sin(torch.pi * y).unsqueeze(1)
    wx = torch.sin(torch.pi * x).unsqueeze(0)
    return torch.clamp(wy * wx, min=1e-4)


def weight_sum(height: int, width: int, work, device=None) -> torch.Tensor:
    """`[H, W]` sum of the blend windows of all tiles in `work` = [(tile, multiplicity)], accumulated in
    `work` order exactly like the reference's `weight_sum += weight` (tile_wrapper.py:358-362)."""
    den = torch.zeros(height, width, dtype=torch.float32, device=device)
    cache = {}
    for (y0, y1, x0, x1), mult in work:
        key = (y1 - y0, x1 - x0)
        if key not in cache:
            cache[key] = blend_weight(key[0], key[1], device=device)
        for _ in range(mult):
            den[y0:y1, x0:x1] += cache[key]
    return den


def pad_to_32(h: int, w: int) -> List[int]:
    """[left, right, top, bottom] of the replicate pad (tile_wrapper.py:226-229, test.py:204-207)."""
    ph = (32 - h % 32) % 32
    pw = (32 - w % 32) % 32
    return [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2]


def shard(items: Sequence, rank: int, world: int) -> List:
    """Round-robin share of `items` for `rank` (tiles differ in cost only at the borders)."""
    return [it for i, it in enumerate(items) if i % world == rank]


def run_tile(model: Callable[..., torch.Tensor], tensors: Sequence[Optional[torch.Tensor]], tile: Tile) -> torch.Tensor:
    """Crop, replicate-pad to /32, run `model`, negate, un-pad (tile_wrapper.py:208-247)."""
    y0, y1, x0, x1 = tile
    pad = pad_to_32(y1 - y0, x1 - x0)
    args = [None if t is None else F.pad(t[:, :, y0:y1, x0:x1], pad, mode="replicate") for t in tensors]
    out = model(*args)
    if isinstance(out, (tuple, list)):
        out = out[0]
    disp = -out
    hd, wd = disp.shape[-2:]
    return disp[..., pad[2]: hd - pad[3], pad[0]: wd - pad[1]]


def tiled_inference(
    model: Callable[..., torch.Tensor],
    left: torch.Tensor,
    right: torch.Tensor,
    mono_left: Optional[torch.Tensor],
    mono_right: Optional[torch.Tensor],
    tile_h: int,
    tile_w: int,
    overlap: int,
    *,
    group=None,
    unique: Optional[bool] = None,
    dst: int = 0,
) -> Optional[torch.Tensor]:
    """Tiled inference, tiles sharded over the ranks of `group` (single process when torch.distributed
    is not initialised).  Returns the stitched `[1,1,H,W]` disparity on rank `dst`, None elsewhere.

    `unique=None` picks the reference's duplicate-keeping enumeration for a single process (bit-exact
    parity with `TileWrapper.forward`) and the de-duplicated one when sharding."""
    import torch.distributed as dist

    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    b, _, height, width = left.shape
    if b != 1:
        raise ValueError("tiled inference supports batch size == 1 (tile_wrapper.py:148-149)")
    if height <= tile_h and width <= tile_w:  # single-shot path of the reference (tile_wrapper.py:151-153)
        if rank != dst:
            return None
        out = model(left, right, mono_left, mono_right)
        out = out[0] if isinstance(out, (tuple, list)) else out
        return -out
    if unique is None:
        unique = world > 1
    if unique:
        work = tile_multiplicity(height, width, tile_h, tile_w, overlap)
    else:
        work = [(t, 1) for t in enumerate_tiles(height, width, tile_h, tile_w, overlap)]
    num = torch.zeros(height, width, dtype=torch.float32, device=left.device)  # sum of disp * w over my tiles
    for (y0, y1, x0, x1), mult in shard(work, rank, world):
        disp = run_tile(model, (left, right, mono_left, mono_right), (y0, y1, x0, x1)).to(num.device)
        wgt = blend_weight(y1 - y0, x1 - x0, device=num.device)
        for _ in range(mult):  # the reference accumulates a repeated tile once per repeat
            num[y0:y1, x0:x1] += disp[0, 0].float() * wgt
    if world > 1:
        dist.reduce(num, dst=dst, op=dist.ReduceOp.SUM, group=group)  # the path's only collective
        if rank != dst:
            return None
    den = weight_sum(height, width, work, device=num.device)  # geometry only: no need to communicate it
    out = torch.where(den > 0, num / torch.clamp(den, min=1e-4), num)
    return out.view(1, 1, height, width)


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Contiguous slice of the batch for `rank` (config 3: 64 pairs -> 8 per GPU)."""
    b = t.shape[0]
    per = math.ceil(b / world)
    return t[rank * per: min(b, (rank + 1) * per)]


def gather_batch(local: torch.Tensor, group=None) -> torch.Tensor:
    """all_gather of per-rank `[B/N, ...]` results back into `[B, ...]` (equal shares)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous(), group=group)
    return torch.cat(parts, dim=0)


def plan_images(images: int, world: int):
    """How `images` full-resolution images are spread over `world` ranks so that a rank only ever holds - and
    communicates - the planes of its own images.  Returns `ranks_of[j]` (the ranks that share image j; the first
    one is its leader), or None when neither number divides the other (then every rank takes tiles of every image
    and the stitch is one global reduce)."""
    if world >= images and world % images == 0:
        per = world // images
        return [list(range(j * per, (j + 1) * per)) for j in range(images)]
    if images % world == 0:
        per = images // world
        return [[j // per] for j in range(images)]
    return None


class ImageStitcher:
    """Stitch of a tile-sharded step over several images with the least data on the wire (config 4, N ranks).

    A global `reduce(sum)` of the `[images, H, W]` accumulator makes every rank push the whole 91 MB (four
    Middlebury images) through the ring although it only touched its own tiles.  Here the tiles are sharded BY IMAGE
    (`plan_images`): the ranks that share an image reduce just that `[H, W]` plane among themselves (a sub-group
    collective; nothing at all when an image belongs to one rank), its leader divides by the weight plane (geometry
    only, formed locally) and sends the finished plane to the gathering rank.  Both phases are asynchronous and
    double-buffered: the sub-group reduce of step k is launched after step k's tiles, its normalise + send when step
    k+1's tiles have been issued, and everything of parity i is waited for when that buffer comes round again.
    """

    def __init__(self, images: int, height: int, width: int, work, device, group=None, dst: int = 0):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.dst = dst
        self.ranks_of = plan_images(images, self.world)
        if self.ranks_of is None:
            raise ValueError(f"{images} images cannot be spread over {self.world} ranks image by image")
        self.images = images
        self.my_images = [j for j, rk in enumerate(self.ranks_of) if self.rank in rk]
        self.slot = {j: k for k, j in enumerate(self.my_images)}      # image -> index in my accumulator
        self.shared = len(self.ranks_of[0]) > 1
        # every rank creates every sub-group, in the same order (torch.distributed requirement)
        self.groups = [dist.new_group(rk) if self.shared else None for rk in self.ranks_of]
        self.lead = [j for j in self.my_images if self.ranks_of[j][0] == self.rank]   # images I finish and send
        n = max(1, len(self.my_images))
        self.acc = [torch.zeros(n, height, width, dtype=torch.float32, device=device) for _ in range(2)]
        self.den = torch.clamp(weight_sum(height, width, work, device=device), min=1e-4) if self.lead else None
        self.norm = [torch.empty(len(self.lead), height, width, dtype=torch.float32, device=device) for _ in range(2)] \
            if self.lead else None
        self.out = torch.zeros(images, height, width, dtype=torch.float32, device=device) if self.rank == dst else None
        self._reduce = [[], []]   # outstanding sub-group reduces of parity i
        self._p2p = [[], []]      # outstanding sends / receives of parity i
        self._stage = [0, 0]      # 0 idle, 1 reduce launched, 2 normalise + send launched

    def units(self, work) -> List:
        """This rank's (image, tile, multiplicity) units: the tiles of its images, round-robin inside an image."""
        mine = []
        for j in self.my_images:
            rk = self.ranks_of[j]
            mine += [(j, t, m) for i, (t, m) in enumerate(work) if i % len(rk) == rk.index(self.rank)]
        return mine

    def buffer(self, i: int) -> torch.Tensor:
        """Accumulator `[my images, H, W]` of parity i (index with `slot[image]`), free to be overwritten."""
        self._finish(i)
        return self.acc[i]

    def launch(self, i: int) -> None:
        """Step k's tiles have been accumulated into buffer i: start its reduce, and move step k-1 one phase on."""
        if self.shared:
            for j in self.my_images:
                k = self.slot[j]
                self._reduce[i].append(self.dist.reduce(self.acc[i][k], dst=self.ranks_of[j][0], op=self.dist.ReduceOp.SUM,
                                                        group=self.groups[j], async_op=True))
        self._stage[i] = 1
        if self._stage[i ^ 1] == 1:
            self._send(i ^ 1)

    def _send(self, i: int) -> None:
        for w in self._reduce[i]:
            w.wait()
        self._reduce[i] = []
        for n, j in enumerate(self.lead):
            torch.div(self.acc[i][self.slot[j]], self.den, out=self.norm[i][n])
            if self.rank == self.dst:
                self.out[j].copy_(self.norm[i][n])
            else:
                self._p2p[i].append(self.dist.isend(self.norm[i][n], dst=self.dst, group=self.group))
        if self.rank == self.dst:
            for j, rk in enumerate(self.ranks_of):
                if rk[0] != self.dst:
                    self._p2p[i].append(self.dist.irecv(self.out[j], src=rk[0], group=self.group))
        self._stage[i] = 2

    def _finish(self, i: int) -> None:
        if self._stage[i] == 1:
            self._send(i)
        for w in self._p2p[i]:
            w.wait()
        self._p2p[i] = []
        self._stage[i] = 0

    def drain(self) -> Optional[torch.Tensor]:
        """Complete everything outstanding (oldest step first); the stitched `[images, H, W]` on the gathering rank."""
        order = (0, 1) if self._stage[0] >= self._stage[1] else (1, 0)
        for i in order:
            self._finish(i)
        return self.out


class PeerStitcher:
    """Stitch of a tile-sharded step over NVLink peer memory instead of an NCCL reduce (config 4).

    Every rank accumulates `sum disp * w` for its tiles into a SYMMETRIC buffer (torch symmetric memory: the same
    allocation is mapped into every process of the node).  `reduce_to(dst_rank)` then lets each rank sum its own
    1/N slice of all N accumulators through peer loads, divide by the weight plane (geometry only, formed locally)
    and store the slice into `dst_rank`'s output buffer: reduce + normalise + gather in ONE kernel per rank
    (`sa_peer_reduce`), bracketed by two device-side barriers.  It runs on a side stream, double-buffered, so the
    next step's tiles compute meanwhile - and unlike an SM-resident NCCL reduce it does not sit on SMs that the
    persistent GEMM kernels of the next tile expect to own.
    """

    def __init__(self, images: int, height: int, width: int, den: torch.Tensor, device, group=None):
        import ctypes as C

        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        from . import _lib

        self._C, self._lib = C, _lib.load()
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.n = images * height * width
        assert self.n % 4 == 0
        self.acc = [symm.empty(images, height, width, dtype=torch.float32, device=device) for _ in range(2)]
        self.acc_h = [symm.rendezvous(t, self.group) for t in self.acc]
        self.out = symm.empty(images, height, width, dtype=torch.float32, device=device)
        self.out_h = symm.rendezvous(self.out, self.group)
        self.den = den.to(device).float().clamp(min=1e-4).unsqueeze(0).expand(images, height, width).contiguous()
        per = (self.n // 4 + self.world - 1) // self.world * 4
        self.lo = min(self.n, self.rank * per)
        self.hi = min(self.n, self.lo + per)
        self.side = torch.cuda.Stream(device=device)
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        for e in self.free:
            e.record(torch.cuda.current_stream(device))

    def buffer(self, i: int) -> torch.Tensor:
        """Accumulator of parity i; waits (on the current stream) until every rank has finished reading it."""
        torch.cuda.current_stream().wait_event(self.free[i])
        return self.acc[i]

    def reduce_to(self, i: int, dst_rank: int = 0) -> None:
        main = torch.cuda.current_stream()
        self.ready[i].record(main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ready[i])
            self.acc_h[i].barrier(channel=0)          # every rank's accumulator i is complete
            if self.hi > self.lo:
                C = self._C
                ptrs = (C.c_void_p * self.world)(*[int(p) + 4 * self.lo for p in self.acc_h[i].buffer_ptrs])
                dst = int(self.out_h.buffer_ptrs[dst_rank]) + 4 * self.lo
                rc = self._lib.sa_peer_reduce(ptrs, self.world, self.den.data_ptr() + 4 * self.lo, dst, self.hi - self.lo,
                                              self.side.cuda_stream)
                if rc != 0:
                    raise RuntimeError("sa_peer_reduce: " + self._lib.sa_last_error().decode())
            self.out_h.barrier(channel=1)             # all slices stored, all peer reads of accumulator i done
            self.free[i].record(self.side)

    def result(self) -> torch.Tensor:
        """The stitched images on the gathering rank (valid after the side stream has been waited for)."""
        return self.out

    def drain(self) -> None:
        torch.cuda.current_stream().wait_stream(self.side)
