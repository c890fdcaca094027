"""Tile geometry / stitching against the golden fixtures of the reference's TileWrapper, and the
sharded (world_size = 2, gloo, CPU) paths of config 3 (batch) and config 4 (tiles)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stereoanywhere_b200 import tiling

T = torch.from_numpy
CASES = ["mid_1984x2880", "def_1984x2880", "sf_544x960", "odd_300x500", "fit_96x96"]


def toy(l, r, ml, mr):
    d = (l.mean(1, keepdim=True) - r.mean(1, keepdim=True)) * 10 + ml * 3 + 0.01 * l.shape[-1]
    return -(d + 0.1 * torch.tanh(mr))


@pytest.mark.parametrize("tag", CASES)
def test_enumeration_matches_reference(golden_tiles, tag):
    g = golden_tiles
    h, w, th, tw, ov = [int(v) for v in g[f"args_{tag}"]]
    got = np.array(tiling.enumerate_tiles(h, w, th, tw, ov), dtype=np.int64).reshape(-1, 4)
    assert np.array_equal(got, g[f"tiles_{tag}"])
    uniq = tiling.enumerate_tiles(h, w, th, tw, ov, unique=True)
    assert len(uniq) == len({tuple(t) for t in g[f"tiles_{tag}"]})
    mult = tiling.tile_multiplicity(h, w, th, tw, ov)
    assert [t for t, _ in mult] == uniq and sum(m for _, m in mult) == len(g[f"tiles_{tag}"])


def test_presets_and_counts(golden_tiles):
    th, tw, ov = tiling.PRESETS["middlebury"]
    assert len(tiling.enumerate_tiles(1984, 2880, th, tw, ov)) == 12
    assert len(tiling.enumerate_tiles(1984, 2880, th, tw, ov, unique=True)) == 10
    th, tw, ov = tiling.PRESETS["default"]
    assert len(tiling.enumerate_tiles(1984, 2880, th, tw, ov, unique=True)) == 48


def test_blend_weight_and_pad(golden_tiles):
    g = golden_tiles
    for (h, w) in [(7, 5), (40, 24), (1, 9)]:
        assert np.array_equal(tiling.blend_weight(h, w).numpy(), g[f"blend_{h}x{w}"])
    assert tiling.pad_to_32(375, 1242) == [3, 3, 4, 5]
    assert tiling.pad_to_32(64, 96) == [0, 0, 0, 0]
    with pytest.raises(ValueError):
        tiling.blend_weight(0, 4)


def test_single_process_stitch_is_bit_exact(golden_tiles):
    g = golden_tiles
    th, tw, ov = [int(v) for v in g["st_args"]]
    out = tiling.tiled_inference(toy, T(g["st_l"]), T(g["st_r"]), T(g["st_ml"]), T(g["st_mr"]), th, tw, ov)
    assert np.array_equal(out.numpy(), g["st_out"])
    # distinct tiles weighted by multiplicity: same image up to fp32 summation order
    out_u = tiling.tiled_inference(toy, T(g["st_l"]), T(g["st_r"]), T(g["st_ml"]), T(g["st_mr"]), th, tw, ov, unique=True)
    assert np.abs(out_u.numpy() - g["st_out"]).max() < 1e-5
    # image that fits one tile takes the single-shot path
    small = tiling.tiled_inference(toy, T(g["st_l"])[..., :64, :64], T(g["st_r"])[..., :64, :64],
                                   T(g["st_ml"])[..., :64, :64], T(g["st_mr"])[..., :64, :64], 80, 96, 24)
    assert torch.equal(small, -toy(T(g["st_l"])[..., :64, :64], T(g["st_r"])[..., :64, :64],
                                   T(g["st_ml"])[..., :64, :64], T(g["st_mr"])[..., :64, :64]))
    with pytest.raises(ValueError):
        tiling.tiled_inference(toy, torch.zeros(2, 3, 200, 200), torch.zeros(2, 3, 200, 200), None, None, 80, 96, 24)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, golden_file, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = dict(np.load(golden_file))
        th, tw, ov = [int(v) for v in g["st_args"]]
        out = tiling.tiled_inference(toy, T(g["st_l"]), T(g["st_r"]), T(g["st_ml"]), T(g["st_mr"]), th, tw, ov)
        assert (out is None) == (rank != 0)
        if rank == 0:
            np.save(os.path.join(out_dir, "stitched.npy"), out.numpy())
        # batch sharding (config 3): every rank owns a contiguous slice, all_gather restores the batch
        full = torch.arange(6 * 1 * 4 * 5, dtype=torch.float32).view(6, 1, 4, 5)
        local = tiling.shard_batch(full, rank, world) * 2.0
        back = tiling.gather_batch(local)
        assert torch.equal(back, full * 2.0)
    finally:
        dist.destroy_process_group()


def test_two_rank_tile_and_batch_sharding(tmp_path, golden_tiles):
    golden_file = os.path.join(os.path.dirname(__file__), "golden", "tiles.npz")
    mp.spawn(_worker, args=(2, _free_port(), golden_file, str(tmp_path)), nprocs=2, join=True)
    out = np.load(tmp_path / "stitched.npy")
    assert np.abs(out - golden_tiles["st_out"]).max() < 1e-5


# ------------------------------------------------------------------------------------------ slot stitch bookkeeping

def test_slot_units_are_sharded_round_robin_and_cover_every_tile():
    """The unit list / shard of `SlotStitcher` (host logic; its kernels are exercised by the `gpu` tests)."""
    h, w = 1984, 2880
    th, tw, ov = tiling.PRESETS["middlebury"]
    work = tiling.tile_multiplicity(h, w, th, tw, ov)
    assert len(work) == 10 and sum(m for _, m in work) == 12      # 12 emitted, 10 distinct
    units = [(img, t, m) for img in range(4) for (t, m) in work]
    for world in (1, 2, 4, 8):
        shares = [[i for i in range(len(units)) if i % world == r] for r in range(world)]
        assert sorted(sum(shares, [])) == list(range(40))
        assert max(map(len, shares)) - min(map(len, shares)) <= 1
    # slot stitch precondition: tile columns start / end on multiples of 4 for every reference preset at config 4
    for name, (ph, pw, po) in tiling.PRESETS.items():
        for (y0, y1, x0, x1), _ in tiling.tile_multiplicity(h, w, ph, pw, po):
            assert x0 % 4 == 0 and (x1 - x0) % 4 == 0, (name, x0, x1)
