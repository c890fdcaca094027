"""The C ABI called directly through ctypes (the INTEGRATION.md recipe): no torch.ops, no CorrBlockB200."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import corr_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def lib():
    import stereoanywhere_b200._lib as L

    return L.load()


def test_pack_and_lookup_through_the_c_abi(lib):
    b, h, w = 2, 6, 40
    gen = torch.Generator().manual_seed(5)
    vol = torch.randn(b, h, w, 1, w, generator=gen)
    x = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w).expand(b, 1, h, w)
    coords = torch.cat([x - torch.rand(b, 1, h, w, generator=gen) * 12, torch.zeros(b, 1, h, w)], 1).contiguous()
    want = O.OracleCorrBlock(vol, num_levels=4, radius=4)(coords)

    dvol, dcoords = vol.to(DEV), coords.to(DEV)
    rowf = lib.sa_packed_row_floats(w)
    assert rowf == (w // 8 + 9) * 32
    packed = torch.empty(b * h * w, rowf, device=DEV)
    out = torch.empty(b, 36, h, w, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    assert lib.sa_pack_pyramid(dvol.data_ptr(), b * h * w, w, None, None, 0.0, 0, packed.data_ptr(), st) == 0
    assert lib.sa_lookup_packed(packed.data_ptr(), None, w, dcoords.data_ptr(), dcoords.stride(0), out.data_ptr(), None,
                                b, h, w, st) == 0
    torch.cuda.synchronize()
    assert float((out.cpu() - want).abs().max()) < 3e-5

    # same through the general entry points: sa_pyramid + sa_lookup
    pit = [w, 20, 12, 8]
    wid = [w, 20, 10, 5]
    lv = [dvol.view(-1, w)] + [torch.empty(b * h * w, p, device=DEV) for p in pit[1:]]
    assert lib.sa_pyramid(lv[0].data_ptr(), b * h * w, w, w, 3, lv[1].data_ptr(), lv[2].data_ptr(), lv[3].data_ptr(),
                          pit[1], pit[2], pit[3], None, None, 0.0, 0, None, st) == 0
    ptrs = (C.c_void_p * 4)(*[t.data_ptr() for t in lv])
    out2 = torch.empty_like(out)
    assert lib.sa_lookup(ptrs, (C.c_int * 4)(*wid), (C.c_int64 * 4)(*pit), 4, 4, dcoords.data_ptr(), dcoords.stride(0),
                         out2.data_ptr(), b, h, w, 0, 0, st) == 0
    torch.cuda.synchronize()
    assert torch.equal(out, out2)


def test_error_codes(lib):
    st = torch.cuda.current_stream().cuda_stream
    t = torch.zeros(64, 64, device=DEV)
    # W3 not a multiple of 8 -> unsupported by the packed family
    assert lib.sa_pack_pyramid(t.data_ptr(), 4, 12, None, None, 0.0, 0, t.data_ptr(), st) == -3
    assert b"multiple of 8" in lib.sa_last_error()
    # misaligned pointer
    assert lib.sa_pack_pyramid(t.data_ptr() + 4, 4, 16, None, None, 0.0, 0, t.data_ptr(), st) == -2
    # tf32 kernel: C must be a multiple of 32
    assert lib.sa_corr_tf32(t.data_ptr(), t.data_ptr(), t.data_ptr(), 1, 24, 1, 8, 8, 1.0, 1.0, None, None, 0.0, None, None,
                            None, 0, 0, 0, st) == -3
    # lookup: more levels than the ABI allows
    assert lib.sa_lookup(None, None, None, 9, 4, None, 0, None, 1, 1, 8, 0, 0, st) == -1
    torch.cuda.synchronize()  # none of the rejected calls launched anything / poisoned the context


def test_streams_are_respected(lib):
    """Launches go to the stream that is passed in: work queued on a side stream is ordered after
    the producer kernel on that stream."""
    import stereoanywhere_b200 as sa

    side = torch.cuda.Stream()
    gen = torch.Generator(device=DEV).manual_seed(9)
    vol = torch.randn(1, 8, 64, 1, 64, device=DEV, generator=gen)
    coords = torch.zeros(1, 2, 8, 64, device=DEV)
    coords[:, 0] = torch.arange(64, device=DEV).float()
    ref = sa.CorrBlockB200(vol)(coords)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        v2 = vol * 2.0
        out = sa.CorrBlockB200(v2)(coords)
    side.synchronize()
    assert torch.allclose(out, 2.0 * ref, atol=1e-6)


def test_stitch_kernels_match_the_reference_tile_wrapper(golden_tiles):
    """sa_stitch_tile + sa_stitch_finish (through `tiling.tiled_inference_b200`, one GPU: local slot pointers - the
    pointers of a multi-GPU run are NVLink peer pointers of the same kind) against the stitched output of the
    reference's real `TileWrapper` around a toy model (tests/golden/tiles.npz), and against the host
    implementation `tiling.tiled_inference`."""
    import torch

    from stereoanywhere_b200 import tiling

    g = golden_tiles
    dev = "cuda:0"
    l, r, ml, mr = (torch.from_numpy(g[k]).to(dev) for k in ("st_l", "st_r", "st_ml", "st_mr"))
    th, tw, ov = (int(v) for v in g["st_args"])

    def toy(l_, r_, ml_, mr_):
        d = (l_.mean(1, keepdim=True) - r_.mean(1, keepdim=True)) * 10 + ml_ * 3 + 0.01 * l_.shape[-1]
        return -(d + 0.1 * torch.tanh(mr_)), None

    # 220 columns, tiles of 96 with stride 72: x0 = 0, 72, 124 - multiples of 4, as the slot stitch requires
    out = tiling.tiled_inference_b200(toy, l, r, ml, mr, th, tw, ov)
    host = tiling.tiled_inference(toy, l, r, ml, mr, th, tw, ov, unique=True)
    assert out.shape == (1, 1, 150, 220)
    assert float((out.cpu() - torch.from_numpy(g["st_out"])).abs().max()) < 1e-5
    assert float((out - host).abs().max()) < 1e-5
    # a reused stitcher (double-buffered slots, several steps) returns the same image every step
    work = tiling.tile_multiplicity(150, 220, th, tw, ov)
    st = tiling.SlotStitcher(1, 150, 220, work, dev)
    for _ in range(5):
        again = tiling.tiled_inference_b200(toy, l, r, ml, mr, th, tw, ov, stitcher=st)
        assert torch.equal(again, out)
    # a geometry the float4 stitch kernels do not cover (218 columns: tiles start at 0, 72, 122) takes the portable
    # stitch instead of raising
    odd = tiling.tiled_inference_b200(toy, l[..., :218], r[..., :218], ml[..., :218], mr[..., :218], th, tw, ov)
    host_odd = tiling.tiled_inference(toy, l[..., :218], r[..., :218], ml[..., :218], mr[..., :218], th, tw, ov, unique=True)
    assert odd.shape == (1, 1, 150, 218) and torch.equal(odd, host_odd)


def test_stitch_error_codes(lib):
    import torch

    t = torch.zeros(64, 64, device="cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    assert lib.sa_stitch_tile(t.data_ptr(), 64, 64, 1, -1.0, 0, 0, 32, 30, t.data_ptr(), 1.0, t.data_ptr(), st) == -3  # tw % 4
    assert lib.sa_stitch_tile(t.data_ptr(), 64, 64, 1, -1.0, 40, 0, 32, 32, t.data_ptr(), 1.0, t.data_ptr(), st) == -1  # does not fit
    assert lib.sa_stitch_tile(None, 64, 64, 1, -1.0, 0, 0, 32, 32, t.data_ptr(), 1.0, t.data_ptr(), st) == -1
