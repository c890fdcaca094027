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


def test_peer_reduce_local_pointers():
    """sa_peer_reduce on one GPU (the pointers of a multi-GPU run are NVLink peer pointers of the same kind)."""
    import ctypes as C

    import torch

    from stereoanywhere_b200 import _lib

    lib = _lib.load()
    gen = torch.Generator(device="cuda:0").manual_seed(3)
    srcs = [torch.randn(4096, device="cuda:0", generator=gen) for _ in range(5)]
    den = torch.rand(4096, device="cuda:0", generator=gen) + 0.5
    dst = torch.empty(4096, device="cuda:0")
    ptrs = (C.c_void_p * 5)(*[t.data_ptr() for t in srcs])
    st = torch.cuda.current_stream().cuda_stream
    assert lib.sa_peer_reduce(ptrs, 5, den.data_ptr(), dst.data_ptr(), 4096, st) == 0
    want = srcs[0].clone()
    for t in srcs[1:]:
        want += t
    assert torch.equal(dst, want / den)
    assert lib.sa_peer_reduce(ptrs, 5, None, dst.data_ptr(), 4096, st) == 0
    assert torch.equal(dst, want)
    assert lib.sa_peer_reduce(ptrs, 17, None, dst.data_ptr(), 4096, st) == -1   # more than 16 sources
    assert lib.sa_peer_reduce(ptrs, 5, None, dst.data_ptr(), 4098, st) == -2    # n % 4
