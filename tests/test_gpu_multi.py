"""Two ranks on two GPUs over NCCL (skipped on a one-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
the collective-free tile stitch (`tiling.SlotStitcher`: NVLink peer stores + signal pads) against the one-rank result, and the
batch path's `all_gather`.  `bench.py` makes the same checks before its timed regions at every N."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _toy(l, r, ml, mr):
    d = (l.mean(1, keepdim=True) - r.mean(1, keepdim=True)) * 10 + ml * 3 + 0.01 * l.shape[-1]
    return -(d + 0.1 * torch.tanh(mr)), None


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from stereoanywhere_b200 import tiling

        g = torch.Generator().manual_seed(7)
        h, w = 150, 220
        l, r = torch.rand(1, 3, h, w, generator=g).to(dev), torch.rand(1, 3, h, w, generator=g).to(dev)
        ml, mr = torch.rand(1, 1, h, w, generator=g).to(dev), torch.rand(1, 1, h, w, generator=g).to(dev)
        work = tiling.tile_multiplicity(h, w, 80, 96, 24)
        st = tiling.SlotStitcher(1, h, w, work, dev)
        outs = [tiling.tiled_inference_b200(_toy, l, r, ml, mr, 80, 96, 24, stitcher=st) for _ in range(5)]   # both parities
        if rank == 0:
            solo = tiling.tiled_inference_b200(_toy, l, r, ml, mr, 80, 96, 24,
                                               stitcher=tiling.SlotStitcher(1, h, w, work, dev, local=True))
            assert all(torch.equal(o, solo) for o in outs), "two ranks != one rank"
            torch.save(solo.cpu(), os.path.join(out_dir, "stitched.pt"))
        else:
            assert all(o is None for o in outs)
        # batch sharding: gather_batch returns every rank's slice in rank order
        local = torch.full((3, 1, 4, 5), float(rank), device=dev) + torch.arange(3, device=dev).view(3, 1, 1, 1) * 0.1
        full = tiling.gather_batch(local)
        want = torch.cat([torch.full((3, 1, 4, 5), float(k), device=dev) + torch.arange(3, device=dev).view(3, 1, 1, 1) * 0.1
                          for k in range(world)])
        assert torch.equal(full, want)
        torch.cuda.synchronize()
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_slot_stitch_and_gather_over_nccl(tmp_path, golden_tiles):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    out = torch.load(tmp_path / "stitched.pt")
    # ... and equals the reference's real TileWrapper output (same inputs as the fixture: seed 7, 150x220)
    assert float((out - torch.from_numpy(golden_tiles["st_out"])).abs().max()) < 1e-5
