#!/usr/bin/env python
"""BASELINE config 5: cost-volume / lookup microbench sweep - C in {128, 256}, W/4 up to 768, 4 levels,
radius 4.  B is chosen so that the level-0 volume is ~256 MB.  Per point: fused correlation +
truncation + pyramid (sa_corr_pack_tf32), mono pack, dual lookup (graph of 8 launches), each as time,
algorithmic GB/s and fraction of the measured HBM peak.  Prints one JSON line per point.
Under torchrun every rank runs the same sweep on its own GPU (replicas); rank 0 prints the max time."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import stereoanywhere_b200 as sa

B_ = sa.CorrBlockB200
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
peak, _ = bench.load_peaks()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
H = 96
points = [(c, w) for c in (128, 256) for w in (128, 192, 256, 312, 384, 512, 640, 768)]
if len(sys.argv) > 1:
    points = points[:: int(sys.argv[1])]


def timed(fn, reps=5):
    ts = []
    for i in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    t = ts[len(ts) // 2]
    if dist is not None:
        tt = torch.tensor([t], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt.item())
    return t, out


for c, w in points:
    b = max(1, round(256e6 / (H * w * w * 4)))
    _, d = bench.make_inputs(b, c, H, w, dev, seed=rank)
    p = b * H * w
    t_corr, fs = timed(lambda: B_.from_features(d["fl"], d["fr"], truncate=(d["tdisp"], d["tconf"], 0.9)))
    t_mono, fm = timed(lambda: B_.from_normals(d["nl"], d["nr"]))
    coords = [d["coords0"] + k * d["delta"] for k in range(8)]
    g = torch.cuda.CUDAGraph()
    B_.lookup_pair(fs, fm, coords[0]); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for k in range(8):
            o = B_.lookup_pair(fs, fm, coords[k])
    t_lk, _ = timed(lambda: g.replay())
    t_lk /= 8
    packed_bytes = p * (w // 8 + 9) * 128
    corr_bytes = 2 * b * c * H * w * 4 + packed_bytes            # bytes the fused kernel really moves
    corr_alg = 2 * b * c * H * w * 4 + 1.875 * p * w * 4         # volume + pooled levels written once
    factored = B_.mono_mode == "factored"
    mono_bytes = (3 * b * H * (w // 8 + 9) * 128 + 3 * b * H * w * 4) if factored else packed_bytes
    lk_alg, lk_real = (464, 432) if factored else (612, 544)    # bytes per pixel of a dual lookup
    line = {"C": c, "W4": w, "B": b, "n_gpus": world, "mono": B_.mono_mode,
            "corr_pack_us": round(t_corr, 1), "corr_pack_gbs": round(corr_bytes / t_corr / 1e3, 0), "corr_pack_frac": round(corr_bytes / t_corr / 1e3 / peak, 3),
            "corr_pack_alg_frac": round(corr_alg / t_corr / 1e3 / peak, 3), "corr_tflops": round(2 * p * w * c / t_corr / 1e6, 1),
            "mono_pack_us": round(t_mono, 1), "mono_pack_frac": round(mono_bytes / t_mono / 1e3 / peak, 3),
            "lookup2_us": round(t_lk, 2), "lookup2_alg_frac": round(lk_alg * p / t_lk / 1e3 / peak, 3),
            "lookup2_real_frac": round(lk_real * p / t_lk / 1e3 / peak, 3)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    del fs, fm, d, g, o
    torch.cuda.empty_cache()
if dist is not None:
    dist.barrier(); dist.destroy_process_group()
