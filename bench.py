#!/usr/bin/env python
"""bench.py - the cost-volume hot path on B200, measured the way BASELINE.json asks.

One "step" = one pass of the whole path over one batch of synthetic stereo pairs:
    stereo correlation (TF32 tcgen05) + mono correlation (x1.73) + truncation product + two
    avg-pooled pyramids + 32 GRU iterations x (stereo lookup + mono lookup).
Default workload = BASELINE.json configs[1]: KITTI-size 375x1242 pairs (padded to 384x1248, quarter
resolution 96x312), batch 8 per GPU, C=256, 4 levels, radius 4, 32 iterations.

    python bench.py                      # N=1, K=10, W=3
    python bench.py --impl reference     # the reference's own CorrBlock1D (oracle/_ref) on the host cores
    torchrun --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs already in HBM);
`e2e` goes through the public CorrBlockB200 API from pinned HOST buffers with the H2D / D2H copies
inside the timed region.  `roofline` is for the dominant kernel (the lookup), timed live with CUDA
events inside the timed region; `cpu_baseline` is the reference's CorrBlock1D timed on this box's host
cores on a bounded sample.  The default line also carries, measured in the same run:
  `variants`  - the README wiring (`--mono aggregated`) and the strict drop-in call sequence (`--variant protocol`);
  `tiled_c4`  - BASELINE config 4 (full-resolution Middlebury, reference tile geometry, tiles sharded over the N
                ranks, collective-free stitch over NVLink peer memory), strong scaling, with the N-rank result
                checked against the 1-rank result before the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B per GPU, C, H/4, W/4)
    "c1_384x512_b1": (1, 256, 96, 128),
    "c2_kitti_375x1242_b8": (8, 256, 96, 312),
    "c3_sceneflow_540x960_b8": (8, 256, 136, 240),
    "c4_middlebury_tile_1120x672_b1": (1, 256, 280, 168),
    "c5_sweep_c256_w768_b1": (1, 256, 96, 768),
}
DEFAULT_WORKLOAD = "c2_kitti_375x1242_b8"
TILED_WORKLOAD = "c4_middlebury_1984x2872_tiled"
ITERS, LEVELS, RADIUS = 32, 4, 4
METRIC = "stereo pairs/sec @375x1242, 32 iters (cost-volume path: corr + pyramid + lookup)"
METRIC_SIZES = {"c1_384x512_b1": "384x512", "c2_kitti_375x1242_b8": "375x1242", "c3_sceneflow_540x960_b8": "540x960",
                "c4_middlebury_tile_1120x672_b1": "1120x672 (one Middlebury tile)", "c5_sweep_c256_w768_b1": "384x3072 (W/4 = 768)"}
UNIT = "pairs/s"


def metric_for(workload):
    return METRIC.replace("375x1242", METRIC_SIZES.get(workload, "375x1242"))


def workload_config(workload, n_gpus):
    """`config` of the JSON line: the workload only, identical in both arms (`--impl b200` / `--impl reference`)."""
    b, c, h, w = WORKLOADS[workload]
    return {"workload": workload, "pairs_per_gpu": b, "C": c, "H4": h, "W4": w, "iters": ITERS, "levels": LEVELS,
            "radius": RADIUS, "l2": "inputs+volumes (>1 GB/step) exceed the 126 MB L2; no explicit flush",
            "parallelism": f"batch-sharded x{n_gpus}, async all_gather of quarter-res disparity per step"}


def tensor_peak_tf32():
    """Dense TF32 tensor peak in TFLOP/s: half the measured bf16 cuBLAS burst figure (MEASURED_PEAKS.json), else
    half the profiling guide's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return round(float(json.load(open(p))["bf16_tflops"]) / 2.0, 1)
    return 795.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------

def make_inputs(b, c, h, w, device, seed=0, pinned=False):
    g = torch.Generator().manual_seed(seed)
    fl = torch.randn(b, c, h, w, generator=g)
    fr = torch.randn(b, c, h, w, generator=g)
    nl = torch.nn.functional.normalize(torch.randn(b, 3, h, w, generator=g), dim=1)
    nr = torch.nn.functional.normalize(torch.randn(b, 3, h, w, generator=g), dim=1)
    x = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w).expand(b, 1, h, w)
    y = torch.arange(h, dtype=torch.float32).view(1, 1, h, 1).expand(b, 1, h, w)
    coords0 = torch.cat([x - torch.rand(b, 1, h, w, generator=g) * (w / 4), y], 1).contiguous()
    # per-iteration update the GRU would produce: a small sub-pixel drift, zero in y
    delta = torch.cat([(torch.rand(b, 1, h, w, generator=g) - 0.5) * 0.5, torch.zeros(b, 1, h, w)], 1).contiguous()
    tdisp = (torch.rand(b, 1, h, w, generator=g) * (w / 4)).contiguous()
    tconf = torch.rand(b, 1, h, w, generator=g).contiguous()
    host = dict(fl=fl, fr=fr, nl=nl, nr=nr, coords0=coords0, delta=delta, tdisp=tdisp, tconf=tconf)
    if pinned:
        host = {k: v.pin_memory() for k, v in host.items()}
    if device is None:
        return host, None
    dev = {k: v.to(device, non_blocking=True) for k, v in host.items()}
    return host, dev


def path_bytes(b, c, h, w):
    """Algorithmic bytes of one step (SURVEY.md 8d / DESIGN.md)."""
    p = b * h * w
    stereo = 2 * b * c * h * w * 4 + p * w * 4 + 0.875 * p * w * 4
    mono = 2 * b * 3 * h * w * 4 + p * w * 4 + 0.875 * p * w * 4
    lookups = ITERS * 612 * p
    return stereo + mono + lookups


# ------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, 100 ms) - evidence that the timed region was not throttled
# ------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# host placement: run on the CPUs next to this rank's GPU before any pinned buffer is allocated
# ------------------------------------------------------------------------------------------

def bind_to_gpu_numa(local):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (first-touch then places the pinned staging
    buffers there).  Returns what was found; a single-node host (or a VM that hides the topology) is reported as is."""
    info = {"numa_node": None, "cpus": None, "bound": False}
    try:
        import pynvml

        pynvml.nvmlInit()
        hdl = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(hdl).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node_file = f"/sys/bus/pci/devices/{bus}/numa_node"
        node = int(open(node_file).read().strip()) if os.path.exists(node_file) else -1
        info["numa_node"] = node
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()] \
            if os.path.isdir("/sys/devices/system/node") else []
        info["host_numa_nodes"] = len(nodes)
        if node >= 0 and len(nodes) > 1:
            cpulist = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
            cpus = set()
            for part in cpulist.split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["cpus"], info["bound"] = cpulist, True
    except Exception as e:  # pragma: no cover - topology files differ between boxes
        info["error"] = f"{type(e).__name__}: {e}"
    return info


# ------------------------------------------------------------------------------------------
# the path on the GPU, through the public API
# ------------------------------------------------------------------------------------------

class GpuPath:
    def __init__(self, sa, d, variant, mono="factored"):
        self.sa, self.d, self.variant, self.mono = sa, d, variant, mono
        self.launches = 0

    def build_stereo(self):
        B, d = self.sa.CorrBlockB200, self.d
        self.launches += 1
        return B.from_features(d["fl"], d["fr"], radius=RADIUS, num_levels=LEVELS, truncate=(d["tdisp"], d["tconf"], 0.9))

    def build_mono(self):
        B, d = self.sa.CorrBlockB200, self.d
        if self.mono == "aggregated":
            # the README's configuration (--use_aggregate_mono_vol): the mono volume is materialised (it feeds the
            # hourglass, stereoanywhere.py:136-165) and the block is built from a dense volume - here the A2 volume
            # itself stands in for the hourglass output, an arbitrary [B,H,W,1,W] tensor
            self.launches += 2
            return B(B.mono_corr(d["nl"], d["nr"]), radius=RADIUS, num_levels=LEVELS)
        if B.mono_mode != "otf":
            self.launches += 1   # packed / factored: one pack kernel; on the fly: nothing to launch
        return B.from_normals(d["nl"], d["nr"], radius=RADIUS, num_levels=LEVELS)

    def build(self):
        sa, d = self.sa, self.d
        B = sa.CorrBlockB200
        if self.variant == "fused":
            # stereo: corr + truncation + pyramid in the GEMM epilogue (one kernel); mono: normals -> packed
            fs = self.build_stereo()
            fm = self.build_mono()
        else:  # strict reference protocol, op for op (stereoanywhere.py:135-136, 203, 253-259)
            vs = B.corr(d["fl"], d["fr"]).squeeze(3).unsqueeze(1)
            vm = 1.73 * B.corr(d["nl"], d["nr"]).squeeze(3).unsqueeze(1)
            t = sa.truncation_mask(d["tdisp"], d["tconf"], 0.9)
            fs = B((t * vs).squeeze(1).unsqueeze(3), radius=RADIUS, num_levels=LEVELS)
            fm = B(vm.squeeze(1).unsqueeze(3), radius=RADIUS, num_levels=LEVELS)
            self.launches += 5
        return fs, fm

    def coords_seq(self):
        """coords of the 32 GRU iterations.  In the model they come out of the update block
        (stereoanywhere.py:280, out of scope); here: coords0 + k * delta, formed on the device."""
        d = self.d
        k = torch.arange(ITERS, device=d["coords0"].device, dtype=torch.float32).view(ITERS, 1, 1, 1, 1)
        return d["coords0"].unsqueeze(0) + k * d["delta"].unsqueeze(0)  # [ITERS,B,2,H,W]

    def lookups(self, fs, fm, seq, timed_events=None):
        B = self.sa.CorrBlockB200
        if timed_events is not None:
            timed_events[0].record()
        s = m = None
        for k in range(ITERS):
            if self.variant == "fused":
                s, m = B.lookup_pair(fs, fm, seq[k])
                self.launches += 1
            else:
                s, m = fs(seq[k]), fm(seq[k])
                self.launches += 2
        if timed_events is not None:
            timed_events[1].record()
        return s, m, seq[ITERS - 1]

    def step(self, timed_events=None):
        seq = self.coords_seq()
        fs, fm = self.build()
        return self.lookups(fs, fm, seq, timed_events)


def _ev():
    return torch.cuda.Event(enable_timing=True)


def measure_path(sa, d, variant, mono, steps, warmup, graph, after_step=None, drain=None, barrier=None, storage="fp32"):
    """Time `steps` steps of the path on device-resident inputs `d` (CUDA events on the current stream).
    Returns ms per step, ms per lookup launch, launches per step, and (fused + graph) the two builders timed alone."""
    B = sa.CorrBlockB200
    saved_mode, saved_storage = B.mono_mode, B.storage
    B.mono_mode = mono if mono != "aggregated" else "packed"
    B.storage = storage
    otf = variant == "fused" and mono == "otf"
    path = GpuPath(sa, d, variant, mono)
    out = None
    try:
        for _ in range(max(warmup, 3)):
            out = path.step()
            if after_step is not None:
                after_step(out)
        g_build = g_look = g_mono = None
        if graph:
            # The step is ~35 launches of 10-400 us: replay it from CUDA graphs (stereo volume + packing; mono
            # packing; the 32 lookups) so that the events between the graphs time each kernel family alone.
            torch.cuda.synchronize()
            seq = path.coords_seq()
            g_build, g_look = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            if otf:
                with torch.cuda.graph(g_build):
                    fs = path.build_stereo()
                fm = path.build_mono()   # holds the normal maps only: no kernel, nothing to capture
            elif variant == "fused":
                g_mono = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_build):
                    fs = path.build_stereo()
                with torch.cuda.graph(g_mono, pool=g_build.pool()):
                    fm = path.build_mono()
            else:
                with torch.cuda.graph(g_build):
                    fs, fm = path.build()
            with torch.cuda.graph(g_look, pool=g_build.pool()):
                g_out = path.lookups(fs, fm, seq)
            for _ in range(2):
                g_build.replay()
                if g_mono is not None:
                    g_mono.replay()
                g_look.replay()
        if drain is not None:
            drain()
        if barrier is not None:
            barrier()
        else:
            torch.cuda.synchronize()
        path.launches = 0
        lk_events = [(_ev(), _ev()) for _ in range(steps)]
        e0, e1 = _ev(), _ev()
        e0.record()
        for k in range(steps):
            if g_build is not None:
                g_build.replay()
                if g_mono is not None:
                    g_mono.replay()
                lk_events[k][0].record()
                g_look.replay()
                lk_events[k][1].record()
                out = g_out
            else:
                out = path.step(lk_events[k])
            if after_step is not None:
                after_step(out)
        if drain is not None:
            drain()
        e1.record()
        if barrier is not None:
            barrier()
        else:
            torch.cuda.synchronize()
        ms_total = e0.elapsed_time(e1)
        per_step = (33 if otf else 35 if mono == "aggregated" else 34) if variant == "fused" else 69
        launches = path.launches if g_build is None else steps * per_step
        lk_ms = sum(a.elapsed_time(bb) for a, bb in lk_events) / steps
        breakdown = None
        if g_build is not None and variant == "fused":
            # per-kernel times of the two builders (one kernel per graph), measured AFTER the timed region so that no
            # extra event sits between the graphs of a timed step: each graph replayed alone, the lookup graph in
            # between so that the volumes of the previous replay are out of L2 as they are in a step
            def alone(g):
                ts = []
                for _ in range(5):
                    g_look.replay()
                    a0, a1 = _ev(), _ev()
                    a0.record(); g.replay(); a1.record(); torch.cuda.synchronize()
                    ts.append(a0.elapsed_time(a1))
                ts.sort()
                return ts[len(ts) // 2]
            breakdown = (alone(g_build), alone(g_mono) if g_mono is not None else None)
        n_lk = ITERS if variant == "fused" else 2 * ITERS
        return {"ms_step": ms_total / steps, "lk_launch_ms": lk_ms / n_lk, "n_lk_launch": n_lk, "launches": launches,
                "breakdown": breakdown}
    finally:
        B.mono_mode, B.storage = saved_mode, saved_storage


def init_dist():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa(local)   # before the first pinned allocation / CUDA context
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    return rank, world, local, dev, dist, numa


def run_gpu(args):
    rank, world, local, dev, dist, numa = init_dist()
    import stereoanywhere_b200 as sa

    sa.CorrBlockB200.precision = args.precision
    otf = args.variant == "fused" and args.mono == "otf"
    factored = args.variant == "fused" and args.mono == "factored"
    b, c, h, w = WORKLOADS[args.workload]
    host, d = make_inputs(b, c, h, w, dev, seed=rank, pinned=True)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput -------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # The path's only collective: gather the quarter-res disparity of every rank.  It is issued
    # asynchronously (NCCL's own stream, double-buffered) so that step k+1 computes while step k's result
    # travels; the buffers of step k are reclaimed - i.e. the gather is waited for - at step k+2, and every
    # outstanding gather is waited for before the closing event of the timed region.
    gather_bufs = [torch.empty((world * b, 1, h, w), dtype=torch.float32, device=dev) for _ in range(2)] if dist is not None else None
    disp_bufs = [torch.empty((b, 1, h, w), dtype=torch.float32, device=dev) for _ in range(2)]
    pending = [None, None]
    step_no = [0]

    def collective(out):
        if dist is None:
            return None
        i = step_no[0] & 1
        step_no[0] += 1
        if pending[i] is not None:
            pending[i].wait()
        disp_bufs[i].copy_(out[2][:, :1])
        pending[i] = dist.all_gather_into_tensor(gather_bufs[i], disp_bufs[i], async_op=True)
        return gather_bufs[i]

    def drain():
        for i in range(2):
            if pending[i] is not None:
                pending[i].wait()
                pending[i] = None

    m = measure_path(sa, d, args.variant, args.mono, args.steps, args.warmup, args.graph, after_step=collective,
                     drain=drain, barrier=barrier, storage=args.storage)
    ms_total, lk_launch_ms, launches, breakdown = m["ms_step"] * args.steps, m["lk_launch_ms"], m["launches"], m["breakdown"]
    n_lk_launch = m["n_lk_launch"]
    clocks = sampler.stop() if rank == 0 else None

    # the collective's result, checked on NCCL hardware: one more gather of this rank's final disparity; rank 0
    # compares every slice it received with what the owning rank sent (each rank ships a checksum of its slice)
    gather_check = None
    if dist is not None:
        last = disp_bufs[(step_no[0] - 1) & 1].clone()
        full = torch.empty((world * b, 1, h, w), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(full, last)
        sums = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(sums, last.double().abs().sum().reshape(1))
        got = full.view(world, -1).double().abs().sum(1)
        gather_check = {"own_slice_equal": bool(torch.equal(full[rank * b:(rank + 1) * b], last)),
                        "max_rel_checksum_diff": float(((got - sums).abs() / sums.clamp(min=1e-30)).max())}
        assert gather_check["own_slice_equal"] and gather_check["max_rel_checksum_diff"] < 1e-12, gather_check

    # ---- the other two wirings of the same workload, same run (driver-measured) -------------------
    variants = None
    if args.extras and args.workload == DEFAULT_WORKLOAD and args.variant == "fused" and args.mono == "factored":
        torch.cuda.empty_cache()
        va = measure_path(sa, d, "fused", "aggregated", 5, 3, args.graph, barrier=barrier)
        torch.cuda.empty_cache()
        vp = measure_path(sa, d, "protocol", "packed", 5, 3, args.graph, barrier=barrier)
        torch.cuda.empty_cache()
        vh = measure_path(sa, d, "fused", "factored", 5, 3, args.graph, barrier=barrier, storage="fp16")
        torch.cuda.empty_cache()
        variants = {
            "aggregated_mono_ms": round(maxr(va["ms_step"]), 4),
            "protocol_ms": round(maxr(vp["ms_step"]), 4),
            "fp16_storage_ms": round(maxr(vh["ms_step"]), 4),
            "fp16_storage": "opt-in (CorrBlockB200.storage = 'fp16'): the headline wiring with the stereo block's packed pyramid "
                            "stored in fp16 (the fp32 values rounded to nearest: +3e-4 of max|V|, inside the TF32 tolerance); "
                            f"corr_pack {round(vh['breakdown'][0] * 1e3, 1) if vh['breakdown'] else None} us, "
                            f"lookup {round(vh['lk_launch_ms'] * 1e3, 2)} us",
            "steps": 5,
            "aggregated_mono": "README wiring (--use_aggregate_mono_vol, stereoanywhere.py:162-165,210,257-259): mono volume "
                               "materialised, mono block built from a dense volume (sa_pack_pyramid), dual packed lookup",
            "protocol": "strict drop-in call sequence (INTEGRATION section 2, fused=False): corr() x2, 1.73*, truncation mask, "
                        "product, two constructors, 64 single lookups",
        }

    # ---- end to end through the public API from pinned host buffers ---------------------------
    # Every step uploads ITS OWN inputs from pinned host memory and reads its result back; the
    # upload of step k+1 runs on a copy stream while step k computes (two device buffer sets).
    sa.CorrBlockB200.mono_mode = args.mono if args.mono != "aggregated" else "packed"
    sa.CorrBlockB200.storage = args.storage
    keys = ["fl", "fr", "nl", "nr", "coords0", "delta", "tdisp", "tconf"]
    sets = [{k: torch.empty_like(d[k]) for k in keys} for _ in range(2)]
    res_s = torch.empty((b, LEVELS * (2 * RADIUS + 1), h, w), dtype=torch.float32).pin_memory()
    res_m = torch.empty_like(res_s).pin_memory()
    h2d = sum(host[k].numel() * 4 for k in keys)
    d2h = res_s.numel() * 4 * 2
    paths = [GpuPath(sa, sset, args.variant, args.mono) for sset in sets]
    copy_stream = torch.cuda.Stream()
    d2h_stream = torch.cuda.Stream()                 # results leave on their own stream: H2D / compute / D2H overlap
    main_stream = torch.cuda.current_stream()
    computed = [torch.cuda.Event() for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]   # upload of set i finished
    freed = [torch.cuda.Event() for _ in range(2)]   # compute on set i finished (buffers reusable)

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[i])
            for k in keys:
                sets[i][k].copy_(host[k], non_blocking=True)
            ready[i].record(copy_stream)

    def e2e_run(n):
        for i in range(2):
            freed[i].record(main_stream)
        upload(0)
        for k in range(n):
            i = k & 1
            if k + 1 < n:
                upload(i ^ 1)
            main_stream.wait_event(ready[i])
            s, mm, _ = paths[i].step()
            freed[i].record(main_stream)
            computed[i].record(main_stream)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(computed[i])
                res_s.copy_(s, non_blocking=True)
                res_m.copy_(mm, non_blocking=True)
                s.record_stream(d2h_stream)
                mm.record_stream(d2h_stream)
        main_stream.wait_stream(d2h_stream)          # the closing event covers the last read-back

    e2e_run(3)
    barrier()
    t0 = time.perf_counter()
    f0, f1 = _ev(), _ev()
    f0.record()
    e2e_run(args.steps)
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    e2e_wall = (time.perf_counter() - t0) * 1e3

    # the ceiling of that number: the same uploads with NO compute, all ranks at once (what the host side - PCIe
    # links, root complexes, host DRAM - delivers to N GPUs simultaneously)
    def copy_only(n):
        with torch.cuda.stream(copy_stream):
            for _ in range(n):
                for k in keys:
                    sets[0][k].copy_(host[k], non_blocking=True)

    copy_only(1)
    barrier()
    c0, c1 = _ev(), _ev()
    with torch.cuda.stream(copy_stream):
        c0.record(copy_stream)
    copy_only(3)
    with torch.cuda.stream(copy_stream):
        c1.record(copy_stream)
    barrier()
    copy_ms = c0.elapsed_time(c1) / 3

    ms_total, e2e_ms, lk_launch_ms, copy_ms = maxr(ms_total), maxr(max(e2e_ms, 0.0)), maxr(lk_launch_ms), maxr(copy_ms)
    ms_step = ms_total / args.steps
    value = world * b / (ms_step / 1e3)
    e2e_value = world * b / (e2e_ms / args.steps / 1e3)

    tiled = None
    if args.extras and args.tiled:
        torch.cuda.empty_cache()
        try:
            tiled = tiled_measure(sa, dev, rank, world, dist, args, steps=max(5, min(args.steps, 20)), warmup=3)
        except Exception as e:   # reported in the line; the headline of this run stands on its own
            tiled = {"error": f"{type(e).__name__}: {e}"}
            print(f"[bench] tiled_c4 failed: {tiled['error']}", file=sys.stderr)

    result = None
    if rank == 0:
        peak, peak_src = load_peaks()
        p = b * h * w
        # algorithmic bytes per pixel and launch: coords 4 + 160 per volume of windows + 144 per volume of
        # outputs.  With the mono volume computed on the fly its 160 B of windows are not memory traffic any more.
        # Factored mono volume: its windows come out of the L2-resident packed right normals; the left normal is read.
        alg_px = (452 if otf else 464 if factored else 612) if args.variant == "fused" else 308
        real_px = (416 if otf else 432 if factored else 544)
        alg = alg_px * p
        achieved = alg / (lk_launch_ms * 1e-3) / 1e9
        s8d = 612 if args.variant == "fused" else 308
        roof = {"bound": "hbm", "kernel": "lookup_packed_kernel" + ("<NV=2>" if args.variant == "fused" else "<NV=1>"),
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": TRAFFIC_BYTES.get((args.workload, args.variant, "packed" if args.mono == "aggregated" else args.mono)),
                "peak_source": peak_src, "algorithmic_bytes_per_pixel": alg_px,
                "algorithmic_bytes_per_launch": alg, "launch_us": round(lk_launch_ms * 1e3, 2),
                # SURVEY 8d counts 612 B/px for the dual lookup (both volumes' windows read from memory); the factored
                # and on-the-fly forms do not read the mono windows, `achieved` above uses their own smaller figure
                "survey_8d_definition": {"bytes_per_pixel": s8d, "gbs": round(s8d * p / (lk_launch_ms * 1e-3) / 1e9, 1),
                                         "frac": round(s8d * p / (lk_launch_ms * 1e-3) / 1e9 / peak, 4),
                                         "note": "informational: SURVEY 8d's per-pixel figure counts mono windows that the "
                                                 "factored / on-the-fly forms never read, so this can exceed 1; `frac` above "
                                                 "is on the bytes this kernel's own algorithm needs"},
                "path_algorithmic_gbs": round(path_bytes(b, c, h, w) / (ms_step * 1e-3) / 1e9, 1)}
        kernels = None
        if breakdown is not None:
            st_us = breakdown[0] * 1e3
            mo_us = breakdown[1] * 1e3 if breakdown[1] is not None else None
            packed = p * (w // 8 + 9) * (128 if args.storage == "fp32" else 64)
            mono_bytes = 3 * b * h * (w // 8 + 9) * 128 + 3 * b * h * w * 4 if factored else p * (w // 8 + 9) * 128
            if args.mono == "aggregated":   # volume written, read again, packed
                mono_bytes = 2 * p * w * 4 + packed
            tf32_peak = tensor_peak_tf32()
            tflops = 2.0 * p * w * c / (st_us * 1e-6) / 1e12
            kernels = {
                "corr_pack_tf32": {"us": round(st_us, 1), "hbm_gbs": round((2 * b * c * h * w * 4 + packed) / st_us / 1e3, 1),
                                   "hbm_frac": round((2 * b * c * h * w * 4 + packed) / st_us / 1e3 / peak, 4),
                                   "tflops": round(tflops, 1), "tensor_peak_tf32": tf32_peak,
                                   "tensor_frac": round(tflops / tf32_peak, 4),
                                   "note": "HBM-bound by the packed write; the tensor pipe is reported, not targeted"},
                "pack_normals": ({"us": round(mo_us, 1), "hbm_gbs": round(mono_bytes / mo_us / 1e3, 1),
                                  "hbm_frac": round(mono_bytes / mo_us / 1e3 / peak, 4),
                                  **({"note": "factored: only the right normal map's rows are packed (launch-bound)"} if factored else {}),
                                  **({"note": "two kernels: A2 volume materialised (stand-in for the hourglass output), "
                                              "then sa_pack_pyramid of the dense volume"} if args.mono == "aggregated" else {})}
                                 if mo_us is not None else
                                 "not launched: the mono lookups are computed from the normal maps inside the lookup kernel"),
                "lookup_packed2": {"us": round(lk_launch_ms * 1e3, 2), "launches": n_lk_launch,
                                   "real_bytes_per_pixel": real_px,
                                   "hbm_gbs_real_bytes": round(real_px * p / (lk_launch_ms * 1e-3) / 1e9, 1),
                                   "hbm_frac_real_bytes": round(real_px * p / (lk_launch_ms * 1e-3) / 1e9 / peak, 4)},
            }
            roof["kernels"] = kernels   # also inside `roofline`: the driver's parser keeps this object whole
        cpu = None if args.no_cpu_baseline else cpu_baseline(args.workload, sample_pairs=args.cpu_pairs)
        copy_gbs = h2d / (copy_ms * 1e-3) / 1e9
        result = {
            "metric": metric_for(args.workload), "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": f"{args.precision} corr (fp32 accumulate), f32 pyramid/lookup",
            "data": "synthetic (seeded N(0,1) features, unit normals, U(0,W/4) disparities)",
            "config": workload_config(args.workload, world),
            "impl_config": {"variant": args.variant, "cuda_graph": bool(args.graph), "storage": args.storage,
                            "mono": ("on the fly: lookups computed from the normal maps inside the lookup kernel, bit-identical to "
                                     "the packed pyramid (no mono volume / pyramid in memory)") if otf else
                                    ("factored: packed pyramid of the right normal map's rows (rank-3 volume, linear pyramid); "
                                     "the lookup combines three lines with the pixel's left normal") if factored else
                                    ("aggregated (README configuration): dense mono volume materialised, block built from a "
                                     "dense volume") if args.mono == "aggregated" else "packed pyramid"},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e2e_ms / args.steps, 4), "wall_ms_per_step": round(e2e_wall / args.steps, 4),
                    # the bound, measured: the same uploads with no compute, all ranks at once (max over ranks)
                    "h2d_only_ms_per_step": round(copy_ms, 4), "h2d_gbs_per_rank": round(copy_gbs, 1),
                    "h2d_gbs_all_ranks": round(copy_gbs * world, 1),
                    "copy_bound_value": round(world * b / (copy_ms / 1e3), 2),
                    "host_binding": numa},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "kernels": kernels,
            "gather_check": gather_check,
            "variants": variants,
            "tiled_c4": tiled,
            "cpu_baseline": cpu,
        }
        emit(result)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return result


# ------------------------------------------------------------------------------------------
# config 4: full-resolution Middlebury, tiles sharded over the ranks (strong scaling)
# ------------------------------------------------------------------------------------------

def tiled_measure(sa, dev, rank, world, dist, args, steps, warmup):
    """BASELINE config 4: `args.images` full-resolution pairs per step, reference tile geometry (`--tile-preset`,
    distinct tiles weighted by multiplicity, mapreduce_v2/tile_wrapper.py:101-120), every tile runs the whole hot path
    at its quarter resolution (one tile at a time, as the reference does), and `tiling.SlotStitcher` stitches without a
    collective: the rank that ran a tile stores the weighted tile straight into rank 0's memory (NVLink peer stores,
    `sa_stitch_tile`), rank 0 sums the slots in enumeration order and normalises (`sa_stitch_finish`).  Path-only:
    a tile's disparity is its final lookup coordinate field (the encoders / GRU that would produce it are out of
    scope).  Before the timed region rank 0 checks the N-rank image against the image it computes alone."""
    from stereoanywhere_b200 import tiling

    B = sa.CorrBlockB200
    saved_mode, B.mono_mode = B.mono_mode, (args.mono if args.mono in ("factored", "packed", "otf") else "packed")
    H, W = 1984, 2880                      # 1984x2872 replicate-padded to /32 (test_mapreduce_v2.py:217-227)
    th, tw, ov = tiling.PRESETS[args.tile_preset]
    work = tiling.tile_multiplicity(H, W, th, tw, ov)
    images = args.images

    def geometry(tile):
        y0, y1, x0, x1 = tile
        pad = tiling.pad_to_32(y1 - y0, x1 - x0)
        return pad, ((y1 - y0 + pad[2] + pad[3]) // 4, (x1 - x0 + pad[0] + pad[1]) // 4)

    def tile_path(d):
        fs = B.from_features(d["fl"], d["fr"], radius=RADIUS, num_levels=LEVELS, truncate=(d["tdisp"], d["tconf"], 0.9))
        fm = B.from_normals(d["nl"], d["nr"], radius=RADIUS, num_levels=LEVELS)
        # coords of the 32 iterations formed up front, as in the batch workloads (GpuPath.coords_seq): in the model
        # they come out of the update block, which is out of scope
        k = torch.arange(ITERS + 1, device=dev, dtype=torch.float32).view(ITERS + 1, 1, 1, 1, 1)
        seq = d["coords0"].unsqueeze(0) + k * d["delta"].unsqueeze(0)
        for i in range(ITERS):
            B.lookup_pair(fs, fm, seq[i])
        return (d["coords0"] - seq[ITERS])[:, :1].contiguous()   # quarter-resolution disparity, positive

    pool = [None]
    runners = {}

    def runner(img, hw):
        """Inputs (seeded by image and tile shape: the same on every rank) + the CUDA graph of one tile run."""
        key = (img, hw)
        if key not in runners:
            d = make_inputs(1, 256, hw[0], hw[1], dev, seed=1000 + 16 * img + (hw[0] * 7 + hw[1]) % 16)[1]
            if not args.graph:
                runners[key] = (None, d, None)
            else:
                for _ in range(2):
                    tile_path(d)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool[0]):
                    q = tile_path(d)
                if pool[0] is None:
                    pool[0] = g.pool()
                runners[key] = (g, d, q)
        return runners[key]

    def run_unit(st, u):
        img, tile, _mult = st.units[u]
        pad, hw = geometry(tile)
        g, d, q = runner(img, hw)
        if g is None:
            q = tile_path(d)
        else:
            g.replay()
        st.add(u, q, up=4, scale=4.0, pad_top=pad[2], pad_left=pad[0])
        return q

    def step(st, units):
        st.begin()
        for u in units:
            run_unit(st, u)
        st.end()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def agree(ok, what):
        """All ranks continue or all ranks raise: a failure on one rank must not leave the others in a barrier."""
        if dist is not None:
            t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = bool(t.item())
        if not ok:
            raise RuntimeError(what)

    try:
        st, err = None, ""
        try:
            st = tiling.SlotStitcher(images, H, W, work, dev)
        except Exception as e:  # symmetric memory unavailable on this box, ...
            err = f"{type(e).__name__}: {e}"
        agree(st is not None, f"SlotStitcher could not be set up on every rank ({err or 'another rank failed'})")
        mine = st.my_units()
        # ---- correctness before speed: N ranks against one rank, and against the host stitch -----------------
        step(st, mine)
        got = st.drain()
        check = None
        if rank == 0:
            got = got.clone()
            torch.cuda.synchronize()
            if world > 1:
                solo = tiling.SlotStitcher(images, H, W, work, dev, local=True)
                step(solo, solo.my_units())
                alone = solo.drain().clone()
            else:
                solo, alone = st, got
            # host stitch (round 1's eager accumulate, the arithmetic of tiling.tiled_inference) of the same tiles
            acc = torch.zeros(images, H, W, device=dev)
            for u in range(len(solo.units)):
                img, (y0, y1, x0, x1), mult = solo.units[u]
                pad, hw = geometry((y0, y1, x0, x1))
                g, d, q = runner(img, hw)
                if g is None:
                    q = tile_path(d)
                else:
                    g.replay()
                full = torch.nn.functional.interpolate(q, scale_factor=4, mode="nearest") * 4.0
                full = full[..., pad[2]: full.shape[-2] - pad[3], pad[0]: full.shape[-1] - pad[1]]
                acc[img, y0:y1, x0:x1].addcmul_(full[0, 0], tiling.blend_weight(y1 - y0, x1 - x0, device=dev) * float(mult))
            host = acc / torch.clamp(tiling.weight_sum(H, W, work, device=dev), min=1e-4)
            check = {"n_rank_vs_1_rank_max_abs": float((got - alone).abs().max()),
                     "vs_host_stitch_max_abs": float((alone - host).abs().max()),
                     "mean_abs_disparity": float(host.abs().mean())}
            del acc, host, alone
            if world > 1:
                del solo
        good = check is None or (check["n_rank_vs_1_rank_max_abs"] <= 1e-4 and
                                 check["vs_host_stitch_max_abs"] <= 1e-3 * max(1.0, check["mean_abs_disparity"]))
        agree(good, f"tile stitch is wrong: {check}")
        barrier()
        for _ in range(max(warmup, 3)):
            step(st, mine)
        st.drain()
        barrier()
        e0, e1 = _ev(), _ev()
        e0.record()
        for _ in range(steps):
            step(st, mine)
        st.drain()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        ms_step = ms / steps
        out = {"metric": "stereo pairs/sec @1984x2872 tiled (cost-volume path per tile, 32 iters)",
               "value": round(images / (ms_step / 1e3), 3), "unit": UNIT, "n_gpus": world, "steps": steps,
               "ms_per_step": round(ms_step, 4), "scaling": "strong", "images_per_step": images,
               "tile_preset": args.tile_preset, "distinct_tiles_per_image": len(work), "tiles_this_rank": len(mine),
               "tiles_per_step": len(st.units), "cuda_graph": bool(args.graph),
               "mono": B.mono_mode,
               "stitch": ("slots over NVLink peer memory: one sa_stitch_tile per tile stores the weighted tile into rank 0, "
                          "sa_stitch_finish sums in enumeration order and normalises; no collective") if world > 1 else
                         "slots (local): sa_stitch_tile per tile + sa_stitch_finish",
               "stitch_check": check}
        del st
        runners.clear()
        torch.cuda.empty_cache()
        return out
    finally:
        B.mono_mode = saved_mode


def run_tiled(args):
    rank, world, local, dev, dist, _numa = init_dist()
    import stereoanywhere_b200 as sa

    sa.CorrBlockB200.precision = args.precision
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t = tiled_measure(sa, dev, rank, world, dist, args, steps=args.steps, warmup=args.warmup)
    if rank == 0:
        clocks = sampler.stop()
        emit({"metric": t["metric"], "value": t["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
              "warmup": max(args.warmup, 3), "ms_per_step": t["ms_per_step"], "higher_is_better": True, "scaling": "strong",
              "vs_baseline": None, "dtype": f"{args.precision} corr, f32 pyramid/lookup", "data": "synthetic",
              "config": {"workload": TILED_WORKLOAD, **{k: t[k] for k in ("images_per_step", "tile_preset",
                         "distinct_tiles_per_image", "tiles_this_rank", "tiles_per_step", "cuda_graph", "mono", "stitch")},
                         "parallelism": f"tiles sharded x{world}, collective-free stitch"},
              "stitch_check": t["stitch_check"], "clocks": clocks})
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
# ncu --set full capture (profiles/); None until a capture for that workload exists.
TRAFFIC_BYTES = {
    # profiles/r1/lookup_tile64_pack_normals_ncu_summary.txt: 60.46 MB read + 13.9 MB written while the kernel runs
    # (the other ~55 MB of its 69 MB output are still dirty in L2 at kernel end and reach HBM later)
    ("c2_kitti_375x1242_b8", "fused", "packed"): 74_360_000,
    # profiles/r2/lookup_final2_ncu_summary.txt (cold capture of one launch): 46.2 MB read (30.7 MB of stereo lines, the
    # left normals, the coords, first touches of the packed right normals) + 13.0-14.3 MB written while the kernel runs
    # (the rest of the 69 MB output is still in L2 at kernel end and reaches HBM during the next launch)
    ("c2_kitti_375x1242_b8", "fused", "factored"): 59_850_000,
}


# ------------------------------------------------------------------------------------------
# CPU: the reference's own classes (oracle/_ref), else the oracle port
# ------------------------------------------------------------------------------------------

def cpu_runner():
    """(callable, kind, description): the reference's own `CorrBlock1D` when its files are present (`/root/reference`
    in the build container, the prebuilt `oracle/_ref` on the GPU box), else the oracle's restatement."""
    from oracle import ref_path

    if ref_path.available():
        return ref_path.run_path_reference, "reference", ("oracle/ref_path.run_path_reference: the reference's own "
                                                          "CorrBlock1D / truncate_corr_volume_v2 (unmodified files)")
    from oracle import corr_oracle as O

    return O.run_path_cpu, "port", "oracle/corr_oracle.run_path_cpu (restatement of the reference ATen op sequence)"


def cpu_once(workload, pairs, fn):
    b, c, h, w = WORKLOADS[workload]
    pairs = min(pairs, b)
    host, _ = make_inputs(pairs, c, h, w, None, seed=0)
    coords = [host["coords0"] + k * host["delta"] for k in range(ITERS)]
    t0 = time.perf_counter()
    with torch.no_grad():
        fn(host["fl"], host["fr"], host["nl"], host["nr"], coords, trunc=(host["tdisp"], host["tconf"], 0.9),
           radius=RADIUS, num_levels=LEVELS)
    return time.perf_counter() - t0, pairs


def cpu_baseline(workload, sample_pairs=8, reps=2, min_seconds=10.0):
    """Bounded CPU sample: whole batches of the workload until >= min_seconds of work (>= reps batches; the
    reference's own block needs ~11 s per KITTI-size batch of 8 on 16 cores: 2 batches)."""
    torch.set_num_threads(os.cpu_count() or 1)
    fn, kind, what = cpu_runner()
    cpu_once(workload, 1, fn)  # warm-up (thread pool, allocator)
    total_t, total_pairs, n = 0.0, 0, 0
    while n < reps or total_t < min_seconds:
        dt, pairs = cpu_once(workload, sample_pairs, fn)
        total_t += dt
        total_pairs += pairs
        n += 1
        if n >= 64:
            break
    return {"value": round(total_pairs / total_t, 4), "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{n} x {pairs} of {WORKLOADS[workload][0]} pairs of {workload}, all {ITERS} iterations, "
                      f"{what}, {total_t:.1f} s of CPU work",
            "host_cpus": os.cpu_count()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    fn, kind, what = cpu_runner()
    b = WORKLOADS[args.workload][0]
    pairs = min(args.cpu_pairs, b)
    for _ in range(args.warmup):
        cpu_once(args.workload, pairs, fn)
    t = 0.0
    for _ in range(args.steps):
        dt, _ = cpu_once(args.workload, pairs, fn)
        t += dt
    value = pairs * args.steps / t
    sample = (f"each step = {pairs} of {b} pairs of {args.workload}, all {ITERS} iterations, {what} on the host cores")
    on_gpu = reference_on_gpu(args, fn) if kind == "reference" else None
    emit({
        "impl": "reference", "metric": metric_for(args.workload), "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t / args.steps * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, args.gpus),
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                         "sample": sample},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        **({"reference_on_b200": on_gpu} if on_gpu else {}),
    })


def reference_on_gpu(args, fn):
    """SURVEY 8d's "real bar", reported next to the CPU number of the reference arm: the SAME unmodified reference
    classes executed by ATen on cuda:0 (device-resident inputs, the whole batch, all 32 iterations, including the
    reference's per-level `torch.unique` host sync).  None when there is no GPU."""
    if not torch.cuda.is_available():
        return None
    try:
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        b, c, h, w = WORKLOADS[args.workload]
        _, d = make_inputs(b, c, h, w, dev, seed=0)
        coords = [d["coords0"] + k * d["delta"] for k in range(ITERS)]

        def step():
            with torch.no_grad():
                fn(d["fl"], d["fr"], d["nl"], d["nr"], coords, trunc=(d["tdisp"], d["tconf"], 0.9), radius=RADIUS,
                   num_levels=LEVELS)

        for _ in range(max(args.warmup, 1)):
            step()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / args.steps
        return {"value": round(b / ms * 1e3, 2), "unit": UNIT, "ms_per_step": round(ms, 3), "pairs_per_step": b,
                "what": "the same unmodified CorrBlock1D / truncate_corr_volume_v2 (oracle/_ref) run by ATen on this B200, "
                        "device-resident inputs - the reference as a GPU user runs it today"}
    except Exception as e:  # the CPU number is the arm's result; this is extra information
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


_RESULT_OUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line when the
    box sets NCCL_DEBUG), so the real stdout is kept aside for `emit` and file descriptor 1 is pointed at stderr."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS) + [TILED_WORKLOAD])
    ap.add_argument("--tile-preset", default="middlebury", help="reference tile preset for the tiled workload")
    ap.add_argument("--images", type=int, default=4, help="full-resolution pairs per step of the tiled workload")
    ap.add_argument("--variant", default="fused", choices=["fused", "protocol"],
                    help="fused: truncate= / mono_corr / lookup_pair entry points; protocol: the reference's exact call sequence")
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"], help="stereo correlation kernel")
    ap.add_argument("--mono", default="factored", choices=["otf", "packed", "factored", "aggregated"],
                    help="mono block of the fused variant: factored (packed right normals, combined inside the lookup), "
                         "the packed pyramid of the volume, lookups computed on the fly from the normals, or 'aggregated': "
                         "the README configuration, block built from a dense (hourglass-output-like) volume")
    ap.add_argument("--storage", default="fp32", choices=["fp32", "fp16", "bf16"],
                    help="storage of the stereo block's packed pyramid (fused variant): fp32 (default) or the opt-in 16-bit modes")
    ap.add_argument("--graph", type=int, default=1, help="replay the step from a CUDA graph in the device-resident run")
    ap.add_argument("--cpu-pairs", type=int, default=8, help="pairs in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", type=int, default=1, help="0: skip the `variants` and `tiled_c4` objects (profiling runs)")
    ap.add_argument("--tiled", type=int, default=1, help="0: skip the `tiled_c4` object of the default line")
    args = ap.parse_args()
    if args.workload == TILED_WORKLOAD:
        if args.impl == "reference":
            args.workload = "c4_middlebury_tile_1120x672_b1"  # one reference tile as the bounded CPU sample
            run_reference(args)
        else:
            run_tiled(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
