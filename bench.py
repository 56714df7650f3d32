#!/usr/bin/env python3
"""bench.py -- registration + fusion of a synthetic 96-well plate (BASELINE.json configs[2], the
configuration north_star's target is quoted on): 96 wells x 3x3 tiles x 2048^2 uint16 x 4 channels.

One step = (a) all-pairs phase-correlation registration of every well on the registration channel
(12 adjacent pairs / well, 1152 pairs) and (b) flatfield-corrected fusion of every well's 4 channels
into its (1, 4, 1, 5734, 5734) canvas in the reference's paste semantics, coordinate placement.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # own arm (CUDA, libstitchb200)
    python bench.py --impl reference ...                             # CPU arm: the oracle port of the reference

Prints ONE JSON line (rank 0).  `value` = canvas Mpx written per second over the whole job with the
plate resident in HBM; `e2e` = the same through the public host-buffer API (pinned host tiles in,
host canvas out, copies inside the timed region).  Multi-GPU: one process per GPU, every rank owns a
whole plate (weak scaling, no data-path collective); time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fused Mpx/s written (registration + fusion step)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--config", type=int, default=2, choices=[0, 1, 2, 3, 4],
                    help="BASELINE.json configs[i] shapes (2 = the headline plate; the others are kept bench lines, see profiles/)")
    ap.add_argument("--wells", type=int, default=None)
    ap.add_argument("--tile", type=int, default=None)
    ap.add_argument("--channels", type=int, default=None)
    ap.add_argument("--grid", type=int, default=None)
    ap.add_argument("--num-z", type=int, default=None)
    ap.add_argument("--blend", choices=["paste", "linear", "feather"], default=None)
    ap.add_argument("--no-flatfield", action="store_true")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="multi-GPU: weak = one plate per rank; strong = ONE plate, its wells split over the ranks")
    ap.add_argument("--no-f64", action="store_true", help="skip the informational float64 registration pass")
    ap.add_argument("--no-affinity", action="store_true", help="do not bind the rank to the CPUs / NUMA node of its GPU")
    ap.add_argument("--e2e-steps", type=int, default=-1, help="steps of the host-buffer leg (default min(steps, 2))")
    ap.add_argument("--host-wells", type=int, default=6, help="distinct wells kept in pinned host memory for e2e")
    ap.add_argument("--fuse-lanes", type=int, default=3, help="library lanes (streams) the per-well fusion launches rotate over")
    ap.add_argument("--per-well-fusion", action="store_true",
                    help="one sb_fuse_region launch per well (round-1 behaviour) instead of one sb_fuse_regions launch per plate")
    ap.add_argument("--no-coordinate-only", action="store_true", help="skip the informational fusion pass without the flat-field")
    ap.add_argument("--e2e-partial-upload", choices=["off", "boxes", "rows"], default="off",
                    help="e2e leg: upload only the pixels (boxes) / rows that can reach the canvas")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-wells", type=int, default=0, help="wells in the CPU sample (0 = one per host core)")
    args = ap.parse_args()
    # BASELINE.json configs: (wells, grid, tile, channels, num_z, blend, flat-field)
    presets = {0: (1, 2, 2048, 1, 1, "linear", False), 1: (1, 5, 2048, 3, 1, "paste", True), 2: (96, 3, 2048, 4, 1, "paste", True),
               3: (384, 2, 2048, 4, 1, "feather", True), 4: (1, 20, 3000, 1, 5, "paste", False)}
    w, g, t, c, z, b, ff = presets[args.config]
    args.wells = w if args.wells is None else args.wells
    args.grid = g if args.grid is None else args.grid
    args.tile = t if args.tile is None else args.tile
    args.channels = c if args.channels is None else args.channels
    args.num_z = z if args.num_z is None else args.num_z
    args.blend = b if args.blend is None else args.blend
    if not ff:
        args.no_flatfield = True
    if args.config != 2:
        args.no_e2e = True                     # the host-buffer leg is defined on the headline plate
    return args


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms.  It is started before the warm-up steps (same load as
    the timed steps; nvidia-smi itself needs ~100 ms to come up) and every sample is stamped with the host time it
    arrived, so the ones inside the timed region can be told apart."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_load0=None, t_timed0=None, t_timed1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, timed = [], [], set(), 0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if t_load0 is not None and not (t_load0 <= ts <= (t_timed1 or ts) + 0.06):
                continue
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            if t_timed0 is not None and t_timed0 <= ts <= t_timed1 + 0.06:
                timed += 1
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_region": timed,
                "window": "warm-up + timed steps (identical load), 50 ms period"}


# ------------------------------------------------------------------------------------------ CPU arm (oracle)
def cpu_sample(spec, plate_tiles_host, flat_host, n_wells, use_flat, workers=None):
    """Time the oracle (NumPy/SciPy restatement of the reference) on `n_wells` wells held in RAM.

    Returns (pairs, px, seconds_registration, seconds_fusion).  File decode is excluded, as in the GPU
    arms.  This is the one place outside tests/ where oracle/ is executed: as the measured CPU arm.
    """
    from oracle import stitch_ref as sr
    ovx, ovy = spec.strip_overlaps()
    xs, ys = spec.stage_positions()
    names = [f"ch{c}" for c in range(spec.channels)]
    st = sr.RegionState(tile_h=spec.tile_h, tile_w=spec.tile_w, pixel_size_um=spec.pixel_size_um,
                        pixel_binning=spec.pixel_binning, monochrome_channels=names, channel_names=names,
                        num_z=spec.num_z, use_registration=False, apply_flatfield=use_flat,
                        flatfields={c: flat_host[c] for c in range(spec.channels)} if use_flat else {})
    from image_stitcher_b200 import geometry as geo
    t_reg = t_fuse = 0.0
    pairs = px = 0
    for w in range(n_wells):
        tiles = plate_tiles_host[w]                       # [rows, cols, C, Z, H, W] uint16
        t0 = time.perf_counter()
        for kind, (r0, c0), (r1, c1) in geo.grid_pairs(spec.rows, spec.cols):
            a, b = tiles[r0, c0, spec.reg_channel, 0], tiles[r1, c1, spec.reg_channel, 0]
            if kind == "h":
                sr.calculate_horizontal_shift(a, b, ovx)
            else:
                sr.calculate_vertical_shift(a, b, ovy)
            pairs += 1
        t_reg += time.perf_counter() - t0
        recs = []
        fovs = sorted(range(spec.rows * spec.cols), key=lambda f: str(f))
        for fov in fovs:
            r, c = divmod(fov, spec.cols)
            for z in range(spec.num_z):
                for ch in range(spec.channels):
                    recs.append(sr.TileRec(x_mm=xs[c], y_mm=ys[r], z_level=z, channel=names[ch],
                                           pixels=tiles[r, c, ch, z], fov=fov))
        t0 = time.perf_counter()
        canvas = sr.stitch_region(st, recs)
        t_fuse += time.perf_counter() - t0
        px += canvas.size
    return pairs, px, t_reg, t_fuse


_POOL_STATE = {}


def _pool_task(i):
    """One well of the CPU sample in a forked worker (arrays inherited, nothing pickled but the timings)."""
    st = _POOL_STATE
    tiles = st["tiles"]
    return cpu_sample(st["spec"], tiles[i % len(tiles)][None], st["flat"], 1, st["use_flat"])


def cpu_pool_sample(spec, tiles, flat, use_flat, n_tasks, procs):
    """`n_tasks` wells (cycling over the distinct wells in `tiles`) over `procs` forked worker processes.
    Returns (pairs, px, wall_seconds, cpu_seconds_registration, cpu_seconds_fusion)."""
    import multiprocessing as mp
    _POOL_STATE.update(spec=spec, tiles=tiles, flat=flat, use_flat=use_flat)
    t0 = time.perf_counter()
    if procs <= 1:
        res = [_pool_task(i) for i in range(n_tasks)]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_pool_task, range(n_tasks), chunksize=1)
    wall = time.perf_counter() - t0
    return (sum(r[0] for r in res), sum(r[1] for r in res), wall, sum(r[2] for r in res), sum(r[3] for r in res))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the box's host cores.  The Python
    reference cannot travel to the GPU box and its third-party stack is not installed, so this is the oracle port
    (oracle/stitch_ref.py + oracle/pcc_ref.py, validated against the reference itself by tests/golden), one well
    per worker process over all host cores.  Each step is a bounded sample: `cores` wells of the named plate."""
    if rank != 0:
        return
    from image_stitcher_b200.plate import PlateSpec
    from oracle import synth
    spec = PlateSpec(wells=args.wells, rows=args.grid, cols=args.grid, tile_h=args.tile, tile_w=args.tile,
                     channels=args.channels, num_z=args.num_z, reg_channel=min(1, args.channels - 1))
    use_flat = not args.no_flatfield
    cores = host_cores()
    n_tasks = args.cpu_sample_wells if args.cpu_sample_wells > 0 else cores
    distinct = min(2, n_tasks)
    # bounded sample: wells generated on the host with the oracle's own generator, reused round-robin by the tasks
    tiles = np.empty((distinct, spec.rows, spec.cols, spec.channels, spec.num_z, spec.tile_h, spec.tile_w), np.uint16)
    for w in range(distinct):
        st, recs, _ = synth.make_region(spec.rows, spec.cols, spec.tile_h, spec.tile_w, seed=w, jitter=3)
        for t in recs:
            r, c = divmod(t.fov, spec.cols)
            for ch in range(spec.channels):
                for z in range(spec.num_z):
                    tiles[w, r, c, ch, z] = t.pixels if ch == spec.reg_channel else (t.pixels // (ch + 2))
    flat = np.stack([synth.vignette(spec.tile_h, spec.tile_w, 0.35, (0.04 * (c + 1), -0.03 * (c + 1)))
                     for c in range(spec.channels)])
    for _ in range(min(args.warmup, 1)):
        cpu_pool_sample(spec, tiles, flat, use_flat, min(n_tasks, cores), cores)
    pairs = px = 0
    wall = t_reg = t_fuse = 0.0
    for _ in range(args.steps):
        p, x, w_s, a, b = cpu_pool_sample(spec, tiles, flat, use_flat, n_tasks, cores)
        pairs += p; px += x; wall += w_s; t_reg += a; t_fuse += b
    val = px / 1e6 / wall
    share_reg = t_reg / (t_reg + t_fuse)
    sample = (f"{n_tasks} of {spec.wells} wells per step ({pairs // args.steps} pairs, {px // args.steps / 1e6:.0f} Mpx), "
              f"one well per worker process over {cores} cores, arrays in RAM; NumPy/SciPy oracle port of the reference")
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mpx/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16 pixels; f64 registration",
        "data": "synthetic",
        "config": {"workload": workload_name(spec, use_flat, args.blend, args.config), "sample": sample},
        "tile_pairs_per_s": pairs / (wall * share_reg), "fusion_mpx_per_s": px / 1e6 / (wall * (1 - share_reg)),
        "cpu_baseline": {"value": val, "unit": "Mpx/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def workload_name(spec, use_flat, blend, config=2, wells_total=None):
    wells = spec.wells if wells_total is None else wells_total
    z = f", {spec.num_z} z-planes" if spec.num_z > 1 else ""
    return (f"{wells}-well plate, {spec.rows}x{spec.cols} tiles/well, {spec.tile_h}x{spec.tile_w} uint16, "
            f"{spec.channels} channels{z}, all-pairs registration on one channel + coordinate-placed {blend} fusion, "
            f"flatfield {'on' if use_flat else 'off'} (BASELINE.json configs[{config}])")


def bind_to_gpu_cpus(local_rank: int):
    """Bind this rank to the CPUs next to its GPU (sysfs local_cpulist of the PCI device) so that its pinned buffers are
    allocated on that NUMA node and its cudaMemcpyAsync submissions do not compete with the other ranks for one core
    set.  When every GPU reports the same CPU list (r1: all eight on NUMA 0 / CPUs 0-31) the list is split evenly over
    the local ranks instead.  Returns a description for the JSON line."""
    try:
        import torch
        world_local = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
        pr = torch.cuda.get_device_properties(local_rank)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        cpus, node = None, None
        base = f"/sys/bus/pci/devices/{bus}"
        if os.path.exists(base + "/local_cpulist"):
            txt = open(base + "/local_cpulist").read().strip()
            cpus = []
            for part in txt.split(","):
                if "-" in part:
                    a, b = part.split("-")
                    cpus += list(range(int(a), int(b) + 1))
                elif part:
                    cpus.append(int(part))
            node = open(base + "/numa_node").read().strip() if os.path.exists(base + "/numa_node") else None
        allowed = sorted(os.sched_getaffinity(0))
        cpus = [c for c in (cpus or allowed) if c in allowed] or allowed
        if world_local > 1:                               # share the list between the local ranks
            per = max(1, len(cpus) // world_local)
            mine = cpus[(local_rank % world_local) * per:(local_rank % world_local + 1) * per] or cpus
        else:
            mine = cpus
        os.sched_setaffinity(0, mine)
        return {"pci": bus, "cpus": f"{mine[0]}-{mine[-1]}" if mine else "", "n_cpus": len(mine), "numa_node": node}
    except Exception as exc:                              # affinity is an optimisation, never a reason to fail
        return {"error": str(exc)}


def pcie_ceiling(ctx, torch, local_rank, dist, mb=512, reps=4):
    """Raw concurrent H2D + D2H of pinned buffers on two streams (all ranks at once): the node's copy ceiling that the
    end-to-end number is reported against.  Returns GB/s per direction for this rank (max-time over ranks)."""
    n = mb << 20
    h_in = ctx.pinned_empty((n,), np.uint8)
    h_out = ctx.pinned_empty((n,), np.uint8)
    d_in = torch.empty(n, dtype=torch.uint8, device=f"cuda:{local_rank}")
    d_out = torch.empty(n, dtype=torch.uint8, device=f"cuda:{local_rank}")
    s1, s2 = torch.cuda.Stream(device=local_rank), torch.cuda.Stream(device=local_rank)
    t_in = torch.from_numpy(h_in)
    t_out = torch.from_numpy(h_out)
    def once():
        with torch.cuda.stream(s1):
            d_in.copy_(t_in, non_blocking=True)
        with torch.cuda.stream(s2):
            t_out.copy_(d_out, non_blocking=True)
    once()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if dist:
        t = torch.tensor([dt], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return n * reps / dt / 1e9


# ------------------------------------------------------------------------------------------ CUDA arm
def run_b200(args, rank, world, local_rank):
    import torch
    from image_stitcher_b200 import _ffi
    from image_stitcher_b200.plate import FuseBatchPlan, FusePlan, PlateSpec, RegisterPlan, make_plate, well_fuse_tiles, well_pairs

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    affinity = None if args.no_affinity else bind_to_gpu_cpus(local_rank)     # before any pinned allocation
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout when the communicator comes up: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    from image_stitcher_b200.shard import wells_for_rank
    strong = args.scaling == "strong" and world > 1
    # weak: every rank owns a whole plate (wells rank * W .. rank * W + W - 1 of an N-plate run); strong: ONE plate,
    # well i on rank i % N (shard.wells_for_rank) -- no data-path collective either way
    my_wells = wells_for_rank(args.wells, world, rank) if strong else [rank * args.wells + w for w in range(args.wells)]
    wells_total = args.wells if strong else args.wells * world
    spec = PlateSpec(wells=len(my_wells), rows=args.grid, cols=args.grid, tile_h=args.tile, tile_w=args.tile,
                     channels=args.channels, num_z=args.num_z, reg_channel=min(1, args.channels - 1), seed=0)
    use_flat = not args.no_flatfield
    blend = _ffi.BLEND_MODES[args.blend]
    ctx = _ffi.Context(local_rank)
    plate = make_plate(spec, device=f"cuda:{local_rank}", with_flat=True, well_ids=my_wells)
    if use_flat:
        for c in range(spec.channels):
            ctx.set_flatfield(c, plate.flat[c], mem=_ffi.SB_MEM_DEVICE)
    Wc, Hc = spec.canvas_size()
    pitch = _ffi.canvas_pitch(Wc)
    planes = spec.channels * spec.num_z
    canvases = torch.empty((spec.wells, planes, Hc, pitch), dtype=torch.int16, device=f"cuda:{local_rank}")
    ovx, ovy = spec.strip_overlaps()
    # real (non-default) streams: the library's lanes launch on them and the timing events are recorded on them.
    # Registration runs on lane 0; the independent per-well fusion launches rotate over `nl` lanes so that the tail
    # of one persistent kernel overlaps the head of the next.  Lane 0's stream brackets everything (fork/join events).
    nl = max(1, min(args.fuse_lanes, ctx.num_lanes))
    streams = [torch.cuda.Stream(device=local_rank) for _ in range(nl)]
    stream = streams[0]
    torch.cuda.synchronize()
    for i, st_ in enumerate(streams):
        ctx.set_lane_stream(i, st_.cuda_stream)

    def dev_ptr(w):
        return lambda r, c, ch, z: plate.pool[w, r, c, ch, z].data_ptr()

    plans, all_pairs = [], []
    for w in range(spec.wells):
        plans.append(FusePlan(ctx, well_fuse_tiles(spec, dev_ptr(w)), (spec.tile_h, spec.tile_w),
                              (spec.channels, spec.num_z, Hc, Wc), canvases[w], tile_mem=_ffi.SB_MEM_DEVICE,
                              out_mem=_ffi.SB_MEM_DEVICE, apply_flatfield=use_flat, blend=blend, blend_ov=(ovx, ovy)))
        all_pairs += well_pairs(spec, dev_ptr(w))[0]
    n_pairs = len(all_pairs)
    px_per_step = spec.wells * planes * Hc * Wc
    batch = None if args.per_well_fusion else FuseBatchPlan(ctx, plans)
    # the pair list marshalled once, like the fuse jobs: a step is then the bare sb_register_pairs call
    reg_plan = RegisterPlan(ctx, all_pairs, (spec.tile_h, spec.tile_w), ovx, ovy, mem=_ffi.SB_MEM_DEVICE, lane=0)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    reg_ms, fuse_ms = [], []
    last_reg = None

    def step(record):
        nonlocal last_reg
        e0, e1, e2 = ev(), ev(), ev()
        e0.record(stream)
        reg_plan.run()
        e1.record(stream)
        if batch is not None:
            batch.run(0)                                 # every well of the plate in one launch, channel by channel
        else:
            for st_ in streams[1:]:
                st_.wait_event(e1)                       # fork: the other lanes start after registration
            for i, p in enumerate(plans):
                p.run(i % nl)
            for st_ in streams[1:]:                      # join: lane 0 waits for the other lanes' last launch
                j = torch.cuda.Event()
                j.record(st_)
                stream.wait_event(j)
        e2.record(stream)
        if record is not None:
            record.append((e0, e1, e2))

    sampler = ClockSampler(local_rank)
    sampler.start()
    step(None)                                           # first touch (allocations inside the library), not sampled
    torch.cuda.synchronize()
    t_load0 = time.perf_counter()
    for _ in range(args.warmup):
        step(None)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    launches0 = ctx.kernel_launches
    torch.cuda.synchronize()
    t_start, t_end = ev(), ev()
    recs = []
    wall0 = time.perf_counter()
    t_start.record(stream)
    for _ in range(args.steps):
        step(recs)
    t_end.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    launches = ctx.kernel_launches - launches0
    last_reg = reg_plan.results()
    clocks = sampler.stop(t_load0, wall0, wall0 + wall)
    total_ms = t_start.elapsed_time(t_end)
    reg_ms = [a.elapsed_time(b) for a, b, _ in recs]
    fuse_ms = [b.elapsed_time(c) for _, b, c in recs]
    px_all, pairs_all = px_per_step, n_pairs            # whole job: every rank's pixels and pairs per step
    if dist:
        t = torch.tensor([total_ms, float(np.sum(reg_ms)), float(np.sum(fuse_ms))], device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, reg_sum, fuse_sum = (float(v) for v in t.tolist())
        l = torch.tensor([launches, px_per_step, n_pairs], device=f"cuda:{local_rank}", dtype=torch.int64)
        dist.all_reduce(l)
        launches, px_all, pairs_all = (int(v) for v in l.tolist())
    else:
        reg_sum, fuse_sum = float(np.sum(reg_ms)), float(np.sum(fuse_ms))

    # parity property at full size: registration recovers the known stage drift of every well
    ok = 0
    for w in range(spec.wells):
        res = last_reg[w * (n_pairs // spec.wells):(w + 1) * (n_pairs // spec.wells)]
        kinds = well_pairs(spec, lambda *a: 0)[1]
        ok += all((r["dy"], r["dx"]) == plate.truth[w][k] for r, k in zip(res, kinds))

    # roofline of the fusion kernel: algorithmic bytes / measured launch time
    peaks = {}
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "fusion_traffic.json")
    if os.path.exists(tpath) and args.blend == "paste" and use_flat and (args.tile, args.grid, args.channels) == (2048, 3, 4):
        with open(tpath) as f:
            tj = json.load(f)
        traffic = float(tj["dram_bytes_read"]) + float(tj["dram_bytes_write"])
    wells_per_launch = spec.wells if batch is not None else 1
    n_fuse = (1 if batch is not None else spec.wells) * args.steps
    fuse_launch_ms = fuse_sum / n_fuse
    if args.blend == "paste":
        alg_bytes = 4.0 * planes * Hc * Wc                      # 2 B winning source px + 2 B written, all px covered
    else:
        alg_bytes = 2.0 * spec.tiles_per_well * spec.tile_h * spec.tile_w + 2.0 * planes * Hc * Wc
    alg_bytes *= wells_per_launch
    if traffic is not None:
        traffic *= wells_per_launch
    achieved = alg_bytes / (fuse_launch_ms * 1e-3) / 1e9
    Sh_h, Sw_h = spec.tile_h - 2 * int(spec.tile_h * 0.25), ovx
    reg_bytes = spec.wells * (spec.rows * spec.cols * 2.0 * spec.tile_h * spec.tile_w) + n_pairs * 4.0 * Sh_h * Sw_h
    reg_achieved = reg_bytes * args.steps / (reg_sum * 1e-3) / 1e9

    out = {
        "metric": METRIC, "value": px_all * args.steps / 1e6 / (total_ms * 1e-3), "unit": "Mpx/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "u16 pixels; f32 flatfield divide; f32 FFT with f64 redo of low-confidence pairs",
        "data": "synthetic",
        "config": {"workload": workload_name(spec, use_flat, args.blend, args.config, wells_total if strong else args.wells),
                   "plates": 1 if strong else world, "wells_per_rank": spec.wells,
                   "sharding": ("one plate, well i on rank i % N (shard.wells_for_rank), no data-path collective" if strong
                                else "one plate per rank, no data-path collective"),
                   "pairs_per_step": pairs_all, "canvas": [planes, Hc, Wc], "strip": [Sh_h, Sw_h],
                   "l2": "inputs (29 GB) and outputs (25 GB) per step exceed L2 (126 MB); no flush needed",
                   "timing": "CUDA events on the launching stream (lane 0 brackets the other lanes with fork/join events), max over ranks",
                   "fuse_lanes": nl},
        "tile_pairs_per_s": pairs_all * args.steps / (reg_sum * 1e-3),
        "fusion_mpx_per_s": px_all * args.steps / 1e6 / (fuse_sum * 1e-3),
        "registration_ms_per_step": reg_sum / args.steps, "fusion_ms_per_step": fuse_sum / args.steps,
        "wall_ms_per_step": wall / args.steps * 1e3,
        "registration_truth_wells_ok": f"{ok}/{spec.wells}",
        "registration_f64_redo_pairs": int(sum(r["precision"] == 1 for r in last_reg)),
        "gpu_launches": int(launches),
        "cpu_affinity": affinity,
        "clocks": clocks,
        "roofline": {"kernel": "paste_rect_kernel" if args.blend == "paste" else "paste_rect_kernel<ROUND> + blend_cells_kernel", "bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs, "traffic": traffic, "traffic_source": "profiles/fusion_traffic.json (ncu --set full, per launch)" if traffic else None,
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": fuse_launch_ms,
                     "regions_per_launch": wells_per_launch,
                     "frac_of_spec_8000": achieved / 8000.0},
        "roofline_registration": {"bound": "hbm", "achieved": reg_achieved, "peak": peak_gbs, "unit": "GB/s",
                                  "frac": reg_achieved / peak_gbs, "algorithmic_bytes_per_step": reg_bytes},
    }

    # ---------------------------------------------------------------- strong scaling: cross-rank equality on a sample
    # rank 0 regenerates the first well of the LAST rank (same global well id -> same pixels), registers and fuses it
    # itself and compares shifts and a canvas checksum with what that rank produced: the sharded plate equals the
    # single-GPU plate well by well (wells are independent; there is nothing to exchange).
    if strong:
        mine = torch.tensor([int(canvases[0].to(torch.int64).sum().item())] +
                            [v for r in last_reg[:n_pairs // spec.wells] for v in (r["dy"], r["dx"])],
                            device=f"cuda:{local_rank}", dtype=torch.int64)
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        if rank == 0:
            wid = wells_for_rank(args.wells, world, world - 1)[0]
            spec1 = PlateSpec(wells=1, rows=spec.rows, cols=spec.cols, tile_h=spec.tile_h, tile_w=spec.tile_w, channels=spec.channels,
                              num_z=spec.num_z, reg_channel=spec.reg_channel, seed=0)
            p1 = make_plate(spec1, device=f"cuda:{local_rank}", with_flat=True, well_ids=[wid])
            ptr1 = lambda r, c, ch, z: p1.pool[0, r, c, ch, z].data_ptr()
            c1 = torch.empty_like(canvases[0])
            torch.cuda.synchronize()
            r1 = ctx.register_pairs(well_pairs(spec1, ptr1)[0], (spec.tile_h, spec.tile_w), ovx, ovy, mem=_ffi.SB_MEM_DEVICE, lane=0)
            ctx.fuse_region(well_fuse_tiles(spec1, ptr1), (spec.tile_h, spec.tile_w), (spec.channels, spec.num_z, Hc, Wc), out=c1,
                            tile_mem=_ffi.SB_MEM_DEVICE, out_mem=_ffi.SB_MEM_DEVICE, apply_flatfield=use_flat, blend=blend,
                            blend_ov=(ovx, ovy), dtype=_ffi.SB_U16)
            torch.cuda.synchronize()
            exp = [int(c1.to(torch.int64).sum().item())] + [v for r in r1 for v in (r["dy"], r["dx"])]
            out["strong_scaling_check"] = {"well": wid, "owner_rank": world - 1,
                                           "equals_single_gpu_result": exp == [int(v) for v in gathered[world - 1].tolist()]}

    # ---------------------------------------------------------------- informational: the two phases overlapped
    # In this workload (coordinate placement) fusion does not depend on the shifts, so a pipeline may run the two phases
    # concurrently: registration (latency bound) on lane 0 and its auxiliary streams, fusion (DRAM bound) on the other
    # lanes.  Reported next to the contract numbers, which keep the phases sequential so that each is measured cleanly.
    if nl >= 2:
        def overlapped_step():
            e0, e1 = ev(), ev()
            e0.record(stream)
            for st_ in streams[1:]:
                st_.wait_event(e0)
            pend = ctx.register_pairs_async(all_pairs, (spec.tile_h, spec.tile_w), ovx, ovy, mem=_ffi.SB_MEM_DEVICE, lane=0)
            for i, p in enumerate(plans):
                p.run(1 + i % (nl - 1))
            for st_ in streams[1:]:
                j = torch.cuda.Event()
                j.record(st_)
                stream.wait_event(j)
            e1.record(stream)
            return pend, e0, e1
        overlapped_step()[0].get()
        torch.cuda.synchronize()
        o_ms, o_ok = [], True
        for _ in range(max(2, min(args.steps, 3))):
            pend, e0, e1 = overlapped_step()
            res_o = pend.get()
            torch.cuda.synchronize()
            o_ms.append(e0.elapsed_time(e1))
            o_ok = o_ok and [(r["dy"], r["dx"]) for r in res_o] == [(r["dy"], r["dx"]) for r in last_reg]
        if dist:
            t = torch.tensor([float(np.mean(o_ms))], device=f"cuda:{local_rank}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            o_mean = float(t.item())
        else:
            o_mean = float(np.mean(o_ms))
        out["concurrent_phases"] = {"ms_per_step": o_mean, "value": px_all / 1e6 / (o_mean * 1e-3), "unit": "Mpx/s",
                                    "same_shifts_as_sequential": bool(o_ok),
                                    "note": "registration (lane 0 + aux streams) and fusion (other lanes) enqueued together; "
                                            "informational, not the contract value (no gain on B200: the registration blocks "
                                            "hold the whole register file, so fusion blocks cannot co-reside)"}

    # ---------------------------------------------------------------- informational: coordinate-only fusion (no flat-field)
    # BASELINE.json names configs[2] "coordinate-only fusion (fusion-kernel bandwidth test)"; the contract value above keeps
    # the flat-field ON (the harder case: + 4 B of field per pixel through L2 and the exact divide).  The same launch without
    # the field is the kernel against the copy roofline alone.
    if use_flat and args.blend == "paste" and batch is not None and not args.no_coordinate_only:
        plans_nf = [FusePlan(ctx, well_fuse_tiles(spec, dev_ptr(w)), (spec.tile_h, spec.tile_w),
                             (spec.channels, spec.num_z, Hc, Wc), canvases[w], tile_mem=_ffi.SB_MEM_DEVICE,
                             out_mem=_ffi.SB_MEM_DEVICE, apply_flatfield=False, blend=blend, blend_ov=(ovx, ovy))
                    for w in range(spec.wells)]
        batch_nf = FuseBatchPlan(ctx, plans_nf)
        for _ in range(2):
            batch_nf.run(0)
        torch.cuda.synchronize()
        nf = []
        for _ in range(max(3, args.steps)):
            a, b = ev(), ev()
            a.record(stream)
            batch_nf.run(0)
            b.record(stream)
            torch.cuda.synchronize()
            nf.append(a.elapsed_time(b))
        nf_ms = float(np.mean(nf))
        if dist:
            t = torch.tensor([nf_ms], device=f"cuda:{local_rank}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            nf_ms = float(t.item())
        nf_gbs = alg_bytes / (nf_ms * 1e-3) / 1e9
        out["fusion_coordinate_only"] = {"ms_per_step": nf_ms, "fusion_mpx_per_s": px_all / 1e6 / (nf_ms * 1e-3),
                                         "achieved": nf_gbs, "unit": "GB/s", "frac": nf_gbs / peak_gbs,
                                         "frac_of_spec_8000": nf_gbs / 8000.0,
                                         "note": "same plate, same launch (sb_fuse_regions), apply_flatfield = False"}
        batch.run(0)                                     # leave the flat-field canvases in place for the checks below
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- e2e: host buffers through the public API
    if not args.no_e2e:
        from image_stitcher_b200.pipeline import WellPipeline
        hw = max(1, min(args.host_wells, spec.wells))
        pipe = WellPipeline(ctx, spec, apply_flatfield=use_flat, blend=args.blend, partial_upload={'off': False, 'boxes': True, 'rows': 'rows'}[args.e2e_partial_upload])
        host_tiles = [ctx.pinned_empty((spec.rows, spec.cols, spec.channels, spec.num_z, spec.tile_h, spec.tile_w),
                                       np.uint16) for _ in range(hw)]
        for i in range(hw):
            host_tiles[i][...] = plate.pool[i].cpu().numpy().view(np.uint16)
        host_out = [ctx.pinned_empty((1, spec.channels, spec.num_z, Hc, Wc), np.uint16) for _ in range(pipe.depth)]
        e2e_steps = args.e2e_steps if args.e2e_steps > 0 else max(1, min(args.steps, 2))
        for i in range(nl):
            ctx.set_lane_stream(i, None)

        def e2e_step():
            results = []
            for w in range(spec.wells):
                results.append(pipe.submit(host_tiles[w % hw], host_out[w % pipe.depth]))
            pipe.drain()
            return results

        e2e_step()                                       # warm-up (allocations, first-touch)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            res = e2e_step()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        if dist:
            t = torch.tensor([e2e_s], device=f"cuda:{local_rank}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        # parity inside the e2e leg: every well's asynchronously delivered shifts equal the generator's ground truth
        kinds_ = well_pairs(spec, lambda *a: 0)[1]
        e2e_reg_ok = sum(all((r["dy"], r["dx"]) == plate.truth[w % hw][k] for r, k in zip(res[w].get(), kinds_))
                         for w in range(spec.wells))
        # check one well against the device-resident result of the same well
        got = host_out[(spec.wells - 1) % pipe.depth][0].reshape(planes, Hc, Wc)
        src_w = (spec.wells - 1) % hw
        exp = canvases[src_w].cpu().numpy().view(np.uint16)[:, :, :Wc]
        # the node's raw copy ceiling with every rank copying both ways at once: what the end-to-end number is bound by
        ceil_gbs = pcie_ceiling(ctx, torch, local_rank, dist)
        h2d_full = int(spec.wells * spec.tiles_per_well * spec.tile_h * spec.tile_w * 2)
        h2d_step = int(spec.wells * pipe.upload_bytes)            # bytes actually copied: paste mode skips pixels a later tile overwrites
        d2h_step = int(px_per_step * 2)
        floor_s = max(h2d_step, d2h_step) / (ceil_gbs * 1e9)      # both directions run concurrently
        out["e2e"] = {"value": px_all * e2e_steps / 1e6 / e2e_s, "unit": "Mpx/s",
                      "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step, "steps": e2e_steps,
                      "h2d_bytes_all_tile_pixels": h2d_full,
                      "h2d_note": ("--e2e-partial-upload: only the part of each tile that can reach the canvas is uploaded (the pixels a "
                                   "later tile overwrites are never read; registration-channel tiles go up whole)"
                                   if h2d_step != h2d_full else "every tile pixel is uploaded (one contiguous copy per well)"),
                      "ms_per_step": e2e_s / e2e_steps * 1e3,
                      "pcie_ceiling": {"gb_per_s_per_direction_per_gpu": ceil_gbs, "ranks_copying_at_once": world,
                                       "ms_per_step_at_ceiling": floor_s * 1e3,
                                       "e2e_frac_of_ceiling": floor_s / (e2e_s / e2e_steps),
                                       "how": "512 MiB pinned H2D and D2H on two streams at once, every rank at the same time, max over ranks"},
                      "tile_pairs_per_s": pairs_all * e2e_steps / e2e_s,
                      "api": "WellPipeline.submit(pinned host tiles) -> host canvas + shifts (sb_memcpy_async / sb_memcpy2d_async + "
                             "sb_register_pairs_async + sb_fuse_region, rotating over 3 lanes)",
                      "registration_truth_wells_ok": f"{e2e_reg_ok}/{spec.wells}",
                      "host_wells_distinct": hw, "matches_device_result": bool(np.array_equal(got, exp))}
    else:
        out["e2e"] = None

    # ---------------------------------------------------------------- float64 registration (informational)
    # the reference's arithmetic is complex128; north_star licenses float32 through its tolerance (bit-exact integer
    # shifts, verified above on every well) -- this is the same-arithmetic figure next to it
    if not args.no_f64:
        for i, st_ in enumerate(streams):
            ctx.set_lane_stream(i, st_.cuda_stream)
        e0, e1 = ev(), ev()
        ctx.register_pairs(all_pairs[:min(len(all_pairs), 96)], (spec.tile_h, spec.tile_w), ovx, ovy, mem=_ffi.SB_MEM_DEVICE,
                           lane=0, precision=_ffi.SB_PREC_F64)                          # warm-up (workspace, twiddles)
        e0.record(stream)
        res64 = ctx.register_pairs(all_pairs, (spec.tile_h, spec.tile_w), ovx, ovy, mem=_ffi.SB_MEM_DEVICE, lane=0,
                                   precision=_ffi.SB_PREC_F64)
        e1.record(stream)
        torch.cuda.synchronize()
        ms64 = e0.elapsed_time(e1)
        same = [(r["dy"], r["dx"], r["coarse"], r["fine"]) for r in res64] == [(r["dy"], r["dx"], r["coarse"], r["fine"]) for r in last_reg]
        out["registration_f64"] = {"ms_per_step": ms64, "tile_pairs_per_s": n_pairs / (ms64 * 1e-3),
                                   "same_indices_and_shifts_as_default_precision": bool(same),
                                   "note": "SB_PREC_F64: the radix FFT engine in complex128 on the GPU (this rank's pairs)"}
        for i in range(nl):
            ctx.set_lane_stream(i, None)

    # ---------------------------------------------------------------- CPU baseline (rank 0, N = 1 only)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)                          # the CPU arm gets every core again
        cores = host_cores()
        n = args.cpu_sample_wells if args.cpu_sample_wells > 0 else cores
        distinct = max(1, min(n, 4))
        tiles = plate.pool[:distinct].cpu().numpy().view(np.uint16)
        flat = plate.flat.cpu().numpy()
        p, x, wall_s, a, b = cpu_pool_sample(spec, tiles, flat, use_flat, n, cores)
        share = a / (a + b)
        out["cpu_baseline"] = {"value": x / 1e6 / wall_s, "unit": "Mpx/s", "cores": cores, "kind": "port",
                               "sample": f"{n} of {spec.wells} wells ({p} pairs, {x / 1e6:.0f} Mpx), one well per worker "
                                         f"process over {cores} cores, arrays in RAM; NumPy/SciPy oracle port of the "
                                         "reference (validated against the reference by tests/golden)",
                               "tile_pairs_per_s": p / (wall_s * share), "fusion_mpx_per_s": x / 1e6 / (wall_s * (1 - share)),
                               "seconds": wall_s, "cpu_seconds": a + b}
    if rank == 0:
        print(json.dumps(out), flush=True)
    ctx.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def run_mosaic(args, rank, world, local_rank):
    """--config 4 --scaling strong: ONE large mosaic (20 x 20 tiles of 3000^2, 5 z-planes) split over the ranks the way
    SURVEY 8e prescribes -- registration by grid-row bands of the pair list (shard.mosaic_pairs_for_rank), fusion by
    (plane, chunk-row) bands of the canvas (shard.fusion_units_for_rank + tiles_for_band), no data-path collective: the
    shifts are all-gathered after the timed region only to be checked.  Every rank generates the same mosaic (seeded);
    rank 0 also fuses plane 0 in ONE call and compares it band by band with what the owners produced.  Works at N = 1
    (all bands on one GPU) -- the base of the strong-scaling ratio."""
    import torch
    from image_stitcher_b200 import _ffi, shard
    from image_stitcher_b200 import geometry as geo
    from image_stitcher_b200.plate import FusePlan, PlateSpec, RegisterPlan, make_plate, well_fuse_tiles

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    affinity = None if args.no_affinity else bind_to_gpu_cpus(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)                                        # NCCL's banner must not land on the JSON line's stdout
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    spec = PlateSpec(wells=1, rows=args.grid, cols=args.grid, tile_h=args.tile, tile_w=args.tile, channels=args.channels,
                     num_z=args.num_z, reg_channel=min(1, args.channels - 1), seed=0)
    use_flat = not args.no_flatfield
    ctx = _ffi.Context(local_rank)
    plate = make_plate(spec, device=dev, with_flat=use_flat, well_ids=[0])
    if use_flat:
        for c in range(spec.channels):
            ctx.set_flatfield(c, plate.flat[c], mem=_ffi.SB_MEM_DEVICE)
    Wc, Hc = spec.canvas_size()
    pitch = _ffi.canvas_pitch(Wc)
    planes = spec.channels * spec.num_z
    ovx, ovy = spec.strip_overlaps()
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.synchronize()
    ctx.set_lane_stream(0, stream.cuda_stream)
    ptr = lambda r, c, ch, z: plate.pool[0, r, c, ch, z].data_ptr()

    # ---- this rank's pairs: grid-row bands of the mosaic's pair list
    mine = shard.mosaic_pairs_for_rank(spec.rows, spec.cols, world, rank)
    reg = [(ptr(r0, c0, spec.reg_channel, 0), ptr(r1, c1, spec.reg_channel, 0),
            _ffi.SB_DIR_HORIZONTAL if kind == "h" else _ffi.SB_DIR_VERTICAL) for kind, (r0, c0), (r1, c1) in mine]
    reg_plan = RegisterPlan(ctx, reg, (spec.tile_h, spec.tile_w), ovx, ovy, mem=_ffi.SB_MEM_DEVICE, lane=0) if reg else None
    # ---- this rank's output bands: (plane, chunk-row) units, each fused from the tiles that reach its rows
    chunk_h = 2048
    tiles = well_fuse_tiles(spec, ptr)
    units = shard.fusion_units_for_rank(planes, Hc, chunk_h, world, rank)
    bands, plans = [], []
    for plane, y0, y1 in units:
        c, z = divmod(plane, spec.num_z)
        sub = [t for t in tiles if t[3] == c and t[4] == z]
        band_tiles = [(t[0], t[1], t[2], 0, 0, *t[5:]) for t in shard.tiles_for_band(sub, spec.tile_h, y0, y1)]
        out = torch.empty((1, y1 - y0, pitch), dtype=torch.int16, device=dev)
        bands.append(out)
        # (a band of channel c fuses as channel 0 of a one-plane job: the field slot is selected per unit below)
        plans.append((c, FusePlan(ctx, band_tiles, (spec.tile_h, spec.tile_w), (1, 1, y1 - y0, Wc), out, tile_mem=_ffi.SB_MEM_DEVICE,
                                  out_mem=_ffi.SB_MEM_DEVICE, apply_flatfield=use_flat and c == 0)))
    if use_flat and spec.channels > 1:
        raise RuntimeError("bench.py: the mosaic split with a flat-field is wired for one channel (BASELINE configs[4] has none)")
    px_mine = sum((y1 - y0) * Wc for _, y0, y1 in units)

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(rec):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record(stream)
        if reg_plan is not None:
            reg_plan.run()
        e1.record(stream)
        for _, pl in plans:
            pl.run(0)
        e2.record(stream)
        if rec is not None:
            rec.append((e0, e1, e2))

    sampler = ClockSampler(local_rank)
    sampler.start()
    step(None)
    torch.cuda.synchronize()
    t_load0 = time.perf_counter()
    for _ in range(args.warmup):
        step(None)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    launches0 = ctx.kernel_launches
    torch.cuda.synchronize()
    t_start, t_end, recs = ev(), ev(), []
    wall0 = time.perf_counter()
    t_start.record(stream)
    for _ in range(args.steps):
        step(recs)
    t_end.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    launches = ctx.kernel_launches - launches0
    clocks = sampler.stop(t_load0, wall0, wall0 + wall)
    total_ms = t_start.elapsed_time(t_end)
    reg_sum = float(sum(a.elapsed_time(b) for a, b, _ in recs))
    fuse_sum = float(sum(b.elapsed_time(c) for _, b, c in recs))
    px_all, pairs_all = px_mine, len(reg)
    if dist:
        t = torch.tensor([total_ms, reg_sum, fuse_sum], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, reg_sum, fuse_sum = (float(v) for v in t.tolist())
        l = torch.tensor([launches, px_mine, len(reg)], device=dev, dtype=torch.int64)
        dist.all_reduce(l)
        launches, px_all, pairs_all = (int(v) for v in l.tolist())

    # ---- checks (outside the timed region): every pair's shift equals the generator's drift; plane 0 of the sharded
    # canvas equals the canvas ONE sb_fuse_region call produces on rank 0, band by band (sum + xor-rotate checksums)
    res = reg_plan.results() if reg_plan is not None else []
    shifts_ok = all((r["dy"], r["dx"]) == plate.truth[0][kind] for r, (kind, _, _) in zip(res, mine))

    def checksum(tensor2d):
        v = tensor2d[:, :Wc].to(torch.int64) & 0xFFFF
        w = torch.arange(1, v.shape[1] + 1, device=v.device, dtype=torch.int64)
        return int(v.sum().item()), int((v * w).sum().item() & 0x7FFFFFFFFFFFFFFF)

    my_sums = [(plane, y0, *checksum(b[0])) for (plane, y0, y1), b in zip(units, bands) if plane == 0]
    ok_t = torch.tensor([int(shifts_ok)], device=dev, dtype=torch.int64)
    gathered = None
    if dist:
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        gathered = [None] * world
        dist.all_gather_object(gathered, my_sums)
    else:
        gathered = [my_sums]
    canvas_ok = None
    if rank == 0:
        sub0 = [(t[0], t[1], t[2], 0, 0, *t[5:]) for t in tiles if t[3] == 0 and t[4] == 0]
        whole = torch.empty((1, Hc, pitch), dtype=torch.int16, device=dev)
        torch.cuda.synchronize()
        ctx.fuse_region(sub0, (spec.tile_h, spec.tile_w), (1, 1, Hc, Wc), out=whole, tile_mem=_ffi.SB_MEM_DEVICE,
                        out_mem=_ffi.SB_MEM_DEVICE, apply_flatfield=use_flat, dtype=_ffi.SB_U16)
        torch.cuda.synchronize()
        n_bands, canvas_ok = 0, True
        for part in gathered:
            for plane, y0, s1, s2 in part:
                y1 = min(y0 + chunk_h, Hc)
                canvas_ok = canvas_ok and checksum(whole[0, y0:y1]) == (s1, s2)
                n_bands += 1
        canvas_ok = bool(canvas_ok and n_bands == -(-Hc // chunk_h))
        del whole

    peaks = {}
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = 4.0 * px_all                                   # 2 B winning source px + 2 B written (all ranks)
    achieved = alg_bytes * args.steps / (fuse_sum * 1e-3) / 1e9
    out = {
        "metric": METRIC, "value": px_all * args.steps / 1e6 / (total_ms * 1e-3), "unit": "Mpx/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": "u16 pixels; f32 FFT with f64 redo of low-confidence pairs", "data": "synthetic",
        "config": {"workload": workload_name(spec, use_flat, args.blend, args.config),
                   "sharding": "ONE mosaic: pairs by grid-row bands (shard.mosaic_pairs_for_rank), canvas by (plane, chunk-row) "
                               f"bands of {chunk_h} rows (shard.fusion_units_for_rank + tiles_for_band); no data-path collective",
                   "pairs_per_step": pairs_all, "canvas": [planes, Hc, Wc], "bands": planes * -(-Hc // chunk_h),
                   "l2": "inputs (36 GB) and outputs (29 GB) per step exceed L2 (126 MB); no flush needed",
                   "timing": "CUDA events on the launching stream, max over ranks"},
        "tile_pairs_per_s": pairs_all * args.steps / (reg_sum * 1e-3) if reg_sum else None,
        "fusion_mpx_per_s": px_all * args.steps / 1e6 / (fuse_sum * 1e-3),
        "registration_ms_per_step": reg_sum / args.steps, "fusion_ms_per_step": fuse_sum / args.steps,
        "wall_ms_per_step": wall / args.steps * 1e3,
        "registration_truth_all_pairs_ok": bool(int(ok_t.item())),
        "plane0_equals_single_call_canvas": canvas_ok,
        "gpu_launches": int(launches), "cpu_affinity": affinity, "clocks": clocks,
        "roofline": {"kernel": "paste_rect_kernel", "bound": "hbm", "achieved": achieved, "peak": peak_gbs * world, "unit": "GB/s",
                     "frac": achieved / (peak_gbs * world), "traffic": None,
                     "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs) x ranks" if "hbm_gbs" in peaks else "fallback 6650 GB/s x ranks",
                     "algorithmic_bytes_per_step": alg_bytes},
        "e2e": None,
    }
    if rank == 0:
        print(json.dumps(out), flush=True)
    ctx.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.scaling == "strong" and args.wells == 1:
        run_mosaic(args, rank, world, local_rank)         # one region: split inside the mosaic (bands), not by wells
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
