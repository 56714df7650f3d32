import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from image_stitcher_b200 import _ffi
from image_stitcher_b200.plate import FusePlan
ctx = _ffi.Context(0)
H = W = 2048; rows = cols = 3; C = 4; step = 1843
Hc = Wc = W + (cols - 1) * step
pitch = _ffi.canvas_pitch(Wc)
n = rows * cols * C
NW = 4
pools = [torch.randint(0, 30000, (n, H, W), dtype=torch.int16, device="cuda") for _ in range(NW)]
outs = [torch.empty((C, Hc, pitch), dtype=torch.int16, device="cuda") for _ in range(NW)]
def job(pool):
    j = []; i = 0
    for r in range(rows):
        for c in range(cols):
            for ch in range(C):
                j.append((pool[i].data_ptr(), c * step, r * step, ch, 0, 0, 0, 0, 0)); i += 1
    return j
stream = torch.cuda.Stream(); torch.cuda.synchronize()
ctx.set_lane_stream(0, stream.cuda_stream)
modes = [(False, 0), (True, 0)] if len(sys.argv) > 1 and sys.argv[1] == "paste" else [(f, b) for f in (False, True) for b in (0, 1, 2)]
for flat, blend in modes:
    ctx.clear_fields()
    if flat:
        ff = np.random.default_rng(0).uniform(0.7, 1.1, (H, W)).astype(np.float32)
        for ch in range(C): ctx.set_flatfield(ch, ff)
    plans = [FusePlan(ctx, job(pools[w]), (H, W), (C, 1, Hc, Wc), outs[w], tile_mem=1, out_mem=1, apply_flatfield=flat, blend=blend, blend_ov=(205, 205)) for w in range(NW)]
    for p in plans: p.run(0)
    torch.cuda.synchronize()
    reps = 5
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(stream)
    for _ in range(reps):
        for p in plans: p.run(0)
    t_host = time.perf_counter() - t0
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * NW)
    px = C * Hc * Wc
    alg = 4 * px if blend == 0 else 2 * n * H * W + 2 * px
    print(f"flat={flat} blend={blend}: {ms:.3f} ms/well (host enqueue {t_host/(reps*NW)*1e3:.3f} ms)  {px/ms/1e3:.0f} Mpx/s  alg {alg/ms/1e6:.0f} GB/s  frac={alg/ms/1e6/6454.6:.3f}")
