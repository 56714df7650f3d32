import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from image_stitcher_b200 import _ffi
ctx = _ffi.Context(0)
H = W = 2048; rows = cols = 3; C = 4; step = 1843
Hc = Wc = W + (cols - 1) * step
pitch = _ffi.canvas_pitch(Wc)
n = rows * cols * C
pool = torch.randint(0, 30000, (n, H, W), dtype=torch.int16, device="cuda")
out = torch.empty((C, Hc, pitch), dtype=torch.int16, device="cuda")
job = []
i = 0
for r in range(rows):
    for c in range(cols):
        for ch in range(C):
            job.append((pool[i].data_ptr(), c * step, r * step, ch, 0, 0, 0, 0, 0)); i += 1
stream = torch.cuda.Stream(); torch.cuda.synchronize()
ctx.set_lane_stream(0, stream.cuda_stream)
def run(flat, blend, reps=10):
    ctx.clear_fields()
    if flat:
        ff = np.random.default_rng(0).uniform(0.7, 1.1, (H, W)).astype(np.float32)
        for ch in range(C): ctx.set_flatfield(ch, ff)
    for _ in range(3):
        ctx.fuse_region(job, (H, W), (C, 1, Hc, Wc), out=out, tile_mem=1, out_mem=1, lane=0, apply_flatfield=flat, blend=blend, blend_ov=(205, 205))
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(reps):
        ctx.fuse_region(job, (H, W), (C, 1, Hc, Wc), out=out, tile_mem=1, out_mem=1, lane=0, apply_flatfield=flat, blend=blend, blend_ov=(205, 205))
    e1.record(stream); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    ms = e0.elapsed_time(e1) / reps
    px = C * Hc * Wc
    alg = 4 * px if blend == 0 else 2 * n * H * W + 2 * px
    print(f"flat={flat} blend={blend}: {ms:.3f} ms/well (wall {wall:.3f})  {px/ms/1e3:.1f} Mpx/s  alg {alg/ms/1e6:.0f} GB/s  frac_of_6454={alg/ms/1e6/6454.6:.3f}")
for flat in (False, True):
    for blend in (0, 1, 2):
        run(flat, blend)
