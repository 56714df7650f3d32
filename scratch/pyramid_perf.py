"""Throughput of sb_pyramid on one plate-well canvas (4 planes of 5734 x 5734 uint16), device to device.
Algorithmic bytes per level: 2 * (h/2 * w  +  h/2 * w/2) per plane (the even rows read, a quarter written)."""
import json
import time

import torch

from image_stitcher_b200 import _ffi

ctx = _ffi.Context(0)
planes, h, w, levels = 4, 5734, 5734, 5
pitch = int(ctx.lib.sb_canvas_pitch(w))
src = torch.randint(0, 65535, (planes, h, pitch), dtype=torch.int32, device="cuda").to(torch.uint16)
total = int(ctx.lib.sb_pyramid_elems(planes, h, w, levels))
out = torch.empty(total, dtype=torch.uint16, device="cuda")
torch.cuda.synchronize()
alg = 0
hh, ww = h, w
for _ in range(1, levels):
    dh, dw = (hh + 1) // 2, (ww + 1) // 2
    alg += 2 * planes * (dh * ww + dh * dw)
    hh, ww = dh, dw
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
times = []
for it in range(13):
    flush.fill_(it)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.pyramid((planes, h, w), levels, src=src, src_mem=_ffi.SB_MEM_DEVICE, src_row_pitch=pitch, dtype=_ffi.SB_U16,
                out=out, out_mem=_ffi.SB_MEM_DEVICE, lane=0)
    ctx.sync(0)
    times.append(time.perf_counter() - t0)
times = sorted(times[3:])
med = times[len(times) // 2]
# correctness at full size against torch slicing
lvl1 = out[:planes * ((h + 1) // 2) * ((w + 1) // 2)].view(planes, (h + 1) // 2, (w + 1) // 2)
ok = bool(torch.equal(lvl1, src[:, ::2, :w:2]))
print(json.dumps({"what": "sb_pyramid 4x5734x5734 u16, levels 1-4, device->device, host-timed incl. launch+sync",
                  "ms": round(med * 1e3, 4), "algorithmic_GB": round(alg / 1e9, 4),
                  "GB_per_s": round(alg / med / 1e9, 1), "level1_equals_slicing": ok}))
