"""Throughput of sb_estimate_flatfield on a reference-sized sample: 32 tiles of 2048 x 2048 uint16 resident on the
device, field written to the device.  Algorithmic bytes = the one pass over the sample (n * H * W * 2) + the field."""
import ctypes as C
import json
import time

import torch

from image_stitcher_b200 import _ffi

ctx = _ffi.Context(0)
n, h, w = 32, 2048, 2048
tiles = torch.randint(200, 4000, (n, h, w), dtype=torch.int32, device="cuda").to(torch.uint16)
field = torch.empty((h, w), dtype=torch.float32, device="cuda")
ptrs = (C.c_void_p * n)(*[tiles[i].data_ptr() for i in range(n)])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
times = []
for it in range(10):
    flush.fill_(it)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rc = ctx.lib.sb_estimate_flatfield(ctx.handle, ptrs, n, h, w, _ffi.SB_U16, _ffi.SB_MEM_DEVICE, 128, 2.0,
                                       C.c_void_p(field.data_ptr()), _ffi.SB_MEM_DEVICE)
    assert rc == 0, ctx.lib.sb_last_error(ctx.handle)
    times.append(time.perf_counter() - t0)
times = sorted(times[3:])
med = times[len(times) // 2]
alg = n * h * w * 2 + h * w * 4
print(json.dumps({"what": "sb_estimate_flatfield, 32 x 2048^2 u16 on the device, grid 128, host-timed whole call (7 launches + sync)",
                  "ms": round(med * 1e3, 3), "algorithmic_GB": round(alg / 1e9, 3), "GB_per_s": round(alg / med / 1e9, 1),
                  "field_mean": round(float(field.mean()), 6)}))
