"""Per-role cycle counters of the tensor-core registration kernels (a -DSB_TC_PROFILE build of the library):

    SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_tcprof.so python scratch/tc_profile.py

Counters of block 0, per mode (0 forward, 1 inverse + argmax, 2 upsampled-DFT rows):
 0 loader wait stg_empty | 1 converter wait stg_full | 2 converter wait ab_empty | 3 MMA wait acc_empty | 4 MMA wait ab_full
 5 MMA wait ab_empty(prev, before the B refill) | 6 epilogue wait acc_full | 7 kernel cycles | 8 tiles | 9 launches
 10 converter compute | 11 converter fence + arrive"""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from image_stitcher_b200 import _ffi
from image_stitcher_b200.plate import PlateSpec, make_plate, well_pairs

spec = PlateSpec(wells=int(os.environ.get('TCP_WELLS', '18')), rows=3, cols=3, tile_h=2048, tile_w=2048, channels=1, reg_channel=0, jitter=3, seed=1)
plate = make_plate(spec, device="cuda:0", with_flat=False)
ctx = _ffi.Context(0)
pairs = []
for w in range(spec.wells):
    pairs += well_pairs(spec, lambda r, c, ch, z, w=w: plate.pool[w, r, c, ch, z].data_ptr())[0]
ovx, ovy = spec.strip_overlaps()
torch.cuda.synchronize()
names = ["ld_wait_stg_empty", "cv_wait_stg_full", "cv_wait_ab_empty", "mma_wait_acc_empty", "mma_wait_ab_full", "mma_wait_ab_empty_prev",
         "ep_wait_acc_full", "kernel", "tiles", "launches", "cv_compute", "cv_fence_arrive"]
for rep in range(2):
    ctx.register_pairs(pairs, (2048, 2048), ovx, ovy, mem=_ffi.SB_MEM_DEVICE, precision=0)
    buf = (C.c_longlong * 48)()
    n = ctx.lib.sb_debug_tc_profile(ctx.handle, buf)
    if n == 0:
        print("not a profile build")
        break
    a = np.array(buf[:]).reshape(3, 16)
    for mode in range(3):
        if a[mode, 9] == 0:
            continue
        tiles, launches = a[mode, 8], a[mode, 9]
        print(f"rep {rep} mode {mode}: launches {launches} tiles(block 0) {tiles} kernel cycles/tile {a[mode, 7] / tiles:.0f}")
        for i, nm in enumerate(names):
            if i in (7, 8, 9):
                continue
            print(f"     {nm:24s} {a[mode, i] / tiles:10.0f} cycles / tile")
ctx.close()
