import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
from conftest import load_golden, SMALL_GOLDENS
from test_fuse_gpu import golden_job
from image_stitcher_b200 import _ffi
ctx = _ffi.Context(0)
for name in SMALL_GOLDENS:
    g, st, tiles, kw = load_golden(name)
    job, cshape = golden_job(g, st, tiles)
    ctx.clear_fields()
    for c, ff in st.flatfields.items():
        ctx.set_flatfield(c, ff)
    out = np.full((1,) + cshape, 0xABCD, np.uint16)
    ctx.fuse_region(job, (st.tile_h, st.tile_w), cshape, out=out, apply_flatfield=st.apply_flatfield)
    exp = g["canvas"]
    bad = np.argwhere(out != exp)
    print(name, "mismatches", len(bad), "of", out.size, "flat dtypes", {c: f.dtype for c, f in st.flatfields.items()})
    if len(bad):
        print(" first:", bad[:8].tolist())
        for b in bad[:8]:
            print("   got", out[tuple(b)], "exp", exp[tuple(b)])
        ys = np.unique(bad[:, 3]); xs = np.unique(bad[:, 4])
        print(" rows", ys[:10], "...", ys[-5:], " cols", xs[:12], "...", xs[-5:], "planes", np.unique(bad[:, 1]), np.unique(bad[:, 2]))
        d = out.astype(int) - exp.astype(int)
        print(" diff stats: min", d.min(), "max", d.max(), "abs==1 frac", (np.abs(d[d != 0]) == 1).mean())
