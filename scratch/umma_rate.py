"""tcgen05.mma issue-rate probe (sb_selftest, SB_SELFTEST_UMMA, arg >= 1000): cycles per kind::tf32 MMA of 128 x N x 8 in
the product's pattern, for the no-swizzle K-major operand layout the kernels use and for a 128-byte-swizzled one."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_stitcher_b200 import _ffi
ctx = _ffi.Context(0)
for cm in (0, 1):
    for layout in (0, 1):
        for two in (0, 1):
            for n in (32, 112):
                for rep in range(2):
                    n_mma, cyc, issue, flag = ctx.selftest(2, 1000 + n + 1000 * layout + 10000 * two + 100000 * cm)
                print(f"commit/round {cm} layout {layout} per_round {'2' if two else '6'} N {n:3d}: {cyc / n_mma:7.1f} cycles / MMA  (issue loop alone {issue / n_mma:6.1f}; flag {flag:#x})")
ctx.close()
