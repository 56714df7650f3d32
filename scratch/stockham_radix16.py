"""Round-2 groundwork: NumPy emulation of the Stockham pass of fft.cuh with a radix-16 butterfly built from two radix-4
stages, to pin the index / twiddle algebra before it is written in CUDA (order 4 . 16 . 16 for
the 1024-point column FFT).

A pass with radix R over a line of N points, Ns = product of the radices already applied, M = N / R:
    for j in [0, M):  k = j % Ns;  x_r = in[j + r M] * W_N^(r k N / (Ns R))          r = 0 .. R-1
                      out[(j // Ns) Ns R + k + q Ns] = sum_r x_r W_R^(r q)            q = 0 .. R-1
Radix 16 in registers, r = 4 r1 + r2, q = q1 + 4 q2:
    y[r2][q1] = sum_r1 x[4 r1 + r2] W_4^(r1 q1)          (four radix-4 butterflies)
    y[r2][q1] *= W_16^(r2 q1)                              (9 non-trivial constants)
    X[q1 + 4 q2] = sum_r2 y[r2][q1] W_4^(r2 q2)           (four radix-4 butterflies)
"""
import numpy as np


def dft4(a, b, c, d, sign):
    """4-point DFT of (a, b, c, d); sign = -1 forward, +1 inverse.  Returns outputs q = 0..3."""
    s02, d02, s13, d13 = a + c, a - c, b + d, b - d
    i_d13 = 1j * sign * d13
    return s02 + s13, d02 + i_d13, s02 - s13, d02 - i_d13


def pass_radix(x, N, Ns, R, sign):
    M = N // R
    tw = np.exp(sign * 2j * np.pi * np.arange(N) / N)
    out = np.empty_like(x)
    tstep = N // (Ns * R)
    for j in range(M):
        k = j % Ns
        xin = [x[j + r * M] * tw[(r * k * tstep) % N] for r in range(R)]
        base = (j // Ns) * Ns * R + k
        if R == 4:
            X = dft4(*xin, sign)
        elif R == 16:
            w16 = np.exp(sign * 2j * np.pi * np.arange(16) / 16)
            y = [dft4(xin[r2], xin[4 + r2], xin[8 + r2], xin[12 + r2], sign) for r2 in range(4)]     # y[r2][q1]
            X = [None] * 16
            for q1 in range(4):
                z = [y[r2][q1] * w16[(r2 * q1) % 16] for r2 in range(4)]
                o = dft4(*z, sign)
                for q2 in range(4):
                    X[q1 + 4 * q2] = o[q2]
        else:
            raise ValueError(R)
        for q in range(R):
            out[base + q * Ns] = X[q]
    return out


def fft_plan(x, radices, sign=-1):
    N, Ns = len(x), 1
    for R in radices:
        x = pass_radix(x, N, Ns, R, sign)
        Ns *= R
    return x


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    x = rng.normal(size=1024) + 1j * rng.normal(size=1024)
    for plan in ([4, 4, 4, 4, 4], [4, 16, 16], [16, 16, 4], [16, 4, 16]):
        f = fft_plan(x.copy(), plan, -1)
        b = fft_plan(f.copy(), plan, +1) / 1024
        print(plan, "forward err", float(np.abs(f - np.fft.fft(x)).max()), "round trip err", float(np.abs(b - x).max()))
    # shared-memory wavefronts of the strided writes (64-bit elements, conflicts counted per half-warp of 16 lanes)
    def write_wavefronts(N, Ns, R):
        M, total = N // R, 0
        for j0 in range(0, M, 16):
            for q in range(R):
                banks = {}
                for j in range(j0, min(j0 + 16, M)):
                    idx = (j // Ns) * Ns * R + j % Ns + q * Ns
                    banks.setdefault((idx * 2) % 32, set()).add(idx)
                total += max(len(v) for v in banks.values())
        return total
    for plan in ([4, 4, 4, 4, 4], [4, 16, 16], [16, 16, 4]):
        Ns, w = 1, []
        for R in plan:
            w.append(write_wavefronts(1024, Ns, R))
            Ns *= R
        print(plan, "write wavefronts per line and pass (conflict-free = 64):", w, "reads: 64 each")
