#include <cstdint>
__device__ __forceinline__ uint64_t pk(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t add2rz(uint64_t a, uint64_t b) { uint64_t d; asm("add.rz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__global__ void k(const float* a, const float* b, float* o) {
    int i = threadIdx.x * 2;
    uint64_t A = pk(a[i], a[i+1]), B = pk(b[i], b[i+1]);
    float r0, r1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(b[i]));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(b[i+1]));
    uint64_t R = pk(r0, r1), ONE = pk(1.f, 1.f), NB = pk(-b[i], -b[i+1]);
    uint64_t E = fma2(NB, R, ONE); R = fma2(R, E, R);
    uint64_t Q = mul2(A, R); uint64_t REM = fma2(NB, Q, A); Q = fma2(R, REM, Q);
    Q = add2rz(Q, pk(8388608.f, 8388608.f));
    Q = add2(Q, A);
    unpk(Q, o[i], o[i+1]);
}
