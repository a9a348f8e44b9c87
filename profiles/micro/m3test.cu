#include <cuda_runtime.h>
typedef unsigned long long u64;
struct F2 { u64 v; };
__device__ __forceinline__ F2 f2(float lo, float hi) { F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float f2_lo(F2 a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); return lo; }
__device__ __forceinline__ float f2_hi(F2 a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); return hi; }
static __constant__ float k_one[2] = {1.0f, 1.0f};
__device__ __forceinline__ F2 f2_mul(F2 a, F2 b) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 f2_add(F2 a, F2 b) { F2 r; F2 one = f2(k_one[0], k_one[0]); asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(one.v), "l"(b.v)); return r; }
struct M3 { float m[9]; };
__device__ __forceinline__ M3 m3_mul2(const M3 &a, const M3 &b) {
	M3 r;
	F2 B0 = f2(b.m[0], b.m[1]), B1 = f2(b.m[3], b.m[4]), B2 = f2(b.m[6], b.m[7]);
#pragma unroll
	for (int i = 0; i < 3; i++) {
		F2 s = f2_add(f2_add(f2_mul(B0, f2(a.m[3 * i], a.m[3 * i])), f2_mul(B1, f2(a.m[3 * i + 1], a.m[3 * i + 1]))), f2_mul(B2, f2(a.m[3 * i + 2], a.m[3 * i + 2])));
		r.m[3 * i] = f2_lo(s);
		r.m[3 * i + 1] = f2_hi(s);
		r.m[3 * i + 2] = __fadd_rn(__fadd_rn(__fmul_rn(b.m[2], a.m[3 * i]), __fmul_rn(b.m[5], a.m[3 * i + 1])), __fmul_rn(b.m[8], a.m[3 * i + 2]));
	}
	return r;
}
__global__ void k(const float *in, float *out) {
	M3 a, b;
	for (int i = 0; i < 9; i++) { a.m[i] = in[threadIdx.x * 18 + i]; b.m[i] = in[threadIdx.x * 18 + 9 + i]; }
	M3 c = m3_mul2(a, b);
	M3 d = m3_mul2(c, a);
	M3 e = m3_mul2(d, c);
	for (int i = 0; i < 9; i++) out[threadIdx.x * 9 + i] = e.m[i];
}
