// Issue rate of the packed FP32x2 instruction forms the solve kernel uses (one SM-filling grid, independent chains).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
static __constant__ float k_unit[2] = {1.0f, -0.0f};

// MODE 0: FMUL2 reg x reg      1: FMUL2 reg x broadcast scalar reg   2: FFMA2 reg, reg, reg
//      3: FFMA2 reg, const-one (uniform), reg  [the add form]        4: scalar FMUL        5: alternating FMUL2 / scalar FADD
//      6: alternating FMUL2 / FFMA2-add                                7: FFMA2 with const -0 addend [mul as fma]
template <int MODE>
__global__ void bench(float *out, int iters, float a, float b) {
	u64 p[8];
	float s[8];
	for (int i = 0; i < 8; i++) { p[i] = pack(threadIdx.x * 0.001f + i, 1.0f + i); s[i] = i + threadIdx.x; }
	const u64 pa = pack(a, b), one = pack(k_unit[0], k_unit[0]), nz = pack(k_unit[1], k_unit[1]);
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int i = 0; i < 8; i++) {
			if (MODE == 0) p[i] = mul2(p[i], pa);
			if (MODE == 1) p[i] = mul2(p[i], pack(a, a));
			if (MODE == 2) p[i] = fma2(p[i], pa, p[(i + 1) & 7]);
			if (MODE == 3) p[i] = fma2(p[i], one, pa);
			if (MODE == 4) { s[i] = __fmul_rn(s[i], a); }
			if (MODE == 5) { p[i] = mul2(p[i], pa); s[i] = __fadd_rn(s[i], a); }
			if (MODE == 6) { p[i] = mul2(p[i], pa); p[i] = fma2(p[i], one, pa); }
			if (MODE == 7) p[i] = fma2(p[i], pa, nz);
		}
	}
	float acc = 0;
	for (int i = 0; i < 8; i++) { float lo, hi; unpack(p[i], lo, hi); acc += lo + hi + s[i]; }
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE>
void run(const char *name, int instr_per_iter) {
	const int iters = 20000;
	float *o; cudaMalloc(&o, 148 * 4 * 512 * sizeof(float));
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	bench<MODE><<<148 * 4, 512>>>(o, 16, 1.0001f, 0.9999f);
	cudaEventRecord(e0);
	bench<MODE><<<148 * 4, 512>>>(o, iters, 1.0001f, 0.9999f);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1); cudaFree(o);
	double instr_per_smsp = 16.0 * instr_per_iter * iters; // 64 warps per SM = 16 per scheduler
	printf("%-52s %7.3f ms   %.3f warp-instr / clk / scheduler (1.965 GHz)\n", name, ms, instr_per_smsp / (ms * 1e-3 * 1.965e9));
}
int main() {
	run<4>("scalar FMUL", 8);
	run<0>("FMUL2 reg x reg", 8);
	run<1>("FMUL2 reg x broadcast", 8);
	run<2>("FFMA2 reg, reg, reg", 8);
	run<3>("FFMA2 reg, uniform one, reg (packed add)", 8);
	run<7>("FFMA2 reg, reg, uniform -0 (packed mul as fma)", 8);
	run<5>("FMUL2 + scalar FADD alternating", 16);
	run<6>("FMUL2 + FFMA2-add alternating", 16);
	return 0;
}
