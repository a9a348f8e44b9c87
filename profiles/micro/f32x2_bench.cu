// Microbenchmark: issue rate of the Blackwell packed FP32x2 instructions (FMUL2 / FADD2) against scalar FMUL / FADD,
// and bit-equality of their per-lane results with __fmul_rn / __fadd_rn.   nvcc -arch=sm_100a -O3 -fmad=false
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#ifdef USE_FMA_FORMS
// a * b == fma(a, b, -0) and a + b == fma(a, 1, b) exactly (one rounding each, signed zeros and NaNs included)
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(0x8000000080000000ull)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(0x3f8000003f800000ull), "l"(b)); return r; }
#else
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
#endif
__device__ __forceinline__ u64 pack(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

template <int MODE> // 0 scalar mul+add alternating (16 chains), 1 packed (8 chains of pairs), 2 scalar dependent chain (latency), 3 packed dependent chain
__global__ void bench(float *out, int iters, float a, float b) {
	float x[16];
	for (int i = 0; i < 16; i++) x[i] = threadIdx.x * 0.001f + i;
	if (MODE == 0) {
		for (int it = 0; it < iters; it++) {
#pragma unroll
			for (int i = 0; i < 16; i++) x[i] = __fadd_rn(__fmul_rn(x[i], a), b);
		}
	} else if (MODE == 1) {
		u64 p[8], pa = pack(a, a), pb = pack(b, b);
		for (int i = 0; i < 8; i++) p[i] = pack(x[2 * i], x[2 * i + 1]);
		for (int it = 0; it < iters; it++) {
#pragma unroll
			for (int i = 0; i < 8; i++) { p[i] = mul2(p[i], pa); asm volatile("" : "+l"(p[i])); p[i] = add2(p[i], pb); }
		}
		for (int i = 0; i < 8; i++) unpack(p[i], x[2 * i], x[2 * i + 1]);
	} else if (MODE == 2) {
		for (int it = 0; it < iters; it++) {
#pragma unroll
			for (int i = 0; i < 16; i++) x[0] = __fadd_rn(__fmul_rn(x[0], a), b);
		}
	} else {
		u64 p = pack(x[0], x[1]), pa = pack(a, a), pb = pack(b, b);
		for (int it = 0; it < iters; it++) {
#pragma unroll
			for (int i = 0; i < 16; i++) { p = mul2(p, pa); asm volatile("" : "+l"(p)); p = add2(p, pb); }
		}
		unpack(p, x[0], x[1]);
	}
	float s = 0;
	for (int i = 0; i < 16; i++) s += x[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void check(const float *a, const float *b, int n, unsigned *bad) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (2 * i + 1 >= n) return;
	u64 pa = pack(a[2 * i], a[2 * i + 1]), pb = pack(b[2 * i], b[2 * i + 1]);
	u64 m = mul2(pa, pb);
	asm volatile("" : "+l"(m));
	u64 s = add2(m, pa);
	float m0, m1, s0, s1;
	unpack(m, m0, m1);
	unpack(s, s0, s1);
	float rm0 = __fmul_rn(a[2 * i], b[2 * i]), rm1 = __fmul_rn(a[2 * i + 1], b[2 * i + 1]);
	float rs0 = __fadd_rn(rm0, a[2 * i]), rs1 = __fadd_rn(rm1, a[2 * i + 1]);
	auto same = [](float x, float y) { return __float_as_uint(x) == __float_as_uint(y) || (x != x && y != y); };
	if (!same(m0, rm0) || !same(m1, rm1) || !same(s0, rs0) || !same(s1, rs1)) atomicAdd(bad, 1u);
}
template <int MODE>
float run(int iters) {
	float *o; cudaMalloc(&o, 148 * 8 * 512 * sizeof(float));
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	bench<MODE><<<148 * 8, 512>>>(o, 16, 1.0001f, 0.5f);
	cudaEventRecord(e0);
	bench<MODE><<<148 * 8, 512>>>(o, iters, 1.0001f, 0.5f);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1); cudaFree(o);
	return ms;
}
int main() {
	const int iters = 20000;
	float t0 = run<0>(iters), t1 = run<1>(iters), t2 = run<2>(iters), t3 = run<3>(iters);
	double lane_ops = 148.0 * 8 * 512 * (double)iters * 32; // mul + add on 16 floats
	printf("scalar FMUL+FADD, 16 chains : %.3f ms  %.2f T lane-op/s\n", t0, lane_ops / t0 / 1e9);
	printf("packed FMUL2+FADD2, 8 chains: %.3f ms  %.2f T lane-op/s\n", t1, lane_ops / t1 / 1e9);
	printf("dependent scalar chain       : %.3f ms  -> %.2f cycles per op at 1.965 GHz (one warp-chain per thread, 16 warps/SM... see code)\n", t2, t2 * 1e-3 * 1.965e9 / (iters * 32.0));
	printf("dependent packed chain       : %.3f ms  -> %.2f\n", t3, t3 * 1e-3 * 1.965e9 / (iters * 32.0));
	// bit-equality on random bit patterns (specials included)
	const int n = 1 << 22;
	unsigned *ha = (unsigned *)malloc(n * 4), *hb = (unsigned *)malloc(n * 4);
	srand(1);
	for (int i = 0; i < n; i++) { ha[i] = (unsigned)rand() ^ ((unsigned)rand() << 16); hb[i] = (unsigned)rand() ^ ((unsigned)rand() << 16); if (i % 7 == 0) hb[i] = (hb[i] & 0x807fffffu) | ((ha[i] & 0x7f800000u)); }
	float *da, *db; unsigned *dbad; cudaMalloc(&da, n * 4); cudaMalloc(&db, n * 4); cudaMalloc(&dbad, 4); cudaMemset(dbad, 0, 4);
	cudaMemcpy(da, ha, n * 4, cudaMemcpyHostToDevice); cudaMemcpy(db, hb, n * 4, cudaMemcpyHostToDevice);
	check<<<(n / 2 + 255) / 256, 256>>>(da, db, n, dbad);
	unsigned bad; cudaMemcpy(&bad, dbad, 4, cudaMemcpyDeviceToHost);
	printf("bit-equality with __fmul_rn / __fadd_rn over %d random pairs: %u mismatches\n", n / 2, bad);
	return bad != 0;
}
