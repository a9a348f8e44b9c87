// mbik_selftest.cu -- device self-test of the guarded sqrt/division groups of mbik_math.cuh against the
// compiler's own correctly rounded __fsqrt_rn / __fdiv_rn (bit-for-bit).  Not part of the solve path.
//   sqrt : EXHAUSTIVE over every float in the guarded range [2^-80, 2^80)
//   div  : every one of the 2^23 divisor mantissas x `rounds` x 64 pseudo-random numerators/exponents/signs
//   vnorm / q_normalized / Basis::get_quaternion helpers: random vectors incl. zero, tiny and huge components,
//          against the unguarded formulation
#include "../../include/mbik.h"
#include "mbik_math.cuh"

#include <cuda_runtime.h>

namespace {
using namespace mbik;

__device__ __forceinline__ uint64_t splitmix(uint64_t x) {
	x += 0x9E3779B97F4A7C15ull;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}

__global__ void sqrt_exhaustive(unsigned long long *bad, unsigned long long *n) {
	const uint32_t lo = kBits2m80, hi = kBits2p80;
	unsigned long long cnt = 0, b = 0;
	for (uint64_t u = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < hi; u += (uint64_t)gridDim.x * blockDim.x) {
		float x = __uint_as_float((uint32_t)u);
		b += __float_as_uint(sqrt_guarded(x)) != __float_as_uint(__fsqrt_rn(x));
		cnt++;
	}
	atomicAdd(bad, b);
	atomicAdd(n, cnt);
}

__global__ void div_sweep(unsigned long long *bad, unsigned long long *n, uint32_t round) {
	unsigned long long cnt = 0, bd = 0;
	for (uint32_t m = blockIdx.x * blockDim.x + threadIdx.x; m < (1u << 23); m += gridDim.x * blockDim.x) {
		uint64_t h = splitmix(((uint64_t)round << 32) | m);
		// divisor: this mantissa, exponent in [-40, 40), random sign
		uint32_t eb = 127 - 40 + (uint32_t)(h % 80);
		float b = __uint_as_float(((uint32_t)(h >> 40) & 0x80000000u) | (eb << 23) | m);
		float y1 = rcp_refined(b);
		for (int k = 0; k < 64; k++) {
			h = splitmix(h);
			uint32_t ea = 127 - 60 + (uint32_t)((h >> 32) % 120);
			uint32_t ma = (uint32_t)h & 0x7fffffu;
			if (k < 8) {
				ma = (k & 1) ? 0x7fffffu - (k >> 1) : (uint32_t)(k >> 1); // mantissa extremes
			} else if (k < 12) {
				ma = m; // numerator mantissa == divisor mantissa (exact quotients)
			}
			float a = __uint_as_float(((uint32_t)(h >> 8) & 0x80000000u) | (ea << 23) | ma);
			bd += __float_as_uint(div_guarded(a, b, y1)) != __float_as_uint(__fdiv_rn(a, b));
			cnt++;
		}
	}
	atomicAdd(bad, bd);
	atomicAdd(n, cnt);
}

// the unguarded formulations (what the host flattener and the CPU reference compute)
__device__ V3 vnorm_plain(V3 a) {
	float l2 = vlen2(a);
	if (l2 == 0.0f) {
		return v3(0.0f, 0.0f, 0.0f);
	}
	float l = __fsqrt_rn(l2);
	return v3(__fdiv_rn(a.x, l), __fdiv_rn(a.y, l), __fdiv_rn(a.z, l));
}
__device__ bool same(float a, float b) { return __float_as_uint(a) == __float_as_uint(b) || (a != a && b != b); }

__global__ void vector_cases(unsigned long long *bad, unsigned long long *n, uint32_t round) {
	uint64_t h = splitmix(((uint64_t)round << 40) ^ ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x));
	unsigned long long cnt = 0, bd = 0;
	for (int k = 0; k < 256; k++) {
		float c[4];
		for (int i = 0; i < 4; i++) {
			h = splitmix(h);
			uint32_t mode = (uint32_t)(h >> 60);
			uint32_t bits = (uint32_t)h;
			if (mode == 0) {
				bits &= 0x80000000u; // +-0
			} else if (mode < 10) {
				bits = (bits & 0x807fffffu) | ((127 - 3 + (uint32_t)((h >> 32) % 6)) << 23); // O(1) values
			} else if (mode < 13) {
				bits = (bits & 0x807fffffu) | ((127 - 100 + (uint32_t)((h >> 32) % 200)) << 23); // wide range
			} // else: raw bits (denormals, inf, NaN, anything)
			c[i] = __uint_as_float(bits);
		}
		V3 a = v3(c[0], c[1], c[2]);
		V3 f = vnorm(a), p = vnorm_plain(a);
		bd += !(same(f.x, p.x) && same(f.y, p.y) && same(f.z, p.z));
		Q4 q = q4(c[0], c[1], c[2], c[3]);
		Q4 qf = q_normalized(q);
		float inv = __fdiv_rn(1.0f, __fsqrt_rn(q_dot(q, q)));
		bd += !(same(qf.x, __fmul_rn(q.x, inv)) && same(qf.y, __fmul_rn(q.y, inv)) && same(qf.z, __fmul_rn(q.z, inv)) && same(qf.w, __fmul_rn(q.w, inv)));
		float s, d;
		sqrt_then_div(c[3], 0.5f, s, d);
		float s2 = __fsqrt_rn(c[3]);
		bd += !(same(s, s2) && same(d, __fdiv_rn(0.5f, s2)));
		cnt += 3;
	}
	atomicAdd(bad, bd);
	atomicAdd(n, cnt);
}

} // namespace

extern "C" {
#pragma GCC visibility push(default)
int mbik_selftest(int32_t device, int32_t rounds, uint64_t *out_checked, uint64_t *out_mismatches) {
	if (!out_checked || !out_mismatches) {
		return MBIK_ERR_INVALID_ARG;
	}
	if (cudaSetDevice(device) != cudaSuccess) {
		return MBIK_ERR_NO_DEVICE;
	}
	unsigned long long *d = nullptr;
	if (cudaMalloc((void **)&d, 2 * sizeof(unsigned long long)) != cudaSuccess) {
		return MBIK_ERR_ALLOC;
	}
	cudaMemset(d, 0, 2 * sizeof(unsigned long long));
	sqrt_exhaustive<<<148 * 8, 256>>>(d, d + 1);
	for (int r = 0; r < (rounds > 0 ? rounds : 1); r++) {
		div_sweep<<<148 * 8, 256>>>(d, d + 1, (uint32_t)r);
		vector_cases<<<148 * 8, 256>>>(d, d + 1, (uint32_t)r);
	}
	unsigned long long h[2] = { 0, 0 };
	cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
	cudaFree(d);
	if (e != cudaSuccess) {
		return MBIK_ERR_CUDA;
	}
	*out_mismatches = h[0];
	*out_checked = h[1];
	return MBIK_OK;
}
#pragma GCC visibility pop
}
