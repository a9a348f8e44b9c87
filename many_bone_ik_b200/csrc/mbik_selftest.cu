// mbik_selftest.cu -- device self-test of the guarded sqrt/division groups of mbik_math.cuh against the
// compiler's own correctly rounded __fsqrt_rn / __fdiv_rn (bit-for-bit).  Not part of the solve path.
//   sqrt : EXHAUSTIVE over every float in the guarded range [2^-80, 2^80)
//   div  : every one of the 2^23 divisor mantissas x `rounds` x 64 pseudo-random numerators/exponents/signs
//   vnorm / q_normalized / Basis::get_quaternion helpers: random vectors incl. zero, tiny and huge components,
//          against the unguarded formulation
#include "../../include/mbik.h"
#include "mbik_kernel_body.cuh"
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

__device__ bool mode_all_ordinary(const float *c) {
	bool ok = true;
	for (int i = 0; i < 4; i++) {
		float a = fabsf(c[i]);
		ok = ok && a >= 0.0625f && a <= 16.0f;
	}
	return ok;
}

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
		// CheckedOps: wherever it reports ok, its values equal the literal IEEE formulation
		{
			CheckedOps co;
			V3 g = co.normalized(a);
			if (co.ok) {
				bd += !(same(g.x, p.x) && same(g.y, p.y) && same(g.z, p.z));
			}
			CheckedOps c2;
			float s3, d3;
			c2.sqrt_then_div(c[3], 0.5f, s3, d3);
			if (c2.ok) {
				bd += !(same(s3, s2) && same(d3, __fdiv_rn(0.5f, s2)));
			}
			CheckedOps c3;
			float q3 = c3.div_const(2.0f, c[2]);
			if (c3.ok) {
				bd += !same(q3, __fdiv_rn(2.0f, c[2]));
			}
			CheckedOps c4;
			float r4 = c4.sqrt(c[1]);
			if (c4.ok) {
				bd += !same(r4, __fsqrt_rn(c[1]));
			}
			// and it must report ok for ordinary operands (otherwise the fast path would never be taken)
			if (mode_all_ordinary(c)) {
				bd += !(co.ok && c3.ok);
			}
		}
		cnt += 7;
	}
	atomicAdd(bad, bd);
	atomicAdd(n, cnt);
}

// ---- stage probes: the kernel's own device functions on caller-supplied inputs ----
// Packed FP32x2 operations (mbik_math.cuh: f2_mul / f2_add / f2_sub and the composites built on them) against the scalar
// individually rounded intrinsics, per lane, on random bit patterns: ordinary values, matched exponents (cancellation),
// zeros of both signs, subnormals, infinities and NaNs.
__device__ float f2_case_value(uint64_t h, int k) {
	uint32_t bits = (uint32_t)(h >> (k * 7));
	switch ((h >> 59) & 7) {
		case 0: return __uint_as_float(bits);                                              // anything, NaN / Inf included
		case 1: return __uint_as_float((bits & 0x807fffffu) | 0x3f800000u);                // [1, 2)
		case 2: return __uint_as_float((bits & 0x807fffffu) | (((bits >> 23) & 7u) + 124u) << 23); // 2^-3 .. 2^4
		case 3: return __uint_as_float(bits & 0x80000000u);                                // +-0
		case 4: return __uint_as_float(bits & 0x807fffffu);                                // subnormal
		case 5: return __uint_as_float((bits & 0x80000000u) | 0x7f800000u);                // +-Inf
		default: return __uint_as_float((bits & 0x807fffffu) | (((bits >> 23) & 63u) + 96u) << 23);
	}
}
__global__ void f2_cases(unsigned long long *bad, unsigned long long *n, uint32_t round) {
	uint64_t h = splitmix(((uint64_t)round << 40) ^ ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) ^ 0xF2F2F2F2ull);
	unsigned long long bd = 0, cnt = 0;
	for (int it = 0; it < 64; it++) {
		h = splitmix(h);
		uint64_t g = splitmix(h ^ 0x5555u);
		float a0 = f2_case_value(h, 0), a1 = f2_case_value(h, 1), b0 = f2_case_value(g, 0), b1 = f2_case_value(g, 1);
		if (it & 1) { // equal magnitudes: exact cancellation and signed-zero results
			b0 = (it & 2) ? a0 : -a0;
		}
		float lo, hi;
		f2_get(f2_mul(f2(a0, a1), f2(b0, b1)), lo, hi);
		bd += !same(lo, __fmul_rn(a0, b0)) + !same(hi, __fmul_rn(a1, b1));
		f2_get(f2_add(f2(a0, a1), f2(b0, b1)), lo, hi);
		bd += !same(lo, __fadd_rn(a0, b0)) + !same(hi, __fadd_rn(a1, b1));
		f2_get(f2_sub(f2(a0, a1), f2(b0, b1)), lo, hi);
		bd += !same(lo, __fsub_rn(a0, b0)) + !same(hi, __fsub_rn(a1, b1));
		// a product feeding a sum must stay two roundings (no contraction into FFMA2)
		f2_get(f2_add(f2_mul(f2(a0, a1), f2(b0, b1)), f2(b1, a0)), lo, hi);
		bd += !same(lo, __fadd_rn(__fmul_rn(a0, b0), b1)) + !same(hi, __fadd_rn(__fmul_rn(a1, b1), a0));
		f2_get(f2_sub(f2(b1, a0), f2_mul(f2(a0, a1), f2(b0, b1))), lo, hi);
		bd += !same(lo, __fsub_rn(b1, __fmul_rn(a0, b0))) + !same(hi, __fsub_rn(a0, __fmul_rn(a1, b1)));
		// composites: Basis product / xform rows, Vector3 dot, against the literal scalar formulas
		M3 A, B;
		uint64_t q = h;
		for (int k = 0; k < 9; k++) {
			q = splitmix(q);
			A.m[k] = f2_case_value(q, 0);
			B.m[k] = f2_case_value(q, 2);
		}
		M3 P = m3_mul(A, B);
		for (int i = 0; i < 3; i++) {
			for (int j = 0; j < 3; j++) {
				float ref = __fadd_rn(__fadd_rn(__fmul_rn(B.m[j], A.m[3 * i]), __fmul_rn(B.m[3 + j], A.m[3 * i + 1])), __fmul_rn(B.m[6 + j], A.m[3 * i + 2]));
				bd += !same(P.m[3 * i + j], ref);
			}
		}
		V3 v = v3(a0, b1, b0), w = m3_xform(A, v);
		float wr[3] = { w.x, w.y, w.z };
		for (int i = 0; i < 3; i++) {
			float ref = __fadd_rn(__fadd_rn(__fmul_rn(A.m[3 * i], v.x), __fmul_rn(A.m[3 * i + 1], v.y)), __fmul_rn(A.m[3 * i + 2], v.z));
			bd += !same(wr[i], ref);
		}
		V3 u = v3(a1, b0, b1);
		bd += !same(vdot(u, v), __fadd_rn(__fadd_rn(__fmul_rn(u.x, v.x), __fmul_rn(u.y, v.y)), __fmul_rn(u.z, v.z)));
		cnt += 10 + 9 + 3 + 1;
	}
	atomicAdd(bad, bd);
	atomicAdd(n, cnt);
}

__global__ void stage_qcp_kernel(int n, const float *moved, const float *target, const double *weight, int translate, int newton_iters, float *out7) {
	if (threadIdx.x != 0 || blockIdx.x != 0) {
		return;
	}
	HeadingAcc A;
	qcp_zero(A.sums);
	A.neg_mc = A.neg_tc = v3(0.0f, 0.0f, 0.0f);
	for (int pass_i = translate ? 0 : 1; pass_i < 2; pass_i++) {
		A.total_w = 0.0;
		A.csum_m = A.csum_t = v3(0.0f, 0.0f, 0.0f);
		for (int i = 0; i < n; i++) {
			V3 mh = v3(moved[3 * i], moved[3 * i + 1], moved[3 * i + 2]);
			V3 th = v3(target[3 * i], target[3 * i + 1], target[3 * i + 2]);
			heading_emit(A, pass_i, translate != 0, th, mh, weight[i], (float)weight[i]);
		}
		if (pass_i == 0) { // same centroid code as solve_body
			V3 moved_center, target_center;
			if (A.total_w > 0.0) {
				moved_center = vdivs(A.csum_m, (float)A.total_w);
				target_center = vdivs(A.csum_t, (float)A.total_w);
			} else {
				moved_center = A.csum_m;
				target_center = A.csum_t;
			}
			A.neg_mc = vmuls(moved_center, -1.0f);
			A.neg_tc = vmuls(target_center, -1.0f);
		}
	}
	Q4 q = (n == 1) ? qcp_rotation_single(A.csum_m, A.csum_t) : qcp_rotation(A.sums, newton_iters);
	V3 t = vsub(vneg(A.neg_tc), vneg(A.neg_mc));
	out7[0] = q.x; out7[1] = q.y; out7[2] = q.z; out7[3] = q.w;
	out7[4] = t.x; out7[5] = t.y; out7[6] = t.z;
}

__global__ void stage_clamp_kernel(int n, const float *quats, const double *cos_half, float *out) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) {
		return;
	}
	Q4 q = clamp_to_cos_half_angle(q4(quats[4 * i], quats[4 * i + 1], quats[4 * i + 2], quats[4 * i + 3]), cos_half[i]);
	out[4 * i] = q.x; out[4 * i + 1] = q.y; out[4 * i + 2] = q.z; out[4 * i + 3] = q.w;
}

__global__ void stage_point_in_limits_kernel(const BlobCone *cones, int n_cones, int n, const float *points, float *out) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) {
		return;
	}
	float in_bounds;
	V3 p = point_in_limits(v3(points[3 * i], points[3 * i + 1], points[3 * i + 2]), cones, n_cones, in_bounds);
	out[4 * i] = p.x; out[4 * i + 1] = p.y; out[4 * i + 2] = p.z; out[4 * i + 3] = in_bounds;
}

} // namespace

namespace mbik {
cudaError_t launch_stage_qcp(int n, const float *d_moved, const float *d_target, const double *d_weight, int translate, int newton_iters, float *d_out7) {
	stage_qcp_kernel<<<1, 32>>>(n, d_moved, d_target, d_weight, translate, newton_iters, d_out7);
	return cudaGetLastError();
}
cudaError_t launch_stage_clamp(int n, const float *d_quats, const double *d_cos_half, float *d_out) {
	stage_clamp_kernel<<<(n + 127) / 128, 128>>>(n, d_quats, d_cos_half, d_out);
	return cudaGetLastError();
}
cudaError_t launch_stage_point_in_limits(const BlobCone *d_cones, int n_cones, int n, const float *d_points, float *d_out) {
	stage_point_in_limits_kernel<<<(n + 127) / 128, 128>>>(d_cones, n_cones, n, d_points, d_out);
	return cudaGetLastError();
}
} // namespace mbik

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
		f2_cases<<<148 * 8, 256>>>(d, d + 1, (uint32_t)r);
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
