// mbik_peak.cu -- FP32 FMA micro-benchmark used as the measured roofline denominator of the solve kernel
// ("builder must re-measure with an FMA micro-benchmark on the box").
// Not part of the solve path.
#include "../../include/mbik.h"

#include <cuda_runtime.h>

namespace {

__global__ void __launch_bounds__(256) fma_peak_kernel(float *out, int iters, float b, float c) {
	float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
#pragma unroll 4
	for (int i = 0; i < iters; i++) {
		a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
		a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

} // namespace

extern "C" {
#pragma GCC visibility push(default)
// Best-of-`reps` FP32 FMA throughput of `device` in TFLOP/s (2 flop per FMA).  Returns MBIK_OK or an error.
int mbik_measure_fp32_tflops(int32_t device, int32_t reps, double *out_tflops) {
	if (!out_tflops) {
		return MBIK_ERR_INVALID_ARG;
	}
	if (cudaSetDevice(device) != cudaSuccess) {
		return MBIK_ERR_NO_DEVICE;
	}
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
		return MBIK_ERR_CUDA;
	}
	const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 1 << 14;
	float *out = nullptr;
	if (cudaMalloc((void **)&out, sizeof(float) * threads * blocks) != cudaSuccess) {
		return MBIK_ERR_ALLOC;
	}
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	double best = 0;
	for (int r = 0; r < (reps > 0 ? reps : 5) + 1; r++) {
		cudaEventRecord(e0);
		fma_peak_kernel<<<blocks, threads>>>(out, iters, 0.999f, 1e-3f);
		cudaEventRecord(e1);
		if (cudaEventSynchronize(e1) != cudaSuccess) {
			cudaFree(out);
			return MBIK_ERR_CUDA;
		}
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		double tf = (double)threads * blocks * iters * 8.0 * 2.0 / (ms * 1e-3) / 1e12;
		if (r > 0 && tf > best) {
			best = tf;
		}
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(out);
	*out_tflops = best;
	return MBIK_OK;
}
#pragma GCC visibility pop
}
