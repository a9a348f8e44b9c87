// mbik_kernel.cu -- variant selection and launch dispatch of the solve kernel.
#include "mbik_kernel.h"

#include <cstdlib>

namespace mbik {

int kernel_variant_for(int n_solved, int max_seg_len, int max_stack, size_t blob_bytes) {
	for (int v = 0; v < kNumVariants; v++) {
		// the small variants keep their scratch transforms in shared memory (up to 150 KiB) next to the rig blob
		if (v <= 2 && blob_bytes > 72 * 1024) {
			continue;
		}
		if (n_solved <= kVariants[v][0] && max_seg_len <= kVariants[v][1] && max_stack <= kVariants[v][2]) {
			return v;
		}
	}
	return -1;
}

int kernel_capacity_of_variant(int v) { return (v >= 0 && v < kNumVariants) ? kVariants[v][0] : -1; }

cudaError_t launch_solve(const SolveArgs &a, int variant, int sm_count, cudaStream_t stream) {
	int threads = kBlockThreads;
	if (a.stabilize) {
		threads = 0;
	} else if (variant <= 1) {
		// Small batches: one CTA per SM with as few warps as cover the batch (a 4096-pose batch runs as 128
		// one-warp CTAs on 128 SMs instead of 11 twelve-warp CTAs on 11 SMs) -- latency, not throughput.
		static const int forced = getenv("MBIK_THREADS") ? atoi(getenv("MBIK_THREADS")) : 0; // tuning knob
		if (forced > 0) {
			threads = forced;
		} else {
			const int cands[] = { 32, 64, 128, 256 };
			for (int c : cands) {
				if ((a.n_poses + c - 1) / c <= (size_t)sm_count) {
					threads = c;
					break;
				}
			}
		}
	}
	switch (variant) {
		case 0:
			return launch_v0(a, threads, stream);
		case 1:
			return launch_v1(a, threads, stream);
		case 2:
			return launch_v2(a, threads, stream);
		case 3:
			return launch_v3(a, threads, stream);
		case 4:
			return launch_v4(a, threads, stream);
		default:
			return cudaErrorInvalidValue;
	}
}

} // namespace mbik
