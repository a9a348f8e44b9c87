// mbik_kernel_sp3.cu -- segment-parallel (small-batch) instantiations of the solve kernel for the size variant
// {64 solved bones, segment 16, stack 8}: one group of 32 poses per CTA, one warp per concurrently solvable segment.
#include "mbik_kernel_body.cuh"

namespace mbik {

cudaError_t launch_sp_v3(const SolveArgs &a, int min_groups_per_sm, cudaStream_t stream) {
	if (a.stabilize) {
		return launch_variant_sp<64, 16, 8, true>(a, min_groups_per_sm, stream);
	}
	return launch_variant_sp<64, 16, 8, false>(a, min_groups_per_sm, stream);
}

} // namespace mbik
