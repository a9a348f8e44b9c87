// mbik_kernel_sp4.cu -- segment-parallel (small-batch) instantiations of the solve kernel for the size variant
// {128 solved bones, segment 128, stack 16}: one group of 32 poses per CTA, one warp per concurrently solvable segment.
#include "mbik_kernel_body.cuh"

namespace mbik {

cudaError_t launch_sp_v4(const SolveArgs &a, int min_groups_per_sm, cudaStream_t stream) {
	if (a.stabilize) {
		return launch_variant_sp<128, 128, 16, true>(a, min_groups_per_sm, stream);
	}
	return launch_variant_sp<128, 128, 16, false>(a, min_groups_per_sm, stream);
}

} // namespace mbik
