// mbik_kernel_l0.cu -- per-pose limit sets (mbik_solve_batch_limits) for the size variant {20 solved bones, segment 4,
// stack 2}: the thread-per-pose kernel with the kusudama data read from the pose's limit-set record (solve_body LIMS).
#include "mbik_kernel_body.cuh"

namespace mbik {

cudaError_t launch_lims_v0(const SolveArgs &a, cudaStream_t stream) {
	if (a.stabilize) {
		return launch_variant_lims<20, 4, 2, kStabBlockThreads, true>(a, stream);
	}
	return launch_variant_lims<20, 4, 2, kBlockThreads>(a, stream);
}

// segment-parallel (small-batch) mapping of the same: one 32-pose group per CTA, one warp per concurrently solvable segment
cudaError_t launch_sp_lims_v0(const SolveArgs &a, int min_groups_per_sm, cudaStream_t stream) {
	return launch_variant_sp<20, 4, 2, false, true>(a, min_groups_per_sm, stream);
}

} // namespace mbik
