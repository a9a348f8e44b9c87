// mbik_kernel_l4.cu -- per-pose limit sets (mbik_solve_batch_limits) for the size variant {128 solved bones, segment 128,
// stack 16}: the thread-per-pose kernel with the kusudama data read from the pose's limit-set record (solve_body LIMS).
// large-rig variant: scalar formulations (see mbik_kernel_v4.cu)
#define MBIK_F2_MAT 0
#define MBIK_F2_VEC 0
#define MBIK_F2_DOT 0
#define MBIK_F2_DIV 0
#include "mbik_kernel_body.cuh"

namespace mbik {

cudaError_t launch_lims_v4(const SolveArgs &a, cudaStream_t stream) {
	if (a.stabilize) {
		return launch_variant_lims<128, 128, 16, kStabBlockThreads, true>(a, stream);
	}
	if (a.use_glw && glw_fits<kBlockThreads>(a)) { // long walks: the streamed-walk instantiation reads the pose's limit-set record too
		const cudaError_t e = launch_variant_glw<128, 128, 16, kBlockThreads, false, true>(a, a.sm_count, stream);
		if (e != cudaErrorMemoryAllocation && e != cudaErrorNotSupported) {
			return e;
		}
		cudaGetLastError();
	}
	return launch_variant_lims<128, 128, 16, kBlockThreads>(a, stream);
}

// segment-parallel (small-batch) mapping of the same: one 32-pose group per CTA, one warp per concurrently solvable segment
cudaError_t launch_sp_lims_v4(const SolveArgs &a, int min_groups_per_sm, cudaStream_t stream) {
	return launch_variant_sp<128, 128, 16, false, true>(a, min_groups_per_sm, stream);
}

} // namespace mbik
