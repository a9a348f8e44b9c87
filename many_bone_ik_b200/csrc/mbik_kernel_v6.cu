// mbik_kernel_v6.cu -- the unbounded-rig variant of the solve kernel (solve_body DYN): rigs beyond every compiled capacity.
// scalar formulations, like the other large-rig variants (see mbik_kernel_v3.cu)
#define MBIK_F2_MAT 0
#define MBIK_F2_VEC 0
#define MBIK_F2_DOT 0
#define MBIK_F2_DIV 0
#include "mbik_kernel_body.cuh"

namespace mbik {

cudaError_t launch_v6(const SolveArgs &a, int sm_count, cudaStream_t stream) {
	if (a.limit_table) {
		return a.stabilize ? launch_variant_dyn<true, true>(a, sm_count, stream) : launch_variant_dyn<false, true>(a, sm_count, stream);
	}
	return a.stabilize ? launch_variant_dyn<true, false>(a, sm_count, stream) : launch_variant_dyn<false, false>(a, sm_count, stream);
}

} // namespace mbik
