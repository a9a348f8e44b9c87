// mbik_kernel_v0.cu -- instantiations of the solve kernel for the size variant {20 solved bones, segment 4, stack 2}.
#include "mbik_kernel_body.cuh"

// experiment knob (profiles/build_variant.py): CTA size of the large-batch instantiation.  640 threads (5 warps per scheduler,
// 96 registers) was re-measured in round 2 after the register diet of the step (profiles/r2_exp_640_threads_lean.log): 27.4 M solves/s
// as is, 29.3 M without the walk's child prefetch, 30.7 M with scalar Vector3 ops on top (112 B / 252 B of spill stores / loads left),
// 28.9 M with every packed composite off (76 B / 164 B) -- against 33.3 M at 512 threads: per full wave the 20-warp builds deliver what
// the 16-warp build does (the 96-register schedule is ~25 % slower per warp), and 2^20 poses quantise worse (11.07 waves -> 12).
#ifndef MBIK_V0_GLW
#define MBIK_V0_GLW 0
#endif
#ifndef MBIK_V0_BIG
#define MBIK_V0_BIG 512
#endif

namespace mbik {

cudaError_t launch_v0(const SolveArgs &a, int threads, cudaStream_t stream) {
	switch (threads) {
		case 0: // stabilisation passes > 0: separate instantiation, the default path pays nothing for it
			return launch_variant<20, 4, 2, kStabBlockThreads, true>(a, stream);
		case 32:
			return launch_variant<20, 4, 2, 32>(a, stream);
		case 64:
			return launch_variant<20, 4, 2, 64>(a, stream);
		case 128:
			return launch_variant<20, 4, 2, 128>(a, stream);
		case 256:
			return launch_variant<20, 4, 2, 256>(a, stream);
		case 512:
#if MBIK_V0_GLW
			// experiment: the streamed-walk instantiation on the small variant (needs thread-local scratch: -DMBIK_SCRATCH_LIMIT_KB=0)
			return launch_variant_glw<20, 4, 2, 512>(a, a.sm_count, stream);
#endif
			return launch_variant<20, 4, 2, MBIK_V0_BIG>(a, stream);
		case 384: // wave-balanced sizes for large batches (launch_solve): the last wave of CTAs is as full as the others
			return launch_variant<20, 4, 2, 384>(a, stream);
		case 448:
			return launch_variant<20, 4, 2, 448>(a, stream);
		default:
			return launch_variant<20, 4, 2, kBlockThreads>(a, stream);
	}
}

} // namespace mbik
