// mbik_kernel_v3.cu -- instantiations of the solve kernel for the size variant {64 solved bones, segment 16, stack 8}.
// The large-rig variants are register-starved at 128 registers (64+ local poses to address): with the packed FP32x2
// composites they spill and run 2-3 % slower (chain64 71.6 vs 69.4 ms, quad80 13.8 vs 13.5 ms per 75776-pose launch),
// so this translation unit keeps the scalar formulations (same bits either way).
// (The streamed-walk instantiations do not spill with them -- and gain nothing: chain64 +0.9 %, quad80 +0.9 %, big_tree240
// -0.5 %, profiles/r2_exp_f2_large_rigs.log -- they wait on the per-pose state, not on issue slots.)
#define MBIK_F2_MAT 0
#define MBIK_F2_VEC 0
#define MBIK_F2_DOT 0
#define MBIK_F2_DIV 0
#include "mbik_kernel_body.cuh"

namespace mbik {

cudaError_t launch_v3(const SolveArgs &a, int threads, cudaStream_t stream) {
	switch (threads) {
		case 0: // stabilisation passes > 0: separate instantiation, the default path pays nothing for it
			if (a.use_glw && ScratchStride<16, 8, kStabBlockThreads>::value == 0 && glw_fits<kStabBlockThreads>(a)) { // long walks: streamed
				const cudaError_t e = launch_variant_glw<64, 16, 8, kStabBlockThreads, true>(a, a.sm_count, stream);
				if (e != cudaErrorMemoryAllocation && e != cudaErrorNotSupported) {
					return e;
				}
				cudaGetLastError(); // no workspace: thread-local state instead
			}
			return launch_variant<64, 16, 8, kStabBlockThreads, true>(a, stream);
		case 32: // small batches: one warp per SM (latency, not throughput)
			return launch_variant<64, 16, 8, 32>(a, stream);
		case 128:
			return launch_variant<64, 16, 8, 128>(a, stream);
		case 384: // wave-balanced sizes for large batches (launch_solve): the last wave of CTAs is as full as the others
			if (a.use_glw && glw_fits<384>(a)) {
				const cudaError_t e = launch_variant_glw<64, 16, 8, 384>(a, a.sm_count, stream);
				if (e != cudaErrorMemoryAllocation && e != cudaErrorNotSupported) {
					return e;
				}
				cudaGetLastError(); // no workspace (pool exhausted / no stream-ordered allocator): thread-local state instead
			}
			return launch_variant<64, 16, 8, 384>(a, stream);
		case 448:
			if (a.use_glw && glw_fits<448>(a)) {
				const cudaError_t e = launch_variant_glw<64, 16, 8, 448>(a, a.sm_count, stream);
				if (e != cudaErrorMemoryAllocation && e != cudaErrorNotSupported) {
					return e;
				}
				cudaGetLastError(); // no workspace (pool exhausted / no stream-ordered allocator): thread-local state instead
			}
			return launch_variant<64, 16, 8, 448>(a, stream);
		default:
			if (a.use_glw && glw_fits<kBlockThreads>(a)) {
				const cudaError_t e = launch_variant_glw<64, 16, 8, kBlockThreads>(a, a.sm_count, stream);
				if (e != cudaErrorMemoryAllocation && e != cudaErrorNotSupported) {
					return e;
				}
				cudaGetLastError(); // no workspace (pool exhausted / no stream-ordered allocator): thread-local state instead
			}
			return launch_variant<64, 16, 8, kBlockThreads>(a, stream);
	}
}

} // namespace mbik
