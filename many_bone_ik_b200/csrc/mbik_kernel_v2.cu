// mbik_kernel_v2.cu -- instantiations of the solve kernel for the size variant {64 solved bones, segment 8, stack 1}.
#include "mbik_kernel_body.cuh"

namespace mbik {

cudaError_t launch_v2(const SolveArgs &a, int threads, cudaStream_t stream) {
	switch (threads) {
		case 0: // stabilisation passes > 0: separate instantiation, the default path pays nothing for it
			return launch_variant<64, 8, 1, kStabBlockThreads, true>(a, stream);
		case 32: // small batches: one warp per SM (latency, not throughput)
			return launch_variant<64, 8, 1, 32>(a, stream);
		case 128:
			return launch_variant<64, 8, 1, 128>(a, stream);
		default:
			return launch_variant<64, 8, 1, kBlockThreads>(a, stream);
	}
}

} // namespace mbik
