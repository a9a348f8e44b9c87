// mbik_kernel_v3.cu -- instantiations of the solve kernel for the size variant {64 solved bones, segment 16, stack 8}.
#include "mbik_kernel_body.cuh"

namespace mbik {

cudaError_t launch_v3(const SolveArgs &a, int threads, cudaStream_t stream) {
	switch (threads) {
		case 0: // stabilisation passes > 0: separate instantiation, the default path pays nothing for it
			return launch_variant<64, 16, 8, kStabBlockThreads, true>(a, stream);
		case 32: // small batches: one warp per SM (latency, not throughput)
			return launch_variant<64, 16, 8, 32>(a, stream);
		case 128:
			return launch_variant<64, 16, 8, 128>(a, stream);
		default:
			return launch_variant<64, 16, 8, kBlockThreads>(a, stream);
	}
}

} // namespace mbik
