// mbik_kernel_v4.cu -- instantiations of the solve kernel for the size variant {128 solved bones, segment 128, stack 16}.
#include "mbik_kernel_body.cuh"

namespace mbik {

cudaError_t launch_v4(const SolveArgs &a, int threads, cudaStream_t stream) {
	switch (threads) {
		case 0: // stabilisation passes > 0: separate instantiation, the default path pays nothing for it
			return launch_variant<128, 128, 16, kStabBlockThreads, true>(a, stream);
		case 32: // small batches: one warp per SM (latency, not throughput)
			return launch_variant<128, 128, 16, 32>(a, stream);
		case 128:
			return launch_variant<128, 128, 16, 128>(a, stream);
		default:
			return launch_variant<128, 128, 16, kBlockThreads>(a, stream);
	}
}

} // namespace mbik
