// mbik_kernel_v0.cu -- instantiations of the solve kernel for the size variant {20 solved bones, segment 4, stack 2}.
#include "mbik_kernel_body.cuh"

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
			return launch_variant<20, 4, 2, 512>(a, stream);
		default:
			return launch_variant<20, 4, 2, kBlockThreads>(a, stream);
	}
}

} // namespace mbik
