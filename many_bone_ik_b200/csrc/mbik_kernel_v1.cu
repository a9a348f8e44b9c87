// mbik_kernel_v1.cu -- instantiations of the solve kernel for the size variant {32 solved bones, segment 8, stack 4}.
#include "mbik_kernel_body.cuh"

namespace mbik {

cudaError_t launch_v1(const SolveArgs &a, int threads, cudaStream_t stream) {
	switch (threads) {
		case 0: // stabilisation passes > 0: separate instantiation, the default path pays nothing for it
			return launch_variant<32, 8, 4, kStabBlockThreads, true>(a, stream);
		case 32:
			return launch_variant<32, 8, 4, 32>(a, stream);
		case 64:
			return launch_variant<32, 8, 4, 64>(a, stream);
		case 128:
			return launch_variant<32, 8, 4, 128>(a, stream);
		case 256:
			return launch_variant<32, 8, 4, 256>(a, stream);
		case 384: // wave-balanced sizes for large batches (launch_solve): the last wave of CTAs is as full as the others
			return launch_variant<32, 8, 4, 384>(a, stream);
		case 448:
			return launch_variant<32, 8, 4, 448>(a, stream);
		default:
			return launch_variant<32, 8, 4, kBlockThreads>(a, stream);
	}
}

} // namespace mbik
