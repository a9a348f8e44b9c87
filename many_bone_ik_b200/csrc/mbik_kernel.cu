// mbik_kernel.cu -- variant selection and launch dispatch of the solve kernel.
#include "mbik_kernel.h"

#include <cstdlib>

namespace mbik {

int kernel_variant_for(int n_solved, int max_seg_len, int max_stack, size_t blob_bytes) {
	for (int v = 0; v < kNumVariants; v++) {
		if (v == kDynVariant) { // nothing compiled fits: the unbounded variant, within the blob's index types
			return (n_solved <= kVariants[v][0] && max_stack <= kVariants[v][2]) ? v : -1;
		}
		// the small variants keep their scratch transforms in shared memory (up to 150 KiB) next to the rig blob
		if (v <= 2 && blob_bytes > 72 * 1024) {
			continue;
		}
		// rigs in the tail layout (constants beyond the shared-memory budget) need the variant that reads the walk list
		// from global memory
		if (v <= 4 && blob_bytes > kResidentBlobBudget) {
			continue;
		}
		if (n_solved <= kVariants[v][0] && max_seg_len <= kVariants[v][1] && max_stack <= kVariants[v][2]) {
			return v;
		}
	}
	return -1;
}

int kernel_capacity_of_variant(int v) { return (v >= 0 && v < kNumVariants) ? kVariants[v][0] : -1; }

// Segment-parallel mapping (mbik_solve_kernel_sp): wins while the batch is small enough that the one-thread-per-pose
// mapping leaves most warp slots of the GPU empty -- its latency is the critical path of the segment tree instead of
// the whole bone list.  Returns 0 = one thread per pose, 1 = segment-parallel with the full register budget (one
// 32-pose group per SM: a 4096-pose batch on 148 SMs), 2 = segment-parallel compiled for 128 registers so that
// floor(16 / roles) groups share an SM; past a few waves of groups the machine is full either way and the lockstep
// kernel's shared instruction stream is the better use of it (measured, humanoid22: 8192 poses 0.81 vs 1.55 ms,
// 16384 poses 1.69 vs 1.58 ms; quad80: 8192 poses 6.3 vs 11.3 ms).
int segment_parallel_choice(const SolveArgs &a, int variant, int sm_count) {
	if (a.sched_mode == 1 || a.sp_roles < (a.sched_mode == 2 ? 1 : 2) || a.sp_roles > kMaxSpRoles || variant == 2 || variant >= 5) {
		return 0;
	}
	const size_t smem = sp_smem_bytes(a);
	if (smem > 227 * 1024) {
		return 0;
	}
	const size_t groups = (a.n_poses + 31) / 32;
	static const int forced = getenv("MBIK_SP_MINB") ? atoi(getenv("MBIK_SP_MINB")) : 0; // tuning knob
	if (forced == 1 || forced == 2) {
		return (a.sched_mode == 2 || a.sp_gain >= 1.25f) ? forced : 0;
	}
	if (groups <= (size_t)sm_count) {
		return (a.sched_mode == 2 || a.sp_gain >= 1.25f) ? 1 : 0;
	}
	const size_t by_regs = 16 / (size_t)a.sp_roles, by_smem = (size_t)(227 * 1024) / (smem + 1024);
	const size_t resident = by_regs < by_smem ? by_regs : by_smem;
	if (a.sched_mode == 2) {
		return resident >= 2 ? 2 : 1;
	}
	if (a.sp_gain < 1.25f) {
		return 0;
	}
	// Larger batches, in units of the time of one wave of full-register groups: the full-register build runs
	// ceil(groups / SMs) waves, the 128-register build packs `resident` groups per SM at ~1.2x per wave (measured,
	// humanoid22: 0.81 vs 0.67 ms), and the one-thread-per-pose mapping takes ~sp_gain such units while its CTAs still
	// fit one wave (measured: humanoid22 1.55 ms up to 16k poses; quad80 11.3 ms).  Segment-parallel only with margin.
	const double waves1 = (double)((groups + sm_count - 1) / sm_count);
	const double waves2 = resident >= 2 ? 1.2 * (double)((groups + resident * sm_count - 1) / (resident * sm_count)) : 1e30;
	const double best = waves1 < waves2 ? waves1 : waves2;
	if (best >= 0.85 * (double)a.sp_gain) {
		return 0;
	}
	return waves1 <= waves2 ? 1 : 2;
}
bool uses_segment_parallel(const SolveArgs &a, int variant, int sm_count) { return segment_parallel_choice(a, variant, sm_count) != 0; }

// shared-memory scratch (segment chain + walk stack columns) a CTA of `threads` threads of `variant` asks for next to the
// rig blob -- the rule of ScratchStride in mbik_kernel_body.cuh
static size_t scratch_smem_bytes(int variant, int threads) {
	const size_t bytes = (size_t)(kVariants[variant][1] + kVariants[variant][2]) * 12 * (size_t)threads * sizeof(float);
	return bytes <= 150 * 1024 ? bytes : 0;
}
static bool cta_fits(const SolveArgs &a, int variant, int threads) {
	const size_t scr = scratch_smem_bytes(variant, threads);
	return scr == 0 ? (size_t)a.blob_bytes <= 227 * 1024 : (((size_t)a.blob_bytes + 127) & ~(size_t)127) + scr <= 227 * 1024;
}

int throughput_block_threads(const SolveArgs &a, int variant, int sm_count) {
	if (a.stabilize) {
		return 0;
	}
	static const int forced = getenv("MBIK_THREADS") ? atoi(getenv("MBIK_THREADS")) : 0; // tuning knob
	if (forced > 0) {
		return forced;
	}
	if (a.limit_table) {
		return kBlockThreads; // limit-set instantiations exist at the full CTA size only
	}
	// Small batches: one CTA per SM with as few warps as cover the batch (a 4096-pose batch runs as 128 one-warp CTAs on
	// 128 SMs instead of 8 sixteen-warp CTAs on 8 SMs) -- latency, not throughput.  Sizes whose shared-memory scratch does
	// not fit beside this rig's blob are skipped (rigs with many cones / unsolved bones on the 64-bone variants).
	const int cands_small[] = { 32, 64, 128, 256 }, cands_large[] = { 32, 128 }; // instantiated CTA sizes per variant
	const int *cands = variant <= 1 ? cands_small : cands_large;
	const int n_cands = variant <= 1 ? 4 : 2;
	for (int i = 0; i < n_cands; i++) {
		if ((a.n_poses + cands[i] - 1) / cands[i] <= (size_t)sm_count && cta_fits(a, variant, cands[i])) {
			return cands[i];
		}
	}
	// Large batches: every CTA is one SM's share of a wave (128 registers per thread: one CTA per SM), so a batch of
	// 1.73 waves of 512-thread CTAs pays for 2.  Wave-balanced size: the smallest instantiated CTA with which the same
	// number of waves covers the batch -- 131 072 poses on 148 SMs run as 2 full waves of 448 threads instead of one
	// full and one 73 % wave of 512 (BASELINE config 3: a 1M batch sharded over 8 GPUs).
	if (variant <= 3) {
		static const bool balanced = !(getenv("MBIK_WAVE_BALANCE") && atoi(getenv("MBIK_WAVE_BALANCE")) == 0);
		if (balanced) {
			const size_t per_wave = (size_t)sm_count * kBlockThreads;
			const size_t waves = (a.n_poses + per_wave - 1) / per_wave;
			const size_t need = (a.n_poses + (size_t)sm_count * waves - 1) / ((size_t)sm_count * waves);
			const int sizes[] = { 384, 448 };
			for (int t : sizes) {
				if ((size_t)t >= need && cta_fits(a, variant, t)) {
					return t;
				}
			}
		}
	}
	return kBlockThreads;
}

cudaError_t launch_solve(const SolveArgs &a, int variant, int sm_count, cudaStream_t stream) {
	if (variant < 0 || variant >= kNumVariants) {
		return cudaErrorInvalidValue;
	}
	const bool lims = a.limit_table != nullptr;
	if (lims && (!a.limit_index || a.n_limit_sets < 1)) {
		return cudaErrorInvalidValue;
	}
	if (variant == kDynVariant) {
		return launch_v6(a, sm_count, stream);
	}
	// segment-parallel mapping with per-pose limit sets: compiled without stabilisation only
	const int sp = (lims && a.stabilize) ? 0 : segment_parallel_choice(a, variant, sm_count);
	if (sp != 0) {
		switch (variant) {
			case 0:
				return lims ? launch_sp_lims_v0(a, sp, stream) : launch_sp_v0(a, sp, stream);
			case 1:
				return lims ? launch_sp_lims_v1(a, sp, stream) : launch_sp_v1(a, sp, stream);
			case 3:
				return lims ? launch_sp_lims_v3(a, sp, stream) : launch_sp_v3(a, sp, stream);
			case 4:
				return lims ? launch_sp_lims_v4(a, sp, stream) : launch_sp_v4(a, sp, stream);
			default:
				break;
		}
	}
	if (lims) {
		SolveArgs b = a;
		b.sm_count = sm_count;
		switch (variant) {
			case 0:
				return launch_lims_v0(a, stream);
			case 1:
				return launch_lims_v1(a, stream);
			case 2:
				return launch_lims_v2(b, stream);
			case 3:
				return launch_lims_v3(b, stream);
			case 4:
				return launch_lims_v4(b, stream);
			default:
				return launch_lims_v5(b, stream);
		}
	}
	const int threads = throughput_block_threads(a, variant, sm_count);
	SolveArgs b = a;
	b.sm_count = sm_count;
	switch (variant) {
		case 0:
			return launch_v0(a, threads, stream);
		case 1:
			return launch_v1(a, threads, stream);
		case 2:
			return launch_v2(b, threads, stream);
		case 3:
			return launch_v3(b, threads, stream);
		case 4:
			return launch_v4(b, threads, stream);
		default:
			return launch_v5(b, threads, stream);
	}
}

} // namespace mbik
