// mbik_kernel.h -- launch interface of the solve kernel (mbik_kernel.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace mbik {

constexpr int kBlockThreads = 384; // poses per CTA: one CTA per SM, all of its warps kept in lockstep
constexpr int kStabBlockThreads = 256; // CTA size of the stabilisation variants
constexpr int kMaxStabEffectors = 32;  // effectors per root-segment list the stabilisation variants can hold

struct SolveArgs {
	const unsigned char *blob; // device copy of the rig blob (16-byte aligned)
	uint32_t blob_bytes;
	int32_t iterations;
	size_t n_poses;
	const float *targets;    // [n_poses][n_pins][12]
	const float *start_pose; // [n_poses][n_bones][12] or nullptr
	float *out_pose;         // [n_poses][n_bones][10]
	float *out_local;        // [n_poses][n_bones][12] or nullptr
	uint32_t *out_status;    // [n_poses] or nullptr
	int32_t stabilize;       // rig has stabilization_passes > 0 (reference src/ik_bone_segment_3d.cpp:163-176)
};

// index of the smallest kernel variant that fits the rig, or -1 if none does
int kernel_variant_for(int n_solved, int max_seg_len, int max_stack);
int kernel_capacity_of_variant(int variant);
cudaError_t launch_solve(const SolveArgs &args, int variant, int sm_count, cudaStream_t stream);

} // namespace mbik
