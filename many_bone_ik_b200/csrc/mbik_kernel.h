// mbik_kernel.h -- launch interface of the solve kernel (mbik_kernel_body.cuh, instantiated in mbik_kernel_v*.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "mbik_blob.h"
#include <stdint.h>

namespace mbik {

constexpr int kBlockThreads = 512; // poses per CTA: one CTA per SM (128 registers per thread), all of its warps kept in lockstep
constexpr int kStabBlockThreads = 512; // CTA size of the stabilisation variants
constexpr int kMinStabEffectors = 32;  // effector slots of the stabilisation variants: max(this, the variant's bone capacity); unbounded variant: workspace

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
	// host-side launch hints (the kernel reads the schedule itself from the blob header)
	int32_t n_solved;        // solved bones of the rig
	int32_t sp_roles;        // warps per pose group of the rig's segment-parallel schedule (<= 1: none)
	int32_t sp_team_bytes;   // shared memory of the heading buffers of its teams (per group)
	float sp_gain;           // estimated serial / critical-path cost of that schedule
	long long *sp_trace;     // debug (MBIK_SP_TRACE=1): per (iteration, phase, warp) start / end clock of CTA 0, or nullptr
	int32_t sched_mode;      // 0 = choose by batch size, 1 = one thread per pose (throughput mapping), 2 = segment-parallel
	// per-pose kusudama limit sets (mbik_solve_batch_limits): n_limit_sets records of limit_stride bytes, each
	// [BlobCone x n_cones][BlobBone x n_solved] built by the host flattener; limit_index[pose] picks one.  nullptr = the
	// rig's own limits from the blob.
	const unsigned char *limit_table = nullptr;
	const int32_t *limit_index = nullptr;
	uint32_t limit_stride = 0;
	int32_t n_limit_sets = 0;
	// unbounded-rig variant only (solve_body DYN): sizes of the rig for the launch loop, and this launch's workspace
	int32_t n_bones = 0, n_pins = 0, max_seg_len = 0, max_stack = 0;
	float *workspace = nullptr;
	// streamed-walk instantiation (solve_body GLW): chosen by the host for rigs with long effector walks; sm_count sizes its workspace
	int32_t use_glw = 0, sm_count = 0;
	int32_t max_list_effs = 0; // longest effector list of the rig (stabilisation scratch of the unbounded variant)
	int32_t newton_iters = 0;  // mbik_solve_params::newton_iters (0 = the reference's QCP: no eigenvalue refinement)
	uint32_t out_flags = 0;    // OUT_* below
};
enum : uint32_t {
	OUT_COMPACT = 1u,          // out_pose = [n_poses][n_solved][10] in bone_list order (MBIK_OUT_SOLVED_ONLY)
	OUT_LOCAL_RECOMPOSED = 2u, // out_local = Skeleton3D::get_bone_pose() after the write-back (MBIK_LOCAL_RECOMPOSED)
};

// Compiled size variants {solved-bone capacity, longest segment, walk-stack depth}; per-pose thread-local state is
// NB*12 + NSEG*12 + NSTK*12 floats, so tight variants keep more of it in L1/L2.
// The last entry is the unbounded variant (solve_body DYN: nothing sized at compile time); its limits are those of the blob's
// index types (int16 bone indices, int8 stack slots).
constexpr int kNumVariants = 7;
constexpr int kDynVariant = 6;
constexpr int kVariants[kNumVariants][3] = { { 20, 4, 2 }, { 32, 8, 4 }, { 64, 8, 1 }, { 64, 16, 8 }, { 128, 128, 16 }, { 256, 256, 32 }, { 16383, 16383, 127 } };

// index of the smallest kernel variant that fits the rig, or -1 if none does
int kernel_variant_for(int n_solved, int max_seg_len, int max_stack, size_t blob_bytes);
int kernel_capacity_of_variant(int variant);
cudaError_t launch_solve(const SolveArgs &args, int variant, int sm_count, cudaStream_t stream);

// one translation unit per variant (parallel compilation); threads == 0 selects the stabilisation instantiation
cudaError_t launch_v0(const SolveArgs &a, int threads, cudaStream_t stream);
cudaError_t launch_v1(const SolveArgs &a, int threads, cudaStream_t stream);
cudaError_t launch_v2(const SolveArgs &a, int threads, cudaStream_t stream);
cudaError_t launch_v3(const SolveArgs &a, int threads, cudaStream_t stream);
cudaError_t launch_v4(const SolveArgs &a, int threads, cudaStream_t stream);
cudaError_t launch_v5(const SolveArgs &a, int threads, cudaStream_t stream);

// per-pose limit sets (mbik_kernel_l*.cu): thread-per-pose mapping, 512-thread CTAs (stabilisation: its own instantiation)
cudaError_t launch_lims_v0(const SolveArgs &a, cudaStream_t stream);
cudaError_t launch_lims_v1(const SolveArgs &a, cudaStream_t stream);
cudaError_t launch_lims_v2(const SolveArgs &a, cudaStream_t stream);
cudaError_t launch_lims_v3(const SolveArgs &a, cudaStream_t stream);
cudaError_t launch_lims_v4(const SolveArgs &a, cudaStream_t stream);
cudaError_t launch_lims_v5(const SolveArgs &a, cudaStream_t stream);
// unbounded-rig variant (mbik_kernel_v6.cu): plain / stabilisation / limit sets / both
cudaError_t launch_v6(const SolveArgs &a, int sm_count, cudaStream_t stream);

// segment-parallel instantiations (mbik_kernel_sp*.cu); min_groups_per_sm 1 = full register budget, 2 = 128 registers
// (several 32-pose groups per SM)
cudaError_t launch_sp_v0(const SolveArgs &a, int min_groups_per_sm, cudaStream_t stream);
cudaError_t launch_sp_v1(const SolveArgs &a, int min_groups_per_sm, cudaStream_t stream);
cudaError_t launch_sp_v3(const SolveArgs &a, int min_groups_per_sm, cudaStream_t stream);
cudaError_t launch_sp_v4(const SolveArgs &a, int min_groups_per_sm, cudaStream_t stream);
// ... with per-pose limit sets (mbik_kernel_l*.cu)
cudaError_t launch_sp_lims_v0(const SolveArgs &a, int min_groups_per_sm, cudaStream_t stream);
cudaError_t launch_sp_lims_v1(const SolveArgs &a, int min_groups_per_sm, cudaStream_t stream);
cudaError_t launch_sp_lims_v3(const SolveArgs &a, int min_groups_per_sm, cudaStream_t stream);
cudaError_t launch_sp_lims_v4(const SolveArgs &a, int min_groups_per_sm, cudaStream_t stream);
// CTA size launch_solve picks for the one-thread-per-pose mapping (0 = the stabilisation instantiation); exported for tests
int throughput_block_threads(const SolveArgs &a, int variant, int sm_count);
// shared memory of one 32-pose group: rig blob + the group's local poses (n_solved x 12 words x 32 lanes) + team buffers
inline size_t sp_smem_bytes(const SolveArgs &a) {
	return (((size_t)a.blob_bytes + 127) & ~(size_t)127) + (size_t)a.n_solved * 12 * 32 * sizeof(float) + (size_t)a.sp_team_bytes;
}
// true if launch_solve would run the segment-parallel mapping for these arguments
bool uses_segment_parallel(const SolveArgs &a, int variant, int sm_count);

// stage probes (mbik_selftest.cu): device pointers in, device pointers out
struct BlobCone;
cudaError_t launch_stage_qcp(int n, const float *d_moved, const float *d_target, const double *d_weight, int translate, int newton_iters, float *d_out7);
cudaError_t launch_stage_clamp(int n, const float *d_quats, const double *d_cos_half, float *d_out);
cudaError_t launch_stage_point_in_limits(const BlobCone *d_cones, int n_cones, int n, const float *d_points, float *d_out);

} // namespace mbik
