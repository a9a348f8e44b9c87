/*
 * mbik.h -- C ABI of the B200-native batched ManyBoneIK solve loop (libmbik.so).
 *
 * This is the drop-in boundary for ONE path of the reference module: the per-frame iterative
 * constrained IK solve, i.e. the body of
 *     ManyBoneIK3D::_process_modification()            reference src/many_bone_ik_3d.cpp:645-694
 *       -> IKBoneSegment3D::segment_solver()           reference src/ik_bone_segment_3d.cpp:210-240
 * The reference has no FFI of its own (the path sits behind a C++ virtual of a Godot engine module,
 * src/many_bone_ik_3d.h:90); these entry points are what a maintainer would bind from that override
 * (see INTEGRATION.md for the stub).  Plain pointers and sizes only; no C++/torch types.
 *
 * Semantics: one call solves `n_poses` INDEPENDENT skeleton poses of the same rig.  For every pose
 * the result equals what the reference computes in the first _process_modification() after
 * _bone_list_changed() rebuilt the rig with the skeleton in `start_pose` (default: the rig's rest
 * pose) and the pin targets in `targets`.
 *
 * All transforms use Godot's in-memory Transform3D layout with real_t = float:
 *   12 floats = basis rows [xx xy xz  yx yy yz  zx zy zz] followed by origin [ox oy oz].
 */
#ifndef MBIK_H
#define MBIK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBIK_VERSION 1

/* status codes (reference error behaviour: never throws, ERR_FAIL_* -> default return;
 * src/ik_bone_segment_3d.cpp:130-133) */
#define MBIK_OK 0
#define MBIK_ERR_INVALID_ARG (-1)
#define MBIK_ERR_CUDA (-2)
#define MBIK_ERR_UNSUPPORTED (-3) /* rig beyond the schedule's index types (16 383 solved bones, walk stack depth 127) */
#define MBIK_ERR_NO_DEVICE (-4)   /* no CUDA device: there is NO CPU fallback */
#define MBIK_ERR_ALLOC (-5)

/* per-pose status bits written to out_status */
#define MBIK_POSE_NONFINITE_RESET 1u /* a solved bone's basis was non-finite and was reset to identity at
                                        write-back (reference src/ik_bone_3d.cpp:174-176) */

/* One row of ManyBoneIK3D::pins (IKEffectorTemplate3D fields, reference
 * src/ik_effector_template_3d.h:40-45). */
typedef struct mbik_pin_desc {
	int32_t bone;                     /* skeleton bone index (stands for the bone name) */
	float weight;                     /* default 0 in the reference: benches must set it */
	float motion_propagation_factor;  /* clamped to [0,1] like IKEffector3D::set_motion_propagation_factor */
	float direction_priorities[3];    /* default (0.2, 0, 0.2) */
} mbik_pin_desc;

/* One entry of ManyBoneIK3D::kusudama_open_cones[i] (Vector4: centre xyz, radius w;
 * reference src/many_bone_ik_3d.h:58). */
typedef struct mbik_cone_desc {
	float center[3];
	float radius;
} mbik_cone_desc;

/* One row of the constraint tables constraint_names / joint_twist / kusudama_open_cone_count
 * (reference src/many_bone_ik_3d.h:53-59). */
typedef struct mbik_constraint_desc {
	int32_t bone;        /* skeleton bone index */
	float twist_from;    /* joint_twist.x = min axial angle */
	float twist_range;   /* joint_twist.y = range */
	int32_t n_cones;     /* kusudama_open_cone_count */
	int32_t cone_offset; /* first cone of this row in mbik_rig_desc::cones */
} mbik_constraint_desc;

/* Everything _bone_list_changed() (reference src/many_bone_ik_3d.cpp:1011-1068) reads. */
typedef struct mbik_rig_desc {
	int32_t n_bones;                          /* Skeleton3D::get_bone_count() */
	const int32_t *parent;                    /* [n_bones], -1 = parentless; children are visited in ascending index */
	const float *rest_local;                  /* [n_bones][12] Skeleton3D::get_bone_pose() at build time */
	int32_t n_pins;
	const mbik_pin_desc *pins;
	int32_t n_constraints;
	const mbik_constraint_desc *constraints;
	const mbik_cone_desc *cones;
	int32_t n_bone_damp;                      /* size of ManyBoneIK3D::bone_damp (indexed by skeleton bone id; may be 0) */
	const float *bone_damp;
	float default_damp;                       /* radians; reference default 0.0872665 (5 deg) */
	int32_t iterations_per_frame;             /* reference default 15 */
	int32_t stabilization_passes;             /* reference default 0 */
	int32_t constraint_mode;                  /* reference default false */
} mbik_rig_desc;

/* Per-call overrides.  Negative values mean "use the rig's".  Zero-initialise the struct (`mbik_solve_params p = {0}`):
 * fields added after `stream` default to the reference's behaviour at 0. */
typedef struct mbik_solve_params {
	int32_t iterations;   /* overrides iterations_per_frame (read every frame in the reference, :685) */
	int32_t device;       /* CUDA device ordinal for this call; -1 = current device */
	uint32_t flags;       /* MBIK_IO_* | MBIK_SCHED_* | MBIK_OUT_* | MBIK_LOCAL_RECOMPOSED */
	void *stream;         /* cudaStream_t for MBIK_IO_DEVICE calls (NULL = default stream) */
	int32_t newton_iters; /* Newton-Raphson steps on the QCP characteristic polynomial before the rotation is read off the
	                         key matrix.  0 (and any negative value) = the reference: its QCP uses the upper bound
	                         (Gt + Gm) / 2 as the eigenvalue and never refines it (reference src/math/qcp.cpp:205, :215), so
	                         ONLY 0 is parity with the reference; > 0 gives the textbook Theobald fit (a different, usually
	                         larger, rotation per bone-step) for callers who want it. */
	int32_t reserved;     /* must be 0 */
} mbik_solve_params;

#define MBIK_IO_HOST 0u   /* buffers are host memory (pinned or pageable); the call copies, solves and returns when done */
#define MBIK_IO_DEVICE 1u /* buffers are device memory on params->device; the call enqueues on params->stream and returns */

/* Output layout.
 *   MBIK_OUT_SOLVED_ONLY   out_pose is [n_poses][n_solved][10]: only the bones of ManyBoneIK3D::bone_list, in bone_list
 *                          order (mbik_rig_get_bone_order) -- exactly the bones _update_skeleton_bones_transform hands to the
 *                          skeleton (reference src/many_bone_ik_3d.cpp:104-116, src/ik_bone_3d.cpp:170-179).  Bones outside
 *                          bone_list are never written by the reference, so a caller that owns a skeleton loses nothing
 *                          and the device -> host copy shrinks by n_bones / n_solved.  Values are bit-identical to the
 *                          same bones' rows of the default layout.
 *   MBIK_LOCAL_RECOMPOSED  out_local holds, for the bones of bone_list, the pose Skeleton3D::get_bone_pose() returns once
 *                          position / rotation / scale are in the skeleton -- Transform3D(Basis(rotation) * diag(scale),
 *                          position), with the non-finite reset -- i.e. what the reference's NEXT frame seeds its IK
 *                          bones from (src/many_bone_ik_3d.cpp:1084, :91-102), instead of the raw IK-bone transforms. */
#define MBIK_OUT_SOLVED_ONLY 8u
#define MBIK_LOCAL_RECOMPOSED 16u

/* Kernel mapping (default: chosen by batch size).  Both mappings run the same per-pose arithmetic in the same order
 * and return identical bits; these flags exist for tests and measurements.
 *   THROUGHPUT        one thread per pose, 512-pose CTAs in lockstep (large batches)
 *   SEGMENT_PARALLEL  32 poses per CTA, one warp per concurrently solvable segment: sibling segments of the segment
 *                     tree are independent in IKBoneSegment3D::segment_solver's post-order recursion (reference
 *                     src/ik_bone_segment_3d.cpp:210-225), so the latency of a small batch is the critical path of
 *                     the tree instead of the whole bone list; ignored for rigs whose tree is a single chain */
#define MBIK_SCHED_THROUGHPUT 2u
#define MBIK_SCHED_SEGMENT_PARALLEL 4u

typedef struct mbik_rig mbik_rig; /* opaque: host schedule + per-device copies */

/* Schedule facts, for tests and callers that size buffers. */
typedef struct mbik_rig_info {
	int32_t n_bones;          /* skeleton bones */
	int32_t n_solved;         /* bones in ManyBoneIK3D::bone_list */
	int32_t n_segments;       /* kept IKBoneSegment3D count */
	int32_t n_steps;          /* bone-steps per iteration (= n_solved) */
	int32_t n_effectors;      /* pinned bones that are solved */
	int32_t n_pins;
	int32_t max_headings;     /* largest heading count of any segment */
	int32_t n_cones;
	int32_t iterations;
	int32_t kernel_capacity;  /* solved-bone capacity of the kernel variant that will run (16383 = the unbounded variant) */
	int64_t rig_blob_bytes;   /* constants staged to shared memory per CTA */
	double flops_per_solve;   /* algorithmic flop floor, SURVEY.md section 8(d) convention */
	int32_t max_segment_len;  /* bones in the longest kept segment */
	int32_t max_walk_stack;   /* branch-point stack depth of the deepest effector walk */
	int32_t sp_roles;         /* segment-parallel schedule: warps per 32-pose group (<= 1: the tree is a chain) */
	int32_t sp_phases;        /* barrier-separated phases per iteration (= height of the segment tree) */
	double sp_gain;           /* estimated serial cost / critical-path cost of that schedule */
} mbik_rig_info;

int mbik_device_count(void);
const char *mbik_strerror(int code);
const char *mbik_last_error(void); /* thread-local detail string of the last failure */

int mbik_rig_create(const mbik_rig_desc *desc, mbik_rig **out_rig);
int mbik_rig_destroy(mbik_rig *rig);
int mbik_rig_get_info(const mbik_rig *rig, mbik_rig_info *out_info);
/* bone_list order (children segments first, tip -> root), reference src/ik_bone_segment_3d.cpp:56-72 */
int mbik_rig_get_bone_order(const mbik_rig *rig, int32_t *out_bones /* [n_solved] */);
/* heading weights of the segment that owns step `step` (reference :281-343); returns count */
int mbik_rig_get_step_weights(const mbik_rig *rig, int32_t step, double *out_weights, int32_t capacity);
/* segment-parallel schedule, one row of 6 per non-empty span: phase, warp (role), first step, end step, team size,
 * member index (0 = the warp that owns the segment; > 0 = heading helper).  Returns the row count. */
int mbik_rig_get_schedule(const mbik_rig *rig, int32_t *out_rows /* [capacity][6] */, int32_t capacity);
/* per solved bone (bone_list order): bone-direction local basis[9], twist-axes local basis[9] (setup constants) */
int mbik_rig_get_bone_frames(const mbik_rig *rig, float *out_dir_basis /* [n_solved][9] */, float *out_twist_basis /* [n_solved][9] */);
/* per cone, flattened in constraint-row order: control point[3], tangent centre 1[3], tangent centre 2[3] */
int mbik_rig_get_cone_geometry(const mbik_rig *rig, float *out /* [n_cones][9] */);

/*
 * The hot path.  Replaces the iteration loop + skeleton read/write of
 * ManyBoneIK3D::_process_modification (reference src/many_bone_ik_3d.cpp:685-693, :91-116).
 *   targets     [n_poses][n_pins][12]  skeleton-space pin targets (IKEffector3D::target_relative_to_skeleton_origin)
 *   start_pose  [n_poses][n_bones][12] local bone poses to seed from, or NULL = rig rest pose
 *   out_pose    [n_poses][n_bones][10] per bone: position xyz, rotation quaternion xyzw, scale xyz -- the three
 *               values IKBone3D::set_skeleton_bone_pose writes (reference src/ik_bone_3d.cpp:170-179);
 *               bones outside bone_list pass through
 *   out_local   [n_poses][n_bones][12] raw local transforms (optional, may be NULL)
 *   out_status  [n_poses] MBIK_POSE_* bits (optional, may be NULL)
 */
int mbik_solve_batch(mbik_rig *rig, const mbik_solve_params *params, size_t n_poses,
		const float *targets, const float *start_pose,
		float *out_pose, float *out_local, uint32_t *out_status);

/* Same, host buffers only, sharded contiguously over `n_devices` GPUs (pose k's result does not depend on
 * the split; no collective).  devices == NULL means 0..n_devices-1. */
int mbik_solve_batch_multi(mbik_rig *rig, const mbik_solve_params *params, size_t n_poses,
		const float *targets, const float *start_pose,
		float *out_pose, float *out_local, uint32_t *out_status,
		const int32_t *devices, int32_t n_devices);

/*
 * Warm-start streaming.  The reference re-seeds its IK bones from the skeleton after
 * every frame (signal modification_processed -> _update_ik_bones_transform, reference src/many_bone_ik_3d.cpp:1084,
 * :91-102), so frame f+1 starts from frame f's solution AS THE SKELETON HOLDS IT: IKBone3D::set_skeleton_bone_pose hands
 * position / rotation quaternion / scale to the skeleton (with the non-finite basis reset, src/ik_bone_3d.cpp:170-179) and
 * Skeleton3D::get_bone_pose() recomposes them, Transform3D(Basis(rotation) * diag(scale), position) (engine
 * scene/3d/skeleton_3d.cpp, not in the module's tree).  A stream keeps that state ON THE DEVICE: the recomposed local
 * transforms of all `n_poses` skeletons live in a device-resident ping-pong pair, each submitted frame uploads only
 * the targets, solves with start = previous frame's recomposed locals, and (optionally) downloads out_pose.  Frame
 * f+1's target upload and frame f's pose download overlap the solves.  Results are bit-identical to calling
 * mbik_solve_batch per frame with MBIK_LOCAL_RECOMPOSED and start_pose = the previous frame's out_local, and to the
 * reference module's own node living across frames (oracle/_ref, tests/test_reference_gpu.py).
 */
typedef struct mbik_stream mbik_stream;
/* initial_pose: host [n_poses][n_bones][12] or NULL (= the rig's rest pose for every skeleton) */
int mbik_stream_create(mbik_rig *rig, int32_t device, size_t n_poses, const float *initial_pose, mbik_stream **out_stream);
/* flags: MBIK_OUT_SOLVED_ONLY = mbik_stream_submit's out_pose is [n_poses][n_solved][10] (bone_list order) */
int mbik_stream_create_ex(mbik_rig *rig, int32_t device, size_t n_poses, const float *initial_pose, uint32_t flags, mbik_stream **out_stream);
int mbik_stream_destroy(mbik_stream *stream);
/* Enqueue one frame and return.  targets: host [n_poses][n_pins][12]; out_pose / out_status: host buffers or NULL
 * (NULL = leave the result on the device).  All three must stay valid until mbik_stream_sync(); use pinned memory
 * for true asynchrony.  iterations < 0 = the rig's iterations_per_frame. */
int mbik_stream_submit(mbik_stream *stream, const float *targets, float *out_pose, uint32_t *out_status, int32_t iterations);
int mbik_stream_sync(mbik_stream *stream); /* wait for every submitted frame; returns the first asynchronous error */
/* Synchronous read-back of the current (recomposed) local transforms [n_poses][n_bones][12] (after all submitted frames):
 * what the next frame will start from. */
int mbik_stream_read_local(mbik_stream *stream, float *out_local);
/* Re-seed every skeleton (NULL = rest pose), like _bone_list_changed() -> _update_ik_bones_transform(). */
int mbik_stream_reset(mbik_stream *stream, const float *initial_pose);
int64_t mbik_stream_frames(const mbik_stream *stream); /* frames submitted so far */

/*
 * Per-pose kusudama limit sets (limits that vary per pose).
 *
 * A limit set is one alternative fill of the rig's constraint tables -- the same rows (same bones, same cone counts),
 * other values: joint_twist from / range, cone centres and radii.  mbik_limit_sets_create runs the reference's
 * constraint authoring -- only that part of the rig flattening, not the rig's topology again -- for every set on the HOST (IKKusudama3D::_update_constraint / set_axial_limits,
 * IKLimitCone3D::update_tangent_handles: reference src/ik_kusudama_3d.cpp:37-115, src/ik_open_cone_3d.cpp:36-180 --
 * the tangent-circle construction goes through libm sin / cos / acos / tan, which only the host evaluates bit-identically
 * to the reference) and uploads the resulting cone / tangent-circle geometry and twist frames as a device table;
 * mbik_solve_batch_limits then solves pose k with set set_index[k].  Result of pose k == mbik_solve_batch on a rig created
 * with that set's constraint values, bit for bit -- in both kernel mappings and with stabilization_passes > 0.
 */
typedef struct mbik_limit_sets mbik_limit_sets;
/* constraints: [n_sets][rig n_constraints] rows in the rig's row order (bone and n_cones must equal the rig's row;
 * cone_offset indexes the set's own block of `cones`); cones: [n_sets][cones_per_set]. */
int mbik_limit_sets_create(mbik_rig *rig, int32_t n_sets, const mbik_constraint_desc *constraints, const mbik_cone_desc *cones,
		int32_t cones_per_set, mbik_limit_sets **out_sets);
/* The authoring of the sets runs on a pool of host threads (one set is one unit; MBIK_AUTHOR_THREADS overrides the
 * thread count, default = all hardware threads).  mbik_limit_sets_create returns when the table is complete;
 * mbik_limit_sets_create_async copies the caller's tables and returns at once -- the authoring overlaps whatever the
 * caller does next (e.g. solving the previous crowd); mbik_solve_batch_limits / mbik_limit_sets_wait block until it is
 * done and report its error, if any. */
int mbik_limit_sets_create_async(mbik_rig *rig, int32_t n_sets, const mbik_constraint_desc *constraints, const mbik_cone_desc *cones,
		int32_t cones_per_set, mbik_limit_sets **out_sets);
int mbik_limit_sets_wait(mbik_limit_sets *sets);
typedef struct mbik_limit_sets_info {
	int32_t n_sets;
	uint32_t bytes_per_set;  /* device table record: n_cones * 160 + n_solved * 208 */
	int64_t table_bytes;
	double author_seconds;   /* wall time of the host authoring of all sets */
	int32_t author_threads;  /* host threads it ran on */
} mbik_limit_sets_info;
int mbik_limit_sets_get_info(mbik_limit_sets *sets, mbik_limit_sets_info *out_info); /* waits for the authoring */
/* The authored geometry of one set, in the layouts of mbik_rig_get_cone_geometry (per cone of the RIG's cone table:
 * control point, tangent centres -- the triples the reference's editor gizmo draws,
 * editor/many_bone_ik_3d_gizmo_plugin.cpp:149-176) and of mbik_rig_get_bone_frames' twist output.  Either output may be
 * NULL.  Waits for the authoring; returns the cone count or a negative error. */
int mbik_limit_sets_get_geometry(mbik_limit_sets *sets, int32_t set, float *out_cones /* [n_cones][9] */, float *out_twist_basis /* [n_solved][9] */);
int mbik_limit_sets_destroy(mbik_limit_sets *sets);
/* set_index: [n_poses] int32, host or device memory like the other buffers (params->flags); values are clamped to
 * [0, n_sets).  Other arguments as mbik_solve_batch. */
int mbik_solve_batch_limits(mbik_rig *rig, mbik_limit_sets *sets, const mbik_solve_params *params, size_t n_poses, const int32_t *set_index,
		const float *targets, const float *start_pose, float *out_pose, float *out_local, uint32_t *out_status);

/* Pinned host allocation helpers (so a C caller can get full H2D/D2H bandwidth without linking CUDA). */
void *mbik_alloc_pinned(size_t bytes);
void mbik_free_pinned(void *p);

/* Device time in milliseconds of the most recent kernel launch of this rig on `device`
 * (cudaEvent pair recorded on the launching stream around the solve kernel). */
int mbik_last_kernel_ms(mbik_rig *rig, int32_t device, float *out_ms);

/* Bench helper (not on the solve path): best-of-`reps` FP32 FMA throughput of `device` in TFLOP/s, the measured
 * denominator of the kernel's FP32 roofline. */
int mbik_measure_fp32_tflops(int32_t device, int32_t reps, double *out_tflops);

/* Device self-test (not on the solve path): the kernel's grouped sqrt/division sequences against the compiler's
 * correctly rounded __fsqrt_rn/__fdiv_rn, bit for bit (sqrt exhaustively over the guarded range; division over all
 * 2^23 divisor mantissas x rounds x 64 numerators), and that the packed FP32x2 operations (two lanes per instruction)
 * equal the scalar individually rounded ones lane by lane, a product feeding a sum included (no FMA contraction).
 * *out_mismatches must come back 0. */
int mbik_selftest(int32_t device, int32_t rounds, uint64_t *out_checked, uint64_t *out_mismatches);

/*
 * Stage probes (not on the solve path): run ONE stage of the kernel's device code on caller-supplied inputs, so that
 * the reference's own unit-test vectors (tests/test_qcp.h, tests/test_ik_kusudama_3d.h) and per-stage differential
 * tests can be checked against the CUDA implementation directly, not only through whole solves.
 *   mbik_stage_qcp              QCP::weighted_superpose (reference src/math/qcp.cpp:220-248): n headings (moved = tip,
 *                               target, double weights) -> out7 = rotation quaternion xyzw + translation xyz
 *   mbik_stage_clamp            IKBoneSegment3D::clamp_to_cos_half_angle (src/ik_bone_segment_3d.cpp:97-112), n quaternions
 *   mbik_stage_point_in_limits  IKKusudama3D::get_local_point_in_limits (src/ik_kusudama_3d.cpp:273-332) with the cones of
 *                               the constraint on skeleton bone `bone` of `rig`; out[n][4] = point xyz + in_bounds
 */
int mbik_stage_qcp(int32_t device, int32_t n, const float *moved, const float *target, const double *weight, int32_t translate, float *out7);
/* the same stage with mbik_solve_params::newton_iters (0 = mbik_stage_qcp) */
int mbik_stage_qcp_newton(int32_t device, int32_t n, const float *moved, const float *target, const double *weight, int32_t translate,
		int32_t newton_iters, float *out7);
int mbik_stage_clamp(int32_t device, int32_t n, const float *quats, const double *cos_half, float *out);
int mbik_stage_point_in_limits(mbik_rig *rig, int32_t device, int32_t bone, int32_t n, const float *points, float *out);

#ifdef __cplusplus
}
#endif
#endif /* MBIK_H */
