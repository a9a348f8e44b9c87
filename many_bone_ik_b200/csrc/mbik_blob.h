// mbik_blob.h -- the flattened rig ("schedule") shared by the host flattener and the solve kernel.
//
// One contiguous, 16-byte aligned blob of rig constants.  The kernel stages it into shared memory
// with one TMA bulk copy per CTA; every lane then reads the same words (broadcast).  Layout:
//   BlobHeader | BlobStep[n_steps] | BlobBone[n_solved] | BlobEff[n_effs] | BlobFk[n_fk] |
//   BlobCone[n_cones] | BlobPass[n_pass] | chain[n_chain] (int16) | list_row[n_solved] (int16) | rest_local[n_bones*12] | BlobSpan[sp_phases*sp_slots*sp_roles] |
//   step_path[n_steps] (int32) | BlobPathRef[] | paths[] (int16)
// Tail layout (large rigs, BlobHeader::resident_bytes < total_bytes): the walk list BlobFk[] -- quadratic in the depth of a
// chain -- moves behind paths[] and is read from global memory by the {256, 256, 32} kernel variant; the segment-parallel
// tables before it are not staged either (that variant has no segment-parallel build).
// Solved bones are renumbered in depth-first preorder ("t index": parents before children, children in
// the reference's ascending order), which is also the order in which a segment's effector list
// enumerates its effectors.
#pragma once
#include <stdint.h>

namespace mbik {

enum : uint32_t {
	STEP_TRANSLATE = 1u,    // bone belongs to a root segment (reference src/ik_bone_segment_3d.cpp:217-223)
	STEP_NODE_PARENT = 2u,  // the bone's IKNode3D has a parent node (false only for non-last skeleton roots)
	STEP_IK_PARENT = 4u,    // the bone has an IKBone3D parent -> constraint snaps apply (:155)
	STEP_SWING = 8u,        // kusudama orientationally constrained
	STEP_TWIST = 16u,       // kusudama axially constrained
	STEP_SEG_ROOT = 32u,    // bone == segment root (previous_deviation reset, :178-180)
	STEP_STABILIZE = 64u,   // segment runs the stabilisation loop (root segments only, many_bone_ik_3d.cpp:1021)
	STEP_SEG_FIRST = 128u,  // first step of its segment (the tip): (re)build the segment's parent-global chain
	STEP_SELF_EFF = 256u,   // the solved bone is itself the first effector of the list (pinned segment tip)
	STEP_PUSH_SELF = 512u,  // the solved bone is a branch point of its walk: its global goes to stack slot 0
	STEP_PLAIN_FRAMES = 1024u, // orientation-axes basis is exactly the identity and the bone-direction basis is finite with
	                           // |entries| <= 2: enables the finite-operand shortcuts of the swing snap (mbik_kernel_body.cuh)
};

// shared-memory budget of the staged rig constants (the rest of the 227 KiB is scratch / alignment)
constexpr uint32_t kResidentBlobBudget = 200u * 1024u;

struct BlobHeader {
	uint32_t magic;       // 'MBIK'
	uint32_t total_bytes; // multiple of 16
	int32_t n_bones, n_solved, n_steps, n_pins, n_effs, n_fk, n_cones, n_pass;
	int32_t iterations, constraint_mode, stabilization_passes, n_chain;
	int32_t max_seg_len, max_stack;
	// segment-parallel schedule (see BlobSpan): sibling segments of the segment tree are independent, so a group of
	// `sp_roles` warps can solve them concurrently; sp_roles <= 1 means the tree offers no parallelism
	int32_t sp_roles, sp_phases, sp_slots;
	int32_t sp_team_bufs;     // heading buffers a group needs (teams that can be active in one phase, 0..kMaxSpTeams)
	int32_t sp_team_headings; // headings per buffer (largest heading list of any team step)
	uint32_t resident_bytes; // bytes the kernel stages into shared memory: total_bytes, or -- "tail layout", rigs whose constants exceed
	                         // the shared-memory budget -- the offset of the tail (segment-parallel tables, walk list) that stays in global memory
	uint32_t off_steps, off_bones, off_effs, off_fk, off_cones, off_pass, off_rest, off_chain, off_sched;
	uint32_t off_step_path, off_path_refs, off_paths; // BlobPathRef tables of the team steps (see BlobSpan)
	uint32_t off_list_row; // int16 list_row[n_solved]: index of solved bone t in bone_list (row of the compact output layout)
	int32_t glw_keep_lo, glw_keep_n; // streamed-walk instantiation: the t range of solved bones kept in L2 with evict_last (BONE_L2_KEEP)
};

struct BlobStep { // 64 bytes
	int32_t bone;    // t index
	int32_t parent;  // t index of the IK parent, -1 for a skeleton root
	uint32_t flags;
	int32_t eff_off, eff_cnt;   // the owning segment's effector list
	int32_t fk_off, fk_cnt;     // walk below `bone` that visits this segment's effector bones in list order
	int32_t cone_off, cone_cnt; // the bone's kusudama cones
	int32_t n_headings;
	double cos_half_damp;       // cos(damp / 2.0), damp per src/ik_bone_segment_3d.cpp:229-237
	int32_t pslot;              // slot of the segment chain buffer holding the global of this bone's parent
	int32_t chain_off, chain_cnt; // STEP_SEG_FIRST: t indices skeleton-root .. parent(tip); the last seg_len are kept
	int32_t seg_len;
};

// BlobBone::flags bit / BlobFk::pad bit 0: the bone is one of the most-read local poses of the rig's effector walks, few enough to stay
// in L2 for a whole resident batch -- the streamed-walk instantiation reads and writes it with an L2 evict_last policy and
// everything else with evict_first (solve_body GLW)
constexpr uint32_t BONE_L2_KEEP = 0x10000u;

struct BlobBone { // per solved bone, t order; 208 bytes; every 3x3 is padded to 12 floats so that it starts on a
                  // 16-byte boundary and is read from shared memory with three 128-bit loads
	int32_t skel_bone;
	int32_t parent;          // t index or -1
	uint32_t flags;          // STEP_NODE_PARENT only
	float twist_cos;         // twist_half_range_half_cos = cos(range / 4)
	float dir_basis[12];     // bone-direction node local basis (origin is zero)
	float orient_basis[12];  // constraint-orientation node local basis (identity after a rebuild)
	float twist_basis[12];   // constraint-twist node local basis (after _update_constraint)
	float twist_center[12];  // Basis(twist_center_rot)
};

struct BlobEff { // one entry per (segment, effector); 80 bytes
	int32_t bone;            // t index of the effector's bone
	int32_t pin;             // row of the pins table = target slot
	float prio[3];           // direction priorities (axis used iff > 0)
	int32_t n_headings;      // 1 + 2 * (#priorities > 0)
	double w_origin;         // heading weight of the origin heading
	double w_axis[3];        // heading weight of each axis pair (0 when unused)
	float w_origin_f;        // (float)w_origin, (float)w_axis[i]: the narrowed weights the reference multiplies
	float w_axis_f[3];       // headings by (real_t w = p_weights->get(index), src/ik_effector_3d.cpp:103)
	int32_t h_off;           // index of the effector's first heading in its segment's heading list
	int32_t pad;
};

// One step of the depth-first walk from the solved bone down to the effectors of its segment's list:
//   run = (src_slot >= 0 ? stack[src_slot] : run) * L[child];  if (push_slot >= 0) stack[push_slot] = run;
//   if (eff >= 0) the headings of effector eff_off + eff are built from `run`.
// A chain needs no stack at all; a branch point is pushed once and re-read for each further child.
struct BlobFk { // 8 bytes
	int16_t child;
	int8_t src_slot, push_slot;
	int16_t eff;
	int16_t pad; // bit 0: the child is a BONE_L2_KEEP bone; bits 1..: length of the plain run that starts here -- consecutive ops
	             // without stack traffic whose children are consecutive t indices, only the last of which may reach an effector
	             // (0: this op reads or writes the stack)
};

struct BlobCone { // 160 bytes
	float cp[3];            // control point (what get_control_point() returns)
	float ncp[3];           // cp.normalized() (closest_to_cone re-normalises, src/ik_open_cone_3d.cpp:360)
	float sin_half_r, cos_half_r; // sin/cos((float)radius * 0.5f) for get_quaternion_axis_angle(axis, radius)
	double radius_cos;      // cos(radius) in double
	// data about the path to the NEXT cone (unused on the last cone)
	double tan_cos;         // cos(tangent circle radius)
	float tc1[3], tc2[3];   // tangent circle centres
	float sin_half_t, cos_half_t; // sin/cos((float)tangent_radius * 0.5f)
	float c1xc2[3];         // cp x next.cp                      (not normalised, :287)
	float c1xt1[3];         // (cp x tc1).normalized()           (:290)
	float t1xc2[3];         // (tc1 x next.cp).normalized()      (:291)
	float t2xc1[3];         // (tc2 x cp).normalized()           (:305)
	float c2xt2[3];         // (next.cp x tc2).normalized()      (:306)
	float pad[5];
};

// Segment-parallel schedule: spans[(phase * sp_slots + slot) * sp_roles + role] = the step range [s0, s1) that warp
// `role` of a pose group runs in that slot of that phase (s0 == s1: idle).  A span is always one whole kept segment
// (segment_solver's post-order recursion, reference src/ik_bone_segment_3d.cpp:210-225: a segment only reads its
// ancestors' bones, which no segment of the same or an earlier phase writes, and the bones of its own subtree, which
// are final once its child segments -- all in earlier phases -- are done).  All roles meet at a barrier after each phase.
constexpr int kMaxSpRoles = 8; // warps per pose group of the segment-parallel kernel

// Teams: when a phase leaves warps idle, they join a busy segment as heading helpers.  Every member of a team runs
// the span's steps; in each step member m walks to effectors m, m + team, ... of the segment's list (path tables
// below), writes their raw tip / target headings to the team's shared-memory buffer `buf`, and the owner (member 0)
// alone folds them into the QCP sums -- in list order, so the arithmetic is unchanged -- and finishes the step.
constexpr int kMaxSpTeams = 2;
struct BlobSpan { // 8 bytes
	int16_t s0, s1;
	int8_t team;   // members of the team running this span (1 = the owner alone, the plain path)
	int8_t member; // 0 = owner
	int8_t buf;    // heading buffer / named barrier of the team
	int8_t pad;
};
// step_path[s] = index of the first BlobPathRef of step s (one per effector of its list) or -1 when s is no team step;
// paths[off .. off+cnt) = t indices of the bones from below the solved bone down to the effector's bone
struct BlobPathRef { // 8 bytes
	int32_t off, cnt;
};

struct BlobPass { // skeleton bones outside bone_list: copied through to the output
	int32_t skel_bone;
};

static_assert(sizeof(BlobStep) == 64, "BlobStep layout");
static_assert(sizeof(BlobFk) == 8, "BlobFk layout");
static_assert(sizeof(BlobBone) == 208, "BlobBone layout");
static_assert(sizeof(BlobEff) == 80, "BlobEff layout");
static_assert(sizeof(BlobCone) == 160, "BlobCone layout");

} // namespace mbik
