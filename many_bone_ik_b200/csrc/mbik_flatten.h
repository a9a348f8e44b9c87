// mbik_flatten.h -- host-side flattener: mbik_rig_desc -> topologically ordered SoA schedule (blob).
//
// Re-derives, on flat arrays, everything ManyBoneIK3D::_bone_list_changed() builds as an object graph
// (reference src/many_bone_ik_3d.cpp:1011-1068): the segment tree, bone_list order, per-segment
// effector lists and heading weights, bone-direction frames, kusudama cone/tangent geometry and twist
// frames.  Runs once per rig on the host; no GPU needed.
#pragma once
#include "../../include/mbik.h"
#include "mbik_blob.h"
#include "mbik_math.cuh"

#include <string>
#include <vector>

namespace mbik {

struct FlatSegment {
	int root_bone = -1, tip_bone = -1; // skeleton ids
	int parent_seg = -1;
	std::vector<int> bones;      // tip -> root (skeleton ids)
	std::vector<int> child_segs; // kept children, ascending skeleton child order
	bool pinned_descendants = false;
	bool kept = true;
	std::vector<int> effectors;  // skeleton ids of the effector bones, reference list order
	std::vector<double> weights; // one per heading
	int stabilize = 0;
};

struct FlatRig {
	// inputs (copied)
	int n_bones = 0;
	std::vector<int> parent;
	std::vector<X34> rest_local;
	std::vector<mbik_pin_desc> pins;
	// derived
	std::vector<FlatSegment> segments;  // creation order; dropped ones have kept=false
	std::vector<int> root_segments;
	std::vector<int> bone_order;        // bone_list (skeleton ids)
	std::vector<int> seg_of_bone;       // skeleton id -> owning kept segment or -1
	std::vector<int> topo;              // t index -> skeleton id
	std::vector<int> t_of_bone;         // skeleton id -> t index or -1
	std::vector<BlobStep> steps;
	std::vector<BlobBone> bones;        // t order
	std::vector<BlobEff> effs;
	std::vector<BlobFk> fk;
	std::vector<int16_t> chain;
	int max_seg_len = 0, max_stack = 0;
	std::vector<BlobCone> cones;        // per constraint row order of appearance on solved bones
	std::vector<int> cone_row_index;    // for mbik_rig_get_cone_geometry: desc cone index -> blob cone index or -1
	std::vector<BlobPass> pass;
	// what author_constraints needs besides the constraint rows
	std::vector<int> ik_parent;             // skeleton id -> IK parent bone (skeleton id) or -1
	std::vector<M3> setup_parent_basis;     // skeleton id -> setup-time global basis of the IK parent (constrained bones only)
	std::vector<unsigned char> blob;
	int max_headings = 0;
	int n_effectors = 0;
	int n_kept_segments = 0;
	int iterations = 15;
	int stabilization_passes = 0;
	double flops_per_solve = 0;
	// segment-parallel schedule (BlobSpan): roles = warps per pose group, cost estimates in arbitrary units
	std::vector<BlobSpan> sched;
	int sp_roles = 0, sp_phases = 0, sp_slots = 0;
	int sp_team_bufs = 0, sp_team_headings = 0;
	std::vector<int32_t> step_path;     // per step: first BlobPathRef or -1
	std::vector<BlobPathRef> path_refs;
	std::vector<int16_t> paths;
	double sp_serial_cost = 0, sp_critical_cost = 0; // all steps on one warp vs the longest role per phase, summed
	std::string error;
};

// Returns MBIK_OK or a negative error code (message in out.error).
int flatten_rig(const mbik_rig_desc *desc, FlatRig &out);

// Side tables of author_constraints, per skeleton bone id
struct ConstraintTables {
	std::vector<char> present;
	std::vector<int> cone_off, cone_cnt; // the bone's range in `cones`
	std::vector<int> cone_row_index;     // desc cone index -> blob cone index or -1
};
// Limit-dependent part of flatten_rig, callable on its own for a rig that is already flattened (`bones`: a copy of
// rig.bones; only its twist fields are written).  `desc` must name the same bones / cone counts in its constraint rows.
int author_constraints(const mbik_rig_desc *desc, const FlatRig &rig, std::vector<BlobBone> &bones, std::vector<BlobCone> &cones, ConstraintTables &tables,
		std::string &error);

// Range-checks every index of the schedule against the capacities {solved bones, segment slots, stack slots} of the
// kernel variant that will run it (the kernel itself does no bounds checking).  Returns true if consistent.
bool validate_schedule(const FlatRig &rig, int cap_bones, int cap_seg, int cap_stack, std::string &error);

} // namespace mbik
