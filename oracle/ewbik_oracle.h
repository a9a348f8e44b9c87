// TEST INFRASTRUCTURE ONLY -- CPU oracle for the ManyBoneIK solve loop (see godot_math.h header).
//
// A function-by-function RESTATEMENT (not a copy) of the reference module's solve path and of the
// setup path that shapes its constants.  Each method cites the reference file:line it follows
// (paths relative to /root/reference).  The object graph, the lazy dirty-flag transform cache and
// the per-step allocations are kept on purpose: this is both the parity checker and the honest
// "reference CPU cost" baseline.
//
// Pinning status: (1) the numeric doctest cases of the reference (tests/test_qcp.h,
// tests/test_ik_node_3d.h, tests/test_ik_kusudama_3d.h) are restated in oracle/kat_main.cpp and
// pass against this code.  (2) The MODULE LOGIC restated here is pinned against the reference's own
// code: oracle/_ref/libmbik_ref.so is every translation unit of /root/reference/src compiled
// unmodified over a stand-in of the Godot engine headers (oracle/godot_shim/, driver
// oracle/ref_harness.cpp, `make ref`); this restatement is bit-identical to it -- end-to-end
// solves, warm-started frames, setup facts, heading weights, stage functions -- on every benchmark,
// edge and random rig (tests/test_reference_cpu.py; frozen as tests/golden/reference_solves.npz).
// (3) What stays UNPINNED is the engine arithmetic underneath both (godot_math.h): the Godot engine
// is not in the reference tree, so both run on the same restatement of core/math.
#pragma once
#include "godot_math.h"

#include <list>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace orc {
using namespace gd;

template <class T>
using Ref = std::shared_ptr<T>;
typedef int BoneId;
typedef std::vector<Vector3> PackedVector3Array;

// ---- src/math/ik_node_3d.{h,cpp} -------------------------------------------------------------
class IKNode3D : public std::enable_shared_from_this<IKNode3D> {
	enum TransformDirty { DIRTY_NONE = 0, DIRTY_VECTORS = 1, DIRTY_LOCAL = 2, DIRTY_GLOBAL = 4 };
	mutable Transform3D global_transform;
	mutable Transform3D local_transform;
	mutable int dirty = DIRTY_NONE;
	std::weak_ptr<IKNode3D> parent;
	std::list<Ref<IKNode3D>> children;

public:
	void _propagate_transform_changed();
	void set_transform(const Transform3D &p_transform);
	void set_global_transform(const Transform3D &p_transform);
	Transform3D get_transform() const;
	Transform3D get_global_transform() const;
	void set_parent(Ref<IKNode3D> p_parent);
	Ref<IKNode3D> get_parent() const { return parent.lock(); }
	Vector3 to_local(const Vector3 &p_global) const;
	Vector3 to_global(const Vector3 &p_local) const;
	void rotate_local_with_global(const Basis &p_basis, bool p_propagate = false);
	void cleanup();
};

// ---- src/ik_ray_3d.{h,cpp} -------------------------------------------------------------------
class IKRay3D {
	Vector3 point_1, point_2, working_vector;

public:
	IKRay3D() {}
	IKRay3D(Vector3 p_p1, Vector3 p_p2);
	Vector3 get_heading();
	void elongate(real_t amt);
	Vector3 get_intersects_plane(Vector3 ta, Vector3 tb, Vector3 tc);
	int intersects_sphere(Vector3 sphereCenter, real_t radius, Vector3 *S1, Vector3 *S2);
	int intersects_sphere(Vector3 rp1, Vector3 rp2, real_t radius, Vector3 *S1, Vector3 *S2);
	Vector3 plane_intersect_test(Vector3 ta, Vector3 tb, Vector3 tc, Vector3 *uvw);
	void set_point_1(Vector3 in) { point_1 = in; }
	void set_point_2(Vector3 in) { point_2 = in; }
	Vector3 get_point_1() { return point_1; }
	Vector3 get_point_2() { return point_2; }
};

// ---- src/ik_open_cone_3d.{h,cpp} -------------------------------------------------------------
class IKKusudama3D;
class IKLimitCone3D {
public:
	Vector3 control_point = Vector3(0, 1, 0);
	double radius_cosine = 0;
	double radius = 0;
	std::weak_ptr<IKKusudama3D> parent_kusudama;
	Vector3 tangent_circle_center_next_1;
	Vector3 tangent_circle_center_next_2;
	double tangent_circle_radius_next = 0;
	double tangent_circle_radius_next_cos = 0;

	void set_attached_to(Ref<IKKusudama3D> k) { parent_kusudama = k; }
	Ref<IKKusudama3D> get_attached_to() { return parent_kusudama.lock(); }
	void update_tangent_handles(Ref<IKLimitCone3D> p_next);
	void set_tangent_circle_radius_next(double rad);
	void set_tangent_circle_center_next_1(Vector3 point) { tangent_circle_center_next_1 = point.normalized(); }
	void set_tangent_circle_center_next_2(Vector3 point) { tangent_circle_center_next_2 = point.normalized(); }
	Vector3 get_control_point() const { return control_point; }
	void set_control_point(Vector3 p_control_point);
	double get_radius() const { return radius; }
	double get_radius_cosine() const { return radius_cosine; }
	void set_radius(double p_radius);
	static Vector3 get_orthogonal(Vector3 p_in);
	Vector3 get_on_great_tangent_triangle(Ref<IKLimitCone3D> next, Vector3 input) const;
	Vector3 closest_to_cone(Vector3 input, std::vector<double> *in_bounds) const;
	Vector3 _closest_cone(Ref<IKLimitCone3D> next, Vector3 input) const;
	Vector3 _get_on_path_sequence(Ref<IKLimitCone3D> next, Vector3 input) const;
	Vector3 get_closest_path_point(Ref<IKLimitCone3D> next, Vector3 input) const;
};

// ---- src/ik_kusudama_3d.{h,cpp} --------------------------------------------------------------
class IKKusudama3D : public std::enable_shared_from_this<IKKusudama3D> {
public:
	std::vector<Ref<IKLimitCone3D>> open_cones;
	Quaternion twist_min_rot;
	Vector3 twist_min_vec, twist_max_vec, twist_center_vec;
	Quaternion twist_center_rot, twist_max_rot;
	real_t twist_half_range_half_cos = 0;
	real_t min_axial_angle = 0.0;
	real_t range_angle = (real_t)Math_TAU;
	bool orientationally_constrained = false;
	bool axially_constrained = false;
	Ref<IKRay3D> bone_ray = Ref<IKRay3D>(new IKRay3D());
	Ref<IKRay3D> constrained_ray = Ref<IKRay3D>(new IKRay3D());
	// instrumentation (not in the reference): counts for the flop model / parity flags
	long n_swing_calls = 0, n_swing_rectified = 0;

	void _update_constraint(Ref<IKNode3D> p_limiting_axes);
	void update_tangent_radii();
	static void get_swing_twist(Quaternion p_rotation, Vector3 p_axis, Quaternion &r_swing, Quaternion &r_twist);
	static Quaternion get_quaternion_axis_angle(const Vector3 &p_axis, real_t p_angle);
	void snap_to_orientation_limit(Ref<IKNode3D> bone_direction, Ref<IKNode3D> to_set, Ref<IKNode3D> limiting_axes, real_t p_dampening, real_t p_cos_half_angle_dampen);
	void set_axial_limits(real_t min_angle, real_t in_range);
	void set_snap_to_twist_limit(Ref<IKNode3D> p_bone_direction, Ref<IKNode3D> p_to_set, Ref<IKNode3D> p_constraint_axes, real_t p_dampening, real_t p_cos_half_dampen);
	Vector3 get_local_point_in_limits(Vector3 in_point, std::vector<double> *in_bounds);
	void add_open_cone(Ref<IKLimitCone3D> p_cone);
	void remove_open_cone(Ref<IKLimitCone3D> limitCone);
	void clear_open_cones() { open_cones.clear(); }
	bool is_axially_constrained() { return axially_constrained; }
	bool is_orientationally_constrained() { return orientationally_constrained; }
	void enable_orientational_limits() { orientationally_constrained = true; }
	void enable_axial_limits() { axially_constrained = true; }
};

// ---- src/math/qcp.{h,cpp} --------------------------------------------------------------------
class QCP {
	double eigenvector_precision = 1E-6;
	PackedVector3Array target, moved;
	std::vector<double> weight;
	double w_sum = 0;
	Vector3 target_center, moved_center;
	double sum_xy = 0, sum_xz = 0, sum_yx = 0, sum_yz = 0, sum_zx = 0, sum_zy = 0;
	double sum_xx_plus_yy = 0, sum_zz = 0, max_eigenvalue = 0, sum_yz_minus_zy = 0, sum_xz_minus_zx = 0, sum_xy_minus_yx = 0;
	double sum_xx_minus_yy = 0, sum_xy_plus_yx = 0, sum_xz_plus_zx = 0;
	double sum_yy = 0, sum_xx = 0, sum_yz_plus_zy = 0;
	bool transformation_calculated = false, inner_product_calculated = false;
	void inner_product(PackedVector3Array &coords1, PackedVector3Array &coords2);
	Quaternion calculate_rotation();
	void set(PackedVector3Array &p_moved, PackedVector3Array &p_target, std::vector<double> &p_weight, bool p_translate);
	static void translate(Vector3 r_translate, PackedVector3Array &r_x);
	Vector3 move_to_weighted_center(PackedVector3Array &r_to_center, std::vector<double> &r_weight);

public:
	QCP(double p_evec_prec) { eigenvector_precision = p_evec_prec; }
	Quaternion weighted_superpose(PackedVector3Array &p_moved, PackedVector3Array &p_target, std::vector<double> &p_weight, bool translate);
	Quaternion get_rotation();
	Vector3 get_translation();
};

// ---- stand-ins for engine objects at the boundary --------------------------------------------
// Skeleton3D (engine, not in tree): only the queries the module makes.
struct Skeleton3D {
	std::vector<int> parent;
	std::vector<Transform3D> pose; // get_bone_pose(): local pose
	int get_bone_count() const { return (int)parent.size(); }
	int get_bone_parent(int b) const { return parent[b]; }
	std::vector<int> get_bone_children(int b) const; // ascending bone index (engine _update_process_order)
	std::vector<int> get_parentless_bones() const; // ascending
	Transform3D get_bone_pose(int b) const { return pose[b]; }
	Transform3D get_bone_global_pose(int b) const; // parent.global * pose
};

// src/ik_effector_template_3d.h:40-45 (data only; the Resource wrapper is out of scope)
struct IKEffectorTemplate3D {
	int bone = -1; // stands for the bone *name*
	real_t motion_propagation_factor = 1.0f;
	real_t weight = 0.0f;
	Vector3 priority_direction = Vector3(0.2f, 0.0f, 0.2f);
};

class IKBone3D;
class ManyBoneIK3D;

// ---- src/ik_effector_3d.{h,cpp} --------------------------------------------------------------
class IKEffector3D {
public:
	std::weak_ptr<IKBone3D> for_bone_w; // Ref<IKBone3D> in the reference (a cycle there; weak here so the graph frees)
	IKBone3D *for_bone = nullptr;
	Transform3D target_relative_to_skeleton_origin;
	real_t weight = 0.0;
	real_t motion_propagation_factor = 0.0;
	Vector3 direction_priorities;
	int pin_index = -1; // which row of the pins table feeds the target (boundary input)

	void set_weight(real_t w) { weight = w; }
	real_t get_weight() const { return weight; }
	void set_direction_priorities(Vector3 p) { direction_priorities = p; }
	Vector3 get_direction_priorities() const { return direction_priorities; }
	float get_motion_propagation_factor() const { return motion_propagation_factor; }
	void set_motion_propagation_factor(float f);
	int32_t update_effector_target_headings(PackedVector3Array *p_headings, int32_t p_index, Ref<IKBone3D> p_for_bone, const std::vector<double> *p_weights) const;
	int32_t update_effector_tip_headings(PackedVector3Array *p_headings, int32_t p_index, Ref<IKBone3D> p_for_bone) const;
};

// ---- src/ik_bone_3d.{h,cpp} ------------------------------------------------------------------
class IKBone3D : public std::enable_shared_from_this<IKBone3D> {
public:
	BoneId bone_id = -1;
	std::weak_ptr<IKBone3D> parent_w; // Ref<> in the reference
	std::vector<Ref<IKBone3D>> children;
	Ref<IKEffector3D> pin;
	float default_dampening = (float)Math_PI;
	float dampening = (float)Math_PI;
	float cos_half_dampen = 0;
	double stiffness = 0.0; // inert in the reference
	Ref<IKKusudama3D> constraint;
	Ref<IKNode3D> constraint_orientation_transform = Ref<IKNode3D>(new IKNode3D());
	Ref<IKNode3D> constraint_twist_transform = Ref<IKNode3D>(new IKNode3D());
	Ref<IKNode3D> godot_skeleton_aligned_transform = Ref<IKNode3D>(new IKNode3D());
	Ref<IKNode3D> bone_direction_transform = Ref<IKNode3D>(new IKNode3D());

	IKBone3D(BoneId p_bone, Skeleton3D *p_skeleton, const Ref<IKBone3D> &p_parent, std::vector<IKEffectorTemplate3D> &p_pins, float p_default_dampening, ManyBoneIK3D *p_many_bone_ik);
	void init_parent(const Ref<IKBone3D> &p_parent) { if (p_parent) set_parent(p_parent); }
	void set_parent(const Ref<IKBone3D> &p_parent);
	Ref<IKBone3D> get_parent() const { return parent_w.lock(); }
	void update_default_bone_direction_transform(Skeleton3D *p_skeleton);
	Ref<IKEffector3D> get_pin() const { return pin; }
	bool is_pinned() const { return (bool)pin; }
	void set_pose(const Transform3D &t) { godot_skeleton_aligned_transform->set_transform(t); }
	Transform3D get_pose() const { return godot_skeleton_aligned_transform->get_transform(); }
	void set_global_pose(const Transform3D &p_transform);
	Transform3D get_global_pose() const { return godot_skeleton_aligned_transform->get_global_transform(); }
	Transform3D get_bone_direction_global_pose() const { return bone_direction_transform->get_global_transform(); }
	void set_initial_pose(Skeleton3D *p_skeleton);
	float get_cos_half_dampen() const { return cos_half_dampen; }
	Ref<IKKusudama3D> get_constraint() const { return constraint; }
	void add_constraint(Ref<IKKusudama3D> c) { constraint = c; }
	Ref<IKNode3D> get_ik_transform() { return godot_skeleton_aligned_transform; }
	Ref<IKNode3D> get_constraint_orientation_transform() { return constraint_orientation_transform; }
	Ref<IKNode3D> get_constraint_twist_transform() { return constraint_twist_transform; }
	Ref<IKNode3D> get_bone_direction_transform() { return bone_direction_transform; }
	bool is_orientationally_constrained() { return constraint ? constraint->is_orientationally_constrained() : false; }
	bool is_axially_constrained() { return constraint ? constraint->is_axially_constrained() : false; }
	BoneId get_bone_id() const { return bone_id; }
};

// ---- src/ik_bone_segment_3d.{h,cpp} ----------------------------------------------------------
class IKBoneSegment3D : public std::enable_shared_from_this<IKBoneSegment3D> {
public:
	Ref<IKBone3D> root;
	Ref<IKBone3D> tip;
	std::vector<Ref<IKBone3D>> bones;
	std::vector<Ref<IKBone3D>> pinned_bones;
	std::vector<Ref<IKBoneSegment3D>> child_segments;
	std::weak_ptr<IKBoneSegment3D> parent_segment; // Ref<> in the reference
	IKBoneSegment3D *root_segment = nullptr;
	std::vector<Ref<IKEffector3D>> effector_list;
	PackedVector3Array target_headings, tip_headings, tip_headings_uniform;
	std::vector<double> heading_weights;
	Skeleton3D *skeleton = nullptr;
	bool pinned_descendants = false;
	double previous_deviation = INFINITY;
	int32_t default_stabilizing_pass_count = 0;
	std::map<BoneId, Ref<IKBone3D>> bone_map;
	const double evec_prec = static_cast<double>(1E-6);
	// instrumentation (not in the reference)
	long n_bone_steps = 0;

	IKBoneSegment3D(Skeleton3D *p_skeleton, BoneId p_root_bone_name, std::vector<IKEffectorTemplate3D> &p_pins, ManyBoneIK3D *p_many_bone_ik, const Ref<IKBoneSegment3D> &p_parent, BoneId p_root, BoneId p_tip, int32_t p_stabilizing_pass_count = 0);
	void post_construct(const Ref<IKBoneSegment3D> &p_parent); // the part of the ctor that needs shared_from_this
	Ref<IKBone3D> get_root() const { return root; }
	Ref<IKBone3D> get_tip() const { return tip; }
	bool is_pinned() const { return tip ? tip->is_pinned() : false; }
	std::vector<Ref<IKBoneSegment3D>> get_child_segments() const { return child_segments; }
	void create_bone_list(std::vector<Ref<IKBone3D>> &p_list, bool p_recursive = false) const;
	void update_pinned_list(std::vector<std::vector<double>> &r_weights);
	static Quaternion clamp_to_cos_half_angle(Quaternion p_quat, double p_cos_half_angle);
	static void recursive_create_headings_arrays_for(Ref<IKBoneSegment3D> p_bone_segment);
	void create_headings_arrays();
	void recursive_create_penalty_array(Ref<IKBoneSegment3D> p_bone_segment, std::vector<std::vector<double>> &r_penalty_array, std::vector<Ref<IKBone3D>> &r_pinned_bones, double p_falloff);
	void segment_solver(const std::vector<float> &p_damp, float p_default_damp, bool p_constraint_mode, int32_t p_current_iteration, int32_t p_total_iteration);
	void generate_default_segments(std::vector<IKEffectorTemplate3D> &p_pins, BoneId p_root_bone, BoneId p_tip_bone, ManyBoneIK3D *p_many_bone_ik);

	void _update_target_headings(Ref<IKBone3D> p_for_bone, std::vector<double> *r_weights, PackedVector3Array *r_htarget);
	void _update_tip_headings(Ref<IKBone3D> p_for_bone, PackedVector3Array *r_heading_tip);
	void _set_optimal_rotation(Ref<IKBone3D> p_for_bone, PackedVector3Array *r_htip, PackedVector3Array *r_htarget, std::vector<double> *r_weights, float p_dampening = -1, bool p_translate = false, bool p_constraint_mode = false, double current_iteration = 0, double total_iterations = 0);
	void _qcp_solver(const std::vector<float> &p_damp, float p_default_damp, bool p_translate, bool p_constraint_mode, int32_t p_current_iteration, int32_t p_total_iterations);
	void _update_optimal_rotation(Ref<IKBone3D> p_for_bone, double p_damp, bool p_translate, bool p_constraint_mode, int32_t current_iteration, int32_t total_iterations);
	float _get_manual_msd(const PackedVector3Array &r_htip, const PackedVector3Array &r_htarget, const std::vector<double> &p_weights);
	bool _is_parent_of_tip(Ref<IKBone3D> p_current_tip, BoneId p_tip_bone);
	bool _has_multiple_children_or_pinned(std::vector<BoneId> &r_children, Ref<IKBone3D> p_current_tip);
	void _process_children(std::vector<BoneId> &r_children, Ref<IKBone3D> p_current_tip, std::vector<IKEffectorTemplate3D> &r_pins, BoneId p_root_bone, BoneId p_tip_bone, ManyBoneIK3D *p_many_bone_ik);
	Ref<IKBone3D> _create_next_bone(BoneId p_bone_id, Ref<IKBone3D> p_current_tip, std::vector<IKEffectorTemplate3D> &p_pins, ManyBoneIK3D *p_many_bone_ik);
	void _finalize_segment(Ref<IKBone3D> p_current_tip);
};

// ---- src/many_bone_ik_3d.{h,cpp}: tables + rebuild + iteration loop + I/O -------------------
struct Vector4f {
	real_t x = 0, y = 0, z = 0, w = 0;
};

class ManyBoneIK3D {
public:
	Skeleton3D skeleton_storage;
	bool is_constraint_mode = false;
	std::vector<Ref<IKBoneSegment3D>> segmented_skeletons;
	int32_t constraint_count = 0, pin_count = 0;
	std::vector<int> constraint_names; // bone ids stand for names
	std::vector<IKEffectorTemplate3D> pins;
	std::vector<Ref<IKBone3D>> bone_list;
	std::vector<float> joint_twist_x, joint_twist_y; // Vector2 joint_twist
	std::vector<float> bone_damp;
	std::vector<std::vector<Vector4f>> kusudama_open_cones;
	std::vector<int> kusudama_open_cone_count;
	int32_t iterations_per_frame = 15;
	float default_damp = Math::deg_to_rad(5.0f);
	Ref<IKNode3D> ik_origin;
	int32_t stabilize_passes = 0;
	std::vector<Transform3D> pin_targets; // boundary input: one per pins row (IKEffector3D::update_target_global_transform result)

	Skeleton3D *get_skeleton() { return &skeleton_storage; }
	float get_iterations_per_frame() const { return iterations_per_frame; }
	real_t get_default_damp() const { return default_damp; }
	bool get_constraint_mode() const { return is_constraint_mode; }

	void _bone_list_changed(); // many_bone_ik_3d.cpp:1011
	void _update_ik_bones_transform(); // :91
	void solve_iterations(); // the loop at :685-692
	// _update_skeleton_bones_transform :104 + IKBone3D::set_skeleton_bone_pose (ik_bone_3d.cpp:170):
	// writes pos3, quat4(xyzw), scale3 per skeleton bone into out[n_bones*10]; unsolved bones pass through.
	void write_skeleton_pose(float *out10, float *out_local12, uint32_t *status);
};

} // namespace orc
