// many_bone_ik_host.hpp -- header-only C++17 host facade over the C ABI (include/mbik.h).
//
// Mirrors the reference module's operator interface for the solve path: the same class names, setter/getter names,
// argument meaning, defaults and error behaviour (ERR_FAIL_* -> silently return a default) as
//     ManyBoneIK3D              reference src/many_bone_ik_3d.{h,cpp}
//     IKEffectorTemplate3D      reference src/ik_effector_template_3d.{h,cpp}
// minus the Godot object model (Variant/ClassDB/Node/Resource), which is out of scope.  What the reference does
// per frame for ONE skeleton (`_process_modification`, src/many_bone_ik_3d.cpp:645-694) this facade does for a
// BATCH of independent poses of the same rig through `process_modification_batch`, which rebuilds the flattened rig
// when dirty (`_bone_list_changed` -> mbik_rig_create) and runs the CUDA solve (mbik_solve_batch).  No solve
// arithmetic lives here, and there is no CPU fallback.
//
// It also implements the reference's dynamic property paths (`_set` / `_get`, src/many_bone_ik_3d.cpp:217-375:
// "pins/<i>/weight", "constraints/<i>/kusudama_open_cone/<j>/radius", ...) and a loader for the `key = value`
// lines of a Godot .tscn node section, so rigs authored in the editor can be fed to the batched solver.
#pragma once

#include "../../include/mbik.h"

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <istream>
#include <sstream>
#include <string>
#include <variant>
#include <vector>

namespace mbik_host {

struct Vector2 {
	float x = 0, y = 0;
};
struct Vector3 {
	float x = 0, y = 0, z = 0;
};
struct Vector4 {
	float x = 0, y = 0, z = 0, w = 0;
};
// Godot Transform3D memory layout: basis rows, then origin
struct Transform3D {
	float basis[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
	float origin[3] = { 0, 0, 0 };
	void to_floats(float *o) const { // the C ABI's 12-float record: basis rows, then origin
		memcpy(o, basis, sizeof(basis));
		memcpy(o + 9, origin, sizeof(origin));
	}
};
using Variant = std::variant<std::monostate, bool, int64_t, double, std::string, Vector2, Vector3, Transform3D>;

// Stand-in for the part of Godot's Skeleton3D the path reads: names, parents, current local poses.
class Skeleton3D {
	std::vector<std::string> names;
	std::vector<int32_t> parents;
	std::vector<Transform3D> poses;

public:
	int32_t add_bone(const std::string &p_name, int32_t p_parent = -1, const Transform3D &p_pose = Transform3D()) {
		names.push_back(p_name);
		parents.push_back(p_parent);
		poses.push_back(p_pose);
		return (int32_t)names.size() - 1;
	}
	int32_t get_bone_count() const { return (int32_t)names.size(); }
	int32_t find_bone(const std::string &p_name) const {
		for (size_t i = 0; i < names.size(); i++) {
			if (names[i] == p_name) {
				return (int32_t)i;
			}
		}
		return -1;
	}
	std::string get_bone_name(int32_t p_bone) const { return (p_bone >= 0 && p_bone < get_bone_count()) ? names[p_bone] : std::string(); }
	int32_t get_bone_parent(int32_t p_bone) const { return (p_bone >= 0 && p_bone < get_bone_count()) ? parents[p_bone] : -1; }
	Transform3D get_bone_pose(int32_t p_bone) const { return (p_bone >= 0 && p_bone < get_bone_count()) ? poses[p_bone] : Transform3D(); }
	void set_bone_pose(int32_t p_bone, const Transform3D &p_pose) {
		if (p_bone >= 0 && p_bone < get_bone_count()) {
			poses[p_bone] = p_pose;
		}
	}
	const std::vector<int32_t> &get_parents() const { return parents; }
};

// reference src/ik_effector_template_3d.h:40-45 (defaults included)
class IKEffectorTemplate3D {
	std::string name;
	std::string target_node;
	float motion_propagation_factor = 1.0f;
	float weight = 0.0f;
	Vector3 priority_direction{ 0.2f, 0.0f, 0.2f };

public:
	std::string get_name() const { return name; }
	void set_name(const std::string &p_name) { name = p_name; }
	std::string get_target_node() const { return target_node; }
	void set_target_node(const std::string &p_path) { target_node = p_path; }
	float get_motion_propagation_factor() const { return motion_propagation_factor; }
	void set_motion_propagation_factor(float p) { motion_propagation_factor = p; }
	float get_weight() const { return weight; }
	void set_weight(float p) { weight = p; }
	Vector3 get_direction_priorities() const { return priority_direction; }
	void set_direction_priorities(Vector3 p) { priority_direction = p; }
};

class ManyBoneIK3D {
	// the tables of reference src/many_bone_ik_3d.h:49-68
	bool is_constraint_mode = false;
	int32_t constraint_count = 0, pin_count = 0, bone_count = 0;
	std::vector<std::string> constraint_names;
	std::vector<IKEffectorTemplate3D> pins;
	std::vector<Vector2> joint_twist;
	std::vector<float> bone_damp;
	std::vector<std::vector<Vector4>> kusudama_open_cones;
	std::vector<int32_t> kusudama_open_cone_count;
	int32_t iterations_per_frame = 15;
	float default_damp = 5.0f * 3.14159265358979323846f / 180.0f; // Math::deg_to_rad(5.0f)
	bool is_dirty = true;
	int32_t ui_selected_bone = -1, stabilize_passes = 0;

	const Skeleton3D *skeleton = nullptr;
	mbik_rig *rig = nullptr;
	int last_error = MBIK_OK;

	static float vlen2(Vector3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }
	static bool is_zero_approx(float s) { return std::fabs(s) < 0.00001f; }

public:
	ManyBoneIK3D() {}
	~ManyBoneIK3D() { mbik_rig_destroy(rig); }
	ManyBoneIK3D(const ManyBoneIK3D &) = delete;
	ManyBoneIK3D &operator=(const ManyBoneIK3D &) = delete;

	// SkeletonModifier3D::get_skeleton / _skeleton_changed (src/many_bone_ik_3d.cpp:1070-1086)
	void set_skeleton(const Skeleton3D *p_skeleton) {
		skeleton = p_skeleton;
		set_dirty();
	}
	const Skeleton3D *get_skeleton() const { return skeleton; }

	void set_dirty() { is_dirty = true; }

	// ---- pins (src/many_bone_ik_3d.cpp:44-89, 442-452, 664-728) ----
	void set_total_effector_count(int32_t p_value) {
		pin_count = p_value;
		pins.resize(p_value < 0 ? 0 : p_value);
		set_dirty();
	}
	int32_t get_effector_count() const { return pin_count; }
	void set_effector_count(int32_t p_pin_count) { pin_count = p_pin_count; }
	void set_effector_bone_name(int32_t p_pin_index, const std::string &p_bone) {
		if (p_pin_index < 0 || p_pin_index >= (int32_t)pins.size()) {
			return;
		}
		pins[p_pin_index].set_name(p_bone);
		set_dirty();
	}
	std::string get_effector_bone_name(int32_t p_effector_index) const {
		return (p_effector_index >= 0 && p_effector_index < (int32_t)pins.size()) ? pins[p_effector_index].get_name() : std::string();
	}
	void set_effector_target_node_path(int32_t p_pin_index, const std::string &p_target_node) {
		if (p_pin_index < 0 || p_pin_index >= (int32_t)pins.size()) {
			return;
		}
		pins[p_pin_index].set_target_node(p_target_node);
		set_dirty();
	}
	std::string get_effector_target_node_path(int32_t p_pin_index) const {
		return (p_pin_index >= 0 && p_pin_index < (int32_t)pins.size()) ? pins[p_pin_index].get_target_node() : std::string();
	}
	void set_pin_weight(int32_t p_pin_index, float p_weight) {
		if (p_pin_index < 0 || p_pin_index >= (int32_t)pins.size()) {
			return;
		}
		pins[p_pin_index].set_weight(p_weight);
		set_dirty();
	}
	float get_pin_weight(int32_t p_pin_index) const { return (p_pin_index >= 0 && p_pin_index < (int32_t)pins.size()) ? pins[p_pin_index].get_weight() : 0.0f; }
	void set_pin_direction_priorities(int32_t p_pin_index, Vector3 p_priority_direction) {
		if (p_pin_index < 0 || p_pin_index >= (int32_t)pins.size()) {
			return;
		}
		pins[p_pin_index].set_direction_priorities(p_priority_direction);
		set_dirty();
	}
	Vector3 get_pin_direction_priorities(int32_t p_pin_index) const {
		return (p_pin_index >= 0 && p_pin_index < (int32_t)pins.size()) ? pins[p_pin_index].get_direction_priorities() : Vector3();
	}
	void set_pin_motion_propagation_factor(int32_t p_effector_index, float p_motion_propagation_factor) {
		if (p_effector_index < 0 || p_effector_index >= (int32_t)pins.size()) {
			return;
		}
		pins[p_effector_index].set_motion_propagation_factor(p_motion_propagation_factor);
		set_dirty();
	}
	float get_pin_motion_propagation_factor(int32_t p_effector_index) const {
		return (p_effector_index >= 0 && p_effector_index < (int32_t)pins.size()) ? pins[p_effector_index].get_motion_propagation_factor() : 0.0f;
	}
	int32_t find_pin(const std::string &p_string) const {
		for (int32_t i = 0; i < pin_count && i < (int32_t)pins.size(); i++) {
			if (pins[i].get_name() == p_string) {
				return i;
			}
		}
		return -1;
	}

	// ---- constraints (src/many_bone_ik_3d.cpp:454-620, 739-754, 975-984) ----
	void _set_constraint_count(int32_t p_count) {
		if (p_count < 0) {
			p_count = 0;
		}
		int32_t old_count = (int32_t)constraint_names.size();
		constraint_count = p_count;
		constraint_names.resize(p_count);
		joint_twist.resize(p_count);
		kusudama_open_cone_count.resize(p_count);
		kusudama_open_cones.resize(p_count);
		for (int32_t i = p_count; i-- > old_count;) {
			constraint_names[i] = std::string();
			kusudama_open_cone_count[i] = 0;
			kusudama_open_cones[i].assign(1, Vector4{ 0, 1, 0, 0.01745f });
			joint_twist[i] = Vector2{ 0, 0.01745f };
		}
		set_dirty();
	}
	void add_constraint() {
		int32_t old_count = constraint_count;
		_set_constraint_count(constraint_count + 1);
		constraint_names[old_count] = std::string();
		kusudama_open_cone_count[old_count] = 0;
		kusudama_open_cones[old_count].assign(1, Vector4{ 0, 1, 0, 3.14159265358979323846f });
		joint_twist[old_count] = Vector2{ 0, 3.14159265358979323846f };
		set_dirty();
	}
	void remove_constraint_at_index(int32_t p_index) {
		if (p_index < 0 || p_index >= constraint_count) {
			return;
		}
		constraint_names.erase(constraint_names.begin() + p_index);
		kusudama_open_cone_count.erase(kusudama_open_cone_count.begin() + p_index);
		kusudama_open_cones.erase(kusudama_open_cones.begin() + p_index);
		joint_twist.erase(joint_twist.begin() + p_index);
		constraint_count--;
		set_dirty();
	}
	int32_t get_constraint_count() const { return constraint_count; }
	void set_constraint_name_at_index(int32_t p_index, const std::string &p_name) {
		if (p_index < 0 || p_index >= (int32_t)constraint_names.size()) {
			return;
		}
		constraint_names[p_index] = p_name;
		set_dirty();
	}
	std::string get_constraint_name(int32_t p_index) const {
		return (p_index >= 0 && p_index < (int32_t)constraint_names.size()) ? constraint_names[p_index] : std::string();
	}
	int32_t find_constraint(const std::string &p_string) const {
		for (int32_t i = 0; i < constraint_count; i++) {
			if (get_constraint_name(i) == p_string) {
				return i;
			}
		}
		return -1;
	}
	Vector2 get_joint_twist(int32_t p_index) const { return (p_index >= 0 && p_index < (int32_t)joint_twist.size()) ? joint_twist[p_index] : Vector2(); }
	void set_joint_twist(int32_t p_index, Vector2 p_to) {
		if (p_index < 0 || p_index >= constraint_count) {
			return;
		}
		joint_twist[p_index] = p_to;
		set_dirty();
	}
	void set_kusudama_open_cone_count(int32_t p_constraint_index, int32_t p_count) {
		if (p_constraint_index < 0 || p_constraint_index >= (int32_t)kusudama_open_cone_count.size() || p_count < 0) {
			return;
		}
		int32_t old_cone_count = (int32_t)kusudama_open_cones[p_constraint_index].size();
		kusudama_open_cone_count[p_constraint_index] = p_count;
		std::vector<Vector4> &cones = kusudama_open_cones[p_constraint_index];
		cones.resize(p_count);
		// new cones point along -Y of the bone-direction transform; before a rebuild that transform is identity
		// (get_direction_transform_of_bone returns Transform3D(), src/many_bone_ik_3d.cpp:806-809)
		for (int32_t cone_i = p_count; cone_i-- > old_cone_count;) {
			cones[cone_i] = Vector4{ -0.0f, -1.0f, -0.0f, 0.0f };
		}
		set_dirty();
	}
	int32_t get_kusudama_open_cone_count(int32_t p_constraint_index) const {
		return (p_constraint_index >= 0 && p_constraint_index < (int32_t)kusudama_open_cone_count.size()) ? kusudama_open_cone_count[p_constraint_index] : 0;
	}
	void set_kusudama_open_cone_center(int32_t p_constraint_index, int32_t p_index, Vector3 p_center) {
		if (p_constraint_index < 0 || p_constraint_index >= (int32_t)kusudama_open_cones.size()) {
			return;
		}
		if (p_index < 0 || p_index >= (int32_t)kusudama_open_cones[p_constraint_index].size()) {
			return;
		}
		Vector4 &cone = kusudama_open_cones[p_constraint_index][p_index];
		if (is_zero_approx(vlen2(p_center))) {
			cone.x = 0;
			cone.y = 1;
			cone.z = 0;
		} else {
			cone.x = p_center.x;
			cone.y = p_center.y;
			cone.z = p_center.z;
		}
		set_dirty();
	}
	Vector3 get_kusudama_open_cone_center(int32_t p_constraint_index, int32_t p_index) const {
		if (p_constraint_index < 0 || p_constraint_index >= (int32_t)kusudama_open_cones.size() || p_index < 0 ||
				p_index >= (int32_t)kusudama_open_cones[p_constraint_index].size()) {
			return Vector3{ 0.0f, 0.0f, 1.0f };
		}
		const Vector4 &c = kusudama_open_cones[p_constraint_index][p_index];
		return Vector3{ c.x, c.y, c.z };
	}
	void set_kusudama_open_cone_radius(int32_t p_constraint_index, int32_t p_index, float p_radius) {
		if (p_constraint_index < 0 || p_constraint_index >= (int32_t)kusudama_open_cone_count.size()) {
			return;
		}
		if (p_index < 0 || p_index >= kusudama_open_cone_count[p_constraint_index] || p_index >= (int32_t)kusudama_open_cones[p_constraint_index].size()) {
			return;
		}
		kusudama_open_cones[p_constraint_index][p_index].w = p_radius;
		set_dirty();
	}
	float get_kusudama_open_cone_radius(int32_t p_constraint_index, int32_t p_index) const {
		if (p_constraint_index < 0 || p_constraint_index >= (int32_t)kusudama_open_cones.size() || p_index < 0 ||
				p_index >= (int32_t)kusudama_open_cones[p_constraint_index].size()) {
			return 6.28318530717958647692f; // Math_TAU
		}
		return kusudama_open_cones[p_constraint_index][p_index].w;
	}
	// set_kusudama_open_cone: centre is normalised here (unlike set_kusudama_open_cone_center), :507-524
	void set_kusudama_open_cone(int32_t p_constraint_index, int32_t p_index, Vector3 p_center, float p_radius) {
		if (p_constraint_index < 0 || p_constraint_index >= (int32_t)kusudama_open_cones.size()) {
			return;
		}
		if (p_index < 0 || p_index >= (int32_t)kusudama_open_cones[p_constraint_index].size()) {
			return;
		}
		if (is_zero_approx(vlen2(p_center))) {
			p_center = Vector3{ 0.0f, 1.0f, 0.0f };
		}
		float l2 = vlen2(p_center);
		float l = std::sqrt(l2);
		Vector3 c = l2 == 0.0f ? Vector3() : Vector3{ p_center.x / l, p_center.y / l, p_center.z / l };
		kusudama_open_cones[p_constraint_index][p_index] = Vector4{ c.x, c.y, c.z, p_radius };
		set_dirty();
	}

	// ---- global settings ----
	float get_default_damp() const { return default_damp; }
	void set_default_damp(float p_default_damp) {
		default_damp = p_default_damp;
		set_dirty();
	}
	float get_iterations_per_frame() const { return (float)iterations_per_frame; }
	void set_iterations_per_frame(float p_iterations_per_frame) { iterations_per_frame = (int32_t)p_iterations_per_frame; }
	bool get_constraint_mode() const { return is_constraint_mode; }
	// read every frame in the reference (:689); baked into the flattened rig here, hence the rebuild
	void set_constraint_mode(bool p_enabled) {
		is_constraint_mode = p_enabled;
		set_dirty();
	}
	void set_stabilization_passes(int32_t p_passes) {
		stabilize_passes = p_passes;
		set_dirty();
	}
	int32_t get_stabilization_passes() const { return stabilize_passes; }
	int32_t get_ui_selected_bone() const { return ui_selected_bone; }
	void set_ui_selected_bone(int32_t p) { ui_selected_bone = p; }
	// _set_bone_count: sizes bone_damp, new entries = default_damp (src/many_bone_ik_3d.cpp:756-764)
	void _set_bone_count(int32_t p_count) {
		if (p_count < 0) {
			p_count = 0;
		}
		bone_damp.resize(p_count);
		for (int32_t i = p_count; i-- > bone_count;) {
			bone_damp[i] = get_default_damp();
		}
		bone_count = p_count;
		set_dirty();
	}
	int32_t get_bone_count() const { return bone_count; }
	// not in the reference (bone_damp has no setter there): lets a caller vary per-bone damping (BASELINE config 4)
	void set_bone_damp(int32_t p_bone_id, float p_damp) {
		if (p_bone_id >= 0 && p_bone_id < (int32_t)bone_damp.size()) {
			bone_damp[p_bone_id] = p_damp;
			set_dirty();
		}
	}
	void reset_constraints() {
		if (skeleton) {
			int32_t saved_pin_count = get_effector_count();
			set_total_effector_count(0);
			set_total_effector_count(saved_pin_count);
			int32_t saved_constraint_count = (int32_t)constraint_names.size();
			_set_constraint_count(0);
			_set_constraint_count(saved_constraint_count);
			_set_bone_count(0);
			_set_bone_count(saved_constraint_count);
		}
		set_dirty();
	}
	void register_skeleton() {
		if (!get_effector_count() && !get_constraint_count()) {
			reset_constraints();
		}
		set_dirty();
	}

	// ---- dynamic properties: ManyBoneIK3D::_set / _get (src/many_bone_ik_3d.cpp:217-375) ----
	static std::string slice(const std::string &s, int idx) {
		size_t b = 0;
		for (int i = 0; i < idx; i++) {
			b = s.find('/', b);
			if (b == std::string::npos) {
				return std::string();
			}
			b++;
		}
		size_t e = s.find('/', b);
		return s.substr(b, e == std::string::npos ? std::string::npos : e - b);
	}
	static double as_number(const Variant &v) {
		if (auto p = std::get_if<double>(&v)) {
			return *p;
		}
		if (auto p = std::get_if<int64_t>(&v)) {
			return (double)*p;
		}
		if (auto p = std::get_if<bool>(&v)) {
			return *p ? 1.0 : 0.0;
		}
		return 0.0;
	}
	static std::string as_string(const Variant &v) {
		if (auto p = std::get_if<std::string>(&v)) {
			return *p;
		}
		return std::string();
	}
	static Vector3 as_vector3(const Variant &v) {
		if (auto p = std::get_if<Vector3>(&v)) {
			return *p;
		}
		return Vector3();
	}

	bool _set(const std::string &name, const Variant &p_value) {
		if (name == "constraint_count") {
			_set_constraint_count((int32_t)as_number(p_value));
			return true;
		} else if (name == "pin_count") {
			set_total_effector_count((int32_t)as_number(p_value));
			return true;
		} else if (name == "bone_count") { // read-only in the reference's _set; accepted here so saved scenes round-trip
			_set_bone_count((int32_t)as_number(p_value));
			return true;
		} else if (name == "iterations_per_frame") {
			set_iterations_per_frame((float)as_number(p_value));
			return true;
		} else if (name == "default_damp") {
			set_default_damp((float)as_number(p_value));
			return true;
		} else if (name == "constraint_mode") {
			set_constraint_mode(as_number(p_value) != 0.0);
			return true;
		} else if (name == "stabilization_passes") {
			set_stabilization_passes((int32_t)as_number(p_value));
			return true;
		} else if (name == "ui_selected_bone") {
			set_ui_selected_bone((int32_t)as_number(p_value));
			return true;
		} else if (name.rfind("bone_damp/", 0) == 0) { // extension (the reference has no bone_damp property)
			set_bone_damp(atoi(slice(name, 1).c_str()), (float)as_number(p_value));
			return true;
		} else if (name.rfind("pins/", 0) == 0) {
			int index = atoi(slice(name, 1).c_str());
			std::string what = slice(name, 2);
			if (index >= (int)pins.size()) {
				set_total_effector_count(constraint_count); // sic: the reference passes constraint_count here (:308)
			}
			if (what == "bone_name") {
				set_effector_bone_name(index, as_string(p_value));
				return true;
			} else if (what == "target_node") {
				set_effector_target_node_path(index, as_string(p_value));
				return true;
			} else if (what == "target_static") {
				if (as_number(p_value) != 0.0) {
					set_effector_target_node_path(index, std::string());
				}
				return true;
			} else if (what == "motion_propagation_factor") {
				set_pin_motion_propagation_factor(index, (float)as_number(p_value));
				return true;
			} else if (what == "weight") {
				set_pin_weight(index, (float)as_number(p_value));
				return true;
			} else if (what == "direction_priorities") {
				set_pin_direction_priorities(index, as_vector3(p_value));
				return true;
			}
		} else if (name.rfind("constraints/", 0) == 0) {
			int index = atoi(slice(name, 1).c_str());
			std::string what = slice(name, 2);
			std::string begins = "constraints/" + std::to_string(index) + "/kusudama_open_cone/";
			if (index >= (int)constraint_names.size()) {
				_set_constraint_count(constraint_count);
			}
			if (what == "bone_name") {
				set_constraint_name_at_index(index, as_string(p_value));
				return true;
			} else if (what == "twist_from") {
				Vector2 t = get_joint_twist(index);
				set_joint_twist(index, Vector2{ (float)as_number(p_value), t.y });
				return true;
			} else if (what == "twist_range") {
				Vector2 t = get_joint_twist(index);
				set_joint_twist(index, Vector2{ t.x, (float)as_number(p_value) });
				return true;
			} else if (what == "kusudama_open_cone_count") {
				set_kusudama_open_cone_count(index, (int32_t)as_number(p_value));
				return true;
			} else if (name.rfind(begins, 0) == 0) {
				int cone_index = atoi(slice(name, 3).c_str());
				std::string cone_what = slice(name, 4);
				if (cone_what == "center") {
					set_kusudama_open_cone_center(index, cone_index, as_vector3(p_value));
					return true;
				} else if (cone_what == "radius") {
					set_kusudama_open_cone_radius(index, cone_index, (float)as_number(p_value));
					return true;
				}
			} else if (what == "bone_direction" || what == "kusudama_orientation" || what == "kusudama_twist") {
				// Stored transforms are pushed into the live object graph by the reference and DROPPED by the next
				// dirty rebuild (src/many_bone_ik_3d.cpp:1017-1018); the batched solve always runs on a fresh rebuild.
				return true;
			}
		}
		return false;
	}

	bool _get(const std::string &name, Variant &r_ret) const {
		if (name == "constraint_count") {
			r_ret = (int64_t)get_constraint_count();
			return true;
		} else if (name == "pin_count") {
			r_ret = (int64_t)get_effector_count();
			return true;
		} else if (name == "bone_count") {
			r_ret = (int64_t)get_bone_count();
			return true;
		} else if (name.rfind("pins/", 0) == 0) {
			int index = atoi(slice(name, 1).c_str());
			std::string what = slice(name, 2);
			if (index < 0 || index >= (int)pins.size()) {
				return false;
			}
			if (what == "bone_name") {
				r_ret = pins[index].get_name();
				return true;
			} else if (what == "target_node") {
				r_ret = pins[index].get_target_node();
				return true;
			} else if (what == "target_static") {
				r_ret = pins[index].get_target_node().empty();
				return true;
			} else if (what == "motion_propagation_factor") {
				r_ret = (double)get_pin_motion_propagation_factor(index);
				return true;
			} else if (what == "weight") {
				r_ret = (double)get_pin_weight(index);
				return true;
			} else if (what == "direction_priorities") {
				r_ret = get_pin_direction_priorities(index);
				return true;
			}
		} else if (name.rfind("constraints/", 0) == 0) {
			int index = atoi(slice(name, 1).c_str());
			std::string what = slice(name, 2);
			if (index < 0 || index >= constraint_count) {
				return false;
			}
			std::string begins = "constraints/" + std::to_string(index) + "/kusudama_open_cone";
			if (what == "bone_name") {
				r_ret = constraint_names[index];
				return true;
			} else if (what == "twist_start") { // sic: _get exposes twist_start/twist_end, _set takes twist_from/twist_range
				r_ret = (double)get_joint_twist(index).x;
				return true;
			} else if (what == "twist_end") {
				r_ret = (double)get_joint_twist(index).y;
				return true;
			} else if (what == "kusudama_open_cone_count") {
				r_ret = (int64_t)get_kusudama_open_cone_count(index);
				return true;
			} else if (name.rfind(begins, 0) == 0) {
				int cone_index = atoi(slice(name, 3).c_str());
				std::string cone_what = slice(name, 4);
				if (cone_what == "center") {
					r_ret = get_kusudama_open_cone_center(index, cone_index);
					return true;
				} else if (cone_what == "radius") {
					r_ret = (double)get_kusudama_open_cone_radius(index, cone_index);
					return true;
				}
			}
		}
		return false;
	}

	// Parse the right-hand side of a .tscn property line: 1.5, 3, true, "str", &"name", NodePath("..."),
	// Vector2(..), Vector3(..), Transform3D(12 numbers: basis columns-in-rows as Godot prints them, then origin).
	static Variant parse_value(std::string v) {
		auto trim = [](std::string &s) {
			size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
			s = (a == std::string::npos) ? std::string() : s.substr(a, b - a + 1);
		};
		trim(v);
		if (v == "true") {
			return true;
		}
		if (v == "false") {
			return false;
		}
		auto numbers = [](const std::string &s) {
			std::vector<double> out;
			std::string t = s.substr(s.find('(') + 1);
			for (char &c : t) {
				if (c == ',' || c == ')') {
					c = ' ';
				}
			}
			std::istringstream is(t);
			double d;
			while (is >> d) {
				out.push_back(d);
			}
			return out;
		};
		if (v.rfind("Vector3(", 0) == 0) {
			auto n = numbers(v);
			n.resize(3, 0.0);
			return Vector3{ (float)n[0], (float)n[1], (float)n[2] };
		}
		if (v.rfind("Vector2(", 0) == 0) {
			auto n = numbers(v);
			n.resize(2, 0.0);
			return Vector2{ (float)n[0], (float)n[1] };
		}
		if (v.rfind("Transform3D(", 0) == 0) {
			auto n = numbers(v);
			n.resize(12, 0.0);
			Transform3D t;
			for (int k = 0; k < 9; k++) { // Godot prints the basis column by column (x axis, y axis, z axis)
				t.basis[(k % 3) * 3 + k / 3] = (float)n[k];
			}
			for (int i = 0; i < 3; i++) {
				t.origin[i] = (float)n[9 + i];
			}
			return t;
		}
		size_t q0 = v.find('"');
		if (q0 != std::string::npos) { // "str", &"name", NodePath("path")
			size_t q1 = v.rfind('"');
			return v.substr(q0 + 1, q1 > q0 ? q1 - q0 - 1 : 0);
		}
		if (v.find_first_of(".eE") == std::string::npos && v.find("inf") == std::string::npos && v.find("nan") == std::string::npos) {
			char *end = nullptr;
			long long i = strtoll(v.c_str(), &end, 10);
			if (end && *end == 0 && !v.empty()) {
				return (int64_t)i;
			}
		}
		return atof(v.c_str());
	}

	// Feed `key = value` lines (a [node ... type="ManyBoneIK3D"] section of a .tscn, or any subset) through _set.
	// Lines that are not properties of this class are ignored; returns the number of properties applied.
	int load_properties(std::istream &in) {
		int applied = 0;
		std::string line;
		while (std::getline(in, line)) {
			size_t eq = line.find('=');
			if (line.empty() || line[0] == '[' || line[0] == ';' || eq == std::string::npos) {
				continue;
			}
			std::string key = line.substr(0, eq), val = line.substr(eq + 1);
			size_t a = key.find_first_not_of(" \t"), b = key.find_last_not_of(" \t");
			if (a == std::string::npos) {
				continue;
			}
			key = key.substr(a, b - a + 1);
			if (_set(key, parse_value(val))) {
				applied++;
			}
		}
		return applied;
	}
	int load_properties(const std::string &text) {
		std::istringstream is(text);
		return load_properties(is);
	}

	// ---- the rebuild: ManyBoneIK3D::_bone_list_changed (src/many_bone_ik_3d.cpp:1011-1068) ----
	// Resolves bone names against the skeleton and hands the tables to mbik_rig_create.  Returns an MBIK_* code.
	int _bone_list_changed() {
		if (!skeleton) {
			return last_error = MBIK_ERR_INVALID_ARG;
		}
		const int32_t nb = skeleton->get_bone_count();
		std::vector<float> rest((size_t)nb * 12);
		for (int32_t b = 0; b < nb; b++) {
			Transform3D t = skeleton->get_bone_pose(b);
			memcpy(&rest[(size_t)b * 12], t.basis, sizeof(float) * 9);
			memcpy(&rest[(size_t)b * 12 + 9], t.origin, sizeof(float) * 3);
		}
		std::vector<mbik_pin_desc> pd(pins.size());
		for (size_t i = 0; i < pins.size(); i++) {
			// a pin whose bone name is empty or unknown never matches a bone (src/ik_bone_3d.cpp:209-222)
			pd[i].bone = skeleton->find_bone(pins[i].get_name());
			pd[i].weight = pins[i].get_weight();
			pd[i].motion_propagation_factor = pins[i].get_motion_propagation_factor();
			Vector3 p = pins[i].get_direction_priorities();
			pd[i].direction_priorities[0] = p.x;
			pd[i].direction_priorities[1] = p.y;
			pd[i].direction_priorities[2] = p.z;
		}
		std::vector<mbik_constraint_desc> cd((size_t)constraint_count);
		std::vector<mbik_cone_desc> cones;
		for (int32_t i = 0; i < constraint_count; i++) {
			cd[i].bone = skeleton->find_bone(constraint_names[i]);
			cd[i].twist_from = joint_twist[i].x;
			cd[i].twist_range = joint_twist[i].y;
			cd[i].n_cones = kusudama_open_cone_count[i];
			cd[i].cone_offset = (int32_t)cones.size();
			for (int32_t j = 0; j < cd[i].n_cones; j++) {
				Vector4 c = j < (int32_t)kusudama_open_cones[i].size() ? kusudama_open_cones[i][j] : Vector4{ 0, 1, 0, 0 };
				mbik_cone_desc m;
				m.center[0] = c.x;
				m.center[1] = c.y;
				m.center[2] = c.z;
				m.radius = c.w;
				cones.push_back(m);
			}
		}
		mbik_rig_desc d;
		memset(&d, 0, sizeof(d));
		d.n_bones = nb;
		d.parent = skeleton->get_parents().data();
		d.rest_local = rest.data();
		d.n_pins = (int32_t)pd.size();
		d.pins = pd.data();
		d.n_constraints = constraint_count;
		d.constraints = cd.data();
		d.cones = cones.data();
		d.n_bone_damp = (int32_t)bone_damp.size();
		d.bone_damp = bone_damp.data();
		d.default_damp = default_damp;
		d.iterations_per_frame = iterations_per_frame;
		d.stabilization_passes = stabilize_passes;
		d.constraint_mode = is_constraint_mode ? 1 : 0;
		mbik_rig_destroy(rig);
		rig = nullptr;
		return last_error = mbik_rig_create(&d, &rig);
	}

	// ---- the hot path: ManyBoneIK3D::_process_modification for a batch of poses (:645-694) ----
	//   targets     [n_poses][pin_count][12]   skeleton-space target of every pin row
	//   start_pose  [n_poses][n_bones][12] or nullptr (= the skeleton's current bone poses, re-read at every call)
	//   out_pose    [n_poses][n_bones][10]     position, rotation quaternion, scale per bone
	// Early-outs of the reference (no skeleton, no pins, no named pin) leave the outputs untouched and return MBIK_OK.
	int process_modification_batch(size_t n_poses, const float *targets, const float *start_pose, float *out_pose, float *out_local = nullptr,
			uint32_t *out_status = nullptr, const mbik_solve_params *params = nullptr) {
		if (!skeleton || get_effector_count() == 0) {
			return MBIK_OK;
		}
		bool has_pins = false;
		for (const IKEffectorTemplate3D &pin : pins) {
			if (!pin.get_name().empty()) {
				has_pins = true;
				break;
			}
		}
		if (!has_pins) {
			return MBIK_OK;
		}
		if (is_dirty || !rig) {
			int rc = _bone_list_changed();
			if (rc != MBIK_OK) {
				return rc;
			}
			is_dirty = false;
		}
		mbik_solve_params p = {};
		if (params) {
			p = *params;
		} else {
			p.iterations = -1;
			p.device = -1;
			p.flags = MBIK_IO_HOST;
		}
		// iterations_per_frame and constraint_mode are read every frame without a rebuild in the reference (:685-689)
		if (p.iterations < 0) {
			p.iterations = iterations_per_frame;
		}
		// The reference re-reads the skeleton every frame (_update_ik_bones_transform -> IKBone3D::set_initial_pose,
		// :91-102, src/ik_bone_3d.cpp:161-168), and Skeleton3D::set_bone_pose does not mark the IK dirty: with no explicit
		// start poses every pose of the batch starts from the skeleton's CURRENT bone poses, not from the poses captured
		// at the last rebuild.  (Host buffers only: a device-buffer call has nowhere to put them.)
		std::vector<float> current;
		if (!start_pose && !(p.flags & MBIK_IO_DEVICE)) {
			const int32_t nb = skeleton->get_bone_count();
			current.resize(n_poses * (size_t)nb * 12);
			for (int32_t b = 0; b < nb && n_poses > 0; b++) {
				skeleton->get_bone_pose(b).to_floats(&current[(size_t)b * 12]);
			}
			for (size_t k = 1; k < n_poses; k++) {
				memcpy(&current[k * (size_t)nb * 12], current.data(), sizeof(float) * 12 * (size_t)nb);
			}
			start_pose = current.data();
		}
		return last_error = mbik_solve_batch(rig, &p, n_poses, targets, start_pose, out_pose, out_local, out_status);
	}

	mbik_rig *get_rig() const { return rig; }
	int get_last_error() const { return last_error; }
	// get_bone_list(): solved skeleton bone ids in bone_list order (children segments first, tip -> root)
	std::vector<int32_t> get_bone_list() const {
		std::vector<int32_t> out;
		mbik_rig_info info;
		if (rig && mbik_rig_get_info(rig, &info) == MBIK_OK) {
			out.resize(info.n_solved);
			mbik_rig_get_bone_order(rig, out.data());
		}
		return out;
	}
};

} // namespace mbik_host
