// godot_module_binding.h -- the in-tree binding of libmbik.so for the reference module,
// as compilable code: ManyBoneIK3D::_process_modification() (reference src/many_bone_ik_3d.cpp:645-694) re-implemented
// on top of the C ABI in include/mbik.h.
//
// How a maintainer uses it: include this header at the end of src/many_bone_ik_3d.cpp (after "many_bone_ik_3d.h"), add
//     friend struct mbik_godot::Binding;   mbik_godot::Binding gpu;
// to class ManyBoneIK3D and make the body of _process_modification() `gpu.process_modification(*this);`.
// Nothing else in the module changes: properties, pins / constraints tables, serialisation and the editor gizmo keep
// working on the same members.  The header needs only the engine types the module already includes (Vector, Ref,
// Transform3D, Skeleton3D, Node3D) and "mbik.h".
//
// This file is exercised for real: the repository's test infrastructure compiles it against the reference's own sources
// (over a stand-in of the engine headers) and tests/test_reference_gpu.py checks that a scene driven through this binding
// leaves bit-identical bone poses on the Skeleton3D as the reference's own _process_modification().
#pragma once
#include "mbik.h"

namespace mbik_godot {

struct Binding {
	mbik_rig *rig = nullptr; // rebuilt whenever the node is dirty (replaces the IKBoneSegment3D graph)
	Vector<int32_t> bone_order; // ManyBoneIK3D::bone_list as skeleton bone ids
	Vector<float> targets, start, out;
	int last_error = MBIK_OK;

	~Binding() { mbik_rig_destroy(rig); }

	static void xform_to_floats(const Transform3D &t, float *o) { // Godot's own memory layout, real_t = float
		for (int r = 0; r < 3; r++) {
			for (int c = 0; c < 3; c++) {
				o[r * 3 + c] = t.basis.rows[r][c];
			}
		}
		o[9] = t.origin.x;
		o[10] = t.origin.y;
		o[11] = t.origin.z;
	}

	// replaces ManyBoneIK3D::_bone_list_changed() (src/many_bone_ik_3d.cpp:1011-1068)
	template <typename IK>
	void rebuild(IK &ik) {
		Skeleton3D *skeleton = ik.get_skeleton();
		const int nb = skeleton->get_bone_count();
		Vector<int32_t> parent;
		Vector<float> rest;
		parent.resize(nb);
		rest.resize(nb * 12);
		for (int b = 0; b < nb; b++) {
			parent.write[b] = skeleton->get_bone_parent(b);
			xform_to_floats(skeleton->get_bone_pose(b), rest.ptrw() + b * 12);
		}
		Vector<mbik_pin_desc> pd;
		pd.resize(ik.pins.size());
		for (int i = 0; i < ik.pins.size(); i++) {
			mbik_pin_desc p = {};
			p.bone = -1;
			if (ik.pins[i].is_valid()) {
				Vector3 pr = ik.pins[i]->get_direction_priorities();
				String name = ik.pins[i]->get_name();
				p.bone = name.is_empty() ? -1 : skeleton->find_bone(name);
				p.weight = ik.pins[i]->get_weight();
				p.motion_propagation_factor = ik.pins[i]->get_motion_propagation_factor();
				p.direction_priorities[0] = pr.x;
				p.direction_priorities[1] = pr.y;
				p.direction_priorities[2] = pr.z;
			}
			pd.write[i] = p;
		}
		Vector<mbik_constraint_desc> cd;
		Vector<mbik_cone_desc> cones;
		for (int i = 0; i < ik.constraint_count; i++) {
			mbik_constraint_desc c = {};
			c.bone = skeleton->find_bone(ik.constraint_names[i]);
			if (c.bone < 0) {
				continue; // _bone_list_changed never finds an IK bone for it (:1038-1043)
			}
			c.twist_from = ik.joint_twist[i].x;
			c.twist_range = ik.joint_twist[i].y;
			c.n_cones = ik.kusudama_open_cone_count[i];
			c.cone_offset = (int32_t)cones.size();
			for (int j = 0; j < c.n_cones; j++) {
				const Vector4 &k = ik.kusudama_open_cones[i][j];
				mbik_cone_desc cone = { { k.x, k.y, k.z }, k.w };
				cones.push_back(cone);
			}
			cd.push_back(c);
		}
		mbik_rig_desc d = {};
		d.n_bones = nb;
		d.parent = parent.ptr();
		d.rest_local = rest.ptr();
		d.n_pins = (int32_t)pd.size();
		d.pins = pd.ptr();
		d.n_constraints = (int32_t)cd.size();
		d.constraints = cd.ptr();
		d.cones = cones.ptr();
		d.n_bone_damp = (int32_t)ik.bone_damp.size();
		d.bone_damp = ik.bone_damp.ptr();
		d.default_damp = ik.default_damp;
		d.iterations_per_frame = ik.iterations_per_frame;
		d.stabilization_passes = ik.stabilize_passes;
		d.constraint_mode = ik.is_constraint_mode ? 1 : 0;
		mbik_rig_destroy(rig);
		rig = nullptr;
		bone_order.clear();
		last_error = mbik_rig_create(&d, &rig);
		if (last_error != MBIK_OK) {
			ERR_PRINT(mbik_last_error());
			rig = nullptr;
			return;
		}
		mbik_rig_info info;
		mbik_rig_get_info(rig, &info);
		bone_order.resize(info.n_solved);
		mbik_rig_get_bone_order(rig, bone_order.ptrw());
	}

	// replaces the body of ManyBoneIK3D::_process_modification() (src/many_bone_ik_3d.cpp:645-694)
	template <typename IK>
	void process_modification(IK &ik) {
		Skeleton3D *skeleton = ik.get_skeleton();
		if (!skeleton) { // :646-648
			return;
		}
		if (ik.get_effector_count() == 0) { // :649-651
			return;
		}
		if (ik.is_dirty || !rig) { // :652-658
			ik.is_dirty = false;
			rebuild(ik);
		}
		if (!rig) {
			return;
		}
		bool has_pins = false; // :669-677
		for (int i = 0; i < ik.pins.size(); i++) {
			if (ik.pins[i].is_valid() && !ik.pins[i]->get_name().is_empty()) {
				has_pins = true;
				break;
			}
		}
		if (!has_pins || !ik.is_enabled() || !ik.is_visible()) { // :675-683
			return;
		}
		const int nb = skeleton->get_bone_count(), np = (int)ik.pins.size();
		targets.resize(np * 12);
		start.resize(nb * 12);
		out.resize(nb * 10);
		for (int i = 0; i < np; i++) { // IKEffector3D::update_target_global_transform (src/ik_effector_3d.cpp:77-84)
			Transform3D x;
			if (ik.pins[i].is_valid()) {
				Node3D *t = Object::cast_to<Node3D>(ik.get_node_or_null(ik.pins[i]->get_target_node()));
				if (t && t->is_visible_in_tree()) {
					x = skeleton->get_global_transform().affine_inverse() * t->get_global_transform();
				}
			}
			xform_to_floats(x, targets.ptrw() + i * 12);
		}
		for (int b = 0; b < nb; b++) { // IKBone3D::set_initial_pose (src/ik_bone_3d.cpp:161-168)
			xform_to_floats(skeleton->get_bone_pose(b), start.ptrw() + b * 12);
		}
		mbik_solve_params p = {};
		p.iterations = (int32_t)ik.get_iterations_per_frame(); // read every frame (:685)
		p.device = -1;
		p.flags = MBIK_IO_HOST;
		last_error = mbik_solve_batch(rig, &p, 1, targets.ptr(), start.ptr(), out.ptrw(), nullptr, nullptr);
		if (last_error != MBIK_OK) {
			ERR_PRINT_ONCE(mbik_last_error()); // never throws; the pose is left untouched, like ERR_FAIL_*
			return;
		}
		for (int k = (int)bone_order.size(); k-- > 0;) { // _update_skeleton_bones_transform (:104-116, src/ik_bone_3d.cpp:170-179)
			const int b = bone_order[k];
			const float *o = out.ptr() + b * 10;
			skeleton->set_bone_pose_position(b, Vector3(o[0], o[1], o[2]));
			skeleton->set_bone_pose_rotation(b, Quaternion(o[3], o[4], o[5], o[6]));
			skeleton->set_bone_pose_scale(b, Vector3(o[7], o[8], o[9]));
		}
		ik.update_gizmos();
	}
};

} // namespace mbik_godot
