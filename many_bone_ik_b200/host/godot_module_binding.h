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

#include <string.h>

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

	// everything _bone_list_changed() reads, as the flat tables of mbik_rig_desc (the Vectors own what `d` points at)
	struct Desc {
		Vector<int32_t> parent;
		Vector<float> rest;
		Vector<mbik_pin_desc> pd;
		Vector<mbik_constraint_desc> cd;
		Vector<mbik_cone_desc> cones;
		Vector<float> damp;
		mbik_rig_desc d = {};
	};
	template <typename IK>
	static void build_desc(IK &ik, Desc &D) {
		Skeleton3D *skeleton = ik.get_skeleton();
		const int nb = skeleton->get_bone_count();
		Vector<int32_t> &parent = D.parent;
		Vector<float> &rest = D.rest;
		Vector<mbik_pin_desc> &pd = D.pd;
		Vector<mbik_constraint_desc> &cd = D.cd;
		Vector<mbik_cone_desc> &cones = D.cones;
		mbik_rig_desc &d = D.d;
		parent.resize(nb);
		rest.resize(nb * 12);
		for (int b = 0; b < nb; b++) {
			parent.write[b] = skeleton->get_bone_parent(b);
			xform_to_floats(skeleton->get_bone_pose(b), rest.ptrw() + b * 12);
		}
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
		cd.clear();
		cones.clear();
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
		D.damp.resize(ik.bone_damp.size());
		for (int i = 0; i < ik.bone_damp.size(); i++) {
			D.damp.write[i] = ik.bone_damp[i];
		}
		d = mbik_rig_desc{};
		d.n_bones = nb;
		d.parent = parent.ptr();
		d.rest_local = rest.ptr();
		d.n_pins = (int32_t)pd.size();
		d.pins = pd.ptr();
		d.n_constraints = (int32_t)cd.size();
		d.constraints = cd.ptr();
		d.cones = cones.ptr();
		d.n_bone_damp = (int32_t)D.damp.size();
		d.bone_damp = D.damp.ptr();
		d.default_damp = ik.default_damp;
		d.iterations_per_frame = ik.iterations_per_frame;
		d.stabilization_passes = ik.stabilize_passes;
		d.constraint_mode = ik.is_constraint_mode ? 1 : 0;
	}

	// replaces ManyBoneIK3D::_bone_list_changed() (src/many_bone_ik_3d.cpp:1011-1068)
	template <typename IK>
	void rebuild(IK &ik) {
		Desc D;
		build_desc(ik, D);
		const mbik_rig_desc &d = D.d;
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

	// the early-outs of _process_modification that come after the rebuild (:669-683)
	template <typename IK>
	static bool frame_is_live(IK &ik) {
		bool has_pins = false; // :669-677
		for (int i = 0; i < ik.pins.size(); i++) {
			if (ik.pins[i].is_valid() && !ik.pins[i]->get_name().is_empty()) {
				has_pins = true;
				break;
			}
		}
		return has_pins && ik.is_enabled() && ik.is_visible(); // :675-683
	}
	// this frame's inputs of one node: np target transforms and nb start poses
	template <typename IK>
	static void gather_inputs(IK &ik, float *targets12, float *start12) {
		Skeleton3D *skeleton = ik.get_skeleton();
		const int nb = skeleton->get_bone_count(), np = (int)ik.pins.size();
		for (int i = 0; i < np; i++) { // IKEffector3D::update_target_global_transform (src/ik_effector_3d.cpp:77-84)
			Transform3D x;
			if (ik.pins[i].is_valid()) {
				Node3D *t = Object::cast_to<Node3D>(ik.get_node_or_null(ik.pins[i]->get_target_node()));
				if (t && t->is_visible_in_tree()) {
					x = skeleton->get_global_transform().affine_inverse() * t->get_global_transform();
				}
			}
			xform_to_floats(x, targets12 + i * 12);
		}
		for (int b = 0; b < nb; b++) { // IKBone3D::set_initial_pose (src/ik_bone_3d.cpp:161-168)
			xform_to_floats(skeleton->get_bone_pose(b), start12 + b * 12);
		}
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
		if (!frame_is_live(ik)) {
			return;
		}
		const int nb = skeleton->get_bone_count(), np = (int)ik.pins.size();
		targets.resize(np * 12);
		start.resize(nb * 12);
		out.resize(nb * 10);
		gather_inputs(ik, targets.ptrw(), start.ptrw());
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

	int crowd_group = -1; // CrowdBinding: the rig group this node belongs to (-1: not resolved / dirty)
};

// Deferred ("crowd") mode of the binding.  One ManyBoneIK3D node is one skeleton, so a scene of k characters that share a
// rig costs k one-pose launches per frame through Binding -- each ~0.6 ms of launch + copy latency for work the GPU does k
// at a time for the same price.  CrowdBinding gathers instead: every node's _process_modification() only ENQUEUES its
// frame (same early-outs, same rebuild-when-dirty, targets and start poses read at that moment), nodes whose rig
// description is byte-identical share one mbik_rig, and flush() -- called once per frame after all skeletons have been
// processed, e.g. from a scene-level manager node at NOTIFICATION_INTERNAL_PROCESS -- issues ONE mbik_solve_batch per rig
// for all of its nodes (MBIK_OUT_SOLVED_ONLY: only bone_list bones come back) and writes position / rotation / scale to
// every skeleton exactly like _update_skeleton_bones_transform (:104-116).  The only semantic difference from Binding is
// WHEN the skeleton receives the pose (at flush(), not inside the node's own _process_modification()).
//
//     // scene-level singleton
//     mbik_godot::CrowdBinding<ManyBoneIK3D> crowd;
//     void ManyBoneIK3D::_process_modification() { crowd.enqueue(*this, gpu /* this node's Binding member */); }
//     // once per frame, after the skeletons:   crowd.flush();
template <typename IK>
struct CrowdBinding {
	struct Group {
		mbik_rig *rig = nullptr;
		Vector<uint8_t> key; // the serialised rig description this rig was created from
		Vector<int32_t> bone_order;
		int n_bones = 0, n_pins = 0;
		Vector<IK *> nodes;          // this frame's members, in enqueue order
		Vector<int32_t> iterations;  // their iterations_per_frame, read at enqueue like the reference does every frame (:685)
		Vector<float> targets, start, out;
	};
	Vector<Group *> groups;
	int last_error = MBIK_OK;
	int launches_last_flush = 0; // mbik_solve_batch calls of the last flush() (one per rig unless iterations differ)

	~CrowdBinding() {
		for (int g = 0; g < groups.size(); g++) {
			mbik_rig_destroy(groups[g]->rig);
			delete groups[g];
		}
	}

	static void append_bytes(Vector<uint8_t> &k, const void *p, size_t n) {
		const int at = k.size();
		k.resize(at + (int)n);
		if (n) {
			memcpy(k.ptrw() + at, p, n);
		}
	}
	static void desc_key(const Binding::Desc &D, Vector<uint8_t> &k) {
		const mbik_rig_desc &d = D.d;
		const int32_t head[6] = { d.n_bones, d.n_pins, d.n_constraints, d.n_bone_damp, d.stabilization_passes, d.constraint_mode };
		append_bytes(k, head, sizeof(head));
		append_bytes(k, &d.default_damp, sizeof(float));
		append_bytes(k, D.parent.ptr(), sizeof(int32_t) * (size_t)D.parent.size());
		append_bytes(k, D.rest.ptr(), sizeof(float) * (size_t)D.rest.size());
		append_bytes(k, D.pd.ptr(), sizeof(mbik_pin_desc) * (size_t)D.pd.size());
		append_bytes(k, D.cd.ptr(), sizeof(mbik_constraint_desc) * (size_t)D.cd.size());
		append_bytes(k, D.cones.ptr(), sizeof(mbik_cone_desc) * (size_t)D.cones.size());
		append_bytes(k, D.damp.ptr(), sizeof(float) * (size_t)D.damp.size());
	}

	// resolves (creating on first sight) the group of a rebuilt node; -1 if mbik_rig_create failed
	int group_of(IK &ik) {
		Binding::Desc D;
		Binding::build_desc(ik, D);
		Vector<uint8_t> key;
		desc_key(D, key);
		for (int g = 0; g < groups.size(); g++) {
			if (groups[g]->key.size() == key.size() && memcmp(groups[g]->key.ptr(), key.ptr(), (size_t)key.size()) == 0) {
				return g;
			}
		}
		mbik_rig *rig = nullptr;
		last_error = mbik_rig_create(&D.d, &rig);
		if (last_error != MBIK_OK) {
			ERR_PRINT(mbik_last_error());
			return -1;
		}
		Group *G = new Group();
		G->rig = rig;
		G->key = key;
		G->n_bones = D.d.n_bones;
		G->n_pins = D.d.n_pins;
		mbik_rig_info info;
		mbik_rig_get_info(rig, &info);
		G->bone_order.resize(info.n_solved);
		mbik_rig_get_bone_order(rig, G->bone_order.ptrw());
		groups.push_back(G);
		return groups.size() - 1;
	}

	// replaces the body of ManyBoneIK3D::_process_modification(): everything but the solve and the write-back
	void enqueue(IK &ik, Binding &state) {
		Skeleton3D *skeleton = ik.get_skeleton();
		if (!skeleton || ik.get_effector_count() == 0) { // :646-651
			return;
		}
		if (ik.is_dirty || state.crowd_group < 0) { // :652-658
			ik.is_dirty = false;
			state.crowd_group = group_of(ik);
		}
		if (state.crowd_group < 0 || !Binding::frame_is_live(ik)) {
			return;
		}
		Group &G = *groups[state.crowd_group];
		const int at = G.nodes.size();
		G.nodes.push_back(&ik);
		G.iterations.push_back((int32_t)ik.get_iterations_per_frame());
		G.targets.resize((at + 1) * G.n_pins * 12);
		G.start.resize((at + 1) * G.n_bones * 12);
		Binding::gather_inputs(ik, G.targets.ptrw() + at * G.n_pins * 12, G.start.ptrw() + at * G.n_bones * 12);
	}

	// one launch per rig (per distinct iterations_per_frame among its nodes), then the write-back of every node
	void flush() {
		launches_last_flush = 0;
		for (int g = 0; g < groups.size(); g++) {
			Group &G = *groups[g];
			const int n = G.nodes.size(), ns = G.bone_order.size();
			if (n == 0) {
				continue;
			}
			G.out.resize(n * ns * 10);
			Vector<char> done;
			done.resize(n);
			memset(done.ptrw(), 0, (size_t)n);
			for (int first = 0; first < n; first++) {
				if (done[first]) {
					continue;
				}
				// the nodes of this rig that share `first`'s iteration count: normally all of them, in place
				int same = 0;
				for (int k = first; k < n; k++) {
					same += (!done[k] && G.iterations[k] == G.iterations[first]) ? 1 : 0;
				}
				mbik_solve_params p = {};
				p.iterations = G.iterations[first];
				p.device = -1;
				p.flags = MBIK_IO_HOST | MBIK_OUT_SOLVED_ONLY;
				int rc;
				if (first == 0 && same == n) {
					rc = mbik_solve_batch(G.rig, &p, (size_t)n, G.targets.ptr(), G.start.ptr(), G.out.ptrw(), nullptr, nullptr);
					memset(done.ptrw(), 1, (size_t)n);
				} else {
					Vector<float> t, s, o;
					Vector<int> who;
					for (int k = first; k < n; k++) {
						if (!done[k] && G.iterations[k] == G.iterations[first]) {
							who.push_back(k);
							done.write[k] = 1;
						}
					}
					const int m = who.size();
					t.resize(m * G.n_pins * 12);
					s.resize(m * G.n_bones * 12);
					o.resize(m * ns * 10);
					for (int j = 0; j < m; j++) {
						memcpy(t.ptrw() + j * G.n_pins * 12, G.targets.ptr() + who[j] * G.n_pins * 12, sizeof(float) * 12 * (size_t)G.n_pins);
						memcpy(s.ptrw() + j * G.n_bones * 12, G.start.ptr() + who[j] * G.n_bones * 12, sizeof(float) * 12 * (size_t)G.n_bones);
					}
					rc = mbik_solve_batch(G.rig, &p, (size_t)m, t.ptr(), s.ptr(), o.ptrw(), nullptr, nullptr);
					for (int j = 0; j < m && rc == MBIK_OK; j++) {
						memcpy(G.out.ptrw() + who[j] * ns * 10, o.ptr() + j * ns * 10, sizeof(float) * 10 * (size_t)ns);
					}
				}
				launches_last_flush++;
				if (rc != MBIK_OK) {
					last_error = rc;
					ERR_PRINT_ONCE(mbik_last_error()); // the poses of this rig's nodes are left untouched, like ERR_FAIL_*
					G.nodes.clear();
					break;
				}
			}
			for (int k = 0; k < G.nodes.size(); k++) {
				IK &ik = *G.nodes[k];
				Skeleton3D *skeleton = ik.get_skeleton();
				const float *rows = G.out.ptr() + k * ns * 10;
				for (int i = ns; i-- > 0;) { // _update_skeleton_bones_transform (:104-116, src/ik_bone_3d.cpp:170-179)
					const int b = G.bone_order[i];
					const float *o = rows + i * 10;
					skeleton->set_bone_pose_position(b, Vector3(o[0], o[1], o[2]));
					skeleton->set_bone_pose_rotation(b, Quaternion(o[3], o[4], o[5], o[6]));
					skeleton->set_bone_pose_scale(b, Vector3(o[7], o[8], o[9]));
				}
				ik.update_gizmos();
			}
			G.nodes.clear();
			G.iterations.clear();
		}
	}
};

} // namespace mbik_godot
