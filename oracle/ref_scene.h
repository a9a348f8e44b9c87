// TEST INFRASTRUCTURE ONLY -- the headless scene that drives the reference module's own code (see ref_harness.cpp for
// what is and is not "the reference" here).  Shared by oracle/ref_harness.cpp (libmbik_ref.so: the reference alone) and
// oracle/ref_binding_harness.cpp (libmbik_ref_binding.so: the same scene with ManyBoneIK3D::_process_modification
// replaced by the libmbik.so binding of many_bone_ik_b200/host/godot_module_binding.h).
#pragma once
#include <atomic>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "godot_shim/godot_shim.h"

// The harness reads two private tables the module offers no accessor for (bone_damp: written only by
// _set_bone_count with the default value, src/many_bone_ik_3d.cpp:756-763; heading_weights: facts export), and the
// binding is written as member code of ManyBoneIK3D.  Access specifiers do not change layout or
// mangling, so these TUs see the same classes the module TUs define.
#define private public
#define protected public
#include "ik_bone_segment_3d.h"
#include "ik_kusudama_3d.h"
#include "ik_open_cone_3d.h"
#include "many_bone_ik_3d.h"
#include "math/qcp.h"
#undef private
#undef protected

#include "../include/mbik.h"

namespace ref_scene {

inline Transform3D load_xform(const float *p) {
	Transform3D t;
	for (int r = 0; r < 3; r++) {
		for (int c = 0; c < 3; c++) {
			t.basis.rows[r][c] = p[r * 3 + c];
		}
	}
	t.origin = Vector3(p[9], p[10], p[11]);
	return t;
}

inline void store_xform(float *o, const Transform3D &t) {
	for (int r = 0; r < 3; r++) {
		for (int c = 0; c < 3; c++) {
			o[r * 3 + c] = t.basis.rows[r][c];
		}
	}
	o[9] = t.origin.x;
	o[10] = t.origin.y;
	o[11] = t.origin.z;
}

inline String bone_name(int b) {
	return String("bone_") + itos(b);
}

// One headless scene: Skeleton3D + ManyBoneIK3D (child of the skeleton) + one target Node3D per pin
// (children of the modifier, addressed by relative node paths "target_<i>").
struct RefScene {
	Skeleton3D *skeleton = nullptr;
	ManyBoneIK3D *ik = nullptr;
	std::vector<Node3D *> targets;
	std::vector<Transform3D> solved_local; // IK bone poses captured when "modification_processed" fires
	std::vector<char> solved;

	// bones the frame wrote, when the node under test does not keep an IK bone list (the GPU binding)
	std::function<bool(std::vector<char> &)> solved_override;

	RefScene(const mbik_rig_desc *d, int iterations, ManyBoneIK3D *p_node = nullptr) {
		skeleton = new Skeleton3D();
		for (int b = 0; b < d->n_bones; b++) {
			skeleton->add_bone(bone_name(b));
		}
		for (int b = 0; b < d->n_bones; b++) {
			skeleton->set_bone_parent(b, d->parent[b]);
			Transform3D rest = load_xform(d->rest_local + 12 * b);
			skeleton->set_bone_rest(b, rest);
			skeleton->set_bone_pose(b, rest); // the rig is built with the skeleton in its rest pose (mbik.h)
		}
		ik = p_node ? p_node : new ManyBoneIK3D();
		solved_local.resize((size_t)d->n_bones);
		solved.assign((size_t)d->n_bones, 0);
		// connected before the module's own handler, so it sees the solved IK bones before they are re-seeded
		Callable capture;
		capture.object = this;
		capture.method_id = "harness_capture";
		capture.fn = [this]() { capture_locals(); };
		ik->connect(SNAME("modification_processed"), capture);
		ik->shim_attach_skeleton(skeleton); // -> _skeleton_changed: signal connections + first _bone_list_changed

		for (int p = 0; p < d->n_pins; p++) {
			Node3D *t = new Node3D();
			t->set_name(String("target_") + itos(p));
			ik->add_child(t);
			targets.push_back(t);
		}
		// configuration through the property paths a saved scene uses (ManyBoneIK3D::_set, :296-375)
		ik->set("pin_count", d->n_pins);
		for (int p = 0; p < d->n_pins; p++) {
			const mbik_pin_desc &pd = d->pins[p];
			String base = String("pins/") + itos(p) + "/";
			ik->set(StringName(base + "bone_name"), pd.bone >= 0 ? bone_name(pd.bone) : String());
			ik->set(StringName(base + "target_node"), NodePath(String("target_") + itos(p)));
			ik->set(StringName(base + "motion_propagation_factor"), pd.motion_propagation_factor);
			ik->set(StringName(base + "weight"), pd.weight);
			ik->set(StringName(base + "direction_priorities"),
					Vector3(pd.direction_priorities[0], pd.direction_priorities[1], pd.direction_priorities[2]));
		}
		ik->set("constraint_count", d->n_constraints);
		for (int c = 0; c < d->n_constraints; c++) {
			const mbik_constraint_desc &cd = d->constraints[c];
			String base = String("constraints/") + itos(c) + "/";
			ik->set(StringName(base + "bone_name"), bone_name(cd.bone));
			ik->set(StringName(base + "twist_from"), cd.twist_from);
			ik->set(StringName(base + "twist_range"), cd.twist_range);
			ik->set(StringName(base + "kusudama_open_cone_count"), cd.n_cones);
			for (int j = 0; j < cd.n_cones; j++) {
				const mbik_cone_desc &cone = d->cones[cd.cone_offset + j];
				String cbase = base + "kusudama_open_cone/" + itos(j) + "/";
				ik->set(StringName(cbase + "center"), Vector3(cone.center[0], cone.center[1], cone.center[2]));
				ik->set(StringName(cbase + "radius"), cone.radius);
			}
		}
		ik->set_default_damp(d->default_damp);
		ik->set_iterations_per_frame((float)(iterations >= 0 ? iterations : d->iterations_per_frame));
		ik->set_stabilization_passes(d->stabilization_passes);
		ik->set_constraint_mode(d->constraint_mode != 0);
		// bone_damp: no public writer in the module (see the note at the includes)
		ik->bone_damp.resize(d->n_bone_damp);
		for (int i = 0; i < d->n_bone_damp; i++) {
			ik->bone_damp.write[i] = d->bone_damp[i];
		}
		ik->set_dirty();
		// the rebuild happens inside the next frame (is_dirty, :655-658) with the skeleton still in its rest pose;
		// run that frame now so that every later frame only re-seeds and solves
		ik->process_modification();
	}

	~RefScene() {
		for (Node3D *t : targets) {
			delete t;
		}
		delete ik;
		delete skeleton;
	}

	void capture_locals() {
		std::fill(solved.begin(), solved.end(), 0);
		Vector<Ref<IKBone3D>> list = ik->get_bone_list();
		for (int i = 0; i < list.size(); i++) {
			Ref<IKBone3D> b = list[i];
			if (b.is_null() || b->get_bone_id() < 0) {
				continue;
			}
			solved_local[(size_t)b->get_bone_id()] = b->get_pose();
			solved[(size_t)b->get_bone_id()] = 1;
		}
	}

	// keep_skeleton: do not touch the skeleton's bone poses before the frame -- the node lives on from the previous frame,
	// its IK bones re-seed from what the skeleton holds after the previous write-back (get_bone_pose(): recomposed from the
	// position / rotation / scale IKBone3D::set_skeleton_bone_pose wrote).  out_skeleton12 (nullable): get_bone_pose() of
	// every bone after the frame.
	void solve(const mbik_rig_desc *d, const float *targets12, const float *start12, float *out10, float *out_local12, uint32_t *status,
			bool keep_skeleton = false, float *out_skeleton12 = nullptr) {
		begin_frame(d, targets12, start12, keep_skeleton);
		end_frame(d, out10, out_local12, status, out_skeleton12);
	}

	std::vector<Transform3D> start; // the skeleton's bone poses at the start of the current frame
	bool wrote = false;

	// everything up to and including SkeletonModifier3D::process_modification() ...
	void begin_frame(const mbik_rig_desc *d, const float *targets12, const float *start12, bool keep_skeleton = false) {
		start.resize((size_t)d->n_bones);
		for (int b = 0; b < d->n_bones; b++) {
			if (keep_skeleton) {
				start[(size_t)b] = skeleton->get_bone_pose(b);
				continue;
			}
			start[(size_t)b] = load_xform(start12 ? start12 + 12 * b : d->rest_local + 12 * b);
			skeleton->set_bone_pose(b, start[(size_t)b]);
		}
		for (int p = 0; p < d->n_pins; p++) {
			targets[(size_t)p]->set_global_transform(load_xform(targets12 + 12 * p));
		}
		// the frame boundary: the previous frame's "modification_processed" re-seeds IK bones and targets
		ik->emit_signal(SNAME("modification_processed"));
		std::fill(solved.begin(), solved.end(), 0);
		wrote = frame_will_write(d);
		ik->process_modification();
	}

	// ... and the read-back of what the frame left on the skeleton (a deferred binding writes between the two)
	void end_frame(const mbik_rig_desc *d, float *out10, float *out_local12, uint32_t *status, float *out_skeleton12 = nullptr) {
		bool have_locals = true;
		if (solved_override) {
			have_locals = false;
			wrote = solved_override(solved);
		}
		uint32_t st = 0;
		for (int b = 0; b < d->n_bones; b++) {
			bool is_solved = wrote && solved[(size_t)b];
			Transform3D local = (is_solved && have_locals) ? solved_local[(size_t)b] : start[(size_t)b];
			if (out_local12) {
				store_xform(out_local12 + 12 * b, local);
			}
			if (have_locals && !local.basis.is_finite()) {
				st |= MBIK_POSE_NONFINITE_RESET;
			}
			Vector3 pos;
			Quaternion rot;
			Vector3 scl;
			if (is_solved) {
				// exactly what IKBone3D::set_skeleton_bone_pose handed to the skeleton (src/ik_bone_3d.cpp:170-179)
				pos = skeleton->get_bone_pose_position(b);
				rot = skeleton->get_bone_pose_rotation(b);
				scl = skeleton->get_bone_pose_scale(b);
			} else {
				// bones the solver does not own pass through (mbik.h): the same decomposition of the start pose
				Basis basis = local.basis;
				if (!basis.is_finite()) {
					basis = Basis();
				}
				pos = local.origin;
				rot = basis.get_rotation_quaternion();
				scl = basis.get_scale();
			}
			float *o = out10 + 10 * b;
			o[0] = pos.x; o[1] = pos.y; o[2] = pos.z;
			o[3] = rot.x; o[4] = rot.y; o[5] = rot.z; o[6] = rot.w;
			o[7] = scl.x; o[8] = scl.y; o[9] = scl.z;
		}
		if (status) {
			*status = st;
		}
		if (out_skeleton12) {
			for (int b = 0; b < d->n_bones; b++) {
				store_xform(out_skeleton12 + 12 * b, skeleton->get_bone_pose(b));
			}
		}
	}

	// _process_modification returns before the write-back when there is no pin (or none with a bone name)
	// (:649-651, :671-677); then the skeleton keeps the start pose.
	bool frame_will_write(const mbik_rig_desc *d) const {
		for (int p = 0; p < d->n_pins; p++) {
			if (d->pins[p].bone >= 0) {
				return true;
			}
		}
		return false;
	}
};

} // namespace ref_scene
