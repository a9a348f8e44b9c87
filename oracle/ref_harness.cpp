// TEST INFRASTRUCTURE ONLY -- C entry points of oracle/_ref/libmbik_ref.so: the reference module's OWN solver
// (every translation unit under /root/reference/src, compiled unmodified from where it lies) driven headless
// over the engine stand-in oracle/godot_shim/.  Loaded with ctypes by tests/ (to pin the restatement oracle and,
// through the committed fixtures it generates, the CUDA path) and by bench.py's cpu_baseline / --impl reference
// legs.  The product (many_bone_ik_b200/) never links or loads it.
//
// What is the reference here and what is not:
//   * the module logic -- segment construction, effector weights, QCP, damping, kusudama snaps, IKNode3D caches,
//     property-path configuration (_set), the per-frame entry _process_modification -- is the reference's code;
//   * the engine underneath (core/math, containers, object model, Skeleton3D) is the stand-in: core/math is the
//     same restatement the oracle uses (oracle/godot_math.h), so this library pins the oracle's reading of the
//     MODULE, not of the engine's arithmetic.
//
// Driving sequence per pose (the engine's frame, src/many_bone_ik_3d.cpp:645-694, :1070-1086):
//   skeleton bone poses <- start pose; target nodes <- targets; emit "modification_processed" (the module's own
//   connection re-seeds the IK bones and targets, :1084 -> :91-102); SkeletonModifier3D::process_modification()
//   (-> _process_modification(): rebuild if dirty, iterations x segment_solver, write-back).
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "godot_shim/godot_shim.h"

// The harness reads two private tables the module offers no accessor for (bone_damp: written only by
// _set_bone_count with the default value, src/many_bone_ik_3d.cpp:756-763; heading_weights: facts export).
// Access specifiers do not change layout or mangling, so this TU sees the same classes the module TUs define.
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

namespace {

Transform3D load_xform(const float *p) {
	Transform3D t;
	for (int r = 0; r < 3; r++) {
		for (int c = 0; c < 3; c++) {
			t.basis.rows[r][c] = p[r * 3 + c];
		}
	}
	t.origin = Vector3(p[9], p[10], p[11]);
	return t;
}

void store_xform(float *o, const Transform3D &t) {
	for (int r = 0; r < 3; r++) {
		for (int c = 0; c < 3; c++) {
			o[r * 3 + c] = t.basis.rows[r][c];
		}
	}
	o[9] = t.origin.x;
	o[10] = t.origin.y;
	o[11] = t.origin.z;
}

String bone_name(int b) {
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

	RefScene(const mbik_rig_desc *d, int iterations) {
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
		ik = new ManyBoneIK3D();
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

	void solve(const mbik_rig_desc *d, const float *targets12, const float *start12, float *out10, float *out_local12, uint32_t *status) {
		std::vector<Transform3D> start((size_t)d->n_bones);
		for (int b = 0; b < d->n_bones; b++) {
			start[(size_t)b] = load_xform(start12 ? start12 + 12 * b : d->rest_local + 12 * b);
			skeleton->set_bone_pose(b, start[(size_t)b]);
		}
		for (int p = 0; p < d->n_pins; p++) {
			targets[(size_t)p]->set_global_transform(load_xform(targets12 + 12 * p));
		}
		// the frame boundary: the previous frame's "modification_processed" re-seeds IK bones and targets
		ik->emit_signal(SNAME("modification_processed"));
		std::fill(solved.begin(), solved.end(), 0);
		bool wrote = frame_will_write(d);
		ik->process_modification();
		uint32_t st = 0;
		for (int b = 0; b < d->n_bones; b++) {
			bool is_solved = wrote && solved[(size_t)b];
			Transform3D local = is_solved ? solved_local[(size_t)b] : start[(size_t)b];
			if (out_local12) {
				store_xform(out_local12 + 12 * b, local);
			}
			if (!local.basis.is_finite()) {
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

void solve_range(const mbik_rig_desc *d, size_t begin, size_t end, const float *targets, const float *start_pose, float *out_pose,
		float *out_local, uint32_t *out_status, int iterations, bool rebuild_each) {
	RefScene *scene = nullptr;
	for (size_t k = begin; k < end; k++) {
		if (!scene || rebuild_each) {
			delete scene;
			scene = new RefScene(d, iterations);
		}
		scene->solve(d, targets + k * (size_t)d->n_pins * 12, start_pose ? start_pose + k * (size_t)d->n_bones * 12 : nullptr,
				out_pose + k * (size_t)d->n_bones * 10, out_local ? out_local + k * (size_t)d->n_bones * 12 : nullptr,
				out_status ? out_status + k : nullptr);
	}
	delete scene;
}

} // namespace

extern "C" {

// Same contract as orc_solve_batch (oracle_capi.cpp): flags bit0 = fresh scene per pose; otherwise one
// long-lived scene per worker thread (the reference's node living across frames).
int ref_solve_batch(const mbik_rig_desc *d, size_t n_poses, const float *targets, const float *start_pose, float *out_pose,
		float *out_local, uint32_t *out_status, int iterations, int n_threads, unsigned flags) {
	if (!d || !targets || !out_pose) {
		return -1;
	}
	bool rebuild_each = (flags & 1u) != 0;
	if (n_threads <= 1 || n_poses < 2) {
		solve_range(d, 0, n_poses, targets, start_pose, out_pose, out_local, out_status, iterations, rebuild_each);
		return 0;
	}
	std::vector<std::thread> th;
	for (int t = 0; t < n_threads; t++) {
		size_t b = n_poses * (size_t)t / (size_t)n_threads, e = n_poses * (size_t)(t + 1) / (size_t)n_threads;
		th.emplace_back(solve_range, d, b, e, targets, start_pose, out_pose, out_local, out_status, iterations, rebuild_each);
	}
	for (auto &t : th) {
		t.join();
	}
	return 0;
}

// Setup facts of the rig the reference builds (for the flattener / oracle cross-checks): solve order of the
// bones, segment count, bone-direction and twist-axes bases per solved bone.  Returns the number of solved bones.
int ref_rig_facts(const mbik_rig_desc *d, int32_t *bone_order, int32_t capacity, int32_t *n_segments, float *dir_basis, float *twist_basis) {
	if (!d) {
		return -1;
	}
	RefScene scene(d, 0);
	Vector<Ref<IKBone3D>> list = scene.ik->get_bone_list();
	int n = (int)list.size();
	for (int i = 0; i < n && i < capacity; i++) {
		if (bone_order) {
			bone_order[i] = list[i]->get_bone_id();
		}
		Transform3D dt = list[i]->get_bone_direction_transform()->get_transform();
		Transform3D tt = list[i]->get_constraint_twist_transform()->get_transform();
		for (int r = 0; r < 3; r++) {
			for (int c = 0; c < 3; c++) {
				if (dir_basis) {
					dir_basis[i * 9 + r * 3 + c] = dt.basis.rows[r][c];
				}
				if (twist_basis) {
					twist_basis[i * 9 + r * 3 + c] = tt.basis.rows[r][c];
				}
			}
		}
	}
	if (n_segments) {
		int cnt = 0;
		std::vector<Ref<IKBoneSegment3D>> stack;
		for (Ref<IKBoneSegment3D> s : scene.ik->get_segmented_skeletons()) {
			stack.push_back(s);
		}
		while (!stack.empty()) {
			Ref<IKBoneSegment3D> s = stack.back();
			stack.pop_back();
			cnt++;
			for (Ref<IKBoneSegment3D> c : s->get_child_segments()) {
				stack.push_back(c);
			}
		}
		*n_segments = cnt;
	}
	return n;
}

// heading weights of the segment that owns bone_list[step]; returns their count (or -1)
int ref_step_weights(const mbik_rig_desc *d, int32_t step, double *out, int32_t capacity) {
	if (!d) {
		return -1;
	}
	RefScene scene(d, 0);
	Vector<Ref<IKBone3D>> list = scene.ik->get_bone_list();
	if (step < 0 || step >= (int)list.size()) {
		return -1;
	}
	Ref<IKBone3D> bone = list[step];
	std::vector<Ref<IKBoneSegment3D>> stack;
	for (Ref<IKBoneSegment3D> s : scene.ik->get_segmented_skeletons()) {
		stack.push_back(s);
	}
	while (!stack.empty()) {
		Ref<IKBoneSegment3D> s = stack.back();
		stack.pop_back();
		for (int i = 0; i < s->bones.size(); i++) {
			if (s->bones[i] == bone) {
				int n = (int)s->heading_weights.size();
				for (int k = 0; k < n && k < capacity; k++) {
					out[k] = s->heading_weights[k];
				}
				return n;
			}
		}
		for (Ref<IKBoneSegment3D> c : s->get_child_segments()) {
			stack.push_back(c);
		}
	}
	return -1;
}

// cone geometry after setup, constraint-row order: control point, tangent circle centres 1 and 2 per cone
int ref_cone_geometry(const mbik_rig_desc *d, float *out, int32_t capacity_cones) {
	if (!d) {
		return -1;
	}
	RefScene scene(d, 0);
	Vector<Ref<IKBone3D>> list = scene.ik->get_bone_list();
	int k = 0;
	for (int ci = 0; ci < d->n_constraints; ci++) {
		Ref<IKBone3D> bone;
		for (int i = 0; i < list.size(); i++) {
			if (list[i]->get_bone_id() == d->constraints[ci].bone) {
				bone = list[i];
				break;
			}
		}
		for (int j = 0; j < d->constraints[ci].n_cones; j++, k++) {
			if (k >= capacity_cones) {
				continue;
			}
			float *o = out + k * 9;
			if (bone.is_null() || bone->get_constraint().is_null() || j >= (int)bone->get_constraint()->open_cones.size()) {
				for (int q = 0; q < 9; q++) {
					o[q] = 0;
				}
				continue;
			}
			Ref<IKLimitCone3D> c = bone->get_constraint()->open_cones[j];
			Vector3 cp = c->get_control_point(), t1 = c->get_tangent_circle_center_next_1(), t2 = c->get_tangent_circle_center_next_2();
			o[0] = cp.x; o[1] = cp.y; o[2] = cp.z;
			o[3] = t1.x; o[4] = t1.y; o[5] = t1.z;
			o[6] = t2.x; o[7] = t2.y; o[8] = t2.z;
		}
	}
	return k;
}

// ---- stage-level entry points (same signatures as the orc_* ones) --------------------------------
void ref_qcp_weighted_superpose(const float *moved, const float *target, const double *weight, int n, int translate, float *out7) {
	PackedVector3Array m, t;
	Vector<double> w;
	m.resize(n);
	t.resize(n);
	w.resize(n);
	for (int i = 0; i < n; i++) {
		m.write[i] = Vector3(moved[3 * i], moved[3 * i + 1], moved[3 * i + 2]);
		t.write[i] = Vector3(target[3 * i], target[3 * i + 1], target[3 * i + 2]);
		w.write[i] = weight[i];
	}
	QCP qcp(1e-6);
	Quaternion q = qcp.weighted_superpose(m, t, w, translate != 0);
	Vector3 tr = qcp.get_translation();
	out7[0] = q.x; out7[1] = q.y; out7[2] = q.z; out7[3] = q.w;
	out7[4] = tr.x; out7[5] = tr.y; out7[6] = tr.z;
}

void ref_kusudama_point_in_limits(const float *cones, int n_cones, const float *point, float *out4) {
	Ref<IKKusudama3D> k;
	k.instantiate();
	for (int i = 0; i < n_cones; i++) {
		Ref<IKLimitCone3D> c;
		c.instantiate();
		c->set_attached_to(k);
		c->set_radius(MAX(1.0e-38, cones[4 * i + 3]));
		c->set_control_point(Vector3(cones[4 * i], cones[4 * i + 1], cones[4 * i + 2]).normalized());
		k->add_open_cone(c);
	}
	Vector<double> bounds;
	bounds.resize(2);
	bounds.write[0] = 0;
	bounds.write[1] = 0;
	Vector3 r = k->get_local_point_in_limits(Vector3(point[0], point[1], point[2]), &bounds);
	out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = (float)bounds[0];
}

void ref_clamp_to_cos_half_angle(const float *q4, double cos_half, float *out4) {
	Quaternion q = IKBoneSegment3D::clamp_to_cos_half_angle(Quaternion(q4[0], q4[1], q4[2], q4[3]), cos_half);
	out4[0] = q.x; out4[1] = q.y; out4[2] = q.z; out4[3] = q.w;
}

void ref_swing_twist_y(const float *q4, float *out8) {
	Quaternion s, t;
	IKKusudama3D::get_swing_twist(Quaternion(q4[0], q4[1], q4[2], q4[3]), Vector3(0, 1, 0), s, t);
	out8[0] = s.x; out8[1] = s.y; out8[2] = s.z; out8[3] = s.w;
	out8[4] = t.x; out8[5] = t.y; out8[6] = t.z; out8[7] = t.w;
}

int ref_hardware_threads(void) {
	unsigned n = std::thread::hardware_concurrency();
	return n ? (int)n : 1;
}

} // extern "C"
