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
#include "ref_scene.h"

using namespace ref_scene;

namespace {

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

// A node living across frames: per pose one scene; frame 0 starts from start_pose (or the rest pose), every later frame
// from whatever the skeleton holds after the previous frame's write-back -- nothing re-seeds it from outside.
//   targets [n_frames][n_poses][n_pins][12]; out_pose [n_frames][n_poses][n_bones][10]; out_local (raw IK-bone transforms
//   captured at modification_processed) and out_skeleton (get_bone_pose() after the frame) [n_frames][n_poses][n_bones][12],
//   both nullable; out_status [n_frames][n_poses] nullable.
int ref_solve_frames(const mbik_rig_desc *d, size_t n_poses, int n_frames, const float *targets, const float *start_pose, float *out_pose,
		float *out_local, float *out_skeleton, uint32_t *out_status, int iterations, int n_threads) {
	if (!d || !targets || !out_pose || n_frames < 1) {
		return -1;
	}
	auto range = [&](size_t begin, size_t end) {
		for (size_t k = begin; k < end; k++) {
			RefScene scene(d, iterations);
			for (int f = 0; f < n_frames; f++) {
				const size_t row = (size_t)f * n_poses + k;
				scene.solve(d, targets + row * (size_t)d->n_pins * 12, start_pose ? start_pose + k * (size_t)d->n_bones * 12 : nullptr,
						out_pose + row * (size_t)d->n_bones * 10, out_local ? out_local + row * (size_t)d->n_bones * 12 : nullptr,
						out_status ? out_status + row : nullptr, f > 0, out_skeleton ? out_skeleton + row * (size_t)d->n_bones * 12 : nullptr);
			}
		}
	};
	if (n_threads <= 1 || n_poses < 2) {
		range(0, n_poses);
		return 0;
	}
	std::vector<std::thread> th;
	for (int t = 0; t < n_threads; t++) {
		th.emplace_back(range, n_poses * (size_t)t / (size_t)n_threads, n_poses * (size_t)(t + 1) / (size_t)n_threads);
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
