// TEST INFRASTRUCTURE ONLY -- C entry points of the CPU oracle (liboracle.so), loaded with ctypes by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
// Uses include/mbik.h only for the plain rig-description structs (the product never links this).
#include "../include/mbik.h"
#include "ewbik_oracle.h"

#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

using namespace orc;

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

// Fill the ManyBoneIK3D tables from the plain description (the job of the property setters,
// reference src/many_bone_ik_3d.cpp:436-620, which are out of scope as Godot Variant plumbing).
void configure(ManyBoneIK3D &ik, const mbik_rig_desc *d) {
	Skeleton3D &sk = ik.skeleton_storage;
	sk.parent.assign(d->parent, d->parent + d->n_bones);
	sk.pose.resize(d->n_bones);
	for (int b = 0; b < d->n_bones; b++) {
		sk.pose[b] = load_xform(d->rest_local + 12 * b);
	}
	ik.pins.resize(d->n_pins);
	ik.pin_count = d->n_pins;
	for (int i = 0; i < d->n_pins; i++) {
		ik.pins[i].bone = d->pins[i].bone;
		ik.pins[i].weight = d->pins[i].weight;
		ik.pins[i].motion_propagation_factor = d->pins[i].motion_propagation_factor;
		ik.pins[i].priority_direction = Vector3(d->pins[i].direction_priorities[0], d->pins[i].direction_priorities[1], d->pins[i].direction_priorities[2]);
	}
	ik.constraint_count = d->n_constraints;
	ik.constraint_names.resize(d->n_constraints);
	ik.joint_twist_x.resize(d->n_constraints);
	ik.joint_twist_y.resize(d->n_constraints);
	ik.kusudama_open_cone_count.resize(d->n_constraints);
	ik.kusudama_open_cones.resize(d->n_constraints);
	for (int i = 0; i < d->n_constraints; i++) {
		const mbik_constraint_desc &c = d->constraints[i];
		ik.constraint_names[i] = c.bone;
		ik.joint_twist_x[i] = c.twist_from;
		ik.joint_twist_y[i] = c.twist_range;
		ik.kusudama_open_cone_count[i] = c.n_cones;
		ik.kusudama_open_cones[i].resize(c.n_cones);
		for (int j = 0; j < c.n_cones; j++) {
			const mbik_cone_desc &cd = d->cones[c.cone_offset + j];
			// set_kusudama_open_cone_center (many_bone_ik_3d.cpp:586-601): zero-ish centre -> (0,1,0)
			Vector3 ctr(cd.center[0], cd.center[1], cd.center[2]);
			if (Math::is_zero_approx(ctr.length_squared())) {
				ctr = Vector3(0, 1, 0);
			}
			ik.kusudama_open_cones[i][j].x = ctr.x;
			ik.kusudama_open_cones[i][j].y = ctr.y;
			ik.kusudama_open_cones[i][j].z = ctr.z;
			ik.kusudama_open_cones[i][j].w = cd.radius;
		}
	}
	ik.bone_damp.assign(d->bone_damp, d->bone_damp + d->n_bone_damp);
	ik.default_damp = d->default_damp;
	ik.iterations_per_frame = d->iterations_per_frame;
	ik.stabilize_passes = d->stabilization_passes;
	ik.is_constraint_mode = d->constraint_mode != 0;
	ik.pin_targets.assign(d->n_pins, Transform3D());
}

void solve_range(const mbik_rig_desc *d, size_t begin, size_t end, const float *targets, const float *start_pose,
		float *out_pose, float *out_local, uint32_t *out_status, int iterations, bool rebuild_each, long long *counters) {
	ManyBoneIK3D *ik = nullptr;
	long long steps = 0, swing_calls = 0, swing_rect = 0;
	auto harvest = [&](ManyBoneIK3D *k) {
		for (auto &b : k->bone_list) {
			swing_calls += b->constraint->n_swing_calls;
			swing_rect += b->constraint->n_swing_rectified;
		}
		std::vector<Ref<IKBoneSegment3D>> stack(k->segmented_skeletons.begin(), k->segmented_skeletons.end());
		while (!stack.empty()) {
			Ref<IKBoneSegment3D> s = stack.back();
			stack.pop_back();
			steps += s->n_bone_steps;
			for (auto &c : s->child_segments) {
				stack.push_back(c);
			}
		}
	};
	for (size_t k = begin; k < end; k++) {
		if (!ik || rebuild_each) {
			if (ik) {
				harvest(ik);
				delete ik;
			}
			ik = new ManyBoneIK3D();
			configure(*ik, d);
			if (iterations >= 0) {
				ik->iterations_per_frame = iterations;
			}
			// rig is always built with the skeleton in its REST pose (mbik.h contract)
			ik->_bone_list_changed();
		}
		for (int p = 0; p < d->n_pins; p++) {
			ik->pin_targets[p] = load_xform(targets + (k * d->n_pins + p) * 12);
		}
		Skeleton3D &sk = ik->skeleton_storage;
		for (int b = 0; b < d->n_bones; b++) {
			sk.pose[b] = load_xform(start_pose ? start_pose + (k * d->n_bones + b) * 12 : d->rest_local + 12 * b);
		}
		ik->_update_ik_bones_transform(); // the modification_processed re-seed (many_bone_ik_3d.cpp:1084, :91-102)
		// early-outs of _process_modification (many_bone_ik_3d.cpp:649-651, :671-680): no pins at all, or no pin
		// with a non-empty bone name (bone < 0 stands for the empty name) -> the skeleton is left untouched
		bool has_pins = false;
		for (int p = 0; p < d->n_pins; p++) {
			has_pins = has_pins || d->pins[p].bone >= 0;
		}
		if (d->n_pins > 0 && has_pins) {
			ik->solve_iterations();
		}
		ik->write_skeleton_pose(out_pose + k * d->n_bones * 10, out_local ? out_local + k * d->n_bones * 12 : nullptr,
				out_status ? out_status + k : nullptr);
	}
	if (ik) {
		harvest(ik);
		delete ik;
	}
	if (counters) {
		counters[0] = steps;
		counters[1] = swing_calls;
		counters[2] = swing_rect;
	}
}

} // namespace

extern "C" {

// flags bit0: rebuild the object graph for every pose (exactly "fresh _bone_list_changed + one frame");
// otherwise one graph per worker thread is reused (the reference's long-lived node), which is what the
// CPU baseline times.  counters (nullable): [bone_steps, swing_calls, swing_rectified].
int orc_solve_batch(const mbik_rig_desc *d, size_t n_poses, const float *targets, const float *start_pose,
		float *out_pose, float *out_local, uint32_t *out_status, int iterations, int n_threads, unsigned flags, long long *counters) {
	if (!d || !targets || !out_pose) {
		return -1;
	}
	bool rebuild_each = (flags & 1u) != 0;
	if (n_threads <= 1 || n_poses < 2) {
		solve_range(d, 0, n_poses, targets, start_pose, out_pose, out_local, out_status, iterations, rebuild_each, counters);
		return 0;
	}
	std::vector<std::thread> th;
	std::vector<long long> cnt((size_t)n_threads * 3, 0);
	for (int t = 0; t < n_threads; t++) {
		size_t b = n_poses * (size_t)t / (size_t)n_threads, e = n_poses * (size_t)(t + 1) / (size_t)n_threads;
		th.emplace_back(solve_range, d, b, e, targets, start_pose, out_pose, out_local, out_status, iterations, rebuild_each, &cnt[(size_t)t * 3]);
	}
	for (auto &t : th) {
		t.join();
	}
	if (counters) {
		counters[0] = counters[1] = counters[2] = 0;
		for (int t = 0; t < n_threads; t++) {
			for (int j = 0; j < 3; j++) {
				counters[j] += cnt[(size_t)t * 3 + j];
			}
		}
	}
	return 0;
}

// Setup facts of the rebuilt rig, for cross-checking the product's flattener.
// bone_order[n_solved]; returns n_solved (or -1).  All output pointers nullable.
int orc_rig_facts(const mbik_rig_desc *d, int32_t *bone_order, int32_t capacity, int32_t *n_segments,
		float *dir_basis /*[n_solved][9]*/, float *twist_basis /*[n_solved][9]*/) {
	ManyBoneIK3D ik;
	configure(ik, d);
	ik._bone_list_changed();
	int n = (int)ik.bone_list.size();
	if (bone_order) {
		for (int i = 0; i < n && i < capacity; i++) {
			bone_order[i] = ik.bone_list[i]->get_bone_id();
		}
	}
	if (n_segments) {
		int cnt = 0;
		std::vector<Ref<IKBoneSegment3D>> stack(ik.segmented_skeletons.begin(), ik.segmented_skeletons.end());
		while (!stack.empty()) {
			Ref<IKBoneSegment3D> s = stack.back();
			stack.pop_back();
			cnt++;
			for (auto &c : s->child_segments) {
				stack.push_back(c);
			}
		}
		*n_segments = cnt;
	}
	for (int i = 0; i < n && i < capacity; i++) {
		Transform3D dt = ik.bone_list[i]->bone_direction_transform->get_transform();
		Transform3D tt = ik.bone_list[i]->constraint_twist_transform->get_transform();
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
	return n;
}

// heading weights of the segment owning bone_list[step]; returns count
int orc_step_weights(const mbik_rig_desc *d, int32_t step, double *out, int32_t capacity) {
	ManyBoneIK3D ik;
	configure(ik, d);
	ik._bone_list_changed();
	if (step < 0 || step >= (int)ik.bone_list.size()) {
		return -1;
	}
	Ref<IKBone3D> bone = ik.bone_list[step];
	std::vector<Ref<IKBoneSegment3D>> stack(ik.segmented_skeletons.begin(), ik.segmented_skeletons.end());
	while (!stack.empty()) {
		Ref<IKBoneSegment3D> s = stack.back();
		stack.pop_back();
		for (auto &b : s->bones) {
			if (b == bone) {
				int n = (int)s->heading_weights.size();
				for (int i = 0; i < n && i < capacity; i++) {
					out[i] = s->heading_weights[i];
				}
				return n;
			}
		}
		for (auto &c : s->child_segments) {
			stack.push_back(c);
		}
	}
	return -1;
}

// cone geometry after setup, constraint-row order: cp[3], tan1[3], tan2[3] per cone; returns cone count
int orc_cone_geometry(const mbik_rig_desc *d, float *out, int32_t capacity_cones) {
	ManyBoneIK3D ik;
	configure(ik, d);
	ik._bone_list_changed();
	int k = 0;
	for (int ci = 0; ci < d->n_constraints; ci++) {
		Ref<IKBone3D> bone;
		for (auto &b : ik.bone_list) {
			if (b->get_bone_id() == d->constraints[ci].bone) {
				bone = b;
				break;
			}
		}
		for (int j = 0; j < d->constraints[ci].n_cones; j++, k++) {
			if (k >= capacity_cones) {
				continue;
			}
			float *o = out + k * 9;
			if (!bone || j >= (int)bone->constraint->open_cones.size()) {
				for (int q = 0; q < 9; q++) {
					o[q] = 0;
				}
				continue;
			}
			Ref<IKLimitCone3D> c = bone->constraint->open_cones[j];
			o[0] = c->control_point.x; o[1] = c->control_point.y; o[2] = c->control_point.z;
			o[3] = c->tangent_circle_center_next_1.x; o[4] = c->tangent_circle_center_next_1.y; o[5] = c->tangent_circle_center_next_1.z;
			o[6] = c->tangent_circle_center_next_2.x; o[7] = c->tangent_circle_center_next_2.y; o[8] = c->tangent_circle_center_next_2.z;
		}
	}
	return k;
}

// ---- stage-level entry points for golden vectors ------------------------------------------------
// QCP::weighted_superpose: moved/target [n][3], weight [n]; out quat xyzw + translation xyz
void orc_qcp_weighted_superpose(const float *moved, const float *target, const double *weight, int n, int translate, float *out7) {
	PackedVector3Array m(n), t(n);
	std::vector<double> w(weight, weight + n);
	for (int i = 0; i < n; i++) {
		m[i] = Vector3(moved[3 * i], moved[3 * i + 1], moved[3 * i + 2]);
		t[i] = Vector3(target[3 * i], target[3 * i + 1], target[3 * i + 2]);
	}
	QCP qcp(1e-6);
	Quaternion q = qcp.weighted_superpose(m, t, w, translate != 0);
	Vector3 tr = qcp.get_translation();
	out7[0] = q.x; out7[1] = q.y; out7[2] = q.z; out7[3] = q.w;
	out7[4] = tr.x; out7[5] = tr.y; out7[6] = tr.z;
}

// IKKusudama3D::get_local_point_in_limits on a kusudama built like _bone_list_changed builds it
// (cones [n][4] = centre xyz + radius); out: point xyz, in_bounds[0]
void orc_kusudama_point_in_limits(const float *cones, int n_cones, const float *point, float *out4) {
	Ref<IKKusudama3D> k(new IKKusudama3D());
	for (int i = 0; i < n_cones; i++) {
		Ref<IKLimitCone3D> c(new IKLimitCone3D());
		c->set_attached_to(k);
		c->set_radius(std::max(1.0e-38, (double)cones[4 * i + 3]));
		c->set_control_point(Vector3(cones[4 * i], cones[4 * i + 1], cones[4 * i + 2]).normalized());
		k->add_open_cone(c);
	}
	std::vector<double> bounds(2, 0.0);
	Vector3 r = k->get_local_point_in_limits(Vector3(point[0], point[1], point[2]), &bounds);
	out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = (float)bounds[0];
}

// IKBoneSegment3D::clamp_to_cos_half_angle
void orc_clamp_to_cos_half_angle(const float *q4, double cos_half, float *out4) {
	Quaternion q = IKBoneSegment3D::clamp_to_cos_half_angle(Quaternion(q4[0], q4[1], q4[2], q4[3]), cos_half);
	out4[0] = q.x; out4[1] = q.y; out4[2] = q.z; out4[3] = q.w;
}

// IKKusudama3D::get_swing_twist about +Y: out swing xyzw, twist xyzw
void orc_swing_twist_y(const float *q4, float *out8) {
	Quaternion s, t;
	IKKusudama3D::get_swing_twist(Quaternion(q4[0], q4[1], q4[2], q4[3]), Vector3(0, 1, 0), s, t);
	out8[0] = s.x; out8[1] = s.y; out8[2] = s.z; out8[3] = s.w;
	out8[4] = t.x; out8[5] = t.y; out8[6] = t.z; out8[7] = t.w;
}

// Engine-math shim probes (oracle/godot_math.h): the Godot core/math functions the solve path calls, exposed so that
// tests can check their mathematical properties independently (the engine source is not in the reference tree).
//   op 0: Basis(Quaternion in[0..3])                         -> out[0..8]
//   op 1: Basis(in[0..8]).get_quaternion()                   -> out[0..3]
//   op 2: Basis(in[0..8]).get_rotation_quaternion()          -> out[0..3]
//   op 3: Basis(in[0..8]).orthonormalized()                  -> out[0..8]
//   op 4: Basis(in[0..8]).inverse()                          -> out[0..8]
//   op 5: Quaternion(Vector3 in[0..2], Vector3 in[3..5])     -> out[0..3]   (shortest arc)
//   op 6: Quaternion(in[0..3]).xform(Vector3 in[4..6])       -> out[0..2]
//   op 7: Basis(in[0..8]).slerp(Basis(in[9..17]), in[18])    -> out[0..8]
//   op 8: Quaternion(Vector3 axis in[0..2], angle in[3])     -> out[0..3]
//   op 9: Transform3D(in[0..11]).affine_inverse()            -> out[0..11]
//   op 10: Basis(in[0..8]) * Basis(in[9..17])                -> out[0..8]
//   op 11: Basis(in[0..8]).get_scale()                       -> out[0..2]
//   op 12: Basis(Quaternion(in[0..3]), scale in[4..6])       -> out[0..8]   (set_quaternion_scale)
// What Skeleton3D::get_bone_pose() returns once out10 = (position, rotation quaternion, scale) is in the skeleton:
// Transform3D(Basis(rotation, scale), position) for n records (engine update_pose_cache; see oracle/godot_shim Skeleton3D).
void orc_recompose_pose(size_t n, const float *out10, float *local12) {
	for (size_t i = 0; i < n; i++) {
		const float *o = out10 + 10 * i;
		Basis b(Quaternion(o[3], o[4], o[5], o[6]), Vector3(o[7], o[8], o[9]));
		float *l = local12 + 12 * i;
		for (int r = 0; r < 3; r++) {
			for (int c = 0; c < 3; c++) {
				l[r * 3 + c] = b.rows[r][c];
			}
		}
		l[9] = o[0]; l[10] = o[1]; l[11] = o[2];
	}
}

int orc_math_probe(int op, const float *in, float *out) {
	auto load_b = [](const float *p) {
		Basis b;
		for (int r = 0; r < 3; r++) {
			for (int c = 0; c < 3; c++) {
				b.rows[r][c] = p[r * 3 + c];
			}
		}
		return b;
	};
	auto store_b = [](const Basis &b, float *p) {
		for (int r = 0; r < 3; r++) {
			for (int c = 0; c < 3; c++) {
				p[r * 3 + c] = b.rows[r][c];
			}
		}
	};
	auto store_q = [](const Quaternion &q, float *p) {
		p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w;
	};
	switch (op) {
		case 12: store_b(Basis(Quaternion(in[0], in[1], in[2], in[3]), Vector3(in[4], in[5], in[6])), out); return 0;
		case 0: store_b(Basis(Quaternion(in[0], in[1], in[2], in[3])), out); return 0;
		case 1: store_q(load_b(in).get_quaternion(), out); return 0;
		case 2: store_q(load_b(in).get_rotation_quaternion(), out); return 0;
		case 3: store_b(load_b(in).orthonormalized(), out); return 0;
		case 4: store_b(load_b(in).inverse(), out); return 0;
		case 5: store_q(Quaternion(Vector3(in[0], in[1], in[2]), Vector3(in[3], in[4], in[5])), out); return 0;
		case 6: {
			Vector3 v = Quaternion(in[0], in[1], in[2], in[3]).xform(Vector3(in[4], in[5], in[6]));
			out[0] = v.x; out[1] = v.y; out[2] = v.z;
			return 0;
		}
		case 7: store_b(load_b(in).slerp(load_b(in + 9), in[18]), out); return 0;
		case 8: store_q(Quaternion(Vector3(in[0], in[1], in[2]), in[3]), out); return 0;
		case 9: {
			Transform3D t = load_xform(in).affine_inverse();
			store_b(t.basis, out);
			out[9] = t.origin.x; out[10] = t.origin.y; out[11] = t.origin.z;
			return 0;
		}
		case 10: store_b(load_b(in) * load_b(in + 9), out); return 0;
		case 11: {
			Vector3 v = load_b(in).get_scale();
			out[0] = v.x; out[1] = v.y; out[2] = v.z;
			return 0;
		}
		default: return -1;
	}
}

int orc_hardware_threads(void) {
	unsigned n = std::thread::hardware_concurrency();
	return n ? (int)n : 1;
}

} // extern "C"
