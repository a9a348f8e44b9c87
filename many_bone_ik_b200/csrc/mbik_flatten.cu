// mbik_flatten.cu -- host-side flattener (see mbik_flatten.h).  Pure host code; compiled by nvcc only so
// that it shares mbik_math.cuh with the kernel.  Citations are to /root/reference.
#include "mbik_flatten.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>

namespace mbik {
namespace {

const double kPi = 3.1415926535897932384626433833;

X34 load_x34(const float *p) {
	X34 t;
	for (int i = 0; i < 9; i++) {
		t.b.m[i] = p[i];
	}
	t.o = v3(p[9], p[10], p[11]);
	return t;
}

// IKKusudama3D::get_quaternion_axis_angle (src/ik_kusudama_3d.cpp:417-427): divides by |axis|^2
Q4 quat_axis_angle_len2(V3 axis, float angle) {
	float d = vlen2(axis);
	if (d == 0) {
		return q4(0, 0, 0, 1);
	}
	float sin_angle = sinf(r_mul(angle, 0.5f));
	float cos_angle = cosf(r_mul(angle, 0.5f));
	float s = r_div(sin_angle, d);
	return q4(r_mul(axis.x, s), r_mul(axis.y, s), r_mul(axis.z, s), cos_angle);
}

// engine Basis(axis, angle) -- only used by _update_constraint through Vector3::rotated
M3 m3_axis_angle(V3 a, float angle) {
	V3 sq = v3(r_mul(a.x, a.x), r_mul(a.y, a.y), r_mul(a.z, a.z));
	float cosine = cosf(angle);
	M3 r;
	r.m[0] = r_add(sq.x, r_mul(cosine, r_sub(1.0f, sq.x)));
	r.m[4] = r_add(sq.y, r_mul(cosine, r_sub(1.0f, sq.y)));
	r.m[8] = r_add(sq.z, r_mul(cosine, r_sub(1.0f, sq.z)));
	float sine = sinf(angle);
	float t = r_sub(1.0f, cosine);
	float xyzt = r_mul(r_mul(a.x, a.y), t);
	float zyxs = r_mul(a.z, sine);
	r.m[1] = r_sub(xyzt, zyxs);
	r.m[3] = r_add(xyzt, zyxs);
	xyzt = r_mul(r_mul(a.x, a.z), t);
	zyxs = r_mul(a.y, sine);
	r.m[2] = r_add(xyzt, zyxs);
	r.m[6] = r_sub(xyzt, zyxs);
	xyzt = r_mul(r_mul(a.y, a.z), t);
	zyxs = r_mul(a.x, sine);
	r.m[5] = r_sub(xyzt, zyxs);
	r.m[7] = r_add(xyzt, zyxs);
	return r;
}

float godot_acosf(float x) { return x < -1.0f ? (float)kPi : (x > 1.0f ? 0.0f : acosf(x)); }

// ---- IKRay3D pieces used by the tangent-circle construction (src/ik_ray_3d.cpp) -------------------------
struct Ray {
	V3 p1, p2;
};
// :64-73
void ray_elongate(Ray &r, float amt) {
	V3 mid = vmuls(vadd(r.p1, r.p2), 0.5f);
	V3 h1 = vsub(r.p1, mid), h2 = vsub(r.p2, mid);
	V3 a1 = vmuls(vnorm(h1), amt), a2 = vmuls(vnorm(h2), amt);
	r.p1 = vadd(vadd(h1, a1), mid);
	r.p2 = vadd(vadd(h2, a2), mid);
}
// :75-85 + plane_intersect_test :146-166
V3 ray_intersects_plane(const Ray &r, V3 ta, V3 tb, V3 tc) {
	ta = vsub(ta, r.p1);
	tb = vsub(tb, r.p1);
	tc = vsub(tc, r.p1);
	V3 u = vsub(tb, ta), v = vsub(tc, ta);
	V3 n = vnorm(vcross(u, v));
	V3 dir = vsub(r.p2, r.p1);
	V3 w0 = vsub(v3(0, 0, 0), ta);
	float a = -vdot(n, w0);
	float b = vdot(n, dir);
	float rr = r_div(a, b);
	V3 I = vmuls(dir, rr);
	return vadd(I, r.p1);
}
// :87-94 + :112-144 with the sphere at the origin, radius 1
void ray_intersects_unit_sphere(const Ray &r, V3 &S1, V3 &S2) {
	V3 c = v3(0, 0, 0);
	V3 rp1 = vsub(r.p1, c), rp2 = vsub(r.p2, c);
	S1 = v3(0, 0, 0);
	S2 = v3(0, 0, 0);
	V3 e = vnorm(vsub(rp2, rp1));
	V3 h = vsub(v3(0, 0, 0), rp1);
	float lf = vdot(e, h);
	float radpow = r_mul(1.0f, 1.0f);
	float hdh = vlen2(h);
	float lfpow = r_mul(lf, lf);
	float s = r_add(r_sub(radpow, hdh), lfpow);
	if (!(s < 0.0f)) {
		s = r_sqrt(s);
		if (lf < s) {
			if (r_add(lf, s) >= 0) {
				s = -s;
			}
		}
		S1 = vadd(vmuls(e, r_sub(lf, s)), rp1);
		S2 = vadd(vmuls(e, r_add(lf, s)), rp1);
	}
	S1 = vadd(S1, c);
	S2 = vadd(S2, c);
}

// IKLimitCone3D::get_orthogonal (src/ik_open_cone_3d.cpp:267-283)
V3 cone_get_orthogonal(V3 p) {
	float threshold = r_mul(vlen(p), 0.6f);
	if (threshold > 0.f) {
		if (fabsf(p.x) <= threshold) {
			float inv = r_div(1.f, r_sqrt(r_add(r_mul(p.y, p.y), r_mul(p.z, p.z))));
			return v3(0.f, r_mul(inv, p.z), r_mul(-inv, p.y));
		} else if (fabsf(p.y) <= threshold) {
			float inv = r_div(1.f, r_sqrt(r_add(r_mul(p.x, p.x), r_mul(p.z, p.z))));
			return v3(r_mul(-inv, p.z), 0.f, r_mul(inv, p.x));
		}
		float inv = r_div(1.f, r_sqrt(r_add(r_mul(p.x, p.x), r_mul(p.y, p.y))));
		return v3(r_mul(inv, p.y), r_mul(-inv, p.x), 0.f);
	}
	return v3(0, 0, 0);
}

struct HostCone {
	V3 cp;
	double radius, radius_cos;
	V3 tc1 = v3(0, 0, 0), tc2 = v3(0, 0, 0);
	double tan_r = 0, tan_cos = 0;
};

// IKLimitCone3D::update_tangent_handles (src/ik_open_cone_3d.cpp:36-120)
void update_tangent_handles(HostCone &A, const HostCone &B) {
	double radA = A.radius, radB = B.radius;
	V3 a = A.cp, b = B.cp;
	V3 arc_normal = vnorm(vcross(a, b));
	double tRadius = (kPi - (radA + radB)) / 2;
	double bA = radA + tRadius, bB = radB + tRadius;

	V3 scaledAxisA = vmuls(a, (float)cos(bA));
	V3 planeDir1A = q_xform(quat_axis_angle_len2(arc_normal, (float)bA), a);
	V3 planeDir2A = q_xform(quat_axis_angle_len2(a, (float)(kPi / 2)), planeDir1A);

	V3 scaledAxisB = vmuls(b, (float)cos(bB));
	V3 planeDir1B = q_xform(quat_axis_angle_len2(arc_normal, (float)bB), b);
	V3 planeDir2B = q_xform(quat_axis_angle_len2(b, (float)(kPi / 2)), planeDir1B);

	Ray r1B{ planeDir1B, scaledAxisB }, r2B{ planeDir1B, planeDir2B };
	ray_elongate(r1B, 99);
	ray_elongate(r2B, 99);
	V3 i1 = ray_intersects_plane(r1B, scaledAxisA, planeDir1A, planeDir2A);
	V3 i2 = ray_intersects_plane(r2B, scaledAxisA, planeDir1A, planeDir2A);
	Ray ir{ i1, i2 };
	ray_elongate(ir, 99);
	V3 s1, s2;
	ray_intersects_unit_sphere(ir, s1, s2);
	A.tc1 = vnorm(s1);
	A.tc2 = vnorm(s2);
	A.tan_r = tRadius;
	A.tan_cos = cos(tRadius);
	if (f_is_zero_approx(vlen2(A.tc1))) {
		A.tc1 = vnorm(cone_get_orthogonal(A.cp));
	}
	if (f_is_zero_approx(vlen2(A.tc2))) {
		A.tc2 = vnorm(cone_get_orthogonal(vmuls(A.tc1, -1.0f)));
	}
}

// IKLimitCone3D::set_control_point (src/ik_open_cone_3d.cpp:159-166)
V3 set_control_point(V3 p) {
	if (f_is_zero_approx(vlen2(p))) {
		return v3(0, 1, 0);
	}
	return vnorm(p);
}

template <class T>
uint32_t append_section(std::vector<unsigned char> &blob, const std::vector<T> &v) {
	while (blob.size() % 16) {
		blob.push_back(0);
	}
	uint32_t off = (uint32_t)blob.size();
	if (!v.empty()) {
		const unsigned char *p = reinterpret_cast<const unsigned char *>(v.data());
		blob.insert(blob.end(), p, p + sizeof(T) * v.size());
	}
	return off;
}

} // namespace

// The limit-dependent part of the flattening: kusudama cone / tangent-circle geometry and twist frames of every constraint
// row (many_bone_ik_3d.cpp:1037-1067), written into the twist fields of `bones` ([t]; the other fields are left alone) and
// `cones`.  Everything it needs from the rig besides the rows is the topology (R.topo, R.t_of_bone, R.ik_parent) and the
// setup-time global basis of each constrained bone's parent (R.setup_parent_basis) -- which is why per-pose limit sets are
// authored by this function alone, not by re-flattening the rig (mbik_limit_sets_create: 1 / 15 of the work per set).
int author_constraints(const mbik_rig_desc *d, const FlatRig &R, std::vector<BlobBone> &bones, std::vector<BlobCone> &cones, ConstraintTables &T,
		std::string &error) {
	struct BoneConstraint {
		bool present = false;
		std::vector<HostCone> cones;
		std::vector<int> desc_cone_index;
		Q4 twist_center_rot = q4(0, 0, 0, 1);
		float twist_cos = 0;
	};
	const int nb = R.n_bones, ns = (int)R.topo.size();
	std::vector<BoneConstraint> cons(nb);
	std::vector<M3> twist_basis(nb, m3_identity());
	int n_desc_cones = 0;
	for (int ci = 0; ci < d->n_constraints; ci++) {
		n_desc_cones = std::max(n_desc_cones, d->constraints[ci].cone_offset + std::max(0, d->constraints[ci].n_cones));
	}
	T.cone_row_index.assign(n_desc_cones, -1);
	for (int ci = 0; ci < d->n_constraints; ci++) {
		const mbik_constraint_desc &cd = d->constraints[ci];
		int b = cd.bone;
		if (b < 0 || b >= nb || R.t_of_bone[b] < 0) {
			continue; // not in bone_list: the row is ignored
		}
		if (cd.n_cones < 0 || (cd.n_cones > 0 && !d->cones)) {
			error = "bad cone table";
			return MBIK_ERR_INVALID_ARG;
		}
		BoneConstraint bc;
		bc.present = true;
		bc.cones.reserve((size_t)cd.n_cones);
		bc.desc_cone_index.reserve((size_t)cd.n_cones);
		auto update_tangent_radii = [&]() { // src/ik_kusudama_3d.cpp:91-101
			for (size_t i = 0; i + 1 < bc.cones.size(); i++) {
				update_tangent_handles(bc.cones[i], bc.cones[i + 1]);
			}
		};
		for (int j = 0; j < cd.n_cones; j++) {
			const mbik_cone_desc &c = d->cones[cd.cone_offset + j];
			V3 ctr = v3(c.center[0], c.center[1], c.center[2]);
			if (f_is_zero_approx(vlen2(ctr))) {
				ctr = v3(0, 1, 0); // set_kusudama_open_cone_center (many_bone_ik_3d.cpp:586-601)
			}
			HostCone hc;
			hc.radius = std::max(1.0e-38, (double)c.radius);
			hc.radius_cos = cos(hc.radius);
			hc.cp = set_control_point(vnorm(ctr));
			bc.cones.push_back(hc);
			bc.desc_cone_index.push_back(cd.cone_offset + j);
			// add_open_cone (src/ik_kusudama_3d.cpp:160-166) updates the tangent radii here, after every cone.  The handles are
			// a pure function of the two cones' control points and radii and are all rewritten by the update that ends
			// _update_constraint below, after the control points were re-normalised: only that one is evaluated.
		}
		// set_axial_limits (src/ik_kusudama_3d.cpp:103-115)
		{
			V3 y_axis = v3(0, 1, 0), z_axis = v3(0, 0, 1);
			Q4 twist_min_rot = quat_axis_angle_len2(y_axis, cd.twist_from);
			V3 twist_min_vec = vnorm(q_xform(twist_min_rot, z_axis));
			V3 twist_center_vec = vnorm(q_xform(twist_min_rot, twist_min_vec));
			bc.twist_center_rot = q_shortest_arc(z_axis, twist_center_vec);
			bc.twist_cos = cosf(r_div(cd.twist_range, 4.0f));
		}
		// _update_constraint(twist node) (src/ik_kusudama_3d.cpp:37-89)
		{
			std::vector<V3> directions;
			if (bc.cones.size() == 1) {
				directions.push_back(bc.cones[0].cp);
			} else {
				for (size_t i = 0; i + 1 < bc.cones.size(); i++) {
					V3 a = bc.cones[i].cp, n = bc.cones[i + 1].cp;
					Q4 q = q_shortest_arc(a, n);
					V3 axis;
					if (fabsf(q.w) > r_sub(1.0f, kCmpEps)) {
						axis = v3(q.x, q.y, q.z);
					} else {
						float r = r_div(1.0f, r_sqrt(r_sub(1.0f, r_mul(q.w, q.w))));
						axis = v3(r_mul(q.x, r), r_mul(q.y, r), r_mul(q.z, r));
					}
					float full = r_mul(2.0f, godot_acosf(q.w));
					double angle = (double)full / 2.0;
					V3 half = m3_xform(m3_axis_angle(axis, (float)angle), a);
					half = vmuls(half, full);
					half = vnorm(half);
					directions.push_back(half);
				}
			}
			V3 new_y = v3(0, 0, 0);
			for (V3 dv : directions) {
				new_y = vadd(new_y, dv);
			}
			if (!directions.empty()) {
				new_y = vdivs(new_y, (float)directions.size());
				new_y = vnorm(new_y);
			}
			if (R.ik_parent[b] >= 0) { // the twist node's parent is the parent bone's aligned node (src/ik_bone_3d.cpp:52-54)
				M3 P = R.setup_parent_basis[b];
				M3 Gtw = m3_mul(P, twist_basis[b]);
				Q4 q = q_shortest_arc(vnorm(m3_col(Gtw, 1)), vnorm(m3_xform(Gtw, new_y)));
				twist_basis[b] = m3_mul(m3_mul(m3_mul(m3_inverse(P), m3_from_quat(q)), P), twist_basis[b]);
			}
			for (auto &c : bc.cones) {
				c.cp = set_control_point(vnorm(c.cp));
			}
			update_tangent_radii();
		}
		cons[b] = std::move(bc);
	}

	// ---- emit the limit fields of the per-bone constants, and the cones ----
	T.present.assign(nb, 0);
	T.cone_off.assign(nb, 0);
	T.cone_cnt.assign(nb, 0);
	cones.clear();
	for (int t = 0; t < ns; t++) {
		int b = R.topo[t];
		BlobBone &B = bones[t];
		B.twist_cos = cons[b].twist_cos;
		memcpy(B.twist_basis, twist_basis[b].m, sizeof(float) * 9);
		M3 tc = m3_from_quat(cons[b].twist_center_rot);
		memcpy(B.twist_center, tc.m, sizeof(float) * 9);
		if (cons[b].present) {
			T.present[b] = 1;
			T.cone_off[b] = (int)cones.size();
			T.cone_cnt[b] = (int)cons[b].cones.size();
			for (size_t i = 0; i < cons[b].cones.size(); i++) {
				const HostCone &c = cons[b].cones[i];
				BlobCone bc;
				memset(&bc, 0, sizeof(bc));
				V3 ncp = vnorm(c.cp);
				bc.cp[0] = c.cp.x; bc.cp[1] = c.cp.y; bc.cp[2] = c.cp.z;
				bc.ncp[0] = ncp.x; bc.ncp[1] = ncp.y; bc.ncp[2] = ncp.z;
				float rf = (float)c.radius;
				bc.sin_half_r = sinf(r_mul(rf, 0.5f));
				bc.cos_half_r = cosf(r_mul(rf, 0.5f));
				bc.radius_cos = c.radius_cos;
				if (i + 1 < cons[b].cones.size()) {
					const HostCone &n = cons[b].cones[i + 1];
					bc.tan_cos = c.tan_cos;
					bc.tc1[0] = c.tc1.x; bc.tc1[1] = c.tc1.y; bc.tc1[2] = c.tc1.z;
					bc.tc2[0] = c.tc2.x; bc.tc2[1] = c.tc2.y; bc.tc2[2] = c.tc2.z;
					float tf = (float)c.tan_r;
					bc.sin_half_t = sinf(r_mul(tf, 0.5f));
					bc.cos_half_t = cosf(r_mul(tf, 0.5f));
					V3 a = vcross(c.cp, n.cp);
					V3 e1 = vnorm(vcross(c.cp, c.tc1)), e2 = vnorm(vcross(c.tc1, n.cp));
					V3 e3 = vnorm(vcross(c.tc2, c.cp)), e4 = vnorm(vcross(n.cp, c.tc2));
					bc.c1xc2[0] = a.x; bc.c1xc2[1] = a.y; bc.c1xc2[2] = a.z;
					bc.c1xt1[0] = e1.x; bc.c1xt1[1] = e1.y; bc.c1xt1[2] = e1.z;
					bc.t1xc2[0] = e2.x; bc.t1xc2[1] = e2.y; bc.t1xc2[2] = e2.z;
					bc.t2xc1[0] = e3.x; bc.t2xc1[1] = e3.y; bc.t2xc1[2] = e3.z;
					bc.c2xt2[0] = e4.x; bc.c2xt2[1] = e4.y; bc.c2xt2[2] = e4.z;
				}
				T.cone_row_index[cons[b].desc_cone_index[i]] = (int)cones.size();
				cones.push_back(bc);
			}
		}
	}
	return MBIK_OK;
}

int flatten_rig(const mbik_rig_desc *d, FlatRig &R) {
	if (!d || d->n_bones <= 0 || !d->parent || !d->rest_local) {
		R.error = "rig description needs n_bones > 0, parent[] and rest_local[]";
		return MBIK_ERR_INVALID_ARG;
	}
	if (d->n_pins < 0 || d->n_constraints < 0 || d->n_bone_damp < 0) {
		R.error = "negative table count";
		return MBIK_ERR_INVALID_ARG;
	}
	for (int ci = 0; ci < d->n_constraints && d->constraints; ci++) {
		const mbik_constraint_desc &cd = d->constraints[ci];
		if (cd.n_cones < 0 || cd.cone_offset < 0 || (cd.n_cones > 0 && !d->cones)) {
			R.error = "constraint row " + std::to_string(ci) + ": negative n_cones / cone_offset, or cones is NULL";
			return MBIK_ERR_INVALID_ARG;
		}
	}
	if ((d->n_pins > 0 && !d->pins) || (d->n_constraints > 0 && !d->constraints) || (d->n_bone_damp > 0 && !d->bone_damp)) {
		R.error = "null table with non-zero count";
		return MBIK_ERR_INVALID_ARG;
	}
	const int nb = d->n_bones;
	R.n_bones = nb;
	R.parent.assign(d->parent, d->parent + nb);
	R.rest_local.resize(nb);
	for (int b = 0; b < nb; b++) {
		if (R.parent[b] >= nb || R.parent[b] == b) {
			R.error = "bad parent index";
			return MBIK_ERR_INVALID_ARG;
		}
		R.rest_local[b] = load_x34(d->rest_local + 12 * b);
	}
	R.pins.assign(d->pins, d->pins + d->n_pins);
	for (auto &p : R.pins) {
		// IKEffector3D::set_motion_propagation_factor clamps to [0,1] (src/ik_effector_3d.cpp:173-175)
		double v = p.motion_propagation_factor;
		p.motion_propagation_factor = (float)(v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v));
	}
	R.iterations = d->iterations_per_frame;

	// Skeleton3D queries: children ascending, parentless ascending (engine process order)
	std::vector<std::vector<int>> sk_children(nb);
	std::vector<int> sk_roots;
	for (int b = 0; b < nb; b++) {
		if (R.parent[b] >= 0) {
			sk_children[R.parent[b]].push_back(b);
		} else {
			sk_roots.push_back(b);
		}
	}
	{ // reject cycles
		for (int b = 0; b < nb; b++) {
			int hops = 0;
			for (int p = R.parent[b]; p >= 0; p = R.parent[p]) {
				if (++hops > nb) {
					R.error = "parent table has a cycle";
					return MBIK_ERR_INVALID_ARG;
				}
			}
		}
	}
	if (sk_roots.empty()) {
		R.error = "skeleton has no parentless bone";
		return MBIK_ERR_INVALID_ARG;
	}

	// first pin row naming a bone wins (IKBone3D ctor, src/ik_bone_3d.cpp:209-222)
	std::vector<int> pin_of_bone(nb, -1);
	for (int i = 0; i < (int)R.pins.size(); i++) {
		int b = R.pins[i].bone;
		if (b >= 0 && b < nb && pin_of_bone[b] < 0) {
			pin_of_bone[b] = i;
		}
	}

	// ---- IK bone graph: every IKBone3D the reference creates, including those of dropped segments ----
	std::vector<char> ik_exists(nb, 0);
	std::vector<int> ik_parent(nb, -1);
	std::vector<std::vector<int>> ik_children(nb); // creation order (IKBone3D::set_parent push_back, :46-55)
	std::vector<float> ik_default_damp(nb, 0.f);

	// ---- segment generation (src/ik_bone_segment_3d.cpp:247-264, :352-427) ----
	std::function<int(int, int)> build_segment = [&](int root_bone, int parent_seg) -> int {
		int si = (int)R.segments.size();
		R.segments.emplace_back();
		R.segments[si].root_bone = root_bone;
		R.segments[si].parent_seg = parent_seg;
		ik_exists[root_bone] = 1;
		ik_default_damp[root_bone] = (float)kPi; // segment roots are created with Math_PI (:252)
		if (parent_seg >= 0) {
			int ptip = R.segments[parent_seg].tip_bone;
			ik_parent[root_bone] = ptip;
			ik_children[ptip].push_back(root_bone);
		}
		int current_tip = root_bone;
		while (true) { // generate_default_segments with p_tip_bone == -1
			const std::vector<int> &children = sk_children[current_tip];
			bool pinned = pin_of_bone[current_tip] >= 0;
			if (children.empty() || children.size() > 1 || pinned) {
				// _process_children
				R.segments[si].tip_bone = current_tip;
				for (int child : children) {
					int ci = build_segment(child, si);
					if (R.segments[ci].pinned_descendants) {
						R.segments[si].pinned_descendants = true;
						R.segments[si].child_segs.push_back(ci);
					} else {
						R.segments[ci].kept = false;
					}
				}
				break;
			}
			// _create_next_bone
			int next = children[0];
			ik_exists[next] = 1;
			ik_default_damp[next] = d->default_damp;
			ik_parent[next] = current_tip;
			ik_children[current_tip].push_back(next);
			current_tip = next;
		}
		// _finalize_segment
		FlatSegment &S = R.segments[si];
		S.tip_bone = current_tip;
		if (pin_of_bone[current_tip] >= 0) {
			S.pinned_descendants = true;
		}
		for (int b = current_tip;; b = ik_parent[b]) {
			S.bones.push_back(b);
			if (b == root_bone) {
				break;
			}
		}
		return si;
	};

	std::vector<char> node_has_parent(nb, 1); // IKNode3D parent of the aligned node
	for (size_t ri = 0; ri < sk_roots.size(); ri++) {
		int si = build_segment(sk_roots[ri], -1);
		R.segments[si].stabilize = d->stabilization_passes; // only root segments receive it (many_bone_ik_3d.cpp:1021)
		R.root_segments.push_back(si);
		// `ik_origin.instantiate()` per root frees the previous origin node, un-parenting the earlier roots
		node_has_parent[sk_roots[ri]] = (ri + 1 == sk_roots.size()) ? 1 : 0;
	}
	// dropped segments: mark whole subtrees as not kept
	{
		std::function<void(int, bool)> mark = [&](int si, bool kept) {
			R.segments[si].kept = kept;
			(void)kept;
		};
		// a dropped segment's descendants were never linked into a kept parent's child_segs; their `kept`
		// flag stays true from construction, so recompute reachability from the roots instead.
		for (auto &s : R.segments) {
			s.kept = false;
		}
		std::function<void(int)> reach = [&](int si) {
			R.segments[si].kept = true;
			for (int c : R.segments[si].child_segs) {
				reach(c);
			}
		};
		for (int si : R.root_segments) {
			reach(si);
		}
		(void)mark;
	}

	// ---- bone_list (create_bone_list recursive, :56-72) ----
	R.seg_of_bone.assign(nb, -1);
	std::function<void(int)> list_bones = [&](int si) {
		for (int c : R.segments[si].child_segs) {
			list_bones(c);
		}
		for (int b : R.segments[si].bones) {
			R.bone_order.push_back(b);
			R.seg_of_bone[b] = si;
		}
	};
	for (int si : R.root_segments) {
		list_bones(si);
	}
	const int ns = (int)R.bone_order.size();
	R.n_kept_segments = 0;
	for (auto &s : R.segments) {
		R.n_kept_segments += s.kept ? 1 : 0;
	}

	// ---- effector lists (update_pinned_list, :74-88) and heading weights (:281-343) ----
	std::function<void(int)> pinned_list = [&](int si) {
		FlatSegment &S = R.segments[si];
		for (int c : S.child_segs) {
			pinned_list(c);
		}
		bool pinned = pin_of_bone[S.tip_bone] >= 0;
		if (pinned) {
			S.effectors.push_back(S.tip_bone);
		}
		double mpf = pinned ? (double)R.pins[pin_of_bone[S.tip_bone]].motion_propagation_factor : 1.0;
		if (mpf > 0.0) {
			for (int c : S.child_segs) {
				S.effectors.insert(S.effectors.end(), R.segments[c].effectors.begin(), R.segments[c].effectors.end());
			}
		}
	};
	for (int si : R.root_segments) {
		pinned_list(si);
	}
	std::function<void(int, std::vector<double> &, double)> penalty = [&](int si, std::vector<double> &out, double falloff) {
		if (falloff <= 0.0) {
			return;
		}
		double current_falloff = 1.0;
		FlatSegment &S = R.segments[si];
		if (pin_of_bone[S.tip_bone] >= 0) {
			const mbik_pin_desc &pin = R.pins[pin_of_bone[S.tip_bone]];
			double weight = pin.weight;
			out.push_back(weight * falloff);
			float pm = std::max(std::max(pin.direction_priorities[0], pin.direction_priorities[1]), pin.direction_priorities[2]);
			double max_pin_weight = pm;
			max_pin_weight = max_pin_weight == 0.0 ? 1.0 : max_pin_weight;
			for (int i = 0; i < 3; ++i) {
				double priority = pin.direction_priorities[i];
				if (priority > 0.0) {
					double sub = weight * (priority / max_pin_weight) * falloff;
					out.push_back(sub);
					out.push_back(sub);
				}
			}
			current_falloff = pin.motion_propagation_factor;
		}
		for (int c : S.child_segs) {
			penalty(c, out, falloff * current_falloff);
		}
	};
	for (auto &S : R.segments) {
		if (!S.kept) {
			continue;
		}
		int si = (int)(&S - &R.segments[0]);
		penalty(si, S.weights, 1.0);
		int h = 0;
		for (int e : S.effectors) {
			const mbik_pin_desc &pin = R.pins[pin_of_bone[e]];
			h += 1;
			for (int a = 0; a < 3; a++) {
				h += pin.direction_priorities[a] > 0.0 ? 2 : 0;
			}
		}
		if (h != (int)S.weights.size()) {
			R.error = "effector list and heading weights disagree (motion_propagation_factor underflow?)";
			return MBIK_ERR_UNSUPPORTED;
		}
		R.max_headings = std::max(R.max_headings, h);
	}

	// ---- topological numbering of solved bones ----
	R.t_of_bone.assign(nb, -1);
	{
		std::vector<char> solved(nb, 0);
		for (int b : R.bone_order) {
			solved[b] = 1;
		}
		std::function<void(int)> visit = [&](int b) {
			if (!solved[b]) {
				return;
			}
			R.t_of_bone[b] = (int)R.topo.size();
			R.topo.push_back(b);
			for (int c : ik_children[b]) {
				visit(c);
			}
		};
		for (int r : sk_roots) {
			visit(r);
		}
		if ((int)R.topo.size() != ns) {
			R.error = "internal: topological order incomplete";
			return MBIK_ERR_UNSUPPORTED;
		}
		for (int b = 0; b < nb; b++) {
			if (!solved[b]) {
				R.pass.push_back(BlobPass{ b });
			}
		}
	}

	// ---- setup-time poses: _update_ik_bones_transform (many_bone_ik_3d.cpp:91-102) seeds bone_list members
	//      from the skeleton; IK bones of dropped segments keep an identity local pose ----
	std::vector<X34> ik_local(nb, x_identity());
	for (int b : R.bone_order) {
		ik_local[b] = R.rest_local[b];
	}
	std::function<X34(int)> ik_global = [&](int b) -> X34 {
		if (ik_parent[b] >= 0) {
			return x_mul(ik_global(ik_parent[b]), ik_local[b]);
		}
		if (node_has_parent[b]) {
			return x_mul(x_identity(), ik_local[b]); // child of the identity ik_origin node
		}
		return ik_local[b];
	};

	// ---- bone direction frames (IKBone3D::update_default_bone_direction_transform, src/ik_bone_3d.cpp:57-93),
	//      evaluated in bone_list order: children before parents ----
	std::vector<M3> dir_basis(nb, m3_identity());
	auto dir_global_basis = [&](int b) { return m3_mul(ik_global(b).b, dir_basis[b]); };
	for (int b : R.bone_order) {
		V3 centroid = v3(0, 0, 0);
		int child_count = 0;
		for (int c : ik_children[b]) {
			centroid = vadd(centroid, ik_global(c).o);
			child_count++;
		}
		if (child_count > 0) {
			centroid = vdivs(centroid, (float)child_count);
		} else {
			// reference: loops over Skeleton3D children (none, or IK children would exist) then divides by 0
			for (int c : sk_children[b]) {
				X34 g = R.rest_local[c];
				for (int p = R.parent[c]; p >= 0; p = R.parent[p]) {
					g = x_mul(R.rest_local[p], g); // not reached in practice; kept for completeness
				}
				centroid = vadd(centroid, g.o);
			}
			centroid = vdivs(centroid, (float)sk_children[b].size());
		}
		X34 G = ik_global(b);
		centroid = vsub(centroid, G.o);
		if (f_is_zero_approx(vlen2(centroid))) {
			int src = ik_parent[b] >= 0 ? ik_parent[b] : b;
			centroid = m3_col(dir_global_basis(src), 1);
		}
		if (!f_is_zero_approx(vlen2(centroid)) && (!ik_children[b].empty() || !sk_children[b].empty())) {
			centroid = vnorm(centroid);
			V3 bone_direction = vnorm(m3_col(dir_global_basis(b), 1));
			M3 rot = m3_from_quat(q_shortest_arc(centroid, bone_direction));
			// rotate_local_with_global on the direction node, whose parent is the aligned node (always present)
			dir_basis[b] = m3_mul(m3_mul(m3_mul(m3_inverse(G.b), rot), G.b), dir_basis[b]);
		}
	}

	// ---- constraints (many_bone_ik_3d.cpp:1037-1067): per-bone constants, cones ----
	R.ik_parent = ik_parent;
	R.setup_parent_basis.assign(nb, m3_identity());
	for (int ci = 0; ci < d->n_constraints; ci++) {
		const int b = d->constraints[ci].bone;
		if (b >= 0 && b < nb && R.t_of_bone[b] >= 0 && ik_parent[b] >= 0) {
			R.setup_parent_basis[b] = ik_global(ik_parent[b]).b;
		}
	}
	R.bones.resize(ns);
	for (int t = 0; t < ns; t++) {
		int b = R.topo[t];
		BlobBone &B = R.bones[t];
		memset(&B, 0, sizeof(B));
		B.skel_bone = b;
		B.parent = ik_parent[b] >= 0 ? R.t_of_bone[ik_parent[b]] : -1;
		B.flags = node_has_parent[b] ? STEP_NODE_PARENT : 0;
		M3 ident = m3_identity();
		memcpy(B.dir_basis, dir_basis[b].m, sizeof(float) * 9);
		memcpy(B.orient_basis, ident.m, sizeof(float) * 9);
	}
	ConstraintTables cons_tables;
	{
		int rc = author_constraints(d, R, R.bones, R.cones, cons_tables, R.error);
		if (rc != MBIK_OK) {
			return rc;
		}
	}
	R.cone_row_index = cons_tables.cone_row_index;
	const std::vector<char> &cons_present = cons_tables.present;
	const std::vector<int> &cone_off = cons_tables.cone_off, &cone_cnt = cons_tables.cone_cnt;

	// ---- per-segment effector entries ----
	std::vector<int> seg_eff_off(R.segments.size(), 0);
	std::vector<char> is_effector(nb, 0);
	for (size_t si = 0; si < R.segments.size(); si++) {
		FlatSegment &S = R.segments[si];
		if (!S.kept) {
			continue;
		}
		seg_eff_off[si] = (int)R.effs.size();
		int h = 0;
		for (int e : S.effectors) {
			const mbik_pin_desc &pin = R.pins[pin_of_bone[e]];
			BlobEff E;
			memset(&E, 0, sizeof(E));
			E.bone = R.t_of_bone[e];
			E.pin = pin_of_bone[e];
			E.h_off = h;
			E.w_origin = S.weights[h++];
			E.w_origin_f = (float)E.w_origin;
			E.n_headings = 1;
			for (int a = 0; a < 3; a++) {
				E.prio[a] = pin.direction_priorities[a];
				if (pin.direction_priorities[a] > 0.0) {
					E.w_axis[a] = S.weights[h];
					E.w_axis_f[a] = (float)E.w_axis[a];
					h += 2;
					E.n_headings += 2;
				}
			}
			is_effector[e] = 1;
			R.effs.push_back(E);
		}
	}
	for (int b = 0; b < nb; b++) {
		R.n_effectors += is_effector[b];
	}

	// ---- steps in bone_list order, with the FK refresh list of each ----
	double flops_iter = 0;
	int n_constrained = 0;
	for (int b : R.bone_order) {
		FlatSegment &S = R.segments[R.seg_of_bone[b]];
		BlobStep st;
		memset(&st, 0, sizeof(st));
		st.bone = R.t_of_bone[b];
		st.parent = ik_parent[b] >= 0 ? R.t_of_bone[ik_parent[b]] : -1;
		bool translate = S.parent_seg < 0;
		st.flags = (translate ? STEP_TRANSLATE : 0) | (node_has_parent[b] ? STEP_NODE_PARENT : 0) | (ik_parent[b] >= 0 ? STEP_IK_PARENT : 0) |
				(cons_present[b] ? (STEP_SWING | STEP_TWIST) : 0) | (b == S.root_bone ? STEP_SEG_ROOT : 0) | (S.stabilize > 0 ? STEP_STABILIZE : 0);
		{
			const BlobBone &BB = R.bones[R.t_of_bone[b]];
			const M3 ident = m3_identity();
			bool plain = true;
			for (int k = 0; k < 9; k++) {
				plain = plain && BB.orient_basis[k] == ident.m[k] && fabsf(BB.dir_basis[k]) <= 2.0f; // NaN fails the <=
			}
			if (plain) {
				st.flags |= STEP_PLAIN_FRAMES;
			}
		}
		st.eff_off = seg_eff_off[R.seg_of_bone[b]];
		st.eff_cnt = (int)S.effectors.size();
		st.cone_off = cone_off[b];
		st.cone_cnt = cone_cnt[b];
		st.n_headings = (int)S.weights.size();
		// damp selection, src/ik_bone_segment_3d.cpp:217-237
		float default_damp = translate ? (float)kPi : d->default_damp;
		float damp = default_damp;
		if (b < d->n_bone_damp) {
			damp = translate ? (float)kPi : d->bone_damp[b];
		}
		if (default_damp < damp) {
			damp = default_damp;
		}
		st.cos_half_damp = cos((double)damp / 2.0);
		// Walk below `bone`: union of the paths bone -> effector bones (exclusive of `bone`) in depth-first
		// preorder (= t order).  Effector bones are met in the order of the reference's effector list, so the
		// kernel can build headings on the fly; globals are kept only at branch points (a small stack).
		std::vector<char> mark(nb, 0);
		for (int e : S.effectors) {
			for (int x = e; x != b && x >= 0; x = ik_parent[x]) {
				mark[x] = 1;
			}
		}
		std::vector<int> marked_children(nb, 0);
		for (int x = 0; x < nb; x++) {
			if (mark[x]) {
				marked_children[ik_parent[x]]++;
			}
		}
		st.fk_off = (int)R.fk.size();
		{
			std::vector<int> slot_of(nb, -1); // stack slot holding the global of a branch point
			int depth = 0;                    // slots in use
			int prev = b;                     // bone whose global is in the running register
			int next_eff = 0;
			if (!S.effectors.empty() && S.effectors[0] == b) {
				st.flags |= STEP_SELF_EFF;
				next_eff = 1;
			}
			std::vector<int> open_branch; // branch points currently on the stack, innermost last
			// the start bone is a branch point too when it has >= 2 marked children: it lives in slot 0
			if (marked_children[b] >= 2) {
				slot_of[b] = depth++; // pushed by the kernel before the walk
				open_branch.push_back(b);
				st.flags |= STEP_PUSH_SELF;
			}
			for (int t = 0; t < ns; t++) {
				int x = R.topo[t];
				if (!mark[x]) {
					continue;
				}
				int par = ik_parent[x];
				BlobFk op;
				memset(&op, 0, sizeof(op));
				op.child = (int16_t)t;
				op.src_slot = -1;
				op.push_slot = -1;
				op.eff = -1;
				if (par != prev) {
					// returning to a branch point: everything pushed below it is dead
					while (!open_branch.empty() && open_branch.back() != par) {
						slot_of[open_branch.back()] = -1;
						open_branch.pop_back();
						depth--;
					}
					if (open_branch.empty() || slot_of[par] < 0) {
						R.error = "internal: walk returned to a bone that is not on the branch stack";
						return MBIK_ERR_UNSUPPORTED;
					}
					op.src_slot = (int8_t)slot_of[par];
				}
				if (marked_children[x] >= 2) {
					slot_of[x] = depth++;
					open_branch.push_back(x);
					op.push_slot = (int8_t)slot_of[x];
					if (depth > 100) {
						R.error = "walk stack too deep";
						return MBIK_ERR_UNSUPPORTED;
					}
				}
				R.max_stack = std::max(R.max_stack, depth);
				if (next_eff < (int)S.effectors.size() && S.effectors[next_eff] == x) {
					op.eff = (int16_t)next_eff++;
				}
				prev = x;
				R.fk.push_back(op);
			}
			R.max_stack = std::max(R.max_stack, depth);
			if (next_eff != (int)S.effectors.size()) {
				R.error = "internal: effector list is not in depth-first order";
				return MBIK_ERR_UNSUPPORTED;
			}
		}
		st.fk_cnt = (int)R.fk.size() - st.fk_off;
		// segment chain: the globals of the parents of this segment's bones, rebuilt at the segment's first step
		{
			int seg_len = (int)S.bones.size();
			st.seg_len = seg_len;
			R.max_seg_len = std::max(R.max_seg_len, seg_len);
			// position of b in the segment counted from the root end
			int pos_from_tip = (int)(std::find(S.bones.begin(), S.bones.end(), b) - S.bones.begin());
			st.pslot = seg_len - 1 - pos_from_tip;
			if (b == S.tip_bone) {
				st.flags |= STEP_SEG_FIRST;
				std::vector<int> up;
				for (int x = ik_parent[S.tip_bone]; x >= 0; x = ik_parent[x]) {
					up.push_back(x);
				}
				st.chain_off = (int)R.chain.size();
				for (auto it2 = up.rbegin(); it2 != up.rend(); ++it2) {
					R.chain.push_back((int16_t)R.t_of_bone[*it2]);
				}
				st.chain_cnt = (int)up.size();
			}
		}
		R.steps.push_back(st);

		// algorithmic flop floor, SURVEY.md section 8(d)
		double A = 0;
		for (int e : S.effectors) {
			int a = 0;
			for (int k = 0; k < 3; k++) {
				a += R.pins[pin_of_bone[e]].direction_priorities[k] > 0.0 ? 1 : 0;
			}
			A += 16 + 39 * a;
		}
		double H = st.n_headings;
		double fs = A + 34 * H + (translate ? 20 * H + 8 : 0) + 82 + 40 + 150 + (translate ? 3 : 0) + 63.0 * st.eff_cnt;
		if ((st.flags & STEP_IK_PARENT) && cons_present[b]) {
			fs += 82 + 14.0 * st.cone_cnt + 400;
			n_constrained++;
		}
		flops_iter += fs;
	}
	flops_iter += 63.0 * (ns + 3.0 * n_constrained + R.n_effectors);
	R.flops_per_solve = flops_iter * std::max(0, R.iterations);

	if (ns > 16383) {
		R.error = "too many solved bones";
		return MBIK_ERR_UNSUPPORTED;
	}

	// ---- segment-parallel schedule (BlobSpan): phase = height of the segment above its deepest leaf segment, so every
	//      child segment sits in an earlier phase than its parent (the post-order of segment_solver, :210-225); within a
	//      phase the segments go to the `roles` warps of a pose group longest-first onto the least loaded warp ----
	{
		const int nseg = (int)R.segments.size();
		std::vector<int> s0(nseg, -1), s1(nseg, -1), height(nseg, 0);
		std::vector<double> cost(nseg, 0.0);
		for (int s = 0; s < (int)R.steps.size(); s++) {
			const int si = R.seg_of_bone[R.bone_order[s]];
			if (s0[si] < 0) {
				s0[si] = s;
			}
			s1[si] = s + 1;
			const BlobStep &st = R.steps[s];
			const double passes = (st.flags & STEP_TRANSLATE) ? 2.0 : 1.0;
			const bool snaps = (st.flags & STEP_IK_PARENT) && (st.flags & (STEP_SWING | STEP_TWIST));
			cost[si] += (snaps ? 3000.0 : 1600.0) + passes * (450.0 * st.eff_cnt + 70.0 * st.fk_cnt) + 14.0 * st.cone_cnt;
		}
		bool contiguous = true;
		for (int si = 0; si < nseg; si++) {
			if (R.segments[si].kept && s0[si] >= 0 && s1[si] - s0[si] != (int)R.segments[si].bones.size()) {
				contiguous = false; // cannot happen: bone_list lists a segment's bones consecutively
			}
		}
		std::function<int(int)> seg_height = [&](int si) {
			int h = 0;
			for (int c : R.segments[si].child_segs) {
				h = std::max(h, 1 + seg_height(c));
			}
			height[si] = h;
			return h;
		};
		int n_phases = 0;
		for (int si : R.root_segments) {
			n_phases = std::max(n_phases, 1 + seg_height(si));
		}
		R.sp_serial_cost = 0;
		for (int si = 0; si < nseg; si++) {
			R.sp_serial_cost += cost[si];
		}
		R.sp_roles = R.sp_phases = R.sp_slots = 0;
		R.sp_critical_cost = R.sp_serial_cost;
		R.sched.clear();
		if (contiguous && n_phases > 0 && ns > 0 && ns < 32767) {
			std::vector<std::vector<int>> by_phase(n_phases);
			int width = 1;
			for (int si = 0; si < nseg; si++) {
				if (R.segments[si].kept && s0[si] >= 0) {
					by_phase[height[si]].push_back(si);
				}
			}
			for (auto &v : by_phase) {
				width = std::max(width, (int)v.size());
			}
			const int roles = std::min(width, kMaxSpRoles);
			std::vector<std::vector<std::vector<int>>> assign(n_phases, std::vector<std::vector<int>>(roles));
			int slots = 1;
			double critical = 0;
			for (int ph = 0; ph < n_phases; ph++) {
				std::vector<int> order = by_phase[ph];
				std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
				std::vector<double> load(roles, 0.0);
				for (int si : order) {
					int r = (int)(std::min_element(load.begin(), load.end()) - load.begin());
					assign[ph][r].push_back(si);
					load[r] += cost[si];
					slots = std::max(slots, (int)assign[ph][r].size());
				}
				critical += *std::max_element(load.begin(), load.end());
			}
			R.sp_roles = roles;
			R.sp_phases = n_phases;
			R.sp_slots = slots;
			R.sched.assign((size_t)n_phases * slots * roles, BlobSpan{ 0, 0, 1, 0, 0, 0 });
			for (int ph = 0; ph < n_phases; ph++) {
				for (int r = 0; r < roles; r++) {
					for (size_t k = 0; k < assign[ph][r].size(); k++) {
						const int si = assign[ph][r][k];
						R.sched[((size_t)ph * slots + k) * roles + r] = BlobSpan{ (int16_t)s0[si], (int16_t)s1[si], 1, 0, 0, 0 };
					}
				}
			}
			// Teams: warps a phase leaves idle become heading helpers of its busiest multi-effector segments (at most
			// kMaxSpTeams per phase; the stabilisation loop and constraint mode keep the plain path).
			R.step_path.assign(R.steps.size(), -1);
			R.path_refs.clear();
			R.paths.clear();
			R.sp_team_bufs = R.sp_team_headings = 0;
			critical = 0;
			const bool teams_allowed = d->stabilization_passes <= 0 && !d->constraint_mode;
			// Estimated cost of a span run by a team of k warps (instruction-equivalents of the owner's critical path):
			// the walk products (80 each) and raw headings (140 per effector) are what the members share, the
			// QCP accumulation (51 per heading, +13 for the centroid pass) stays with the owner; a team step pays two
			// barriers and a shared-memory round trip of the headings.
			auto span_cost = [&](int si, int k) {
				double total = 0;
				for (int s = s0[si]; s < s1[si]; s++) {
					const BlobStep &st = R.steps[s];
					const bool tr = (st.flags & STEP_TRANSLATE) != 0;
					const bool snaps = (st.flags & STEP_IK_PARENT) && (st.flags & (STEP_SWING | STEP_TWIST));
					double c = (snaps ? 3000.0 : 1600.0) + 14.0 * st.cone_cnt + st.n_headings * (51.0 + (tr ? 13.0 : 0.0));
					if (k <= 1) {
						c += 80.0 * st.fk_cnt * ((tr && st.eff_cnt > 16) ? 2.0 : 1.0) + 140.0 * st.eff_cnt * (tr ? 2.0 : 1.0);
					} else {
						std::vector<double> load(k, 0.0);
						const int b = R.bone_order[s];
						int e_i = 0;
						for (int e : R.segments[si].effectors) {
							int plen = 0;
							for (int x = e; x != b && x >= 0; x = ik_parent[x]) {
								plen++;
							}
							load[e_i % k] += 80.0 * plen + 160.0;
							e_i++;
						}
						c += 150.0 + 6.0 * st.n_headings + *std::max_element(load.begin(), load.end());
					}
					total += c;
				}
				return total;
			};
			R.sp_serial_cost = 0;
			for (int si = 0; si < nseg; si++) {
				if (R.segments[si].kept && s0[si] >= 0) {
					R.sp_serial_cost += span_cost(si, 1);
				}
			}
			for (int ph = 0; ph < n_phases; ph++) {
				std::vector<int> idle, busy;
				bool single_slot = true;
				for (int r = 0; r < roles; r++) {
					if (assign[ph][r].empty()) {
						idle.push_back(r);
					} else {
						busy.push_back(r);
						single_slot = single_slot && assign[ph][r].size() == 1;
					}
				}
				std::vector<int> team_of_role(roles, 1); // owner role -> team size
				std::vector<std::vector<int>> helpers(roles);
				static const long team_phase_mask = getenv("MBIK_SP_TEAM_PHASES") ? strtol(getenv("MBIK_SP_TEAM_PHASES"), nullptr, 0) : -1; // tuning knob
				// Teams only in phases with ONE busy segment: every warp of a group streams the ~100 KB step body through the
				// SM's instruction caches, and warps at different places of it compete for that fetch path -- measured on
				// quad80, helpers next to a second busy segment slowed every warp of the SM by ~30 % (the team's own span
				// got 28 % shorter, the batch 20 % slower), whereas helpers of a lone segment run in step with each other
				// and with their owner (humanoid22: spine and hips phases, -10 % latency).
				static const bool teams_beside_busy = getenv("MBIK_SP_TEAMS_BESIDE_BUSY") != nullptr; // tuning knob
				if (teams_allowed && single_slot && !idle.empty() && (busy.size() == 1 || teams_beside_busy) && ((team_phase_mask >> (ph < 60 ? ph : 60)) & 1)) {
					// give idle warps, one at a time, to the span on the phase's critical path while that shortens it
					std::vector<double> cur(roles, 0.0);
					for (int r : busy) {
						cur[r] = span_cost(assign[ph][r][0], 1);
					}
					int n_teams = 0;
					while (!idle.empty()) {
						int crit = busy[0];
						for (int r : busy) {
							if (cur[r] > cur[crit]) {
								crit = r;
							}
						}
						const int si = assign[ph][crit][0];
						const int E = R.steps[s0[si]].eff_cnt;
						if (team_of_role[crit] >= E || (team_of_role[crit] == 1 && n_teams >= kMaxSpTeams)) {
							break;
						}
						// a second member may not pay for the barriers where a third does: look ahead over the sizes reachable
						int best_k = -1;
						double best_c = cur[crit] * 0.97;
						for (int k = team_of_role[crit] + 1; k <= std::min(E, team_of_role[crit] + (int)idle.size()); k++) {
							const double c = span_cost(si, k);
							if (c < best_c) {
								best_c = c;
								best_k = k;
							}
						}
						if (best_k < 0) {
							break;
						}
						if (team_of_role[crit] == 1) {
							n_teams++;
						}
						while (team_of_role[crit] < best_k) {
							// prefer a warp whose scheduler (warp index mod 4) no busy warp of this phase uses: two warps
							// on one scheduler share its L0 instruction cache, and they run different code
							size_t pick = 0;
							int pick_clash = 1 << 30;
							for (size_t i = 0; i < idle.size(); i++) {
								int clash = 0;
								for (int r : busy) {
									clash += (r % 4 == idle[i] % 4) ? (r == crit ? 1 : 4) : 0;
								}
								for (int r = 0; r < roles; r++) {
									for (int h : helpers[r]) {
										clash += (h % 4 == idle[i] % 4) ? 1 : 0;
									}
								}
								if (clash < pick_clash) {
									pick_clash = clash;
									pick = i;
								}
							}
							helpers[crit].push_back(idle[pick]);
							idle.erase(idle.begin() + pick);
							team_of_role[crit]++;
						}
						cur[crit] = best_c;
					}
				}
				int n_bufs = 0;
				double phase_cost = 0;
				for (int r : busy) {
					double load = 0;
					for (int si : assign[ph][r]) {
						load += span_cost(si, 1);
					}
					if (team_of_role[r] > 1) {
						const int si = assign[ph][r][0];
						const int team = team_of_role[r], buf = n_bufs++;
						BlobSpan &own = R.sched[((size_t)ph * slots + 0) * roles + r];
						own.team = (int8_t)team;
						own.buf = (int8_t)buf;
						for (size_t m = 0; m < helpers[r].size(); m++) {
							R.sched[((size_t)ph * slots + 0) * roles + helpers[r][m]] = BlobSpan{ own.s0, own.s1, (int8_t)team, (int8_t)(m + 1), (int8_t)buf, 0 };
						}
						load = span_cost(si, team);
						for (int s = s0[si]; s < s1[si]; s++) {
							const BlobStep &st = R.steps[s];
							const FlatSegment &S = R.segments[si];
							R.sp_team_headings = std::max(R.sp_team_headings, st.n_headings);
							R.step_path[s] = (int32_t)R.path_refs.size();
							const int b = R.bone_order[s];
							for (int e : S.effectors) {
								std::vector<int> up;
								for (int x = e; x != b && x >= 0; x = ik_parent[x]) {
									up.push_back(x);
								}
								BlobPathRef pr;
								pr.off = (int32_t)R.paths.size();
								pr.cnt = (int32_t)up.size();
								for (auto it2 = up.rbegin(); it2 != up.rend(); ++it2) {
									R.paths.push_back((int16_t)R.t_of_bone[*it2]);
								}
								R.path_refs.push_back(pr);
							}
						}
					}
					phase_cost = std::max(phase_cost, load);
				}
				R.sp_team_bufs = std::max(R.sp_team_bufs, n_bufs);
				critical += phase_cost;
			}
			R.sp_critical_cost = critical;
		}
	}

	// ---- assemble the blob ----
	BlobHeader hdr;
	memset(&hdr, 0, sizeof(hdr));
	hdr.magic = 0x4B49424Du;
	hdr.n_bones = nb;
	hdr.n_solved = ns;
	hdr.n_steps = (int)R.steps.size();
	hdr.n_pins = (int)R.pins.size();
	hdr.n_effs = (int)R.effs.size();
	hdr.n_fk = (int)R.fk.size();
	hdr.n_cones = (int)R.cones.size();
	hdr.n_pass = (int)R.pass.size();
	hdr.iterations = R.iterations;
	hdr.constraint_mode = d->constraint_mode;
	hdr.stabilization_passes = d->stabilization_passes;
	R.stabilization_passes = d->stabilization_passes;
	hdr.n_chain = (int)R.chain.size();
	hdr.max_seg_len = R.max_seg_len;
	hdr.max_stack = R.max_stack;
	hdr.sp_roles = R.sp_roles;
	hdr.sp_phases = R.sp_phases;
	hdr.sp_slots = R.sp_slots;
	hdr.sp_team_bufs = R.sp_team_bufs;
	hdr.sp_team_headings = R.sp_team_headings;
	std::vector<float> rest(nb * 12);
	memcpy(rest.data(), d->rest_local, sizeof(float) * 12 * nb);
	int hdr_keep_lo = 0, hdr_keep_n = 0;
	{
		// L2-keep set of the streamed-walk instantiation: the most-read local poses (effector walks + segment-chain refreshes per
		// iteration), as many as fit 24 MB of L2 for a resident batch of 148 x 512 poses (48 B each): 6 bones.  Measured
		// (chain64, 75 776 poses, profiles/r2_exp_glw_keep_mb.log): what pays is the evict_first policy on everything else --
		// 59.0 ms without policies, 53.4 ms with an empty keep set, 53.0 / 53.1 / 54.1 / 55.7 / 57.0 ms keeping 24 / 40 / 56 / 72 /
		// 88 MB: the streamed poses no longer push the kernel's code, the targets and the thread-local scratch out of L2
		std::vector<double> reads(R.bones.size(), 0.0);
		for (const BlobStep &st : R.steps) {
			const double passes = ((st.flags & STEP_TRANSLATE) && st.eff_cnt > 16) ? 2.0 : 1.0;
			for (int k = 0; k < st.fk_cnt; k++) {
				reads[(size_t)R.fk[(size_t)st.fk_off + k].child] += passes;
			}
			if (st.flags & STEP_SEG_FIRST) {
				for (int k = 0; k < st.chain_cnt; k++) {
					reads[(size_t)R.chain[(size_t)st.chain_off + k]] += 1.0;
				}
			}
		}
		// a contiguous t range (chains: the deepest bones), so that the kernel tests membership with one compare
		static const double keep_mb = getenv("MBIK_GLW_KEEP_MB") ? atof(getenv("MBIK_GLW_KEEP_MB")) : 24.0; // tuning knob
		const int keep = (int)(keep_mb * 1.0e6 / (48.0 * 148 * 512));
		int best_lo = 0, best_n = 0;
		double best_sum = 0.0;
		for (int lo = 0; lo < (int)reads.size() && keep > 0; lo++) {
			double sum = 0.0;
			for (int n = 1; n <= keep && lo + n <= (int)reads.size(); n++) {
				sum += reads[(size_t)(lo + n - 1)];
				if (sum > best_sum && reads[(size_t)(lo + n - 1)] >= 4.0) {
					best_sum = sum;
					best_lo = lo;
					best_n = n;
				}
			}
		}
		hdr_keep_lo = best_lo;
		hdr_keep_n = best_n;
		for (int t = best_lo; t < best_lo + best_n; t++) {
			R.bones[(size_t)t].flags |= BONE_L2_KEEP;
		}
		// plain runs of the walk lists (BlobFk::pad >> 1), per step from the back
		for (const BlobStep &st : R.steps) {
			int next_run = 0;
			for (int k = st.fk_cnt - 1; k >= 0; k--) {
				BlobFk &op = R.fk[(size_t)st.fk_off + k];
				int run = 0;
				if (op.src_slot < 0 && op.push_slot < 0) {
					run = 1;
					if (k + 1 < st.fk_cnt && op.eff < 0 && next_run > 0 && R.fk[(size_t)st.fk_off + k + 1].child == op.child + 1) {
						run = next_run + 1;
					}
				}
				run = run > 16000 ? 16000 : run;
				op.pad = (int16_t)(((R.bones[(size_t)op.child].flags & BONE_L2_KEEP) ? 1 : 0) | (run << 1));
				next_run = run;
			}
		}
	}
	hdr.glw_keep_lo = hdr_keep_lo;
	hdr.glw_keep_n = hdr_keep_n;
	std::vector<int16_t> list_row(ns, 0); // t index -> position in bone_list
	for (size_t i = 0; i < R.bone_order.size(); i++) {
		list_row[(size_t)R.t_of_bone[R.bone_order[i]]] = (int16_t)i;
	}
	auto assemble = [&](bool tail_layout) {
		R.blob.clear();
		R.blob.resize(sizeof(BlobHeader), 0);
		hdr.off_steps = append_section(R.blob, R.steps);
		hdr.off_bones = append_section(R.blob, R.bones);
		hdr.off_effs = append_section(R.blob, R.effs);
		if (!tail_layout) {
			hdr.off_fk = append_section(R.blob, R.fk);
		}
		hdr.off_cones = append_section(R.blob, R.cones);
		hdr.off_pass = append_section(R.blob, R.pass);
		hdr.off_chain = append_section(R.blob, R.chain);
		hdr.off_list_row = append_section(R.blob, list_row);
		hdr.off_rest = append_section(R.blob, rest);
		hdr.off_sched = append_section(R.blob, R.sched);
		hdr.resident_bytes = hdr.off_sched; // tail layout: nothing from here on is staged into shared memory
		hdr.off_step_path = append_section(R.blob, R.step_path);
		hdr.off_path_refs = append_section(R.blob, R.path_refs);
		hdr.off_paths = append_section(R.blob, R.paths);
		if (tail_layout) {
			hdr.off_fk = append_section(R.blob, R.fk);
		}
		while (R.blob.size() % 16) {
			R.blob.push_back(0);
		}
		if (!tail_layout) {
			hdr.resident_bytes = (uint32_t)R.blob.size();
		}
	};
	assemble(false);
	// the walk list (quadratic in chain depth) and the segment-parallel tables stay in global memory when the blob would not
	// fit shared memory -- and already when a rig of the {256, 256, 32} variant would leave no room for the streamed-walk
	// instantiation's cp.async ring (144 KiB at 512 threads)
	if (R.blob.size() > kResidentBlobBudget || (ns > 128 && R.blob.size() > (size_t)(227 - 144 - 1) * 1024)) {
		assemble(true);
	}
	hdr.total_bytes = (uint32_t)R.blob.size();
	memcpy(R.blob.data(), &hdr, sizeof(hdr));
	if (getenv("MBIK_DEBUG_WALK_READS")) {
		// analysis aid: how often each solved bone's local pose is read by the effector walks of one iteration
		std::vector<double> reads(R.bones.size(), 0.0);
		double total = 0;
		for (const BlobStep &st : R.steps) {
			const double passes = (st.flags & STEP_TRANSLATE) ? 2.0 : 1.0;
			for (int k = 0; k < st.fk_cnt; k++) {
				reads[(size_t)R.fk[(size_t)st.fk_off + k].child] += passes;
				total += passes;
			}
		}
		std::vector<double> sorted = reads;
		std::sort(sorted.begin(), sorted.end(), [](double a, double b) { return a > b; });
		fprintf(stderr, "[mbik] walk reads per iteration: %.0f over %zu solved bones; share of the hottest 4 / 7 / 10 / 14 bones:", total, reads.size());
		for (int m : { 4, 7, 10, 14 }) {
			double acc = 0;
			for (int i = 0; i < m && i < (int)sorted.size(); i++) {
				acc += sorted[(size_t)i];
			}
			fprintf(stderr, " %.1f%%", total > 0 ? 100.0 * acc / total : 0.0);
		}
		fprintf(stderr, "\n");
	}
	return MBIK_OK;
}

bool validate_schedule(const FlatRig &R, int cap_bones, int cap_seg, int cap_stack, std::string &error) {
	const int ns = (int)R.bones.size(), n_fk = (int)R.fk.size(), n_effs = (int)R.effs.size(), n_cones = (int)R.cones.size();
	const int n_chain = (int)R.chain.size(), n_pins = (int)R.pins.size();
	auto bad = [&](const char *what, int step) {
		error = std::string("inconsistent schedule: ") + what + " (step " + std::to_string(step) + ")";
		return false;
	};
	if (ns > cap_bones || (int)R.steps.size() != ns) {
		return bad("solved-bone count", -1);
	}
	for (int t = 0; t < ns; t++) {
		if (R.bones[t].skel_bone < 0 || R.bones[t].skel_bone >= R.n_bones || R.bones[t].parent < -1 || R.bones[t].parent >= ns) {
			return bad("bone table", t);
		}
	}
	for (const BlobPass &p : R.pass) {
		if (p.skel_bone < 0 || p.skel_bone >= R.n_bones) {
			return bad("pass-through table", -1);
		}
	}
	for (const BlobEff &e : R.effs) {
		if (e.bone < 0 || e.bone >= ns || e.pin < 0 || e.pin >= n_pins) {
			return bad("effector table", -1);
		}
	}
	for (int s = 0; s < (int)R.steps.size(); s++) {
		const BlobStep &S = R.steps[s];
		if (S.bone < 0 || S.bone >= ns || S.parent < -1 || S.parent >= ns) {
			return bad("bone index", s);
		}
		if (S.seg_len < 1 || S.seg_len > cap_seg || S.pslot < 0 || S.pslot >= S.seg_len) {
			return bad("segment slot", s);
		}
		if (S.eff_off < 0 || S.eff_cnt < 0 || S.eff_off + S.eff_cnt > n_effs || S.fk_off < 0 || S.fk_cnt < 0 || S.fk_off + S.fk_cnt > n_fk) {
			return bad("effector / walk range", s);
		}
		if (S.cone_off < 0 || S.cone_cnt < 0 || S.cone_off + S.cone_cnt > n_cones) {
			return bad("cone range", s);
		}
		if (S.flags & STEP_SEG_FIRST) {
			if (S.chain_off < 0 || S.chain_cnt < 0 || S.chain_off + S.chain_cnt > n_chain || S.chain_cnt - S.seg_len < -1) {
				return bad("ancestor chain", s);
			}
			for (int k = 0; k < S.chain_cnt; k++) {
				int t = R.chain[S.chain_off + k];
				if (t < 0 || t >= ns) {
					return bad("ancestor chain entry", s);
				}
			}
		}
		if ((S.flags & STEP_SELF_EFF) && S.eff_cnt < 1) {
			return bad("self effector", s);
		}
		if ((S.flags & STEP_PUSH_SELF) && cap_stack < 1) {
			return bad("walk stack", s);
		}
		for (int k = 0; k < S.fk_cnt; k++) {
			const BlobFk &op = R.fk[S.fk_off + k];
			if (op.child < 0 || op.child >= ns || op.src_slot < -1 || op.src_slot >= cap_stack || op.push_slot < -1 || op.push_slot >= cap_stack ||
					op.eff < -1 || op.eff >= S.eff_cnt) {
				return bad("walk op", s);
			}
		}
	}
	// segment-parallel schedule: every step in exactly one span, every span one whole segment (tip .. segment root), and
	// every segment in a later phase than all segments below it
	if (R.sp_roles > 0) {
		const int n_steps = (int)R.steps.size();
		if (R.sp_roles > kMaxSpRoles || R.sp_phases < 1 || R.sp_slots < 1 || (int)R.sched.size() != R.sp_phases * R.sp_slots * R.sp_roles) {
			return bad("segment-parallel table shape", -1);
		}
		std::vector<int> phase_of_step(n_steps, -1);
		if (R.sp_team_bufs < 0 || R.sp_team_bufs > kMaxSpTeams || (int)R.step_path.size() != n_steps) {
			return bad("team table shape", -1);
		}
		for (int ph = 0; ph < R.sp_phases; ph++) {
			std::vector<int> members_seen(kMaxSpTeams, 0), team_size(kMaxSpTeams, 0), team_s0(kMaxSpTeams, -1);
			for (int k = 0; k < R.sp_slots * R.sp_roles; k++) {
				const BlobSpan sp = R.sched[(size_t)ph * R.sp_slots * R.sp_roles + k];
				if (sp.s0 < 0 || sp.s1 < sp.s0 || sp.s1 > n_steps || sp.team < 1 || sp.member < 0 || sp.member >= sp.team) {
					return bad("segment-parallel span", sp.s0);
				}
				if (sp.s0 == sp.s1) {
					continue;
				}
				if (!(R.steps[sp.s0].flags & STEP_SEG_FIRST) || !(R.steps[sp.s1 - 1].flags & STEP_SEG_ROOT)) {
					return bad("segment-parallel span is not a whole segment", sp.s0);
				}
				if (sp.team > 1) {
					// every member of a team carries the same span; each member index exactly once (a missing member
					// would leave the others waiting at the team barrier)
					if (sp.buf < 0 || sp.buf >= R.sp_team_bufs || k >= R.sp_roles /* teams live in slot 0 */ || sp.team > R.sp_roles) {
						return bad("team span", sp.s0);
					}
					if (team_size[sp.buf] == 0) {
						team_size[sp.buf] = sp.team;
						team_s0[sp.buf] = sp.s0;
					} else if (team_size[sp.buf] != sp.team || team_s0[sp.buf] != sp.s0) {
						return bad("team members disagree", sp.s0);
					}
					if (members_seen[sp.buf] & (1 << sp.member)) {
						return bad("duplicate team member", sp.s0);
					}
					members_seen[sp.buf] |= 1 << sp.member;
					for (int s = sp.s0; s < sp.s1; s++) {
						const BlobStep &S = R.steps[s];
						const int ref = R.step_path[s];
						if (ref < 0 || ref + S.eff_cnt > (int)R.path_refs.size() || S.n_headings > R.sp_team_headings || S.eff_cnt < 2) {
							return bad("team step path table", s);
						}
						int h = 0;
						for (int e = 0; e < S.eff_cnt; e++) {
							const BlobPathRef pr = R.path_refs[ref + e];
							const BlobEff &E = R.effs[S.eff_off + e];
							if (pr.off < 0 || pr.cnt < 0 || pr.off + pr.cnt > (int)R.paths.size() || E.h_off != h) {
								return bad("team step path", s);
							}
							h += E.n_headings;
							int at = S.bone; // the path descends from the solved bone to the effector's bone
							for (int q = 0; q < pr.cnt; q++) {
								const int t = R.paths[pr.off + q];
								if (t < 0 || t >= ns || R.bones[t].parent != at) {
									return bad("team step path entry", s);
								}
								at = t;
							}
							if (at != E.bone) {
								return bad("team step path end", s);
							}
						}
						if (h != S.n_headings) {
							return bad("team step heading count", s);
						}
					}
					if (sp.member != 0) {
						continue; // helpers repeat the owner's span
					}
				}
				for (int s = sp.s0; s < sp.s1; s++) {
					if (phase_of_step[s] >= 0 || (s > sp.s0 && (R.steps[s].flags & STEP_SEG_FIRST))) {
						return bad("segment-parallel span overlap", s);
					}
					phase_of_step[s] = ph;
				}
			}
			for (int b = 0; b < kMaxSpTeams; b++) {
				if (team_size[b] > 0 && members_seen[b] != (1 << team_size[b]) - 1) {
					return bad("incomplete team", team_s0[b]);
				}
			}
		}
		std::vector<int> phase_of_bone(ns, -1);
		for (int s = 0; s < n_steps; s++) {
			if (phase_of_step[s] < 0) {
				return bad("step missing from the segment-parallel schedule", s);
			}
			phase_of_bone[R.steps[s].bone] = phase_of_step[s];
		}
		for (int s = 0; s < n_steps; s++) {
			// a bone-step reads the bones on its effector walk (its own segment or segments below it) and its ancestors
			for (int k = 0; k < R.steps[s].fk_cnt; k++) {
				if (phase_of_bone[R.fk[R.steps[s].fk_off + k].child] > phase_of_step[s]) {
					return bad("segment-parallel phase order (walk)", s);
				}
			}
			for (int p = R.steps[s].parent; p >= 0; p = R.bones[p].parent) {
				if (phase_of_bone[p] < phase_of_step[s]) {
					return bad("segment-parallel phase order (ancestor)", s);
				}
			}
		}
	}
	return true;
}

} // namespace mbik
