// TEST INFRASTRUCTURE ONLY -- see ewbik_oracle.h.  Restatement of the reference solve + setup path.
#include "ewbik_oracle.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace orc {

// =================================================================================================
// IKNode3D -- src/math/ik_node_3d.cpp
// =================================================================================================

// ik_node_3d.cpp:33-49
void IKNode3D::_propagate_transform_changed() {
	for (Ref<IKNode3D> &transform : children) {
		if (transform) {
			transform->_propagate_transform_changed();
		}
	}
	dirty |= DIRTY_GLOBAL;
}

// ik_node_3d.cpp:56-67 : local.basis = P^-1 * R * P * local.basis ; dirties only this node unless p_propagate
void IKNode3D::rotate_local_with_global(const Basis &p_basis, bool p_propagate) {
	Ref<IKNode3D> parent_ik_node = parent.lock();
	if (!parent_ik_node) {
		return;
	}
	const Basis new_rot = parent_ik_node->get_global_transform().basis;
	local_transform.basis = new_rot.inverse() * p_basis * new_rot * local_transform.basis;
	dirty |= DIRTY_GLOBAL;
	if (p_propagate) {
		_propagate_transform_changed();
	}
}

// ik_node_3d.cpp:69-75
void IKNode3D::set_transform(const Transform3D &p_transform) {
	if (local_transform != p_transform) {
		local_transform = p_transform;
		dirty |= DIRTY_VECTORS;
		_propagate_transform_changed();
	}
}

// ik_node_3d.cpp:77-83
void IKNode3D::set_global_transform(const Transform3D &p_transform) {
	Ref<IKNode3D> ik_node = parent.lock();
	Transform3D xform = ik_node ? ik_node->get_global_transform().affine_inverse() * p_transform : p_transform;
	local_transform = xform;
	dirty |= DIRTY_VECTORS;
	_propagate_transform_changed();
}

// ik_node_3d.cpp:85-91 (DIRTY_LOCAL is never set anywhere in the module, so _update_local_transform is dead)
Transform3D IKNode3D::get_transform() const {
	return local_transform;
}

// ik_node_3d.cpp:93-113 (disable_scale is never enabled on the solve path)
Transform3D IKNode3D::get_global_transform() const {
	if (dirty & DIRTY_GLOBAL) {
		Ref<IKNode3D> ik_node = parent.lock();
		if (ik_node) {
			global_transform = ik_node->get_global_transform() * local_transform;
		} else {
			global_transform = local_transform;
		}
		dirty &= ~DIRTY_GLOBAL;
	}
	return global_transform;
}

// ik_node_3d.cpp:123-132 (note: does not recompute the local transform)
void IKNode3D::set_parent(Ref<IKNode3D> p_parent) {
	Ref<IKNode3D> self = shared_from_this();
	if (p_parent) {
		p_parent->children.remove(self);
	}
	parent = p_parent;
	if (p_parent) {
		p_parent->children.push_back(self);
	}
	_propagate_transform_changed();
}

// ik_node_3d.cpp:138-144
Vector3 IKNode3D::to_local(const Vector3 &p_global) const {
	return get_global_transform().affine_inverse().xform(p_global);
}
Vector3 IKNode3D::to_global(const Vector3 &p_local) const {
	return get_global_transform().xform(p_local);
}

void IKNode3D::cleanup() {
	for (Ref<IKNode3D> &child : children) {
		child->parent.reset();
	}
}

// =================================================================================================
// IKRay3D -- src/ik_ray_3d.cpp
// =================================================================================================

// ik_ray_3d.cpp:36-40
IKRay3D::IKRay3D(Vector3 p_p1, Vector3 p_p2) {
	working_vector = p_p1;
	point_1 = p_p1;
	point_2 = p_p2;
}

// ik_ray_3d.cpp:42-45
Vector3 IKRay3D::get_heading() {
	working_vector = point_2;
	return working_vector - point_1;
}

// ik_ray_3d.cpp:64-73
void IKRay3D::elongate(real_t amt) {
	Vector3 midPoint = (point_1 + point_2) * 0.5f;
	Vector3 p1Heading = point_1 - midPoint;
	Vector3 p2Heading = point_2 - midPoint;
	Vector3 p1Add = p1Heading.normalized() * amt;
	Vector3 p2Add = p2Heading.normalized() * amt;
	point_1 = p1Heading + p1Add + midPoint;
	point_2 = p2Heading + p2Add + midPoint;
}

// ik_ray_3d.cpp:75-85
Vector3 IKRay3D::get_intersects_plane(Vector3 ta, Vector3 tb, Vector3 tc) {
	Vector3 uvw;
	Vector3 tta = ta, ttb = tb, ttc = tc;
	tta -= point_1;
	ttb -= point_1;
	ttc -= point_1;
	Vector3 result = plane_intersect_test(tta, ttb, ttc, &uvw);
	return result + point_1;
}

// ik_ray_3d.cpp:87-94
int IKRay3D::intersects_sphere(Vector3 sphereCenter, real_t radius, Vector3 *S1, Vector3 *S2) {
	Vector3 tp1 = point_1 - sphereCenter;
	Vector3 tp2 = point_2 - sphereCenter;
	int result = intersects_sphere(tp1, tp2, radius, S1, S2);
	*S1 += sphereCenter;
	*S2 += sphereCenter;
	return result;
}

// ik_ray_3d.cpp:112-144
int IKRay3D::intersects_sphere(Vector3 rp1, Vector3 rp2, real_t radius, Vector3 *S1, Vector3 *S2) {
	Vector3 direction = rp2 - rp1;
	Vector3 e = direction;
	e.normalize();
	Vector3 h = Vector3(0.0f, 0.0f, 0.0f);
	h = h - rp1;
	real_t lf = e.dot(h);
	real_t radpow = radius * radius;
	real_t hdh = h.length_squared();
	real_t lfpow = lf * lf;
	real_t s = radpow - hdh + lfpow;
	if (s < 0.0f) {
		return 0;
	}
	s = Math::sqrt(s);
	int result = 0;
	if (lf < s) {
		if (lf + s >= 0) {
			s = -s;
			result = 1;
		}
	} else {
		result = 2;
	}
	*S1 = e * (lf - s);
	*S1 += rp1;
	*S2 = e * (lf + s);
	*S2 += rp1;
	return result;
}

// ik_ray_3d.cpp:146-166 (the barycentric() call only fills the unused uvw output; omitted)
Vector3 IKRay3D::plane_intersect_test(Vector3 ta, Vector3 tb, Vector3 tc, Vector3 *uvw) {
	Vector3 u = tb;
	Vector3 v = tc;
	Vector3 n = Vector3(0, 0, 0);
	Vector3 dir = get_heading();
	Vector3 w0 = Vector3(0, 0, 0);
	real_t r, a, b;
	u -= ta;
	v -= ta;
	n = u.cross(v).normalized();
	w0 -= ta;
	a = -(n.dot(w0));
	b = n.dot(dir);
	r = a / b;
	Vector3 I = dir;
	I *= r;
	(void)uvw;
	return I;
}

// =================================================================================================
// IKLimitCone3D -- src/ik_open_cone_3d.cpp
// =================================================================================================

// ik_open_cone_3d.cpp:36-120
void IKLimitCone3D::update_tangent_handles(Ref<IKLimitCone3D> p_next) {
	if (!p_next) {
		return;
	}
	double radA = get_radius();
	double radB = p_next->get_radius();
	Vector3 A = get_control_point();
	Vector3 B = p_next->get_control_point();
	Vector3 arc_normal = A.cross(B).normalized();

	double tRadius = (Math_PI - (radA + radB)) / 2;
	double boundaryPlusTangentRadiusA = radA + tRadius;
	double boundaryPlusTangentRadiusB = radB + tRadius;

	// `Vector3 * double`: scalar narrows to real_t
	Vector3 scaledAxisA = A * (real_t)Math::cos(boundaryPlusTangentRadiusA);
	Quaternion temp_var = IKKusudama3D::get_quaternion_axis_angle(arc_normal, (real_t)boundaryPlusTangentRadiusA);
	Vector3 planeDir1A = temp_var.xform(A);
	Quaternion tempVar2 = IKKusudama3D::get_quaternion_axis_angle(A, (real_t)(Math_PI / 2));
	Vector3 planeDir2A = tempVar2.xform(planeDir1A);

	Vector3 scaledAxisB = B * (real_t)::cos(boundaryPlusTangentRadiusB);
	Quaternion tempVar3 = IKKusudama3D::get_quaternion_axis_angle(arc_normal, (real_t)boundaryPlusTangentRadiusB);
	Vector3 planeDir1B = tempVar3.xform(B);
	Quaternion tempVar4 = IKKusudama3D::get_quaternion_axis_angle(B, (real_t)(Math_PI / 2));
	Vector3 planeDir2B = tempVar4.xform(planeDir1B);

	IKRay3D r1B(planeDir1B, scaledAxisB);
	IKRay3D r2B(planeDir1B, planeDir2B);
	r1B.elongate(99);
	r2B.elongate(99);

	Vector3 intersection1 = r1B.get_intersects_plane(scaledAxisA, planeDir1A, planeDir2A);
	Vector3 intersection2 = r2B.get_intersects_plane(scaledAxisA, planeDir1A, planeDir2A);

	IKRay3D intersectionRay(intersection1, intersection2);
	intersectionRay.elongate(99);

	Vector3 sphereIntersect1;
	Vector3 sphereIntersect2;
	Vector3 sphereCenter;
	intersectionRay.intersects_sphere(sphereCenter, 1.0f, &sphereIntersect1, &sphereIntersect2);

	set_tangent_circle_center_next_1(sphereIntersect1);
	set_tangent_circle_center_next_2(sphereIntersect2);
	set_tangent_circle_radius_next(tRadius);
	if (Math::is_zero_approx(tangent_circle_center_next_1.length_squared())) {
		tangent_circle_center_next_1 = get_orthogonal(control_point).normalized();
	}
	if (Math::is_zero_approx(tangent_circle_center_next_2.length_squared())) {
		tangent_circle_center_next_2 = get_orthogonal(tangent_circle_center_next_1 * -1).normalized();
	}
	// compute_triangles (:142-153) only fills first/second_triangle_next, which nothing on the solve path reads.
}

// ik_open_cone_3d.cpp:122-125
void IKLimitCone3D::set_tangent_circle_radius_next(double rad) {
	tangent_circle_radius_next = rad;
	tangent_circle_radius_next_cos = ::cos(tangent_circle_radius_next);
}

// ik_open_cone_3d.cpp:159-166
void IKLimitCone3D::set_control_point(Vector3 p_control_point) {
	if (Math::is_zero_approx(p_control_point.length_squared())) {
		control_point = Vector3(0, 1, 0);
	} else {
		control_point = p_control_point;
		control_point.normalize();
	}
}

// ik_open_cone_3d.cpp:176-179
void IKLimitCone3D::set_radius(double p_radius) {
	radius = p_radius;
	radius_cosine = ::cos(p_radius);
}

// ik_open_cone_3d.cpp:267-283
Vector3 IKLimitCone3D::get_orthogonal(Vector3 p_in) {
	Vector3 result;
	float threshold = p_in.length() * 0.6f;
	if (threshold > 0.f) {
		if (Math::abs(p_in.x) <= threshold) {
			float inverse = 1.f / Math::sqrt(p_in.y * p_in.y + p_in.z * p_in.z);
			return result = Vector3(0.f, inverse * p_in.z, -inverse * p_in.y);
		} else if (Math::abs(p_in.y) <= threshold) {
			float inverse = 1.f / Math::sqrt(p_in.x * p_in.x + p_in.z * p_in.z);
			return result = Vector3(-inverse * p_in.z, 0.f, inverse * p_in.x);
		}
		float inverse = 1.f / Math::sqrt(p_in.x * p_in.x + p_in.y * p_in.y);
		return result = Vector3(inverse * p_in.y, -inverse * p_in.x, 0.f);
	}
	return result;
}

// ik_open_cone_3d.cpp:285-321
Vector3 IKLimitCone3D::get_on_great_tangent_triangle(Ref<IKLimitCone3D> next, Vector3 input) const {
	if (!next) {
		return input;
	}
	Vector3 c1xc2 = control_point.cross(next->control_point);
	double c1c2dir = input.dot(c1xc2);
	if (c1c2dir < 0.0) {
		Vector3 c1xt1 = control_point.cross(tangent_circle_center_next_1).normalized();
		Vector3 t1xc2 = tangent_circle_center_next_1.cross(next->control_point).normalized();
		if (input.dot(c1xt1) > 0 && input.dot(t1xc2) > 0) {
			double to_next_cos = input.dot(tangent_circle_center_next_1);
			if (to_next_cos > tangent_circle_radius_next_cos) {
				Vector3 plane_normal = tangent_circle_center_next_1.cross(input).normalized();
				plane_normal.normalize();
				Quaternion rotate_about_by = Quaternion(plane_normal, (real_t)tangent_circle_radius_next);
				return rotate_about_by.xform(tangent_circle_center_next_1);
			} else {
				return input;
			}
		} else {
			return Vector3(NAN, NAN, NAN);
		}
	} else {
		Vector3 t2xc1 = tangent_circle_center_next_2.cross(control_point).normalized();
		Vector3 c2xt2 = next->control_point.cross(tangent_circle_center_next_2).normalized();
		if (input.dot(t2xc1) > 0 && input.dot(c2xt2) > 0) {
			if (input.dot(tangent_circle_center_next_2) > tangent_circle_radius_next_cos) {
				Vector3 plane_normal = tangent_circle_center_next_2.cross(input).normalized();
				plane_normal.normalize();
				Quaternion rotate_about_by = Quaternion(plane_normal, (real_t)tangent_circle_radius_next);
				return rotate_about_by.xform(tangent_circle_center_next_2);
			} else {
				return input;
			}
		} else {
			return Vector3(NAN, NAN, NAN);
		}
	}
}

// ik_open_cone_3d.cpp:323-332
Vector3 IKLimitCone3D::_closest_cone(Ref<IKLimitCone3D> next, Vector3 input) const {
	if (!next) {
		return control_point;
	}
	if (input.dot(control_point) > input.dot(next->control_point)) {
		return control_point;
	} else {
		return next->control_point;
	}
}

// ik_open_cone_3d.cpp:358-381
Vector3 IKLimitCone3D::closest_to_cone(Vector3 input, std::vector<double> *in_bounds) const {
	Vector3 normalized_input = input.normalized();
	Vector3 normalized_control_point = get_control_point().normalized();
	if (normalized_input.dot(normalized_control_point) > get_radius_cosine()) {
		if (in_bounds != nullptr) {
			(*in_bounds)[0] = 1.0;
		}
		return Vector3(NAN, NAN, NAN);
	}
	Vector3 axis = normalized_control_point.cross(normalized_input).normalized();
	if (Math::is_zero_approx(axis.length_squared()) || !axis.is_finite()) {
		axis = Vector3(0, 1, 0);
	}
	Quaternion rot_to = IKKusudama3D::get_quaternion_axis_angle(axis, (real_t)get_radius());
	Vector3 axis_control_point = normalized_control_point;
	if (Math::is_zero_approx(axis_control_point.length_squared())) {
		axis_control_point = Vector3(0, 1, 0);
	}
	Vector3 result = rot_to.xform(axis_control_point);
	if (in_bounds != nullptr) {
		(*in_bounds)[0] = -1;
	}
	return result;
}

// ik_open_cone_3d.cpp:391-418 (only reached from the reference's own tests via get_closest_path_point)
Vector3 IKLimitCone3D::_get_on_path_sequence(Ref<IKLimitCone3D> next, Vector3 input) const {
	if (!next) {
		return Vector3(NAN, NAN, NAN);
	}
	Vector3 c1xc2 = get_control_point().cross(next->control_point).normalized();
	double c1c2dir = input.dot(c1xc2);
	if (c1c2dir < 0.0) {
		Vector3 c1xt1 = get_control_point().cross(tangent_circle_center_next_1).normalized();
		Vector3 t1xc2 = tangent_circle_center_next_1.cross(next->get_control_point()).normalized();
		if (input.dot(c1xt1) > 0.0f && input.dot(t1xc2) > 0.0f) {
			IKRay3D tan1ToInput(tangent_circle_center_next_1, input);
			Vector3 result = tan1ToInput.get_intersects_plane(Vector3(0.0f, 0.0f, 0.0f), get_control_point(), next->get_control_point());
			return result.normalized();
		} else {
			return Vector3(NAN, NAN, NAN);
		}
	} else {
		Vector3 t2xc1 = tangent_circle_center_next_2.cross(control_point).normalized();
		Vector3 c2xt2 = next->get_control_point().cross(tangent_circle_center_next_2).normalized();
		if (input.dot(t2xc1) > 0 && input.dot(c2xt2) > 0) {
			IKRay3D tan2ToInput(tangent_circle_center_next_2, input);
			Vector3 result = tan2ToInput.get_intersects_plane(Vector3(0.0f, 0.0f, 0.0f), get_control_point(), next->get_control_point());
			return result.normalized();
		} else {
			return Vector3(NAN, NAN, NAN);
		}
	}
}

// ik_open_cone_3d.cpp:236-248
Vector3 IKLimitCone3D::get_closest_path_point(Ref<IKLimitCone3D> next, Vector3 input) const {
	Vector3 result;
	if (!next) {
		// reference: _closest_cone(Ref(this), input) -> compares input.dot(cp) with itself -> control_point
		result = control_point;
	} else {
		result = _get_on_path_sequence(next, input);
		bool is_number = !(Math::is_nan(result.x) && Math::is_nan(result.y) && Math::is_nan(result.z));
		if (!is_number) {
			result = _closest_cone(next, input);
		}
	}
	return result;
}

// =================================================================================================
// IKKusudama3D -- src/ik_kusudama_3d.cpp
// =================================================================================================

// ik_kusudama_3d.cpp:37-89
void IKKusudama3D::_update_constraint(Ref<IKNode3D> p_limiting_axes) {
	std::vector<Vector3> directions;
	if (open_cones.size() == 1 && open_cones[0] != nullptr) {
		directions.push_back(open_cones[0]->get_control_point());
	} else {
		for (int i = 0; i < (int)open_cones.size() - 1; i++) {
			if (open_cones[i] == nullptr || open_cones[i + 1] == nullptr) {
				continue;
			}
			Vector3 this_control_point = open_cones[i]->get_control_point();
			Vector3 next_control_point = open_cones[i + 1]->get_control_point();
			Quaternion this_to_next = Quaternion(this_control_point, next_control_point);
			Vector3 axis = this_to_next.get_axis();
			double angle = this_to_next.get_angle() / 2.0;
			Vector3 half_angle = this_control_point.rotated(axis, (real_t)angle);
			half_angle *= this_to_next.get_angle();
			half_angle.normalize();
			directions.push_back(half_angle);
		}
	}

	Vector3 new_y;
	for (Vector3 direction_vector : directions) {
		new_y += direction_vector;
	}
	if (!directions.empty()) {
		new_y /= (real_t)directions.size();
		new_y.normalize();
	}

	Transform3D new_y_ray = Transform3D(Basis(), new_y);
	Quaternion old_y_to_new_y = Quaternion(p_limiting_axes->get_global_transform().get_basis().get_column(Vector3::AXIS_Y).normalized(),
			p_limiting_axes->get_global_transform().get_basis().xform(new_y_ray.origin).normalized());
	p_limiting_axes->rotate_local_with_global(old_y_to_new_y);

	for (Ref<IKLimitCone3D> open_cone : open_cones) {
		if (open_cone == nullptr) {
			continue;
		}
		Vector3 control_point = open_cone->get_control_point();
		open_cone->set_control_point(control_point.normalized());
	}
	update_tangent_radii();
}

// ik_kusudama_3d.cpp:91-101
void IKKusudama3D::update_tangent_radii() {
	for (int i = 0; i < (int)open_cones.size(); i++) {
		Ref<IKLimitCone3D> next;
		if (i < (int)open_cones.size() - 1) {
			next = open_cones[i + 1];
		}
		Ref<IKLimitCone3D> cone = open_cones[i];
		cone->update_tangent_handles(next);
	}
}

// ik_kusudama_3d.cpp:103-115
void IKKusudama3D::set_axial_limits(real_t min_angle, real_t in_range) {
	min_axial_angle = min_angle;
	range_angle = in_range;
	Vector3 y_axis = Vector3(0.0f, 1.0f, 0.0f);
	Vector3 z_axis = Vector3(0.0f, 0.0f, 1.0f);
	twist_min_rot = IKKusudama3D::get_quaternion_axis_angle(y_axis, min_axial_angle);
	twist_min_vec = twist_min_rot.xform(z_axis).normalized();
	twist_center_vec = twist_min_rot.xform(twist_min_vec).normalized();
	twist_center_rot = Quaternion(z_axis, twist_center_vec);
	twist_half_range_half_cos = Math::cos(in_range / real_t(4.0));
	twist_max_vec = IKKusudama3D::get_quaternion_axis_angle(y_axis, in_range).xform(twist_min_vec).normalized();
	twist_max_rot = Quaternion(z_axis, twist_max_vec);
}

// ik_kusudama_3d.cpp:117-132
void IKKusudama3D::set_snap_to_twist_limit(Ref<IKNode3D> p_bone_direction, Ref<IKNode3D> p_to_set, Ref<IKNode3D> p_constraint_axes, real_t p_dampening, real_t p_cos_half_dampen) {
	if (!is_axially_constrained()) {
		return;
	}
	Transform3D global_transform_constraint = p_constraint_axes->get_global_transform();
	Transform3D global_transform_to_set = p_to_set->get_global_transform();
	Basis parent_global_inverse = p_to_set->get_parent()->get_global_transform().basis.inverse();
	Basis global_twist_center = global_transform_constraint.basis * twist_center_rot;
	Basis align_rot = (global_twist_center.inverse() * global_transform_to_set.basis).orthonormalized();
	Quaternion twist_rotation, swing_rotation;
	get_swing_twist(align_rot.get_rotation_quaternion(), Vector3(0, 1, 0), swing_rotation, twist_rotation);
	twist_rotation = IKBoneSegment3D::clamp_to_cos_half_angle(twist_rotation, twist_half_range_half_cos);
	Basis recomposition = (global_twist_center * (swing_rotation * twist_rotation)).orthonormalized();
	Basis rotation = parent_global_inverse * recomposition;
	p_to_set->set_transform(Transform3D(rotation, p_to_set->get_transform().origin));
}

// ik_kusudama_3d.cpp:134-158
void IKKusudama3D::get_swing_twist(Quaternion p_rotation, Vector3 p_axis, Quaternion &r_swing, Quaternion &r_twist) {
	if (Math::is_zero_approx(p_axis.length_squared())) {
		r_swing = Quaternion();
		r_twist = Quaternion();
		return;
	}
	Quaternion rotation = p_rotation;
	if (rotation.w < real_t(0.0)) {
		rotation *= -1;
	}
	Vector3 p = p_axis * (rotation.x * p_axis.x + rotation.y * p_axis.y + rotation.z * p_axis.z);
	r_twist = Quaternion(p.x, p.y, p.z, rotation.w).normalized();
	real_t d = Vector3(r_twist.x, r_twist.y, r_twist.z).dot(p_axis);
	if (d < real_t(0.0)) {
		r_twist *= real_t(-1.0);
	}
	r_swing = (rotation * r_twist.inverse()).normalized();
}

// ik_kusudama_3d.cpp:160-166
void IKKusudama3D::add_open_cone(Ref<IKLimitCone3D> p_cone) {
	if (!p_cone || !p_cone->get_attached_to()) {
		return;
	}
	open_cones.push_back(p_cone);
	update_tangent_radii();
}

// ik_kusudama_3d.cpp:168-171
void IKKusudama3D::remove_open_cone(Ref<IKLimitCone3D> limitCone) {
	if (!limitCone) {
		return;
	}
	auto it = std::find(open_cones.begin(), open_cones.end(), limitCone);
	if (it != open_cones.end()) {
		open_cones.erase(it);
	}
}

// ik_kusudama_3d.cpp:273-332
Vector3 IKKusudama3D::get_local_point_in_limits(Vector3 in_point, std::vector<double> *in_bounds) {
	Vector3 point = in_point.normalized();
	real_t closest_cos = -2.0;
	(*in_bounds)[0] = -1;
	Vector3 closest_collision_point = in_point;

	for (int i = 0; i < (int)open_cones.size(); i++) {
		Ref<IKLimitCone3D> cone = open_cones[i];
		Vector3 collision_point = cone->closest_to_cone(point, in_bounds);
		if (Math::is_nan(collision_point.x) || Math::is_nan(collision_point.y) || Math::is_nan(collision_point.z)) {
			(*in_bounds)[0] = 1;
			return point;
		}
		real_t this_cos = collision_point.dot(point);
		if (closest_collision_point.is_zero_approx() || this_cos > closest_cos) {
			closest_collision_point = collision_point;
			closest_cos = this_cos;
		}
	}

	if ((*in_bounds)[0] == -1) {
		for (int i = 0; i < (int)open_cones.size() - 1; i++) {
			Ref<IKLimitCone3D> currCone = open_cones[i];
			Ref<IKLimitCone3D> nextCone = open_cones[i + 1];
			Vector3 collision_point = currCone->get_on_great_tangent_triangle(nextCone, point);
			if (Math::is_nan(collision_point.x)) {
				continue;
			}
			real_t this_cos = collision_point.dot(point);
			if (Math::is_equal_approx(this_cos, real_t(1.0))) {
				(*in_bounds)[0] = 1;
				return point;
			}
			if (this_cos > closest_cos) {
				closest_collision_point = collision_point;
				closest_cos = this_cos;
			}
		}
	}
	return closest_collision_point;
}

// ik_kusudama_3d.cpp:347-376
void IKKusudama3D::snap_to_orientation_limit(Ref<IKNode3D> bone_direction, Ref<IKNode3D> to_set, Ref<IKNode3D> limiting_axes, real_t p_dampening, real_t p_cos_half_angle_dampen) {
	if (!bone_direction || !to_set || !limiting_axes) {
		return;
	}
	std::vector<double> in_bounds;
	in_bounds.resize(1);
	in_bounds[0] = 1.0;
	Vector3 limiting_origin = limiting_axes->get_global_transform().origin;
	Vector3 bone_dir_xform = bone_direction->get_global_transform().xform(Vector3(0.0, 1.0, 0.0));

	bone_ray->set_point_1(limiting_origin);
	bone_ray->set_point_2(bone_dir_xform);

	Vector3 bone_tip = limiting_axes->to_local(bone_ray->get_point_2());
	Vector3 in_limits = get_local_point_in_limits(bone_tip, &in_bounds);
	n_swing_calls++;

	if (in_bounds[0] < 0) {
		n_swing_rectified++;
		constrained_ray->set_point_1(bone_ray->get_point_1());
		constrained_ray->set_point_2(limiting_axes->to_global(in_limits));
		Quaternion rectified_rot = Quaternion(bone_ray->get_heading(), constrained_ray->get_heading());
		to_set->rotate_local_with_global(rectified_rot);
	}
}

// ik_kusudama_3d.cpp:417-427 (divides by length SQUARED; identical to the engine ctor for unit axes)
Quaternion IKKusudama3D::get_quaternion_axis_angle(const Vector3 &p_axis, real_t p_angle) {
	real_t d = p_axis.length_squared();
	if (d == 0) {
		return Quaternion();
	} else {
		real_t sin_angle = Math::sin(p_angle * 0.5f);
		real_t cos_angle = Math::cos(p_angle * 0.5f);
		real_t s = sin_angle / d;
		return Quaternion(p_axis.x * s, p_axis.y * s, p_axis.z * s, cos_angle);
	}
}

// =================================================================================================
// QCP -- src/math/qcp.cpp
// =================================================================================================

// qcp.cpp:44-54
Quaternion QCP::get_rotation() {
	Quaternion result;
	if (!transformation_calculated) {
		if (!inner_product_calculated) {
			inner_product(target, moved);
		}
		result = calculate_rotation();
		transformation_calculated = true;
	}
	return result;
}

// qcp.cpp:56-127
Quaternion QCP::calculate_rotation() {
	Quaternion result;
	if (moved.size() == 1) {
		Vector3 u = moved[0];
		Vector3 v = target[0];
		double norm_product = u.length() * v.length(); // float product, widened
		if (norm_product == 0.0) {
			return Quaternion();
		}
		double dot = u.dot(v);
		if (dot < ((2.0e-15 - 1.0) * norm_product)) {
			Vector3 w = u.normalized();
			result = Quaternion(w.x, w.y, w.z, 0.0f).normalized();
		} else {
			double q0 = Math::sqrt(0.5 * (1.0 + dot / norm_product));
			double coeff = 1.0 / (2.0 * q0 * norm_product);
			Vector3 q = v.cross(u).normalized();
			result = Quaternion((real_t)(coeff * q.x), (real_t)(coeff * q.y), (real_t)(coeff * q.z), (real_t)q0).normalized();
		}
	} else {
		double a13 = -sum_xz_minus_zx;
		double a14 = sum_xy_minus_yx;
		double a21 = sum_yz_minus_zy;
		double a22 = sum_xx_minus_yy - sum_zz - max_eigenvalue;
		double a23 = sum_xy_plus_yx;
		double a24 = sum_xz_plus_zx;
		double a31 = a13;
		double a32 = a23;
		double a33 = sum_yy - sum_xx - sum_zz - max_eigenvalue;
		double a34 = sum_yz_plus_zy;
		double a41 = a14;
		double a42 = a24;
		double a43 = a34;
		double a44 = sum_zz - sum_xx_plus_yy - max_eigenvalue;

		double a3344_4334 = a33 * a44 - a43 * a34;
		double a3244_4234 = a32 * a44 - a42 * a34;
		double a3243_4233 = a32 * a43 - a42 * a33;
		double a3143_4133 = a31 * a43 - a41 * a33;
		double a3144_4134 = a31 * a44 - a41 * a34;
		double a3142_4132 = a31 * a42 - a41 * a32;

		double quaternion_w = a22 * a3344_4334 - a23 * a3244_4234 + a24 * a3243_4233;
		double quaternion_x = -a21 * a3344_4334 + a23 * a3144_4134 - a24 * a3143_4133;
		double quaternion_y = a21 * a3244_4234 - a22 * a3144_4134 + a24 * a3142_4132;
		double quaternion_z = -a21 * a3243_4233 + a22 * a3143_4133 - a23 * a3142_4132;
		double qsqr = quaternion_w * quaternion_w + quaternion_x * quaternion_x + quaternion_y * quaternion_y + quaternion_z * quaternion_z;

		if (qsqr < eigenvector_precision) {
			result = Quaternion();
		} else {
			quaternion_x *= -1;
			quaternion_y *= -1;
			quaternion_z *= -1;
			double min = quaternion_w;
			min = quaternion_x < min ? quaternion_x : min;
			min = quaternion_y < min ? quaternion_y : min;
			min = quaternion_z < min ? quaternion_z : min;
			quaternion_w /= min;
			quaternion_x /= min;
			quaternion_y /= min;
			quaternion_z /= min;
			result = Quaternion((real_t)quaternion_x, (real_t)quaternion_y, (real_t)quaternion_z, (real_t)quaternion_w).normalized();
		}
	}
	return result;
}

// qcp.cpp:129-133
void QCP::translate(Vector3 r_translate, PackedVector3Array &r_x) {
	for (Vector3 &p : r_x) {
		p += r_translate;
	}
}

// qcp.cpp:135-137
Vector3 QCP::get_translation() {
	return target_center - moved_center;
}

// qcp.cpp:139-160 (`Vector3 * double` and `Vector3 /= double` narrow the scalar to real_t)
Vector3 QCP::move_to_weighted_center(PackedVector3Array &r_to_center, std::vector<double> &r_weight) {
	Vector3 center;
	double total_weight = 0;
	bool weight_is_empty = r_weight.empty();
	int size = (int)r_to_center.size();
	for (int i = 0; i < size; i++) {
		if (!weight_is_empty) {
			total_weight += r_weight[i];
			center += r_to_center[i] * (real_t)r_weight[i];
		} else {
			center += r_to_center[i];
			total_weight++;
		}
	}
	if (total_weight > 0) {
		center /= (real_t)total_weight;
	}
	return center;
}

// qcp.cpp:162-218 : float products, double accumulators; lambda = (Gt+Gm)/2 is FINAL (no Newton)
void QCP::inner_product(PackedVector3Array &coords1, PackedVector3Array &coords2) {
	Vector3 weighted_coord1, weighted_coord2;
	double sum_of_squares1 = 0, sum_of_squares2 = 0;
	sum_xx = 0; sum_xy = 0; sum_xz = 0;
	sum_yx = 0; sum_yy = 0; sum_yz = 0;
	sum_zx = 0; sum_zy = 0; sum_zz = 0;

	bool weight_is_empty = weight.empty();
	int size = (int)coords1.size();
	for (int i = 0; i < size; i++) {
		if (!weight_is_empty) {
			weighted_coord1 = (real_t)weight[i] * coords1[i];
			sum_of_squares1 += weighted_coord1.dot(coords1[i]);
		} else {
			weighted_coord1 = coords1[i];
			sum_of_squares1 += weighted_coord1.dot(weighted_coord1);
		}
		weighted_coord2 = coords2[i];
		sum_of_squares2 += weight_is_empty ? weighted_coord2.dot(weighted_coord2) : (weight[i] * weighted_coord2.dot(weighted_coord2));

		sum_xx += (weighted_coord1.x * weighted_coord2.x);
		sum_xy += (weighted_coord1.x * weighted_coord2.y);
		sum_xz += (weighted_coord1.x * weighted_coord2.z);
		sum_yx += (weighted_coord1.y * weighted_coord2.x);
		sum_yy += (weighted_coord1.y * weighted_coord2.y);
		sum_yz += (weighted_coord1.y * weighted_coord2.z);
		sum_zx += (weighted_coord1.z * weighted_coord2.x);
		sum_zy += (weighted_coord1.z * weighted_coord2.y);
		sum_zz += (weighted_coord1.z * weighted_coord2.z);
	}

	double initial_eigenvalue = (sum_of_squares1 + sum_of_squares2) * 0.5;
	sum_xz_plus_zx = sum_xz + sum_zx;
	sum_yz_plus_zy = sum_yz + sum_zy;
	sum_xy_plus_yx = sum_xy + sum_yx;
	sum_yz_minus_zy = sum_yz - sum_zy;
	sum_xz_minus_zx = sum_xz - sum_zx;
	sum_xy_minus_yx = sum_xy - sum_yx;
	sum_xx_plus_yy = sum_xx + sum_yy;
	sum_xx_minus_yy = sum_xx - sum_yy;
	max_eigenvalue = initial_eigenvalue;
	inner_product_calculated = true;
}

// qcp.cpp:220-223
Quaternion QCP::weighted_superpose(PackedVector3Array &p_moved, PackedVector3Array &p_target, std::vector<double> &p_weight, bool translate) {
	set(p_moved, p_target, p_weight, translate);
	return get_rotation();
}

// qcp.cpp:225-248
void QCP::set(PackedVector3Array &p_moved, PackedVector3Array &p_target, std::vector<double> &p_weight, bool p_translate) {
	transformation_calculated = false;
	inner_product_calculated = false;
	moved = p_moved;
	target = p_target;
	weight = p_weight;
	if (p_translate) {
		moved_center = move_to_weighted_center(moved, weight);
		w_sum = 0;
		target_center = move_to_weighted_center(target, weight);
		translate(moved_center * -1, moved);
		translate(target_center * -1, target);
	} else {
		if (!p_weight.empty()) {
			for (size_t i = 0; i < p_weight.size(); i++) {
				w_sum += p_weight[i];
			}
		} else {
			w_sum = (double)p_moved.size();
		}
	}
}

// =================================================================================================
// Skeleton3D stand-in (engine, not in tree)
// =================================================================================================
std::vector<int> Skeleton3D::get_bone_children(int b) const {
	std::vector<int> r;
	for (int i = 0; i < (int)parent.size(); i++) {
		if (parent[i] == b) {
			r.push_back(i);
		}
	}
	return r;
}
std::vector<int> Skeleton3D::get_parentless_bones() const {
	std::vector<int> r;
	for (int i = 0; i < (int)parent.size(); i++) {
		if (parent[i] < 0) {
			r.push_back(i);
		}
	}
	return r;
}
Transform3D Skeleton3D::get_bone_global_pose(int b) const {
	if (parent[b] >= 0) {
		return get_bone_global_pose(parent[b]) * pose[b];
	}
	return pose[b];
}

// =================================================================================================
// IKEffector3D -- src/ik_effector_3d.cpp
// =================================================================================================

// ik_effector_3d.cpp:173-175 (CLAMP(x, 0.0, 1.0) evaluated in double, stored to real_t)
void IKEffector3D::set_motion_propagation_factor(float f) {
	double v = f;
	motion_propagation_factor = (real_t)(v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v));
}

// ik_effector_3d.cpp:90-116 -- note: origin taken from the EFFECTOR's own bone (`for_bone`, :97)
int32_t IKEffector3D::update_effector_target_headings(PackedVector3Array *p_headings, int32_t p_index, Ref<IKBone3D> p_for_bone, const std::vector<double> *p_weights) const {
	int32_t index = p_index;
	Vector3 bone_origin_relative_to_skeleton_origin = for_bone->get_bone_direction_global_pose().origin;
	(*p_headings)[index] = target_relative_to_skeleton_origin.origin - bone_origin_relative_to_skeleton_origin;
	index++;
	Vector3 priority = get_direction_priorities();
	for (int axis = Vector3::AXIS_X; axis <= Vector3::AXIS_Z; ++axis) {
		if (priority[axis] > 0.0) {
			real_t w = (real_t)(*p_weights)[index];
			Vector3 column = target_relative_to_skeleton_origin.basis.get_column(axis);
			(*p_headings)[index] = (column + target_relative_to_skeleton_origin.origin) - bone_origin_relative_to_skeleton_origin;
			(*p_headings)[index] *= Vector3(w, w, w);
			index++;
			(*p_headings)[index] = (target_relative_to_skeleton_origin.origin - column) - bone_origin_relative_to_skeleton_origin;
			(*p_headings)[index] *= Vector3(w, w, w);
			index++;
		}
	}
	return index;
}

// ik_effector_3d.cpp:118-149 -- origin taken from the bone being SOLVED (`p_for_bone`, :125)
int32_t IKEffector3D::update_effector_tip_headings(PackedVector3Array *p_headings, int32_t p_index, Ref<IKBone3D> p_for_bone) const {
	Transform3D tip_xform_relative_to_skeleton_origin = for_bone->get_bone_direction_global_pose();
	Basis tip_basis = tip_xform_relative_to_skeleton_origin.basis;
	Vector3 bone_origin_relative_to_skeleton_origin = p_for_bone->get_bone_direction_global_pose().origin;

	int32_t index = p_index;
	(*p_headings)[index] = tip_xform_relative_to_skeleton_origin.origin - bone_origin_relative_to_skeleton_origin;
	index++;
	double distance = target_relative_to_skeleton_origin.origin.distance_to(bone_origin_relative_to_skeleton_origin);
	double scale_by = distance < 1.0f ? distance : 1.0f; // MIN(distance, 1.0f)
	const Vector3 priority = get_direction_priorities();
	for (int axis = Vector3::AXIS_X; axis <= Vector3::AXIS_Z; ++axis) {
		if (priority[axis] > 0.0) {
			Vector3 column = tip_basis.get_column(axis) * priority[axis];
			(*p_headings)[index] = (column + tip_xform_relative_to_skeleton_origin.origin) - bone_origin_relative_to_skeleton_origin;
			(*p_headings)[index] *= (real_t)scale_by;
			index++;
			(*p_headings)[index] = (tip_xform_relative_to_skeleton_origin.origin - column) - bone_origin_relative_to_skeleton_origin;
			(*p_headings)[index] *= (real_t)scale_by;
			index++;
		}
	}
	return index;
}

// =================================================================================================
// IKBone3D -- src/ik_bone_3d.cpp
// =================================================================================================

// ik_bone_3d.cpp:198-245.  The stiffness / "returnfulness" tables (:225-244) are computed there but
// never read by the solve, so only `dampening` is kept for the record.
IKBone3D::IKBone3D(BoneId p_bone, Skeleton3D *p_skeleton, const Ref<IKBone3D> &p_parent, std::vector<IKEffectorTemplate3D> &p_pins, float p_default_dampening, ManyBoneIK3D *p_many_bone_ik) {
	(void)p_skeleton;
	(void)p_many_bone_ik;
	(void)p_parent; // set_parent needs shared_from_this: done by init_parent() right after construction
	default_dampening = p_default_dampening;
	cos_half_dampen = cos(default_dampening / real_t(2.0));
	bone_id = p_bone;
	for (size_t i = 0; i < p_pins.size(); i++) {
		IKEffectorTemplate3D &elem = p_pins[i];
		if (elem.bone == p_bone && p_bone >= 0) {
			pin = Ref<IKEffector3D>(new IKEffector3D());
			pin->for_bone = this;
			pin->set_motion_propagation_factor(elem.motion_propagation_factor);
			pin->set_weight(elem.weight);
			pin->set_direction_priorities(elem.priority_direction);
			pin->pin_index = (int)i;
			break;
		}
	}
	// :224 bone_direction_transform->set_parent(godot_skeleton_aligned_transform) -- also in init_parent()
	constraint = Ref<IKKusudama3D>(new IKKusudama3D()); // :229-233 default (unconstrained) kusudama
}

// ik_bone_3d.cpp:46-55
void IKBone3D::set_parent(const Ref<IKBone3D> &p_parent) {
	if (!p_parent) {
		return;
	}
	parent_w = p_parent;
	p_parent->children.push_back(shared_from_this());
	godot_skeleton_aligned_transform->set_parent(p_parent->godot_skeleton_aligned_transform);
	constraint_orientation_transform->set_parent(p_parent->godot_skeleton_aligned_transform);
	constraint_twist_transform->set_parent(p_parent->godot_skeleton_aligned_transform);
}

// ik_bone_3d.cpp:57-93
void IKBone3D::update_default_bone_direction_transform(Skeleton3D *p_skeleton) {
	Vector3 child_centroid;
	int child_count = 0;
	for (Ref<IKBone3D> &ik_bone : children) {
		child_centroid += ik_bone->get_ik_transform()->get_global_transform().origin;
		child_count++;
	}
	if (child_count > 0) {
		child_centroid /= (real_t)child_count;
	} else {
		const std::vector<int> bone_children = p_skeleton->get_bone_children(bone_id);
		for (BoneId child_bone_idx : bone_children) {
			child_centroid += p_skeleton->get_bone_global_pose(child_bone_idx).origin;
		}
		child_centroid /= (real_t)bone_children.size();
	}
	const Vector3 godot_bone_origin = godot_skeleton_aligned_transform->get_global_transform().origin;
	child_centroid -= godot_bone_origin;

	Ref<IKBone3D> parent = get_parent();
	if (Math::is_zero_approx(child_centroid.length_squared())) {
		if (parent) {
			child_centroid = parent->get_bone_direction_transform()->get_global_transform().basis.get_column(Vector3::AXIS_Y);
		} else {
			child_centroid = get_bone_direction_transform()->get_global_transform().basis.get_column(Vector3::AXIS_Y);
		}
	}
	if (!Math::is_zero_approx(child_centroid.length_squared()) && (children.size() || p_skeleton->get_bone_children(bone_id).size())) {
		child_centroid.normalize();
		Vector3 bone_direction = bone_direction_transform->get_global_transform().basis.get_column(Vector3::AXIS_Y);
		bone_direction.normalize();
		bone_direction_transform->rotate_local_with_global(Quaternion(child_centroid, bone_direction));
	}
}

// ik_bone_3d.cpp:145-151
void IKBone3D::set_global_pose(const Transform3D &p_transform) {
	godot_skeleton_aligned_transform->set_global_transform(p_transform);
	Transform3D transform = constraint_orientation_transform->get_transform();
	transform.origin = godot_skeleton_aligned_transform->get_transform().origin;
	constraint_orientation_transform->set_transform(transform);
	constraint_orientation_transform->_propagate_transform_changed();
}

// ik_bone_3d.cpp:161-168
void IKBone3D::set_initial_pose(Skeleton3D *p_skeleton) {
	if (bone_id == -1) {
		return;
	}
	Transform3D bone_origin_to_parent_origin = p_skeleton->get_bone_pose(bone_id);
	set_pose(bone_origin_to_parent_origin);
}

// =================================================================================================
// IKBoneSegment3D -- src/ik_bone_segment_3d.cpp
// =================================================================================================

// ik_bone_segment_3d.cpp:56-72
void IKBoneSegment3D::create_bone_list(std::vector<Ref<IKBone3D>> &p_list, bool p_recursive) const {
	if (p_recursive) {
		for (size_t child_i = 0; child_i < child_segments.size(); child_i++) {
			child_segments[child_i]->create_bone_list(p_list, p_recursive);
		}
	}
	Ref<IKBone3D> current_bone = tip;
	std::vector<Ref<IKBone3D>> list;
	while (current_bone) {
		list.push_back(current_bone);
		if (current_bone == root) {
			break;
		}
		current_bone = current_bone->get_parent();
	}
	p_list.insert(p_list.end(), list.begin(), list.end());
}

// ik_bone_segment_3d.cpp:74-88
void IKBoneSegment3D::update_pinned_list(std::vector<std::vector<double>> &r_weights) {
	for (size_t chain_i = 0; chain_i < child_segments.size(); chain_i++) {
		Ref<IKBoneSegment3D> chain = child_segments[chain_i];
		chain->update_pinned_list(r_weights);
	}
	if (is_pinned()) {
		effector_list.push_back(tip->get_pin());
	}
	double motion_propagation_factor = is_pinned() ? tip->get_pin()->motion_propagation_factor : 1.0;
	if (motion_propagation_factor > 0.0) {
		for (Ref<IKBoneSegment3D> child : child_segments) {
			effector_list.insert(effector_list.end(), child->effector_list.begin(), child->effector_list.end());
		}
	}
}

// ik_bone_segment_3d.cpp:90-95 (iteration arguments are NOT forwarded)
void IKBoneSegment3D::_update_optimal_rotation(Ref<IKBone3D> p_for_bone, double p_damp, bool p_translate, bool p_constraint_mode, int32_t current_iteration, int32_t total_iterations) {
	(void)current_iteration;
	(void)total_iterations;
	_update_target_headings(p_for_bone, &heading_weights, &target_headings);
	_update_tip_headings(p_for_bone, &tip_headings);
	_set_optimal_rotation(p_for_bone, &tip_headings, &target_headings, &heading_weights, (float)p_damp, p_translate, p_constraint_mode);
}

// ik_bone_segment_3d.cpp:97-112
Quaternion IKBoneSegment3D::clamp_to_cos_half_angle(Quaternion p_quat, double p_cos_half_angle) {
	if (p_quat.w < 0.0) {
		p_quat = p_quat * -1;
	}
	double previous_coefficient = (1.0 - (p_quat.w * p_quat.w)); // float product, widened
	if (p_cos_half_angle <= p_quat.w || previous_coefficient == 0.0) {
		return p_quat;
	} else {
		double composite_coefficient = Math::sqrt((1.0 - (p_cos_half_angle * p_cos_half_angle)) / previous_coefficient);
		p_quat.w = (real_t)p_cos_half_angle;
		p_quat.x = (real_t)(p_quat.x * composite_coefficient); // float *= double: product in double, narrowed
		p_quat.y = (real_t)(p_quat.y * composite_coefficient);
		p_quat.z = (real_t)(p_quat.z * composite_coefficient);
	}
	return p_quat;
}

// ik_bone_segment_3d.cpp:114-127
float IKBoneSegment3D::_get_manual_msd(const PackedVector3Array &r_htip, const PackedVector3Array &r_htarget, const std::vector<double> &p_weights) {
	float manual_RMSD = 0.0f;
	float w_sum = 0.0f;
	for (size_t i = 0; i < r_htarget.size(); i++) {
		float x_d = r_htarget[i].x - r_htip[i].x;
		float y_d = r_htarget[i].y - r_htip[i].y;
		float z_d = r_htarget[i].z - r_htip[i].z;
		float mag_sq = (float)(p_weights[i] * (x_d * x_d + y_d * y_d + z_d * z_d));
		manual_RMSD += mag_sq;
		w_sum = (float)(w_sum + p_weights[i]);
	}
	manual_RMSD /= w_sum * w_sum;
	return manual_RMSD;
}

static void trace_stage(const char *stage, Ref<IKBone3D> b, long step) {
	const char *e = getenv("ORC_TRACE_BONE");
	if (!e || atoi(e) != b->get_bone_id()) {
		return;
	}
	Transform3D l = b->get_pose();
	Transform3D g = b->get_global_pose();
	Transform3D pg = b->get_parent() ? b->get_parent()->get_global_pose() : Transform3D();
	fprintf(stderr, "[bone %d step %ld] %-6s local det %.6f global det %.6f parent det %.6f | local row0 (%g %g %g)\n", b->get_bone_id(), step, stage, l.basis.determinant(), g.basis.determinant(), pg.basis.determinant(), l.basis.rows[0].x, l.basis.rows[0].y, l.basis.rows[0].z);
}

// ik_bone_segment_3d.cpp:129-181
void IKBoneSegment3D::_set_optimal_rotation(Ref<IKBone3D> p_for_bone, PackedVector3Array *r_htip, PackedVector3Array *r_htarget, std::vector<double> *r_weights, float p_dampening, bool p_translate, bool p_constraint_mode, double current_iteration, double total_iterations) {
	n_bone_steps++;
	_update_target_headings(p_for_bone, &heading_weights, &target_headings);
	Transform3D prev_transform = p_for_bone->get_pose();
	bool got_closer = true;
	double bone_damp = p_for_bone->get_cos_half_dampen();
	int i = 0;
	do {
		_update_tip_headings(p_for_bone, &tip_headings);
		if (!p_constraint_mode) {
			QCP qcp = QCP(evec_prec);
			Basis rotation = qcp.weighted_superpose(*r_htip, *r_htarget, *r_weights, p_translate);
			Vector3 translation = qcp.get_translation();
			double dampening = (p_dampening != -1.0) ? p_dampening : bone_damp;
			rotation = clamp_to_cos_half_angle(rotation.get_rotation_quaternion(), cos(dampening / 2.0));
			if (current_iteration == 0) {
				current_iteration = 0.0001;
			}
			rotation = rotation.slerp(p_for_bone->get_global_pose().basis, (real_t)(static_cast<double>(total_iterations) / current_iteration));
			if (getenv("ORC_TRACE") && (!rotation.is_finite() || !translation.is_finite())) {
				static int once = 0;
				if (!once++) {
					fprintf(stderr, "[orc trace] first non-finite rotation at bone %d step %ld translate=%d H=%zu\n", p_for_bone->get_bone_id(), n_bone_steps, (int)p_translate, r_htip->size());
					for (size_t h = 0; h < r_htip->size(); h++) {
						fprintf(stderr, "  h%zu w=%g tip=(%g %g %g) target=(%g %g %g)\n", h, (*r_weights)[h], (*r_htip)[h].x, (*r_htip)[h].y, (*r_htip)[h].z, (*r_htarget)[h].x, (*r_htarget)[h].y, (*r_htarget)[h].z);
					}
					Transform3D g = p_for_bone->get_global_pose();
					fprintf(stderr, "  global basis row0=(%g %g %g) origin=(%g %g %g)\n", g.basis.rows[0].x, g.basis.rows[0].y, g.basis.rows[0].z, g.origin.x, g.origin.y, g.origin.z);
				}
			}
			p_for_bone->get_ik_transform()->rotate_local_with_global(rotation);
			Transform3D result = Transform3D(p_for_bone->get_global_pose().basis, p_for_bone->get_global_pose().origin + translation);
			p_for_bone->set_global_pose(result);
			trace_stage("qcp", p_for_bone, n_bone_steps);
		}
		bool is_parent_valid = (bool)p_for_bone->get_parent();
		static int trace_once = 0;
		bool tr = getenv("ORC_TRACE") != nullptr && !trace_once;
		real_t det0 = tr ? p_for_bone->get_global_pose().basis.determinant() : 1.0f;
		if (tr && !(Math::abs(det0 - 1.0f) < 1e-2f)) {
			trace_once = 1;
			fprintf(stderr, "[orc trace] det %g after QCP at bone %d step %ld\n", det0, p_for_bone->get_bone_id(), n_bone_steps);
		}
		if (is_parent_valid && p_for_bone->is_orientationally_constrained()) {
			p_for_bone->get_constraint()->snap_to_orientation_limit(p_for_bone->get_bone_direction_transform(), p_for_bone->get_ik_transform(), p_for_bone->get_constraint_orientation_transform(), (real_t)bone_damp, p_for_bone->get_cos_half_dampen());
			trace_stage("swing", p_for_bone, n_bone_steps);
		}
		if (is_parent_valid && p_for_bone->is_axially_constrained()) {
			p_for_bone->get_constraint()->set_snap_to_twist_limit(p_for_bone->get_bone_direction_transform(), p_for_bone->get_ik_transform(), p_for_bone->get_constraint_twist_transform(), (real_t)bone_damp, p_for_bone->get_cos_half_dampen());
			trace_stage("twist", p_for_bone, n_bone_steps);
		}
		if (tr && !trace_once) {
			real_t det1 = p_for_bone->get_global_pose().basis.determinant();
			if (!(Math::abs(det1 - 1.0f) < 1e-2f)) {
				trace_once = 1;
				Transform3D l = p_for_bone->get_pose();
				Transform3D pg = p_for_bone->get_parent() ? p_for_bone->get_parent()->get_global_pose() : Transform3D();
				fprintf(stderr, "[orc trace] det %g (was %g) after snaps at bone %d step %ld; local det %g parent-global det %g\n", det1, det0, p_for_bone->get_bone_id(), n_bone_steps, l.basis.determinant(), pg.basis.determinant());
			}
		}
		if (default_stabilizing_pass_count > 0) {
			_update_tip_headings(p_for_bone, &tip_headings_uniform);
			double current_msd = _get_manual_msd(tip_headings_uniform, target_headings, heading_weights);
			if (current_msd <= previous_deviation * 1.0001) {
				previous_deviation = current_msd;
				got_closer = true;
				break;
			} else {
				got_closer = false;
				p_for_bone->set_pose(prev_transform);
			}
		}
		i++;
	} while (i < default_stabilizing_pass_count && !got_closer);

	if (root == p_for_bone) {
		previous_deviation = INFINITY;
	}
}

// ik_bone_segment_3d.cpp:183-195
void IKBoneSegment3D::_update_target_headings(Ref<IKBone3D> p_for_bone, std::vector<double> *r_weights, PackedVector3Array *r_target_headings) {
	(void)r_weights;
	int32_t last_index = 0;
	for (size_t effector_i = 0; effector_i < effector_list.size(); effector_i++) {
		Ref<IKEffector3D> effector = effector_list[effector_i];
		if (!effector) {
			continue;
		}
		last_index = effector->update_effector_target_headings(r_target_headings, last_index, p_for_bone, &heading_weights);
	}
}

// ik_bone_segment_3d.cpp:197-208
void IKBoneSegment3D::_update_tip_headings(Ref<IKBone3D> p_for_bone, PackedVector3Array *r_heading_tip) {
	int32_t last_index = 0;
	for (size_t effector_i = 0; effector_i < effector_list.size(); effector_i++) {
		Ref<IKEffector3D> effector = effector_list[effector_i];
		if (!effector) {
			continue;
		}
		last_index = effector->update_effector_tip_headings(r_heading_tip, last_index, p_for_bone);
	}
}

// ik_bone_segment_3d.cpp:210-225
void IKBoneSegment3D::segment_solver(const std::vector<float> &p_damp, float p_default_damp, bool p_constraint_mode, int32_t p_current_iteration, int32_t p_total_iteration) {
	for (Ref<IKBoneSegment3D> child : child_segments) {
		if (!child) {
			continue;
		}
		child->segment_solver(p_damp, p_default_damp, p_constraint_mode, p_current_iteration, p_total_iteration);
	}
	bool is_translate = parent_segment.expired();
	if (is_translate) {
		std::vector<float> damp = p_damp;
		std::fill(damp.begin(), damp.end(), (float)Math_PI);
		_qcp_solver(damp, (float)Math_PI, is_translate, p_constraint_mode, p_current_iteration, p_total_iteration);
		return;
	}
	_qcp_solver(p_damp, p_default_damp, is_translate, p_constraint_mode, p_current_iteration, p_total_iteration);
}

// ik_bone_segment_3d.cpp:227-240
void IKBoneSegment3D::_qcp_solver(const std::vector<float> &p_damp, float p_default_damp, bool p_translate, bool p_constraint_mode, int32_t p_current_iteration, int32_t p_total_iterations) {
	for (Ref<IKBone3D> current_bone : bones) {
		float damp = p_default_damp;
		bool is_valid_access = !((int)p_damp.size() < 0 || (current_bone->get_bone_id()) >= (int)(p_damp.size()));
		if (is_valid_access) {
			damp = p_damp[current_bone->get_bone_id()];
		}
		bool is_non_default_damp = p_default_damp < damp;
		if (is_non_default_damp) {
			damp = p_default_damp;
		}
		_update_optimal_rotation(current_bone, damp, p_translate, p_constraint_mode, p_current_iteration, p_total_iterations);
	}
}

// ik_bone_segment_3d.cpp:247-264.  The reference passes the parent *segment* where IKBone3D's ctor
// expects a parent *bone*; Ref<> cross-casting yields null there, so the root bone is parented only
// by the explicit set_parent at :261.
IKBoneSegment3D::IKBoneSegment3D(Skeleton3D *p_skeleton, BoneId p_root_bone_name, std::vector<IKEffectorTemplate3D> &p_pins, ManyBoneIK3D *p_many_bone_ik, const Ref<IKBoneSegment3D> &p_parent, BoneId p_root, BoneId p_tip, int32_t p_stabilizing_pass_count) {
	(void)p_root;
	(void)p_tip;
	skeleton = p_skeleton;
	root = Ref<IKBone3D>(new IKBone3D(p_root_bone_name, p_skeleton, nullptr, p_pins, (float)Math_PI, p_many_bone_ik));
	root->bone_direction_transform->set_parent(root->godot_skeleton_aligned_transform); // ik_bone_3d.cpp:224
	if (p_parent) {
		root_segment = p_parent->root_segment;
	} else {
		root_segment = this;
	}
	root_segment->bone_map[root->get_bone_id()] = root;
	default_stabilizing_pass_count = p_stabilizing_pass_count;
}
void IKBoneSegment3D::post_construct(const Ref<IKBoneSegment3D> &p_parent) {
	if (p_parent) {
		parent_segment = p_parent;
		root->set_parent(p_parent->get_tip());
	}
}

// ik_bone_segment_3d.cpp:281-307
void IKBoneSegment3D::create_headings_arrays() {
	std::vector<std::vector<double>> penalty_array;
	std::vector<Ref<IKBone3D>> new_pinned_bones;
	recursive_create_penalty_array(shared_from_this(), penalty_array, new_pinned_bones, 1.0);
	pinned_bones = new_pinned_bones;
	int32_t total_headings = 0;
	for (const std::vector<double> &current_penalty_array : penalty_array) {
		total_headings += (int32_t)current_penalty_array.size();
	}
	target_headings.assign(total_headings, Vector3());
	tip_headings.assign(total_headings, Vector3());
	tip_headings_uniform.assign(total_headings, Vector3());
	heading_weights.assign(total_headings, 0.0);
	int currentHeading = 0;
	for (const std::vector<double> &current_penalty_array : penalty_array) {
		for (double ad : current_penalty_array) {
			heading_weights[currentHeading] = ad;
			currentHeading++;
		}
	}
}

// ik_bone_segment_3d.cpp:309-343
void IKBoneSegment3D::recursive_create_penalty_array(Ref<IKBoneSegment3D> p_bone_segment, std::vector<std::vector<double>> &r_penalty_array, std::vector<Ref<IKBone3D>> &r_pinned_bones, double p_falloff) {
	if (p_falloff <= 0.0) {
		return;
	}
	double current_falloff = 1.0;
	if (p_bone_segment->is_pinned()) {
		Ref<IKBone3D> current_tip = p_bone_segment->get_tip();
		Ref<IKEffector3D> pin = current_tip->get_pin();
		double weight = pin->get_weight();
		std::vector<double> inner_weight_array;
		inner_weight_array.push_back(weight * p_falloff);

		Vector3 pr = pin->get_direction_priorities();
		double max_pin_weight = std::max(std::max(pr.x, pr.y), pr.z); // MAX on real_t, widened
		max_pin_weight = max_pin_weight == 0.0 ? 1.0 : max_pin_weight;
		for (int i = 0; i < 3; ++i) {
			double priority = pr[i];
			if (priority > 0.0) {
				double sub_target_weight = weight * (priority / max_pin_weight) * p_falloff;
				inner_weight_array.push_back(sub_target_weight);
				inner_weight_array.push_back(sub_target_weight);
			}
		}
		r_penalty_array.push_back(inner_weight_array);
		r_pinned_bones.push_back(current_tip);
		current_falloff = pin->get_motion_propagation_factor();
	}
	for (Ref<IKBoneSegment3D> s : p_bone_segment->get_child_segments()) {
		recursive_create_penalty_array(s, r_penalty_array, r_pinned_bones, p_falloff * current_falloff);
	}
}

// ik_bone_segment_3d.cpp:345-350
void IKBoneSegment3D::recursive_create_headings_arrays_for(Ref<IKBoneSegment3D> p_bone_segment) {
	p_bone_segment->create_headings_arrays();
	for (Ref<IKBoneSegment3D> segments : p_bone_segment->get_child_segments()) {
		recursive_create_headings_arrays_for(segments);
	}
}

// ik_bone_segment_3d.cpp:352-369
void IKBoneSegment3D::generate_default_segments(std::vector<IKEffectorTemplate3D> &p_pins, BoneId p_root_bone, BoneId p_tip_bone, ManyBoneIK3D *p_many_bone_ik) {
	Ref<IKBone3D> current_tip = root;
	std::vector<BoneId> children;
	while (!_is_parent_of_tip(current_tip, p_tip_bone)) {
		children = skeleton->get_bone_children(current_tip->get_bone_id());
		if (children.empty() || _has_multiple_children_or_pinned(children, current_tip)) {
			_process_children(children, current_tip, p_pins, p_root_bone, p_tip_bone, p_many_bone_ik);
			break;
		} else {
			current_tip = _create_next_bone(children[0], current_tip, p_pins, p_many_bone_ik);
		}
	}
	_finalize_segment(current_tip);
}

// ik_bone_segment_3d.cpp:371-373
bool IKBoneSegment3D::_is_parent_of_tip(Ref<IKBone3D> p_current_tip, BoneId p_tip_bone) {
	return skeleton->get_bone_parent(p_current_tip->get_bone_id()) >= p_tip_bone && p_tip_bone != -1;
}

// ik_bone_segment_3d.cpp:375-377
bool IKBoneSegment3D::_has_multiple_children_or_pinned(std::vector<BoneId> &r_children, Ref<IKBone3D> p_current_tip) {
	return r_children.size() > 1 || p_current_tip->is_pinned();
}

// ik_bone_segment_3d.cpp:379-395 (+ _create_child_segment :397-399)
void IKBoneSegment3D::_process_children(std::vector<BoneId> &r_children, Ref<IKBone3D> p_current_tip, std::vector<IKEffectorTemplate3D> &r_pins, BoneId p_root_bone, BoneId p_tip_bone, ManyBoneIK3D *p_many_bone_ik) {
	tip = p_current_tip;
	Ref<IKBoneSegment3D> parent = shared_from_this();
	for (size_t child_i = 0; child_i < r_children.size(); child_i++) {
		BoneId child_bone = r_children[child_i];
		Ref<IKBoneSegment3D> child_segment(new IKBoneSegment3D(skeleton, child_bone, r_pins, p_many_bone_ik, parent, p_root_bone, p_tip_bone));
		child_segment->post_construct(parent);
		child_segment->generate_default_segments(r_pins, p_root_bone, p_tip_bone, p_many_bone_ik);
		if (child_segment->pinned_descendants) {
			pinned_descendants = true;
			child_segments.push_back(child_segment);
		}
	}
}

// ik_bone_segment_3d.cpp:409-427
void IKBoneSegment3D::_finalize_segment(Ref<IKBone3D> p_current_tip) {
	tip = p_current_tip;
	if (tip->is_pinned()) {
		pinned_descendants = true;
	}
	bones.clear();
	create_bone_list(bones, false);
}

// =================================================================================================
// ManyBoneIK3D -- src/many_bone_ik_3d.cpp
// =================================================================================================

// ik_bone_segment_3d.cpp:401-407 (defined here because it needs ManyBoneIK3D::get_default_damp)
Ref<IKBone3D> IKBoneSegment3D::_create_next_bone(BoneId p_bone_id, Ref<IKBone3D> p_current_tip, std::vector<IKEffectorTemplate3D> &p_pins, ManyBoneIK3D *p_many_bone_ik) {
	Ref<IKBone3D> next_bone(new IKBone3D(p_bone_id, skeleton, p_current_tip, p_pins, p_many_bone_ik->get_default_damp(), p_many_bone_ik));
	next_bone->init_parent(p_current_tip); // ik_bone_3d.cpp:206-208
	next_bone->bone_direction_transform->set_parent(next_bone->godot_skeleton_aligned_transform); // ik_bone_3d.cpp:224
	root_segment->bone_map[p_bone_id] = next_bone;
	return next_bone;
}

// many_bone_ik_3d.cpp:91-102 (root first: bone_list is walked backwards)
void ManyBoneIK3D::_update_ik_bones_transform() {
	for (int32_t bone_i = (int32_t)bone_list.size(); bone_i-- > 0;) {
		Ref<IKBone3D> bone = bone_list[bone_i];
		if (!bone) {
			continue;
		}
		bone->set_initial_pose(get_skeleton());
		if (bone->is_pinned()) {
			// IKEffector3D::update_target_global_transform (ik_effector_3d.cpp:77-84): the scene-node lookup is
			// the boundary input; the skeleton-space target arrives in pin_targets.
			int pi = bone->get_pin()->pin_index;
			if (pi >= 0 && pi < (int)pin_targets.size()) {
				bone->get_pin()->target_relative_to_skeleton_origin = pin_targets[pi];
			}
		}
	}
}

// many_bone_ik_3d.cpp:1011-1068
void ManyBoneIK3D::_bone_list_changed() {
	Skeleton3D *skeleton = get_skeleton();
	std::vector<int32_t> roots = skeleton->get_parentless_bones();
	if (roots.empty()) {
		return;
	}
	bone_list.clear();
	segmented_skeletons.clear();
	for (BoneId root_bone_index : roots) {
		Ref<IKBoneSegment3D> segmented_skeleton(new IKBoneSegment3D(skeleton, root_bone_index, pins, this, nullptr, root_bone_index, -1, stabilize_passes));
		segmented_skeleton->post_construct(nullptr);
		// `ik_origin.instantiate()` drops the previous root's origin node: its destructor un-parents the children
		// (ik_node_3d.cpp:146-158), so with several parentless bones only the LAST root keeps an IKNode3D parent.
		if (ik_origin) {
			ik_origin->cleanup();
		}
		ik_origin = Ref<IKNode3D>(new IKNode3D());
		segmented_skeleton->get_root()->get_ik_transform()->set_parent(ik_origin);
		segmented_skeleton->generate_default_segments(pins, root_bone_index, -1, this);
		std::vector<Ref<IKBone3D>> new_bone_list;
		segmented_skeleton->create_bone_list(new_bone_list, true);
		bone_list.insert(bone_list.end(), new_bone_list.begin(), new_bone_list.end());
		std::vector<std::vector<double>> weight_array;
		segmented_skeleton->update_pinned_list(weight_array);
		IKBoneSegment3D::recursive_create_headings_arrays_for(segmented_skeleton);
		segmented_skeletons.push_back(segmented_skeleton);
	}
	_update_ik_bones_transform();
	for (Ref<IKBone3D> &ik_bone_3d : bone_list) {
		ik_bone_3d->update_default_bone_direction_transform(skeleton);
	}
	for (int constraint_i = 0; constraint_i < constraint_count; ++constraint_i) {
		BoneId bone_id = constraint_names[constraint_i];
		for (Ref<IKBone3D> &ik_bone_3d : bone_list) {
			if (ik_bone_3d->get_bone_id() != bone_id) {
				continue;
			}
			Ref<IKKusudama3D> constraint(new IKKusudama3D());
			constraint->enable_orientational_limits();
			int32_t cone_count = kusudama_open_cone_count[constraint_i];
			const std::vector<Vector4f> &cones = kusudama_open_cones[constraint_i];
			for (int32_t cone_i = 0; cone_i < cone_count; ++cone_i) {
				const Vector4f &cone = cones[cone_i];
				Ref<IKLimitCone3D> new_cone(new IKLimitCone3D());
				new_cone->set_attached_to(constraint);
				new_cone->set_radius(std::max(1.0e-38, (double)cone.w)); // MAX(1.0e-38, cone.w)
				new_cone->set_control_point(Vector3(cone.x, cone.y, cone.z).normalized());
				constraint->add_open_cone(new_cone);
			}
			constraint->enable_axial_limits();
			constraint->set_axial_limits(joint_twist_x[constraint_i], joint_twist_y[constraint_i]);
			ik_bone_3d->add_constraint(constraint);
			constraint->_update_constraint(ik_bone_3d->get_constraint_twist_transform());
			break;
		}
	}
}

// many_bone_ik_3d.cpp:685-692
void ManyBoneIK3D::solve_iterations() {
	for (int32_t i = 0; i < get_iterations_per_frame(); i++) {
		for (Ref<IKBoneSegment3D> segmented_skeleton : segmented_skeletons) {
			if (!segmented_skeleton) {
				continue;
			}
			segmented_skeleton->segment_solver(bone_damp, get_default_damp(), get_constraint_mode(), i, (int32_t)get_iterations_per_frame());
		}
	}
}

// many_bone_ik_3d.cpp:104-116 -> ik_bone_3d.cpp:170-179
void ManyBoneIK3D::write_skeleton_pose(float *out10, float *out_local12, uint32_t *status) {
	Skeleton3D *sk = get_skeleton();
	int nb = sk->get_bone_count();
	std::vector<Transform3D> local(nb);
	std::vector<char> solved(nb, 0);
	for (int b = 0; b < nb; b++) {
		local[b] = sk->pose[b];
	}
	uint32_t st = 0;
	for (int32_t bone_i = (int32_t)bone_list.size(); bone_i-- > 0;) {
		Ref<IKBone3D> bone = bone_list[bone_i];
		if (!bone || bone->get_bone_id() == -1) {
			continue;
		}
		local[bone->get_bone_id()] = bone->get_pose();
		solved[bone->get_bone_id()] = 1;
	}
	for (int b = 0; b < nb; b++) {
		Transform3D bone_to_parent = local[b];
		if (out_local12) {
			float *o = out_local12 + b * 12;
			for (int r = 0; r < 3; r++) {
				for (int c = 0; c < 3; c++) {
					o[r * 3 + c] = bone_to_parent.basis.rows[r][c];
				}
			}
			o[9] = bone_to_parent.origin.x;
			o[10] = bone_to_parent.origin.y;
			o[11] = bone_to_parent.origin.z;
		}
		if (!bone_to_parent.basis.is_finite()) {
			bone_to_parent.basis = Basis();
			st |= 1u;
		}
		Quaternion q = bone_to_parent.basis.get_rotation_quaternion();
		Vector3 s = bone_to_parent.basis.get_scale();
		float *o = out10 + b * 10;
		o[0] = bone_to_parent.origin.x;
		o[1] = bone_to_parent.origin.y;
		o[2] = bone_to_parent.origin.z;
		o[3] = q.x;
		o[4] = q.y;
		o[5] = q.z;
		o[6] = q.w;
		o[7] = s.x;
		o[8] = s.y;
		o[9] = s.z;
	}
	if (status) {
		*status = st;
	}
}

} // namespace orc
