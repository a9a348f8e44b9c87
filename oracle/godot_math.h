// TEST INFRASTRUCTURE ONLY -- CPU oracle for the ManyBoneIK solve loop.
// Nothing under oracle/ may be imported, linked or executed by the product path
// (many_bone_ik_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, as the checker.
//
// L0 shim: a restatement of the subset of Godot Engine `core/math` (Vector3, Basis, Quaternion,
// Transform3D, Math::*) that the reference module calls on its solve path.  The engine is an
// UN-VENDORED dependency of /root/reference (no submodule, no version pin; lower bound Godot 4.3
// because the module derives from SkeletonModifier3D, reference src/many_bone_ik_3d.h:42,46), so
// this file restates the published engine algorithms from knowledge of Godot 4.3/4.4
// (core/math/{vector3,basis,quaternion,transform_3d}.{h,cpp}, math_funcs.h) with
//   real_t = float, CMP_EPSILON = 1e-5, UNIT_EPSILON = 1e-3, MATH_CHECKS off (release template).
// PARITY UNPINNED for everything here that the reference's own tests do not touch (see
// SURVEY.md section 8c); pinned pieces: Quaternion::xform + QCP (tests/test_qcp.h:40-113),
// identity/translation Transform3D round trips (tests/test_ik_node_3d.h), one axis-angle rotation
// (tests/test_ik_kusudama_3d.h:127-156) -- asserted by the reference's own 15 doctest cases, which
// run unmodified on the reference's own code over THIS file (oracle/_ref/ref_doctests, 52 checks).
// This header is also the core/math layer of the engine stand-in (oracle/godot_shim/) that the
// reference's sources are compiled against, so the module logic above it is the real thing there.
//
// Evaluation order matters: every expression below is written in the engine's operand order so
// that, compiled with -ffp-contract=off, it is a deterministic IEEE-754 sequence the CUDA kernel
// can reproduce bit for bit.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

namespace gd {

typedef float real_t;
static constexpr double CMP_EPSILON = 0.00001;
static constexpr double UNIT_EPSILON = 0.001;
static constexpr double Math_PI = 3.1415926535897932384626433833;
static constexpr double Math_TAU = 6.2831853071795864769252867666;

#ifndef GD_ARC_LEGACY
#define GD_ARC_LEGACY 0 // 0: Godot >= 4.3 shortest-arc ctor (normalises, parallel short-circuit); 1: 4.0-4.2 ctor
#endif

namespace Math {
inline float sqrt(float x) { return ::sqrtf(x); }
inline double sqrt(double x) { return ::sqrt(x); }
inline float sin(float x) { return ::sinf(x); }
inline double sin(double x) { return ::sin(x); }
inline float cos(float x) { return ::cosf(x); }
inline double cos(double x) { return ::cos(x); }
inline float acos(float x) { return (x < -1.0f) ? (float)Math_PI : (x > 1.0f ? 0.0f : ::acosf(x)); }
inline double acos(double x) { return (x < -1.0) ? Math_PI : (x > 1.0 ? 0.0 : ::acos(x)); }
inline float abs(float x) { return ::fabsf(x); }
inline double abs(double x) { return ::fabs(x); }
inline bool is_nan(float x) { return x != x; }
inline bool is_nan(double x) { return x != x; }
inline bool is_finite(float x) { return std::isfinite(x); }
inline float lerp(float a, float b, float w) { return a + (b - a) * w; }
inline float deg_to_rad(float d) { return d * (float)(Math_PI / 180.0); }
inline bool is_zero_approx(float s) { return abs(s) < (float)CMP_EPSILON; }
inline bool is_zero_approx(double s) { return abs(s) < CMP_EPSILON; }
inline bool is_equal_approx(float a, float b) {
	if (a == b) {
		return true;
	}
	float tolerance = (float)CMP_EPSILON * abs(a);
	if (tolerance < (float)CMP_EPSILON) {
		tolerance = (float)CMP_EPSILON;
	}
	return abs(a - b) < tolerance;
}
} // namespace Math

struct Vector3 {
	enum Axis { AXIS_X, AXIS_Y, AXIS_Z };
	real_t x = 0, y = 0, z = 0;
	Vector3() {}
	Vector3(real_t px, real_t py, real_t pz) : x(px), y(py), z(pz) {}
	real_t &operator[](int i) { return (&x)[i]; }
	const real_t &operator[](int i) const { return (&x)[i]; }
	Vector3 operator+(const Vector3 &v) const { return Vector3(x + v.x, y + v.y, z + v.z); }
	Vector3 operator-(const Vector3 &v) const { return Vector3(x - v.x, y - v.y, z - v.z); }
	Vector3 operator*(const Vector3 &v) const { return Vector3(x * v.x, y * v.y, z * v.z); }
	Vector3 operator*(real_t s) const { return Vector3(x * s, y * s, z * s); }
	Vector3 operator/(real_t s) const { return Vector3(x / s, y / s, z / s); }
	Vector3 operator-() const { return Vector3(-x, -y, -z); }
	Vector3 &operator+=(const Vector3 &v) { x += v.x; y += v.y; z += v.z; return *this; }
	Vector3 &operator-=(const Vector3 &v) { x -= v.x; y -= v.y; z -= v.z; return *this; }
	Vector3 &operator*=(const Vector3 &v) { x *= v.x; y *= v.y; z *= v.z; return *this; }
	Vector3 &operator*=(real_t s) { x *= s; y *= s; z *= s; return *this; }
	Vector3 &operator/=(real_t s) { x /= s; y /= s; z /= s; return *this; }
	bool operator==(const Vector3 &v) const { return x == v.x && y == v.y && z == v.z; }
	bool operator!=(const Vector3 &v) const { return x != v.x || y != v.y || z != v.z; }
	real_t dot(const Vector3 &v) const { return x * v.x + y * v.y + z * v.z; }
	Vector3 cross(const Vector3 &v) const {
		return Vector3((y * v.z) - (z * v.y), (z * v.x) - (x * v.z), (x * v.y) - (y * v.x));
	}
	real_t length_squared() const {
		real_t x2 = x * x;
		real_t y2 = y * y;
		real_t z2 = z * z;
		return x2 + y2 + z2;
	}
	real_t length() const { return Math::sqrt(length_squared()); }
	void normalize() {
		real_t lengthsq = length_squared();
		if (lengthsq == 0) {
			x = y = z = 0;
		} else {
			real_t len = Math::sqrt(lengthsq);
			x /= len;
			y /= len;
			z /= len;
		}
	}
	Vector3 normalized() const {
		Vector3 v = *this;
		v.normalize();
		return v;
	}
	real_t distance_to(const Vector3 &v) const { return (v - *this).length(); }
	bool is_zero_approx() const { return Math::is_zero_approx(x) && Math::is_zero_approx(y) && Math::is_zero_approx(z); }
	bool is_equal_approx(const Vector3 &v) const {
		return Math::is_equal_approx(x, v.x) && Math::is_equal_approx(y, v.y) && Math::is_equal_approx(z, v.z);
	}
	bool is_finite() const { return Math::is_finite(x) && Math::is_finite(y) && Math::is_finite(z); }
	Vector3 get_any_perpendicular() const {
		return cross((Math::abs(x) <= Math::abs(y) && Math::abs(x) <= Math::abs(z)) ? Vector3(1, 0, 0) : Vector3(0, 1, 0)).normalized();
	}
	inline Vector3 rotated(const Vector3 &p_axis, real_t p_angle) const;
};
// `double * Vector3` and `Vector3 * double` both narrow the scalar to real_t first (engine: operator*(real_t)).
inline Vector3 operator*(real_t s, const Vector3 &v) { return v * s; }

struct Quaternion;

struct Basis {
	Vector3 rows[3] = { Vector3(1, 0, 0), Vector3(0, 1, 0), Vector3(0, 0, 1) };
	Basis() {}
	Basis(real_t xx, real_t xy, real_t xz, real_t yx, real_t yy, real_t yz, real_t zx, real_t zy, real_t zz) { set(xx, xy, xz, yx, yy, yz, zx, zy, zz); }
	inline Basis(const Quaternion &q); // implicit, as in the engine
	inline Basis(const Quaternion &q, const Vector3 &p_scale); // set_quaternion_scale (engine core/math/basis.cpp)
	Basis(const Vector3 &p_axis, real_t p_angle) { set_axis_angle(p_axis, p_angle); }
	void set(real_t xx, real_t xy, real_t xz, real_t yx, real_t yy, real_t yz, real_t zx, real_t zy, real_t zz) {
		rows[0] = Vector3(xx, xy, xz);
		rows[1] = Vector3(yx, yy, yz);
		rows[2] = Vector3(zx, zy, zz);
	}
	Vector3 get_column(int i) const { return Vector3(rows[0][i], rows[1][i], rows[2][i]); }
	void set_column(int i, const Vector3 &v) { rows[0][i] = v.x; rows[1][i] = v.y; rows[2][i] = v.z; }
	real_t tdotx(const Vector3 &v) const { return rows[0][0] * v[0] + rows[1][0] * v[1] + rows[2][0] * v[2]; }
	real_t tdoty(const Vector3 &v) const { return rows[0][1] * v[0] + rows[1][1] * v[1] + rows[2][1] * v[2]; }
	real_t tdotz(const Vector3 &v) const { return rows[0][2] * v[0] + rows[1][2] * v[1] + rows[2][2] * v[2]; }
	Vector3 xform(const Vector3 &v) const { return Vector3(rows[0].dot(v), rows[1].dot(v), rows[2].dot(v)); }
	Basis operator*(const Basis &m) const {
		return Basis(m.tdotx(rows[0]), m.tdoty(rows[0]), m.tdotz(rows[0]),
				m.tdotx(rows[1]), m.tdoty(rows[1]), m.tdotz(rows[1]),
				m.tdotx(rows[2]), m.tdoty(rows[2]), m.tdotz(rows[2]));
	}
	void operator*=(const Basis &m) { *this = *this * m; }
	bool operator==(const Basis &b) const { return rows[0] == b.rows[0] && rows[1] == b.rows[1] && rows[2] == b.rows[2]; }
	bool operator!=(const Basis &b) const { return !(*this == b); }
	real_t cofac(int row1, int col1, int row2, int col2) const {
		return rows[row1][col1] * rows[row2][col2] - rows[row1][col2] * rows[row2][col1];
	}
	void invert() {
		real_t co[3] = { cofac(1, 1, 2, 2), cofac(1, 2, 2, 0), cofac(1, 0, 2, 1) };
		real_t det = rows[0][0] * co[0] + rows[0][1] * co[1] + rows[0][2] * co[2];
		real_t s = 1.0f / det;
		set(co[0] * s, cofac(0, 2, 2, 1) * s, cofac(0, 1, 1, 2) * s,
				co[1] * s, cofac(0, 0, 2, 2) * s, cofac(0, 2, 1, 0) * s,
				co[2] * s, cofac(0, 1, 2, 0) * s, cofac(0, 0, 1, 1) * s);
	}
	Basis inverse() const {
		Basis b = *this;
		b.invert();
		return b;
	}
	void orthonormalize() {
		Vector3 x = get_column(0);
		Vector3 y = get_column(1);
		Vector3 z = get_column(2);
		x.normalize();
		y = (y - x * (x.dot(y)));
		y.normalize();
		z = (z - x * (x.dot(z)) - y * (y.dot(z)));
		z.normalize();
		set_column(0, x);
		set_column(1, y);
		set_column(2, z);
	}
	Basis orthonormalized() const {
		Basis b = *this;
		b.orthonormalize();
		return b;
	}
	real_t determinant() const {
		return rows[0][0] * (rows[1][1] * rows[2][2] - rows[2][1] * rows[1][2]) -
				rows[1][0] * (rows[0][1] * rows[2][2] - rows[2][1] * rows[0][2]) +
				rows[2][0] * (rows[0][1] * rows[1][2] - rows[1][1] * rows[0][2]);
	}
	void scale(const Vector3 &s) {
		rows[0][0] *= s.x; rows[0][1] *= s.x; rows[0][2] *= s.x;
		rows[1][0] *= s.y; rows[1][1] *= s.y; rows[1][2] *= s.y;
		rows[2][0] *= s.z; rows[2][1] *= s.z; rows[2][2] *= s.z;
	}
	// engine: scaled() = copy + scale() (rows), scaled_local() = *this * diag(s) (columns),
	// orthogonalize() = orthonormalize keeping the scale.  Referenced by src/math/ik_node_3d.cpp:52,106
	// behind flags the solve path never sets (DIRTY_LOCAL, disable_scale); only oracle/_ref compiles them.
	Basis scaled(const Vector3 &s) const {
		Basis b = *this;
		b.scale(s);
		return b;
	}
	Basis scaled_local(const Vector3 &s) const {
		Basis d;
		d.set(s.x, 0, 0, 0, s.y, 0, 0, 0, s.z);
		return (*this) * d;
	}
	void scale_local(const Vector3 &s) { *this = scaled_local(s); }
	void orthogonalize() {
		Vector3 scl = get_scale();
		orthonormalize();
		scale_local(scl);
	}
	Vector3 get_scale_abs() const {
		return Vector3(Vector3(rows[0][0], rows[1][0], rows[2][0]).length(),
				Vector3(rows[0][1], rows[1][1], rows[2][1]).length(),
				Vector3(rows[0][2], rows[1][2], rows[2][2]).length());
	}
	Vector3 get_scale() const {
		real_t det = determinant();
		real_t det_sign = det > 0 ? 1.0f : (det < 0 ? -1.0f : 0.0f); // SIGN()
		return det_sign * get_scale_abs();
	}
	bool is_finite() const { return rows[0].is_finite() && rows[1].is_finite() && rows[2].is_finite(); }
	inline Quaternion get_quaternion() const;
	inline Quaternion get_rotation_quaternion() const;
	inline Basis slerp(const Basis &p_to, real_t p_weight) const;
	void set_axis_angle(const Vector3 &p_axis, real_t p_angle) {
		Vector3 axis_sq(p_axis.x * p_axis.x, p_axis.y * p_axis.y, p_axis.z * p_axis.z);
		real_t cosine = Math::cos(p_angle);
		rows[0][0] = axis_sq.x + cosine * (1.0f - axis_sq.x);
		rows[1][1] = axis_sq.y + cosine * (1.0f - axis_sq.y);
		rows[2][2] = axis_sq.z + cosine * (1.0f - axis_sq.z);
		real_t sine = Math::sin(p_angle);
		real_t t = 1 - cosine;
		real_t xyzt = p_axis.x * p_axis.y * t;
		real_t zyxs = p_axis.z * sine;
		rows[0][1] = xyzt - zyxs;
		rows[1][0] = xyzt + zyxs;
		xyzt = p_axis.x * p_axis.z * t;
		zyxs = p_axis.y * sine;
		rows[0][2] = xyzt + zyxs;
		rows[2][0] = xyzt - zyxs;
		xyzt = p_axis.y * p_axis.z * t;
		zyxs = p_axis.x * sine;
		rows[1][2] = xyzt - zyxs;
		rows[2][1] = xyzt + zyxs;
	}
};

inline Vector3 Vector3::rotated(const Vector3 &p_axis, real_t p_angle) const {
	return Basis(p_axis, p_angle).xform(*this);
}

struct Quaternion {
	real_t x = 0, y = 0, z = 0, w = 1;
	Quaternion() {}
	// The engine ctor takes real_t; doubles passed by the module narrow here.
	Quaternion(real_t px, real_t py, real_t pz, real_t pw) : x(px), y(py), z(pz), w(pw) {}
	Quaternion(const Basis &b) { *this = b.get_quaternion(); }
	Quaternion(const Vector3 &p_axis, real_t p_angle) {
		real_t d = p_axis.length();
		if (d == 0) {
			x = 0; y = 0; z = 0; w = 0;
		} else {
			real_t sin_angle = Math::sin(p_angle * 0.5f);
			real_t cos_angle = Math::cos(p_angle * 0.5f);
			real_t s = sin_angle / d;
			x = p_axis.x * s;
			y = p_axis.y * s;
			z = p_axis.z * s;
			w = cos_angle;
		}
	}
	// Shortest arc.  Version sensitive (hard part 3).
	Quaternion(const Vector3 &p_v0, const Vector3 &p_v1) {
#if GD_ARC_LEGACY
		Vector3 c = p_v0.cross(p_v1);
		real_t d = p_v0.dot(p_v1);
		if (d < -1.0f + (real_t)CMP_EPSILON) {
			x = 0; y = 1; z = 0; w = 0;
		} else {
			real_t s = Math::sqrt((1.0f + d) * 2.0f);
			real_t rs = 1.0f / s;
			x = c.x * rs; y = c.y * rs; z = c.z * rs; w = s * 0.5f;
		}
#else
		constexpr real_t ALMOST_ONE = 1.0f - (real_t)CMP_EPSILON;
		Vector3 n0 = p_v0.normalized();
		Vector3 n1 = p_v1.normalized();
		real_t d = n0.dot(n1);
		if (Math::abs(d) > ALMOST_ONE) {
			if (d >= 0) {
				return; // identity
			}
			Vector3 axis = n0.get_any_perpendicular();
			x = axis.x; y = axis.y; z = axis.z; w = 0;
		} else {
			Vector3 c = n0.cross(n1);
			real_t s = Math::sqrt((1.0f + d) * 2.0f);
			real_t rs = 1.0f / s;
			x = c.x * rs;
			y = c.y * rs;
			z = c.z * rs;
			w = s * 0.5f;
		}
#endif
	}
	real_t dot(const Quaternion &q) const { return x * q.x + y * q.y + z * q.z + w * q.w; }
	real_t length_squared() const { return dot(*this); }
	real_t length() const { return Math::sqrt(length_squared()); }
	Quaternion operator*(real_t s) const { return Quaternion(x * s, y * s, z * s, w * s); }
	Quaternion operator/(real_t s) const { return *this * (1.0f / s); }
	void operator*=(real_t s) { x *= s; y *= s; z *= s; w *= s; }
	Quaternion operator-() const { return Quaternion(-x, -y, -z, -w); }
	Quaternion normalized() const { return *this / length(); }
	Quaternion inverse() const { return Quaternion(-x, -y, -z, w); }
	Quaternion operator*(const Quaternion &q) const {
		Quaternion r = *this;
		real_t xx = w * q.x + x * q.w + y * q.z - z * q.y;
		real_t yy = w * q.y + y * q.w + z * q.x - x * q.z;
		real_t zz = w * q.z + z * q.w + x * q.y - y * q.x;
		r.w = w * q.w - x * q.x - y * q.y - z * q.z;
		r.x = xx;
		r.y = yy;
		r.z = zz;
		return r;
	}
	Vector3 xform(const Vector3 &v) const {
		Vector3 u(x, y, z);
		Vector3 uv = u.cross(v);
		return v + ((uv * w) + u.cross(uv)) * ((real_t)2);
	}
	Vector3 xform_inv(const Vector3 &v) const { return inverse().xform(v); }
	bool is_finite() const { return Math::is_finite(x) && Math::is_finite(y) && Math::is_finite(z) && Math::is_finite(w); }
	Vector3 get_axis() const {
		if (Math::abs(w) > 1 - (real_t)CMP_EPSILON) {
			return Vector3(x, y, z);
		}
		real_t r = ((real_t)1) / Math::sqrt(1 - w * w);
		return Vector3(x * r, y * r, z * r);
	}
	real_t get_angle() const { return 2 * Math::acos(w); }
	Quaternion slerp(const Quaternion &p_to, real_t p_weight) const {
		Quaternion to1;
		real_t omega, cosom, sinom, scale0, scale1;
		cosom = dot(p_to);
		if (cosom < 0.0f) {
			cosom = -cosom;
			to1 = -p_to;
		} else {
			to1 = p_to;
		}
		if ((1.0f - cosom) > (real_t)CMP_EPSILON) {
			omega = Math::acos(cosom);
			sinom = Math::sin(omega);
			scale0 = Math::sin((1.0f - p_weight) * omega) / sinom;
			scale1 = Math::sin(p_weight * omega) / sinom;
		} else {
			scale0 = 1.0f - p_weight;
			scale1 = p_weight;
		}
		return Quaternion(scale0 * x + scale1 * to1.x, scale0 * y + scale1 * to1.y,
				scale0 * z + scale1 * to1.z, scale0 * w + scale1 * to1.w);
	}
};

inline Basis::Basis(const Quaternion &q) {
	real_t d = q.length_squared();
	real_t s = 2.0f / d;
	real_t xs = q.x * s, ys = q.y * s, zs = q.z * s;
	real_t wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
	real_t xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
	real_t yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
	set(1.0f - (yy + zz), xy - wz, xz + wy,
			xy + wz, 1.0f - (xx + zz), yz - wx,
			xz - wy, yz + wx, 1.0f - (xx + yy));
}

// Basis::set_quaternion_scale: _set_diagonal(p_scale); rotate(p_quaternion);  rotate(q) is *this = Basis(q) * (*this)
// (engine core/math/basis.cpp; what Skeleton3D::Bone::update_pose_cache builds get_bone_pose() from)
inline Basis::Basis(const Quaternion &q, const Vector3 &p_scale) {
	Basis diag(p_scale.x, 0, 0, 0, p_scale.y, 0, 0, 0, p_scale.z);
	*this = Basis(q) * diag;
}

inline Quaternion Basis::get_quaternion() const {
	const Basis &m = *this;
	real_t trace = m.rows[0][0] + m.rows[1][1] + m.rows[2][2];
	real_t temp[4];
	if (trace > 0.0f) {
		real_t s = Math::sqrt(trace + 1.0f);
		temp[3] = (s * 0.5f);
		s = 0.5f / s;
		temp[0] = ((m.rows[2][1] - m.rows[1][2]) * s);
		temp[1] = ((m.rows[0][2] - m.rows[2][0]) * s);
		temp[2] = ((m.rows[1][0] - m.rows[0][1]) * s);
	} else {
		int i = m.rows[0][0] < m.rows[1][1]
				? (m.rows[1][1] < m.rows[2][2] ? 2 : 1)
				: (m.rows[0][0] < m.rows[2][2] ? 2 : 0);
		int j = (i + 1) % 3;
		int k = (i + 2) % 3;
		real_t s = Math::sqrt(m.rows[i][i] - m.rows[j][j] - m.rows[k][k] + 1.0f);
		temp[i] = s * 0.5f;
		s = 0.5f / s;
		temp[3] = (m.rows[k][j] - m.rows[j][k]) * s;
		temp[j] = (m.rows[j][i] + m.rows[i][j]) * s;
		temp[k] = (m.rows[k][i] + m.rows[i][k]) * s;
	}
	return Quaternion(temp[0], temp[1], temp[2], temp[3]);
}

inline Quaternion Basis::get_rotation_quaternion() const {
	Basis m = orthonormalized();
	real_t det = m.determinant();
	if (det < 0) {
		m.scale(Vector3(-1, -1, -1));
	}
	return m.get_quaternion();
}

inline Basis Basis::slerp(const Basis &p_to, real_t p_weight) const {
	Quaternion from(*this);
	Quaternion to(p_to);
	Basis b(from.slerp(to, p_weight));
	b.rows[0] *= Math::lerp(rows[0].length(), p_to.rows[0].length(), p_weight);
	b.rows[1] *= Math::lerp(rows[1].length(), p_to.rows[1].length(), p_weight);
	b.rows[2] *= Math::lerp(rows[2].length(), p_to.rows[2].length(), p_weight);
	return b;
}

struct Transform3D {
	Basis basis;
	Vector3 origin;
	Transform3D() {}
	Transform3D(const Basis &b, const Vector3 &o) : basis(b), origin(o) {}
	Vector3 xform(const Vector3 &v) const {
		return Vector3(basis.rows[0].dot(v) + origin.x, basis.rows[1].dot(v) + origin.y, basis.rows[2].dot(v) + origin.z);
	}
	Transform3D operator*(const Transform3D &t) const {
		Transform3D r = *this;
		r.origin = xform(t.origin);
		r.basis *= t.basis;
		return r;
	}
	void affine_invert() {
		basis.invert();
		origin = basis.xform(-origin);
	}
	Transform3D affine_inverse() const {
		Transform3D r = *this;
		r.affine_invert();
		return r;
	}
	const Basis &get_basis() const { return basis; }
	bool operator==(const Transform3D &t) const { return basis == t.basis && origin == t.origin; }
	bool operator!=(const Transform3D &t) const { return basis != t.basis || origin != t.origin; }
};

} // namespace gd
