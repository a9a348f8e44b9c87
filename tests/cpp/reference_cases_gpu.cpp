// The reference's numeric unit-test cases (reference tests/test_qcp.h, tests/test_ik_kusudama_3d.h), written against
// the C ABI so that they exercise the CUDA stage code: same case names, same inputs, same CHECKs.
// Build: g++ -std=c++17 reference_cases_gpu.cpp -L<pkg> -l:libmbik.so ...; exit code = number of failed checks
// (77 = no CUDA device).
#include "../../many_bone_ik_b200/host/many_bone_ik_host.hpp"

#include <cmath>
#include <cstdio>
#include <vector>

static int g_checks = 0, g_failed = 0;
static const char *g_case = "";
#define TEST_CASE(name) g_case = name;
#define CHECK(cond)                                                            \
	do {                                                                       \
		g_checks++;                                                            \
		if (!(cond)) {                                                         \
			g_failed++;                                                        \
			printf("FAILED %s: %s (line %d)\n", g_case, #cond, __LINE__);      \
		}                                                                      \
	} while (0)

static const double CMP_EPSILON = 0.00001;

struct Vec3 {
	float x, y, z;
};
struct Quat {
	float x, y, z, w;
	Vec3 xform(Vec3 v) const { // Quaternion::xform, float
		Vec3 u{ x, y, z };
		Vec3 uv{ u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x };
		Vec3 uuv{ u.y * uv.z - u.z * uv.y, u.z * uv.x - u.x * uv.z, u.x * uv.y - u.y * uv.x };
		return Vec3{ v.x + (uv.x * w + uuv.x) * 2.0f, v.y + (uv.y * w + uuv.y) * 2.0f, v.z + (uv.z * w + uuv.z) * 2.0f };
	}
	Vec3 xform_inv(Vec3 v) const { return Quat{ -x, -y, -z, w }.xform(v); }
	Quat normalized() const {
		float l = std::sqrt(x * x + y * y + z * z + w * w);
		return Quat{ x / l, y / l, z / l, w / l };
	}
};

// QCP::weighted_superpose + get_translation through the stage probe
static bool weighted_superpose(const std::vector<Vec3> &moved, const std::vector<Vec3> &target, const std::vector<double> &weight, bool translate,
		Quat &rot, Vec3 &translation) {
	float out[7];
	int rc = mbik_stage_qcp(0, (int32_t)moved.size(), &moved[0].x, &target[0].x, weight.data(), translate ? 1 : 0, out);
	rot = Quat{ out[0], out[1], out[2], out[3] };
	translation = Vec3{ out[4], out[5], out[6] };
	return rc == MBIK_OK;
}

// IKKusudama3D with the given open cones on the second bone of a two-bone chain; get_local_point_in_limits
static bool local_point_in_limits(const std::vector<mbik_host::Vector4> &cones, Vec3 point, Vec3 &result, float &in_bounds) {
	using namespace mbik_host;
	Skeleton3D skel;
	Transform3D t;
	skel.add_bone("a", -1, t);
	t.origin[1] = 0.3f;
	skel.add_bone("b", 0, t);
	skel.add_bone("c", 1, t);
	ManyBoneIK3D ik;
	ik.set_skeleton(&skel);
	ik.set_total_effector_count(1);
	ik.set_effector_bone_name(0, "c");
	ik.set_pin_weight(0, 1.0f);
	ik._set_constraint_count(1);
	ik.set_constraint_name_at_index(0, "b");
	ik.set_joint_twist(0, Vector2{ 0.0f, 1.0f });
	ik.set_kusudama_open_cone_count(0, (int32_t)cones.size());
	for (size_t i = 0; i < cones.size(); i++) {
		ik.set_kusudama_open_cone_center(0, (int32_t)i, Vector3{ cones[i].x, cones[i].y, cones[i].z });
		ik.set_kusudama_open_cone_radius(0, (int32_t)i, cones[i].w);
	}
	if (ik._bone_list_changed() != MBIK_OK) {
		return false;
	}
	float out[4];
	int rc = mbik_stage_point_in_limits(ik.get_rig(), 0, skel.find_bone("b"), 1, &point.x, out);
	result = Vec3{ out[0], out[1], out[2] };
	in_bounds = out[3];
	return rc == MBIK_OK;
}

int main() {
	if (mbik_device_count() < 1) {
		printf("no CUDA device\n");
		return 77;
	}
	{
		TEST_CASE("[Modules][QCP] Weighted Superpose") // reference tests/test_qcp.h:40-57
		double epsilon = CMP_EPSILON;
		Quat expected{ 0, 0, (float)(std::sqrt(2.0) / 2), (float)(std::sqrt(2.0) / 2) };
		std::vector<Vec3> moved = { { 4, 5, 6 }, { 7, 8, 9 }, { 1, 2, 3 } };
		std::vector<Vec3> target = moved;
		for (Vec3 &element : target) {
			element = expected.xform(element);
		}
		std::vector<double> weight = { 1.0, 1.0, 1.0 };
		Quat result;
		Vec3 tr;
		CHECK(weighted_superpose(moved, target, weight, false, result, tr));
		CHECK(std::abs(result.x - expected.x) < epsilon);
		CHECK(std::abs(result.y - expected.y) < epsilon);
		CHECK(std::abs(result.z - expected.z) < epsilon);
		CHECK(std::abs(result.w - expected.w) < epsilon);
	}
	{
		TEST_CASE("[Modules][QCP] Weighted Translation") // :59-85
		double epsilon = CMP_EPSILON;
		Quat expected{ 0, 0, 0, 1 };
		std::vector<Vec3> moved = { { 4, 5, 6 }, { 7, 8, 9 }, { 1, 2, 3 } };
		std::vector<Vec3> target = moved;
		Vec3 translation_vector{ 1, 2, 3 };
		for (Vec3 &element : target) {
			element = expected.xform(Vec3{ element.x + translation_vector.x, element.y + translation_vector.y, element.z + translation_vector.z });
		}
		std::vector<double> weight = { 1.0, 1.0, 1.0 };
		Quat result;
		Vec3 tr;
		CHECK(weighted_superpose(moved, target, weight, true, result, tr));
		CHECK(std::abs(result.x - expected.x) < epsilon);
		CHECK(std::abs(result.y - expected.y) < epsilon);
		CHECK(std::abs(result.z - expected.z) < epsilon);
		CHECK(std::abs(result.w - expected.w) < epsilon);
		Vec3 translation_result = expected.xform_inv(tr);
		CHECK(std::abs(translation_result.x - translation_vector.x) < epsilon);
		CHECK(std::abs(translation_result.y - translation_vector.y) < epsilon);
		CHECK(std::abs(translation_result.z - translation_vector.z) < epsilon);
	}
	{
		TEST_CASE("[Modules][QCP] Weighted Translation Shortest Path") // :87-113 (a negative test in the reference)
		double epsilon = CMP_EPSILON;
		Quat expected = Quat{ 1, 2, 3, 4 }.normalized();
		std::vector<Vec3> moved = { { 4, 5, 6 }, { 7, 8, 9 }, { 1, 2, 3 } };
		std::vector<Vec3> target = moved;
		Vec3 translation_vector{ 1, 2, 3 };
		for (Vec3 &element : target) {
			element = expected.xform(Vec3{ element.x + translation_vector.x, element.y + translation_vector.y, element.z + translation_vector.z });
		}
		std::vector<double> weight = { 1.0, 1.0, 1.0 };
		Quat result;
		Vec3 tr;
		CHECK(weighted_superpose(moved, target, weight, true, result, tr));
		CHECK(std::abs(result.x - expected.x) > epsilon);
		CHECK(std::abs(result.y - expected.y) > epsilon);
		CHECK(std::abs(result.z - expected.z) > epsilon);
		CHECK(std::abs(result.w - expected.w) > epsilon);
		Vec3 translation_result = expected.xform_inv(tr);
		CHECK(std::abs(translation_result.x - translation_vector.x) > epsilon);
		CHECK(std::abs(translation_result.y - translation_vector.y) > epsilon);
		CHECK(std::abs(translation_result.z - translation_vector.z) > epsilon);
	}
	const float deg30 = (float)(30.0 * 3.14159265358979323846 / 180.0);
	{
		TEST_CASE("[Modules][ManyBoneIK][IKKusudama3D] point inside a 30 degree cone is returned unchanged") // test_ik_kusudama_3d.h:38-64
		Vec3 p{ 0.0f, 0.1f, 1.0f }, r;
		float b;
		CHECK(local_point_in_limits({ { 0, 0, 1, deg30 } }, p, r, b));
		float l = std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z);
		CHECK(b > 0);
		CHECK(std::abs(r.x - p.x / l) < 1e-6 && std::abs(r.y - p.y / l) < 1e-6 && std::abs(r.z - p.z / l) < 1e-6);
	}
	{
		TEST_CASE("[Modules][ManyBoneIK][IKKusudama3D] point outside a zero-radius cone returns the control point") // :96-124
		Vec3 r;
		float b;
		CHECK(local_point_in_limits({ { 0, 0, 1, 0.0f } }, Vec3{ 1, 0, 0 }, r, b));
		CHECK(b < 0);
		CHECK(std::abs(r.x) < 1e-4 && std::abs(r.y) < 1e-4 && std::abs(r.z - 1.0f) < 1e-4);
	}
	{
		TEST_CASE("[Modules][ManyBoneIK][IKKusudama3D] point outside a 30 degree cone lands on its boundary") // :127-156
		Vec3 r;
		float b;
		CHECK(local_point_in_limits({ { 0, 0, 1, deg30 } }, Vec3{ 1, 0, 0 }, r, b));
		CHECK(b == -1.0f);
		CHECK(std::abs(r.x - 0.5f) < CMP_EPSILON);
		CHECK(std::abs(r.y - 0.0f) < CMP_EPSILON);
		CHECK(std::abs(r.z - 0.8660254f) < CMP_EPSILON);
	}
	printf("%d checks, %d failed\n", g_checks, g_failed);
	return g_failed;
}
