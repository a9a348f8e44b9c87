// Test driver for the C++ host facade (many_bone_ik_b200/host/many_bone_ik_host.hpp).
//   facade_driver <skeleton.txt> <properties.txt> [<targets.bin> <out_pose.bin> <n_poses>]
// skeleton.txt:  n_bones, then per bone "name parent 12 floats" (Transform3D memory layout)
// properties.txt: `key = value` lines exactly as a Godot .tscn stores a ManyBoneIK3D node
// Prints the schedule facts as `key: values` lines; with the optional arguments it also solves the batch through
// ManyBoneIK3D::process_modification_batch and writes out_pose ([n][n_bones][10] float32).
#include "../../many_bone_ik_b200/host/many_bone_ik_host.hpp"

#include <cstdio>
#include <fstream>
#include <iostream>

using namespace mbik_host;

int main(int argc, char **argv) {
	if (argc < 3) {
		fprintf(stderr, "usage: %s skeleton.txt properties.txt [targets.bin out.bin n_poses]\n", argv[0]);
		return 2;
	}
	Skeleton3D skeleton;
	{
		std::ifstream in(argv[1]);
		int n = 0;
		in >> n;
		for (int b = 0; b < n; b++) {
			std::string name;
			int parent;
			Transform3D t;
			in >> name >> parent;
			for (int k = 0; k < 9; k++) {
				in >> t.basis[k];
			}
			for (int k = 0; k < 3; k++) {
				in >> t.origin[k];
			}
			skeleton.add_bone(name, parent, t);
		}
	}
	ManyBoneIK3D ik;
	ik.set_skeleton(&skeleton);
	std::ifstream props(argv[2]);
	int applied = ik.load_properties(props);
	printf("applied: %d\n", applied);
	printf("pin_count: %d\nconstraint_count: %d\n", ik.get_effector_count(), ik.get_constraint_count());
	// reference-style round trip through _get
	Variant v;
	if (ik._get("pins/0/weight", v)) {
		printf("pins/0/weight: %.9g\n", std::get<double>(v));
	}
	if (ik._get("constraints/0/kusudama_open_cone/0/radius", v)) {
		printf("constraints/0/kusudama_open_cone/0/radius: %.9g\n", std::get<double>(v));
	}
	int rc = ik._bone_list_changed();
	printf("rebuild: %d\n", rc);
	if (rc != MBIK_OK) {
		printf("error: %s\n", mbik_last_error());
		return 1;
	}
	mbik_rig_info info;
	mbik_rig_get_info(ik.get_rig(), &info);
	printf("n_bones: %d\nn_solved: %d\nn_segments: %d\nn_effectors: %d\nmax_headings: %d\nn_cones: %d\niterations: %d\n", info.n_bones, info.n_solved,
			info.n_segments, info.n_effectors, info.max_headings, info.n_cones, info.iterations);
	printf("bone_list:");
	for (int32_t b : ik.get_bone_list()) {
		printf(" %d", b);
	}
	printf("\n");
	if (argc >= 6) {
		size_t n = (size_t)atoll(argv[5]);
		std::vector<float> targets(n * ik.get_effector_count() * 12), out(n * info.n_bones * 10);
		std::vector<uint32_t> status(n);
		FILE *f = fopen(argv[3], "rb");
		if (!f || fread(targets.data(), sizeof(float), targets.size(), f) != targets.size()) {
			fprintf(stderr, "cannot read targets\n");
			return 2;
		}
		fclose(f);
		rc = ik.process_modification_batch(n, targets.data(), nullptr, out.data(), nullptr, status.data());
		printf("solve: %d\n", rc);
		if (rc != MBIK_OK) {
			printf("error: %s\n", mbik_last_error());
			return rc == MBIK_ERR_NO_DEVICE ? 3 : 1;
		}
		f = fopen(argv[4], "wb");
		fwrite(out.data(), sizeof(float), out.size(), f);
		fclose(f);
	}
	return 0;
}
