// TEST INFRASTRUCTURE ONLY -- oracle/_ref/libmbik_ref_binding.so: the reference module's own classes with
// ManyBoneIK3D::_process_modification() REPLACED by the libmbik.so binding a maintainer would add
// (many_bone_ik_b200/host/godot_module_binding.h).  Same headless scene, same property-path
// configuration, same per-frame driving as ref_harness.cpp; the only difference is which code solves the frame:
//   libmbik_ref.so          : the reference's CPU solver                (ref_solve_batch)
//   libmbik_ref_binding.so  : the CUDA path through the C ABI, 1 pose   (ref_binding_solve_batch)
// tests/test_reference_gpu.py asserts that both leave bit-identical position / rotation / scale on the Skeleton3D.
// Links libmbik.so (the product) -- which is why it is a separate library from the pure reference build.
#include "ref_scene.h"

#include "../many_bone_ik_b200/host/godot_module_binding.h"

using namespace ref_scene;

namespace {

// the in-tree edit, expressed as a subclass so the reference's translation units stay unmodified: the virtual
// _process_modification (src/many_bone_ik_3d.h:90) is overridden with the binding's body
class ManyBoneIK3DOnGpu : public ManyBoneIK3D {
public:
	mbik_godot::Binding gpu;
	bool frame_wrote = false;
	void _process_modification() override {
		frame_wrote = false;
		gpu.process_modification(*this);
		frame_wrote = gpu.rig != nullptr && gpu.last_error == MBIK_OK;
	}
};

} // namespace

extern "C" {

// Same contract as ref_solve_batch, minus out_local (the binding hands the skeleton position / rotation / scale only).
// Returns 0, or the first mbik error code the binding saw.
int ref_binding_solve_batch(const mbik_rig_desc *d, size_t n_poses, const float *targets, const float *start_pose, float *out_pose,
		uint32_t *out_status, int iterations, unsigned flags) {
	if (!d || !targets || !out_pose) {
		return -1;
	}
	bool rebuild_each = (flags & 1u) != 0;
	RefScene *scene = nullptr;
	ManyBoneIK3DOnGpu *node = nullptr;
	int rc = 0;
	for (size_t k = 0; k < n_poses; k++) {
		if (!scene || rebuild_each) {
			delete scene;
			node = new ManyBoneIK3DOnGpu();
			scene = new RefScene(d, iterations, node);
			scene->solved_override = [node, d](std::vector<char> &solved) {
				std::fill(solved.begin(), solved.end(), 0);
				for (int i = 0; i < node->gpu.bone_order.size(); i++) {
					solved[(size_t)node->gpu.bone_order[i]] = 1;
				}
				return node->frame_wrote;
			};
		}
		scene->solve(d, targets + k * (size_t)d->n_pins * 12, start_pose ? start_pose + k * (size_t)d->n_bones * 12 : nullptr,
				out_pose + k * (size_t)d->n_bones * 10, nullptr, out_status ? out_status + k : nullptr);
		if (node->gpu.last_error != MBIK_OK && rc == 0) {
			rc = node->gpu.last_error;
		}
	}
	delete scene;
	return rc;
}

} // extern "C"
