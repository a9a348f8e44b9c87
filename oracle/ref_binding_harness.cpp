// TEST INFRASTRUCTURE ONLY -- oracle/_ref/libmbik_ref_binding.so: the reference module's own classes with
// ManyBoneIK3D::_process_modification() REPLACED by the libmbik.so binding a maintainer would add
// (many_bone_ik_b200/host/godot_module_binding.h).  Same headless scene, same property-path
// configuration, same per-frame driving as ref_harness.cpp; the only difference is which code solves the frame:
//   libmbik_ref.so          : the reference's CPU solver                (ref_solve_batch)
//   libmbik_ref_binding.so  : the CUDA path through the C ABI, 1 pose   (ref_binding_solve_batch)
// tests/test_reference_gpu.py asserts that both leave bit-identical position / rotation / scale on the Skeleton3D.
// Links libmbik.so (the product) -- which is why it is a separate library from the pure reference build.
#include "ref_scene.h"

#include <algorithm>

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

// Crowd mode (mbik_godot::CrowdBinding): the node only enqueues its frame; the session flushes once per frame.
class ManyBoneIK3DInCrowd : public ManyBoneIK3D {
public:
	mbik_godot::CrowdBinding<ManyBoneIK3D> *crowd = nullptr;
	mbik_godot::Binding gpu; // per-node state of the binding (here: its rig group)
	void _process_modification() override { crowd->enqueue(*this, gpu); }
};

// deep copy of a rig description (the session outlives the caller's buffers)
struct OwnedDesc {
	std::vector<int32_t> parent;
	std::vector<float> rest, damp;
	std::vector<mbik_pin_desc> pins;
	std::vector<mbik_constraint_desc> cons;
	std::vector<mbik_cone_desc> cones;
	mbik_rig_desc d{};
	explicit OwnedDesc(const mbik_rig_desc *s) {
		d = *s;
		parent.assign(s->parent, s->parent + s->n_bones);
		rest.assign(s->rest_local, s->rest_local + (size_t)s->n_bones * 12);
		pins.assign(s->pins, s->pins + s->n_pins);
		cons.assign(s->constraints, s->constraints + s->n_constraints);
		size_t nc = 0;
		for (const mbik_constraint_desc &c : cons) {
			nc = std::max(nc, (size_t)(c.cone_offset + std::max(0, c.n_cones)));
		}
		cones.assign(s->cones, s->cones + nc);
		damp.assign(s->bone_damp, s->bone_damp + s->n_bone_damp);
		d.parent = parent.data();
		d.rest_local = rest.data();
		d.pins = pins.data();
		d.constraints = cons.data();
		d.cones = cones.data();
		d.bone_damp = damp.data();
	}
};

struct CrowdSession {
	mbik_godot::CrowdBinding<ManyBoneIK3D> crowd;
	std::vector<OwnedDesc *> descs;
	struct Member {
		RefScene *scene;
		ManyBoneIK3DInCrowd *node;
		int desc;
		std::vector<float> start; // first frame only
		bool started = false;
	};
	std::vector<Member> members;
	~CrowdSession() {
		for (Member &m : members) {
			delete m.scene;
		}
		for (OwnedDesc *d : descs) {
			delete d;
		}
	}
};

} // namespace

extern "C" {

// ---- a crowd of long-lived nodes, solved through CrowdBinding: one mbik_solve_batch per rig and frame ----
void *ref_crowd_create(void) { return new CrowdSession(); }
void ref_crowd_destroy(void *h) { delete static_cast<CrowdSession *>(h); }
// n_nodes scenes of rig `d`, each with its own Skeleton3D + ManyBoneIK3D; start_pose [n_nodes][n_bones][12] or NULL (rest)
int ref_crowd_add(void *h, const mbik_rig_desc *d, int n_nodes, const float *start_pose, int iterations) {
	CrowdSession *S = static_cast<CrowdSession *>(h);
	if (!S || !d || n_nodes < 1) {
		return -1;
	}
	S->descs.push_back(new OwnedDesc(d));
	const int di = (int)S->descs.size() - 1;
	const mbik_rig_desc *own = &S->descs[(size_t)di]->d;
	for (int k = 0; k < n_nodes; k++) {
		CrowdSession::Member m;
		m.node = new ManyBoneIK3DInCrowd();
		m.node->crowd = &S->crowd;
		m.scene = new RefScene(own, iterations, m.node); // (its constructor runs one frame: enqueued, flushed below)
		m.desc = di;
		if (start_pose) {
			m.start.assign(start_pose + (size_t)k * own->n_bones * 12, start_pose + (size_t)(k + 1) * own->n_bones * 12);
		}
		ManyBoneIK3DInCrowd *node = m.node;
		m.scene->solved_override = [node, S](std::vector<char> &solved) {
			std::fill(solved.begin(), solved.end(), 0);
			if (node->gpu.crowd_group < 0) {
				return false;
			}
			const Vector<int32_t> &order = S->crowd.groups[node->gpu.crowd_group]->bone_order;
			for (int i = 0; i < order.size(); i++) {
				solved[(size_t)order[i]] = 1;
			}
			return true;
		};
		S->members.push_back(m);
	}
	S->crowd.flush(); // the construction frames
	return (int)S->members.size();
}
// One frame of every node.  targets: all nodes' [n_pins][12] records concatenated in node order; out_pose: all nodes'
// [n_bones][10] concatenated; out_status: one word per node (nullable); launches (nullable): mbik_solve_batch calls of the flush.
// Frame 0 of a node starts from its start pose (or rest); later frames from what its skeleton holds.
int ref_crowd_frame(void *h, const float *targets, float *out_pose, uint32_t *out_status, int *launches) {
	CrowdSession *S = static_cast<CrowdSession *>(h);
	if (!S || !targets || !out_pose) {
		return -1;
	}
	const float *t = targets;
	for (CrowdSession::Member &m : S->members) {
		const mbik_rig_desc *d = &S->descs[(size_t)m.desc]->d;
		m.scene->begin_frame(d, t, m.start.empty() ? nullptr : m.start.data(), m.started);
		m.started = true;
		t += (size_t)d->n_pins * 12;
	}
	S->crowd.flush();
	float *o = out_pose;
	size_t k = 0;
	for (CrowdSession::Member &m : S->members) {
		const mbik_rig_desc *d = &S->descs[(size_t)m.desc]->d;
		m.scene->end_frame(d, o, nullptr, out_status ? out_status + k : nullptr);
		o += (size_t)d->n_bones * 10;
		k++;
	}
	if (launches) {
		*launches = S->crowd.launches_last_flush;
	}
	return S->crowd.last_error;
}

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
