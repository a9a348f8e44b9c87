// mbik_capi.cu -- the extern "C" boundary declared in include/mbik.h.
// Host runtime around the solve kernel: rig handles, per-device rig copies and staging buffers,
// host<->device copies on per-device streams, contiguous multi-GPU sharding (one worker thread per
// device, no collective).  There is no CPU fallback: without a CUDA device the solve entry points fail.
#include "../../include/mbik.h"
#include "mbik_flatten.h"
#include "mbik_kernel.h"

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h> // header-only NVTX v3: ranges are no-ops unless a profiler injects itself

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg) {
	g_last_error = msg;
	return code;
}
int cuda_fail(cudaError_t e, const char *what) {
	return fail(MBIK_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// restores the caller's current device when an entry point had to switch (multi-GPU processes: a later call with
// params->device = -1, or the caller's own CUDA work, must not silently move to another GPU)
struct DeviceGuard {
	int prev = -1;
	DeviceGuard() { cudaGetDevice(&prev); }
	~DeviceGuard() {
		int cur = -1;
		if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) {
			cudaSetDevice(prev);
		}
	}
};

// NVTX range of one stage of the host pipeline (tracing): upload / solve / download per chunk
struct NvtxRange {
	explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
	~NvtxRange() { nvtxRangePop(); }
};

constexpr int kLanes = 3; // host-I/O pipeline depth: H2D(i+1) | solve(i) | D2H(i-1) on three streams

// staging of one pipeline lane, grown on demand and reused (no allocation in steady state)
struct Lane {
	cudaStream_t stream = nullptr;
	cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
	float *d_targets = nullptr, *d_start = nullptr, *d_out = nullptr, *d_local = nullptr;
	uint32_t *d_status = nullptr;
	int32_t *d_index = nullptr; // limit-set index per pose (mbik_solve_batch_limits)
	size_t cap_targets = 0, cap_start = 0, cap_out = 0, cap_local = 0, cap_status = 0, cap_index = 0;
};

// per-pose limit sets of one call (device table of this device + the caller's index buffer)
struct LimitArgs {
	const unsigned char *table = nullptr;
	uint32_t stride = 0;
	int32_t n_sets = 0;
	const int32_t *set_index = nullptr; // host or device memory, like the other buffers of the call
};

struct DeviceState {
	int device = -1;
	int sm_count = 0;
	unsigned char *blob = nullptr;
	cudaEvent_t ev_start = nullptr, ev_stop = nullptr; // MBIK_IO_DEVICE launches (on the caller's stream)
	bool timed = false;
	float host_path_kernel_ms = -1.f; // sum over the chunks of the last MBIK_IO_HOST call
	bool last_was_host = false;
	Lane lanes[kLanes];
	std::mutex host_mu; // the staging lanes are per (rig, device): concurrent host-I/O calls on one device take turns
};

} // namespace

// deep copy of the description a rig was created from (limit sets re-run the flattener with other constraint values)
struct DescCopy {
	std::vector<int32_t> parent;
	std::vector<float> rest;
	std::vector<mbik_pin_desc> pins;
	std::vector<mbik_constraint_desc> constraints;
	std::vector<mbik_cone_desc> cones;
	std::vector<float> bone_damp;
	mbik_rig_desc scalars{};
	void assign(const mbik_rig_desc *d) {
		scalars = *d;
		parent.assign(d->parent, d->parent + d->n_bones);
		rest.assign(d->rest_local, d->rest_local + (size_t)d->n_bones * 12);
		pins.assign(d->pins, d->pins + (d->n_pins > 0 ? d->n_pins : 0));
		constraints.assign(d->constraints, d->constraints + (d->n_constraints > 0 ? d->n_constraints : 0));
		size_t n_cones = 0;
		for (const mbik_constraint_desc &c : constraints) {
			if (c.n_cones > 0 && (size_t)(c.cone_offset + c.n_cones) > n_cones) {
				n_cones = (size_t)(c.cone_offset + c.n_cones);
			}
		}
		cones.assign(d->cones, d->cones + n_cones);
		bone_damp.assign(d->bone_damp, d->bone_damp + (d->n_bone_damp > 0 ? d->n_bone_damp : 0));
	}
	mbik_rig_desc view() const {
		mbik_rig_desc d = scalars;
		d.parent = parent.data();
		d.rest_local = rest.data();
		d.pins = pins.data();
		d.constraints = constraints.data();
		d.cones = cones.data();
		d.bone_damp = bone_damp.data();
		return d;
	}
};

struct mbik_rig {
	mbik::FlatRig flat;
	DescCopy desc;
	int n_solved = 0;
	int variant = -1;
	std::mutex mu;
	std::map<int, std::unique_ptr<DeviceState>> devices;
};

// alternative fills of the rig's constraint tables, flattened on the host: n_sets records of `stride` bytes,
// [BlobCone x n_cones][BlobBone x n_solved] each (what solve_body<LIMS> reads instead of the blob's cones / twist frames)
struct mbik_limit_sets {
	mbik_rig *rig = nullptr;
	int32_t n_sets = 0;
	uint32_t stride = 0;
	std::vector<unsigned char> table;
	std::mutex mu;
	std::map<int, unsigned char *> device_tables;
	// authoring (host, thread pool): mbik_limit_sets_create waits for it, mbik_limit_sets_create_async returns while it runs
	std::thread worker;
	std::vector<mbik_constraint_desc> own_constraints; // async: the caller's tables are copied, it may free them at once
	std::vector<mbik_cone_desc> own_cones;
	int rc = MBIK_OK;
	std::string error;
	double author_seconds = 0.0;
	int author_threads = 0;
};

struct mbik_stream {
	mbik_rig *rig = nullptr;
	int device = -1;
	int sm_count = 148;
	size_t n_poses = 0;
	unsigned char *blob = nullptr;    // borrowed from the rig's DeviceState
	float *local[2] = { nullptr, nullptr }; // ping-pong raw local transforms [n][n_bones][12]
	int cur = 0;                      // local[cur] holds the latest state
	float *d_targets[2] = { nullptr, nullptr };
	float *d_out[2] = { nullptr, nullptr };
	uint32_t *d_status[2] = { nullptr, nullptr };
	cudaStream_t s_up = nullptr, s_solve = nullptr, s_down = nullptr;
	cudaEvent_t ev_up[2] = { nullptr, nullptr };     // targets slot uploaded
	cudaEvent_t ev_solved[2] = { nullptr, nullptr }; // solve reading/writing slot finished
	cudaEvent_t ev_down[2] = { nullptr, nullptr };   // out slot downloaded (slot reusable)
	bool slot_used[2] = { false, false };
	int64_t frames = 0;
	uint32_t flags = 0;  // MBIK_OUT_SOLVED_ONLY: out_pose rows of mbik_stream_submit
	size_t out_rows = 0; // bones per pose in out_pose
};

namespace {

void free_device_state(DeviceState &ds) {
	if (ds.device < 0) {
		return;
	}
	cudaSetDevice(ds.device);
	cudaFree(ds.blob);
	if (ds.ev_start) {
		cudaEventDestroy(ds.ev_start);
	}
	if (ds.ev_stop) {
		cudaEventDestroy(ds.ev_stop);
	}
	for (Lane &ln : ds.lanes) {
		cudaFree(ln.d_targets);
		cudaFree(ln.d_start);
		cudaFree(ln.d_out);
		cudaFree(ln.d_local);
		cudaFree(ln.d_status);
		cudaFree(ln.d_index);
		if (ln.ev_start) {
			cudaEventDestroy(ln.ev_start);
		}
		if (ln.ev_stop) {
			cudaEventDestroy(ln.ev_stop);
		}
		if (ln.stream) {
			cudaStreamDestroy(ln.stream);
		}
	}
}

// Get (creating on first use) the per-device copy of the rig.  Caller has set the device.
int get_device_state(mbik_rig *rig, int device, DeviceState **out) {
	std::lock_guard<std::mutex> lock(rig->mu);
	auto it = rig->devices.find(device);
	if (it != rig->devices.end()) {
		*out = it->second.get();
		return MBIK_OK;
	}
	std::unique_ptr<DeviceState> owner(new DeviceState());
	DeviceState &ds = *owner;
	ds.device = device;
	cudaError_t e = cudaMalloc((void **)&ds.blob, rig->flat.blob.size());
	if (e != cudaSuccess) {
		return cuda_fail(e, "cudaMalloc(rig blob)");
	}
	e = cudaMemcpy(ds.blob, rig->flat.blob.data(), rig->flat.blob.size(), cudaMemcpyHostToDevice);
	if (e != cudaSuccess) {
		cudaFree(ds.blob);
		return cuda_fail(e, "cudaMemcpy(rig blob)");
	}
	e = cudaEventCreate(&ds.ev_start);
	if (e == cudaSuccess) {
		e = cudaEventCreate(&ds.ev_stop);
	}
	for (Lane &ln : ds.lanes) {
		if (e == cudaSuccess) {
			e = cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking);
		}
		if (e == cudaSuccess) {
			e = cudaEventCreate(&ln.ev_start);
		}
		if (e == cudaSuccess) {
			e = cudaEventCreate(&ln.ev_stop);
		}
	}
	if (e == cudaSuccess) {
		e = cudaDeviceGetAttribute(&ds.sm_count, cudaDevAttrMultiProcessorCount, device);
	}
	if (e == cudaSuccess) {
		// stream-ordered workspaces (streamed-walk and unbounded instantiations): keep freed blocks in the device's default
		// pool instead of returning them to the driver at every synchronisation
		cudaMemPool_t pool = nullptr;
		if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
			uint64_t keep = UINT64_MAX;
			cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
		}
	}
	if (e != cudaSuccess) {
		free_device_state(ds);
		return cuda_fail(e, "stream/event creation");
	}
	*out = owner.get();
	rig->devices.emplace(device, std::move(owner));
	return MBIK_OK;
}

template <class T>
int ensure_capacity(T **p, size_t *cap, size_t bytes) {
	if (bytes <= *cap) {
		return MBIK_OK;
	}
	if (*p) {
		cudaFree(*p);
		*p = nullptr;
		*cap = 0;
	}
	cudaError_t e = cudaMalloc((void **)p, bytes);
	if (e != cudaSuccess) {
		return cuda_fail(e, "cudaMalloc(staging)");
	}
	*cap = bytes;
	return MBIK_OK;
}

int resolve_device(const mbik_solve_params *params, int *device) {
	int dev = params ? params->device : -1;
	if (dev < 0) {
		cudaError_t e = cudaGetDevice(&dev);
		if (e != cudaSuccess) {
			return fail(MBIK_ERR_NO_DEVICE, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
		}
	}
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count <= 0) {
		return fail(MBIK_ERR_NO_DEVICE, "no CUDA device: mbik has no CPU fallback");
	}
	if (dev >= count) {
		return fail(MBIK_ERR_INVALID_ARG, "device ordinal out of range");
	}
	*device = dev;
	return MBIK_OK;
}

// what launch_solve needs to know about the rig's segment-parallel schedule (mbik_blob.h: BlobSpan)
void set_launch_hints(mbik::SolveArgs &a, const mbik::FlatRig &F, uint32_t flags) {
	a.n_solved = (int32_t)F.bones.size();
	a.n_bones = F.n_bones;
	a.n_pins = (int32_t)F.pins.size();
	a.max_seg_len = F.max_seg_len;
	a.max_stack = F.max_stack;
	a.max_list_effs = 0;
	for (const mbik::FlatSegment &seg : F.segments) {
		a.max_list_effs = std::max(a.max_list_effs, (int32_t)seg.effectors.size());
	}
	a.sp_roles = F.sp_roles;
	a.sp_team_bytes = (int32_t)((size_t)F.sp_team_bufs * F.sp_team_headings * 6 * 32 * sizeof(float));
	a.sp_gain = F.sp_critical_cost > 0 ? (float)(F.sp_serial_cost / F.sp_critical_cost) : 1.0f;
	a.sched_mode = (flags & MBIK_SCHED_THROUGHPUT) ? 1 : ((flags & MBIK_SCHED_SEGMENT_PARALLEL) ? 2 : 0);
	// Streamed-walk instantiation for the large batches of the 64-bone-and-up variants (local poses of 3 KB and more per pose:
	// a resident batch does not fit L2).  Measured against thread-local state (profiles/r2_exp_glw_*.log): chain64 +47 ... +50 %,
	// chain150 +52 %, chain200 +49 %, quad80 +9 ... +11 %.  MBIK_GLW=0 / 1 overrides (A/B).
	{
		static const int forced = getenv("MBIK_GLW") ? atoi(getenv("MBIK_GLW")) : -1;
		a.use_glw = (forced >= 0 ? forced != 0 : true) ? 1 : 0;
	}
	a.sp_trace = nullptr;
}

// debug knob MBIK_SP_TRACE=1: per (iteration, phase, warp) busy cycles of CTA 0 of one device-buffer launch, to stderr
void sp_trace_launch(mbik::SolveArgs a, const mbik::FlatRig &F, int variant, int sm_count, cudaStream_t stream) {
	if (!mbik::uses_segment_parallel(a, variant, sm_count) || F.sp_phases < 1) {
		return;
	}
	const size_t n = (size_t)a.iterations * F.sp_phases * F.sp_roles * 2;
	long long *d = nullptr;
	if (n == 0 || cudaMalloc(&d, n * sizeof(long long)) != cudaSuccess) {
		return;
	}
	cudaMemset(d, 0, n * sizeof(long long));
	a.sp_trace = d;
	mbik::launch_solve(a, variant, sm_count, stream);
	std::vector<long long> h(n);
	cudaStreamSynchronize(stream);
	cudaMemcpy(h.data(), d, n * sizeof(long long), cudaMemcpyDeviceToHost);
	cudaFree(d);
	{
		// whole-kernel view: per iteration, first start .. last end of every phase
		fprintf(stderr, "[mbik sp trace] per iteration: cycles from the first phase start to the last phase end, and gaps between phases\n ");
		long long prev_end = -1;
		for (int i2 = 0; i2 < a.iterations; i2++) {
			long long first = -1, last = -1;
			for (int ph = 0; ph < F.sp_phases; ph++) {
				for (int r = 0; r < F.sp_roles; r++) {
					const long long s = h[(((size_t)i2 * F.sp_phases + ph) * F.sp_roles + r) * 2], e = h[(((size_t)i2 * F.sp_phases + ph) * F.sp_roles + r) * 2 + 1];
					first = (first < 0 || s < first) ? s : first;
					last = e > last ? e : last;
				}
			}
			fprintf(stderr, " it%d %lld(gap %lld)", i2, last - first, prev_end < 0 ? 0 : first - prev_end);
			prev_end = last;
		}
		fprintf(stderr, "\n");
	}
	int it = atoi(getenv("MBIK_SP_TRACE")); // iteration whose phases are listed
	it = it < 0 ? 0 : (it >= a.iterations ? a.iterations - 1 : it);
	fprintf(stderr, "[mbik sp trace] iteration %d, cycles per phase and warp (end - start; start relative to the phase's first warp)\n", it);
	for (int ph = 0; ph < F.sp_phases; ph++) {
		long long t0 = -1;
		for (int r = 0; r < F.sp_roles; r++) {
			const long long s = h[(((size_t)it * F.sp_phases + ph) * F.sp_roles + r) * 2];
			t0 = (t0 < 0 || s < t0) ? s : t0;
		}
		fprintf(stderr, "  phase %d:", ph);
		for (int r = 0; r < F.sp_roles; r++) {
			const long long s = h[(((size_t)it * F.sp_phases + ph) * F.sp_roles + r) * 2], e = h[(((size_t)it * F.sp_phases + ph) * F.sp_roles + r) * 2 + 1];
			fprintf(stderr, "  w%d %lld(+%lld)", r, e - s, s - t0);
		}
		fprintf(stderr, "\n");
	}
}

// one shard on one device; host or device buffers
int solve_on_device(mbik_rig *rig, int device, uint32_t flags, cudaStream_t user_stream, int iterations, int newton_iters, size_t n_poses,
		const float *targets, const float *start_pose, float *out_pose, float *out_local, uint32_t *out_status, const LimitArgs *lim = nullptr) {
	if (n_poses == 0) {
		return MBIK_OK;
	}
	DeviceGuard guard;
	cudaError_t e = cudaSetDevice(device);
	if (e != cudaSuccess) {
		return cuda_fail(e, "cudaSetDevice");
	}
	DeviceState *ds = nullptr;
	int rc = get_device_state(rig, device, &ds);
	if (rc != MBIK_OK) {
		return rc;
	}
	const mbik::FlatRig &F = rig->flat;
	const size_t nb = (size_t)F.n_bones, np = F.pins.size();
	mbik::SolveArgs a;
	a.blob = ds->blob;
	a.blob_bytes = reinterpret_cast<const mbik::BlobHeader *>(F.blob.data())->resident_bytes; // what the kernel stages into shared memory
	a.iterations = iterations;
	a.n_poses = n_poses;
	a.stabilize = F.stabilization_passes > 0 ? 1 : 0;
	a.newton_iters = newton_iters > 0 ? newton_iters : 0;
	a.out_flags = ((flags & MBIK_OUT_SOLVED_ONLY) ? mbik::OUT_COMPACT : 0u) | ((flags & MBIK_LOCAL_RECOMPOSED) ? mbik::OUT_LOCAL_RECOMPOSED : 0u);
	const size_t out_rows = (flags & MBIK_OUT_SOLVED_ONLY) ? F.bones.size() : (size_t)F.n_bones; // rows of out_pose per pose
	set_launch_hints(a, F, flags);

	if (lim) {
		a.limit_table = lim->table;
		a.limit_stride = lim->stride;
		a.n_limit_sets = lim->n_sets;
		a.limit_index = lim->set_index; // device path; the host path points it at each lane's staging copy
	}

	if (flags & MBIK_IO_DEVICE) {
		a.targets = targets;
		a.start_pose = start_pose;
		a.out_pose = out_pose;
		a.out_local = out_local;
		a.out_status = out_status;
		static const bool trace = getenv("MBIK_SP_TRACE") != nullptr;
		if (trace) {
			sp_trace_launch(a, F, rig->variant, ds->sm_count, user_stream);
		}
		NvtxRange range("mbik solve (device buffers)");
		cudaEventRecord(ds->ev_start, user_stream);
		e = mbik::launch_solve(a, rig->variant, ds->sm_count, user_stream);
		cudaEventRecord(ds->ev_stop, user_stream);
		ds->timed = true;
		ds->last_was_host = false;
		if (e != cudaSuccess) {
			return cuda_fail(e, "kernel launch");
		}
		return MBIK_OK;
	}

	// Host buffers: the batch is cut into chunks of whole kernel waves and streamed through kLanes staging
	// lanes, so the H2D copy of chunk i+1, the solve of chunk i and the D2H copy of chunk i-1 overlap.
	std::lock_guard<std::mutex> host_lock(ds->host_mu);
	const size_t wave = (size_t)(ds->sm_count > 0 ? ds->sm_count : 148) * mbik::kBlockThreads;
	size_t chunk = wave; // one kernel wave per chunk: shortest pipeline fill / drain (measured: 28.8 vs 27.7 M solves/s at two waves)
	if (const char *env = getenv("MBIK_CHUNK_POSES")) {
		size_t v = (size_t)atoll(env);
		if (v > 0) {
			chunk = v;
		}
	}
	const size_t n_chunks = (n_poses + chunk - 1) / chunk;
	std::vector<int> chunks_of_lane[kLanes];
	for (size_t c = 0; c < n_chunks; c++) {
		Lane &ln = ds->lanes[c % kLanes];
		const size_t b0 = c * chunk, cn = (b0 + chunk <= n_poses) ? chunk : n_poses - b0;
		const size_t bt = cn * np * 12 * sizeof(float), bs = cn * nb * 12 * sizeof(float);
		const size_t bo = cn * out_rows * 10 * sizeof(float), bl = cn * nb * 12 * sizeof(float), bst = cn * sizeof(uint32_t);
		if ((rc = ensure_capacity(&ln.d_targets, &ln.cap_targets, bt ? bt : 16)) != MBIK_OK ||
				(rc = ensure_capacity(&ln.d_out, &ln.cap_out, bo)) != MBIK_OK ||
				(start_pose && (rc = ensure_capacity(&ln.d_start, &ln.cap_start, bs)) != MBIK_OK) ||
				(out_local && (rc = ensure_capacity(&ln.d_local, &ln.cap_local, bl)) != MBIK_OK) ||
				(out_status && (rc = ensure_capacity(&ln.d_status, &ln.cap_status, bst)) != MBIK_OK) ||
				(lim && (rc = ensure_capacity(&ln.d_index, &ln.cap_index, cn * sizeof(int32_t))) != MBIK_OK)) {
			for (Lane &l2 : ds->lanes) {
				cudaStreamSynchronize(l2.stream);
			}
			return rc;
		}
		cudaStream_t st = ln.stream;
		nvtxRangePushA("mbik H2D chunk");
		if (bt) {
			cudaMemcpyAsync(ln.d_targets, targets + b0 * np * 12, bt, cudaMemcpyHostToDevice, st);
		}
		if (start_pose) {
			cudaMemcpyAsync(ln.d_start, start_pose + b0 * nb * 12, bs, cudaMemcpyHostToDevice, st);
		}
		if (lim) {
			cudaMemcpyAsync(ln.d_index, lim->set_index + b0, cn * sizeof(int32_t), cudaMemcpyHostToDevice, st);
			a.limit_index = ln.d_index;
		}
		nvtxRangePop();
		a.n_poses = cn;
		a.targets = ln.d_targets;
		a.start_pose = start_pose ? ln.d_start : nullptr;
		a.out_pose = ln.d_out;
		a.out_local = out_local ? ln.d_local : nullptr;
		a.out_status = out_status ? ln.d_status : nullptr;
		const bool last_of_lane = c + kLanes >= n_chunks;
		nvtxRangePushA("mbik solve chunk");
		if (last_of_lane) {
			cudaEventRecord(ln.ev_start, st);
		}
		e = mbik::launch_solve(a, rig->variant, ds->sm_count, st);
		if (last_of_lane) {
			cudaEventRecord(ln.ev_stop, st);
		}
		nvtxRangePop();
		if (e != cudaSuccess) {
			for (Lane &l2 : ds->lanes) {
				cudaStreamSynchronize(l2.stream);
			}
			return cuda_fail(e, "kernel launch");
		}
		nvtxRangePushA("mbik D2H chunk");
		cudaMemcpyAsync(out_pose + b0 * out_rows * 10, ln.d_out, bo, cudaMemcpyDeviceToHost, st);
		if (out_local) {
			cudaMemcpyAsync(out_local + b0 * nb * 12, ln.d_local, bl, cudaMemcpyDeviceToHost, st);
		}
		if (out_status) {
			cudaMemcpyAsync(out_status + b0, ln.d_status, bst, cudaMemcpyDeviceToHost, st);
		}
		nvtxRangePop();
	}
	NvtxRange drain("mbik drain (stream synchronize)");
	cudaError_t first_err = cudaSuccess;
	for (Lane &ln : ds->lanes) {
		cudaError_t e2 = cudaStreamSynchronize(ln.stream);
		if (e2 != cudaSuccess && first_err == cudaSuccess) {
			first_err = e2;
		}
	}
	if (first_err != cudaSuccess) {
		return cuda_fail(first_err, "solve (stream synchronize)");
	}
	// kernel time of this call = the last chunk of each lane that ran one (an estimate when n_chunks > kLanes)
	{
		float total = 0.f;
		size_t timed_chunks = 0;
		for (size_t c = (n_chunks > (size_t)kLanes ? n_chunks - kLanes : 0); c < n_chunks; c++) {
			float ms = 0.f;
			Lane &ln = ds->lanes[c % kLanes];
			if (cudaEventElapsedTime(&ms, ln.ev_start, ln.ev_stop) == cudaSuccess) {
				total += ms;
				timed_chunks++;
			}
		}
		ds->host_path_kernel_ms = timed_chunks ? total * (float)n_chunks / (float)timed_chunks : -1.f;
		ds->last_was_host = true;
	}
	return MBIK_OK;
}

int check_solve_args(mbik_rig *rig, const mbik_solve_params *params, size_t n_poses, const float *targets, float *out_pose, int *iterations) {
	if (!rig) {
		return fail(MBIK_ERR_INVALID_ARG, "rig is NULL");
	}
	if (n_poses > 0 && (!out_pose || (!targets && !rig->flat.pins.empty()))) {
		return fail(MBIK_ERR_INVALID_ARG, "targets/out_pose must not be NULL");
	}
	int it = (params && params->iterations >= 0) ? params->iterations : rig->flat.iterations;
	if (it < 0) {
		it = 0;
	}
	// early-outs of ManyBoneIK3D::_process_modification (reference src/many_bone_ik_3d.cpp:649-651, :671-680): no pins,
	// or no pin with a non-empty bone name (bone < 0) -> nothing is solved, every bone passes through
	bool has_pins = false;
	for (const mbik_pin_desc &p : rig->flat.pins) {
		has_pins = has_pins || p.bone >= 0;
	}
	if (!has_pins) {
		it = 0;
	}
	*iterations = it;
	return MBIK_OK;
}

} // namespace

extern "C" {
#pragma GCC visibility push(default)

int mbik_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		return 0;
	}
	return n;
}

const char *mbik_strerror(int code) {
	switch (code) {
		case MBIK_OK: return "ok";
		case MBIK_ERR_INVALID_ARG: return "invalid argument";
		case MBIK_ERR_CUDA: return "CUDA error";
		case MBIK_ERR_UNSUPPORTED: return "unsupported rig";
		case MBIK_ERR_NO_DEVICE: return "no CUDA device (no CPU fallback)";
		case MBIK_ERR_ALLOC: return "allocation failure";
		default: return "unknown error";
	}
}

const char *mbik_last_error(void) { return g_last_error.c_str(); }

int mbik_rig_create(const mbik_rig_desc *desc, mbik_rig **out_rig) {
	if (!desc || !out_rig) {
		return fail(MBIK_ERR_INVALID_ARG, "desc/out_rig is NULL");
	}
	*out_rig = nullptr;
	mbik_rig *rig = new (std::nothrow) mbik_rig();
	if (!rig) {
		return fail(MBIK_ERR_ALLOC, "out of memory");
	}
	int rc = mbik::flatten_rig(desc, rig->flat);
	if (rc != MBIK_OK) {
		std::string msg = rig->flat.error;
		delete rig;
		return fail(rc, msg);
	}
	rig->desc.assign(desc);
	rig->n_solved = (int)rig->flat.bone_order.size();
	rig->variant = mbik::kernel_variant_for(rig->n_solved, rig->flat.max_seg_len, rig->flat.max_stack, rig->flat.blob.size());
	if (rig->variant == 5 && reinterpret_cast<const mbik::BlobHeader *>(rig->flat.blob.data())->resident_bytes > mbik::kResidentBlobBudget) {
		rig->variant = mbik::kDynVariant; // constants beyond the shared-memory budget even without the walk list: read in place
	}
	if (rig->variant < 0) {
		delete rig;
		return fail(MBIK_ERR_UNSUPPORTED, "rig exceeds the index types of the schedule (16383 solved bones, walk stack depth 127)");
	}
	{
		std::string verr;
		if (!mbik::validate_schedule(rig->flat, mbik::kVariants[rig->variant][0], mbik::kVariants[rig->variant][1], mbik::kVariants[rig->variant][2], verr)) {
			delete rig;
			return fail(MBIK_ERR_UNSUPPORTED, verr);
		}
	}
	if (desc->stabilization_passes > 0 && rig->variant != mbik::kDynVariant) {
		// the stabilisation variants keep one pre-step tip origin per effector of a list in max(32, bone capacity) slots; a list
		// has at most one effector per solved bone, so this cannot trip -- checked because the kernel does not
		const int cap = std::max(mbik::kMinStabEffectors, mbik::kVariants[rig->variant][0]);
		for (const mbik::FlatSegment &seg : rig->flat.segments) {
			if ((int)seg.effectors.size() > cap) {
				delete rig;
				return fail(MBIK_ERR_UNSUPPORTED, "internal: effector list longer than the kernel variant's bone capacity");
			}
		}
	}
	if (rig->variant != mbik::kDynVariant && reinterpret_cast<const mbik::BlobHeader *>(rig->flat.blob.data())->resident_bytes > mbik::kResidentBlobBudget) {
		delete rig;
		return fail(MBIK_ERR_UNSUPPORTED, "rig constants exceed the shared-memory budget (200 KiB without the walk list)");
	}
	*out_rig = rig;
	return MBIK_OK;
}

int mbik_rig_destroy(mbik_rig *rig) {
	if (!rig) {
		return MBIK_OK;
	}
	for (auto &kv : rig->devices) {
		free_device_state(*kv.second);
	}
	delete rig;
	return MBIK_OK;
}

int mbik_rig_get_info(const mbik_rig *rig, mbik_rig_info *o) {
	if (!rig || !o) {
		return fail(MBIK_ERR_INVALID_ARG, "NULL argument");
	}
	const mbik::FlatRig &F = rig->flat;
	o->n_bones = F.n_bones;
	o->n_solved = rig->n_solved;
	o->n_segments = F.n_kept_segments;
	o->n_steps = (int)F.steps.size();
	o->n_effectors = F.n_effectors;
	o->n_pins = (int)F.pins.size();
	o->max_headings = F.max_headings;
	o->n_cones = (int)F.cones.size();
	o->iterations = F.iterations;
	o->kernel_capacity = mbik::kernel_capacity_of_variant(rig->variant);
	o->rig_blob_bytes = (int64_t)F.blob.size();
	o->flops_per_solve = F.flops_per_solve;
	o->max_segment_len = F.max_seg_len;
	o->max_walk_stack = F.max_stack;
	o->sp_roles = F.sp_roles;
	o->sp_phases = F.sp_phases;
	o->sp_gain = F.sp_critical_cost > 0 ? F.sp_serial_cost / F.sp_critical_cost : 1.0;
	return MBIK_OK;
}

int mbik_rig_get_bone_order(const mbik_rig *rig, int32_t *out_bones) {
	if (!rig || !out_bones) {
		return fail(MBIK_ERR_INVALID_ARG, "NULL argument");
	}
	for (size_t i = 0; i < rig->flat.bone_order.size(); i++) {
		out_bones[i] = rig->flat.bone_order[i];
	}
	return MBIK_OK;
}

int mbik_rig_get_schedule(const mbik_rig *rig, int32_t *out_rows, int32_t capacity) {
	if (!rig) {
		return fail(MBIK_ERR_INVALID_ARG, "NULL argument");
	}
	const mbik::FlatRig &F = rig->flat;
	int n = 0;
	for (int ph = 0; ph < F.sp_phases; ph++) {
		for (int slot = 0; slot < F.sp_slots; slot++) {
			for (int r = 0; r < F.sp_roles; r++) {
				const mbik::BlobSpan sp = F.sched[((size_t)ph * F.sp_slots + slot) * F.sp_roles + r];
				if (sp.s0 == sp.s1) {
					continue;
				}
				if (out_rows && n < capacity) {
					int32_t *o = out_rows + (size_t)n * 6;
					o[0] = ph; o[1] = r; o[2] = sp.s0; o[3] = sp.s1; o[4] = sp.team; o[5] = sp.member;
				}
				n++;
			}
		}
	}
	return n;
}

int mbik_rig_get_step_weights(const mbik_rig *rig, int32_t step, double *out_weights, int32_t capacity) {
	if (!rig || step < 0 || step >= (int)rig->flat.bone_order.size()) {
		return fail(MBIK_ERR_INVALID_ARG, "bad step");
	}
	const mbik::FlatSegment &S = rig->flat.segments[rig->flat.seg_of_bone[rig->flat.bone_order[step]]];
	for (int i = 0; i < (int)S.weights.size() && i < capacity && out_weights; i++) {
		out_weights[i] = S.weights[i];
	}
	return (int)S.weights.size();
}

int mbik_rig_get_bone_frames(const mbik_rig *rig, float *out_dir_basis, float *out_twist_basis) {
	if (!rig) {
		return fail(MBIK_ERR_INVALID_ARG, "NULL argument");
	}
	const mbik::FlatRig &F = rig->flat;
	for (size_t i = 0; i < F.bone_order.size(); i++) {
		const mbik::BlobBone &B = F.bones[F.t_of_bone[F.bone_order[i]]];
		if (out_dir_basis) {
			memcpy(out_dir_basis + 9 * i, B.dir_basis, sizeof(float) * 9);
		}
		if (out_twist_basis) {
			memcpy(out_twist_basis + 9 * i, B.twist_basis, sizeof(float) * 9);
		}
	}
	return MBIK_OK;
}

int mbik_rig_get_cone_geometry(const mbik_rig *rig, float *out) {
	if (!rig || !out) {
		return fail(MBIK_ERR_INVALID_ARG, "NULL argument");
	}
	const mbik::FlatRig &F = rig->flat;
	for (size_t i = 0; i < F.cone_row_index.size(); i++) {
		float *o = out + 9 * i;
		int k = F.cone_row_index[i];
		if (k < 0) {
			memset(o, 0, sizeof(float) * 9);
			continue;
		}
		const mbik::BlobCone &c = F.cones[k];
		memcpy(o, c.cp, sizeof(float) * 3);
		memcpy(o + 3, c.tc1, sizeof(float) * 3);
		memcpy(o + 6, c.tc2, sizeof(float) * 3);
	}
	return (int)F.cone_row_index.size();
}

int mbik_solve_batch(mbik_rig *rig, const mbik_solve_params *params, size_t n_poses, const float *targets, const float *start_pose,
		float *out_pose, float *out_local, uint32_t *out_status) {
	int iterations = 0;
	int rc = check_solve_args(rig, params, n_poses, targets, out_pose, &iterations);
	if (rc != MBIK_OK) {
		return rc;
	}
	int device = -1;
	if ((rc = resolve_device(params, &device)) != MBIK_OK) {
		return rc;
	}
	uint32_t flags = params ? params->flags : MBIK_IO_HOST;
	cudaStream_t stream = params ? (cudaStream_t)params->stream : nullptr;
	return solve_on_device(rig, device, flags, stream, iterations, params ? params->newton_iters : 0, n_poses, targets, start_pose, out_pose, out_local, out_status);
}

namespace {

// Runs the reference's constraint authoring (author_constraints: IKKusudama3D::_update_constraint / set_axial_limits,
// IKLimitCone3D::update_tangent_handles on the host libm) for every set, on a pool of host threads: sets are independent.
int author_limit_sets(mbik_limit_sets *ls, const mbik_constraint_desc *constraints, const mbik_cone_desc *cones, int32_t cones_per_set) {
	mbik_rig *rig = ls->rig;
	const mbik::FlatRig &F = rig->flat;
	const size_t n_rows = rig->desc.constraints.size();
	const int32_t n_sets = ls->n_sets;
	const size_t cone_bytes = F.cones.size() * sizeof(mbik::BlobCone), bone_bytes = F.bones.size() * sizeof(mbik::BlobBone);
	int n_threads = (int)std::thread::hardware_concurrency();
	if (const char *env = getenv("MBIK_AUTHOR_THREADS")) {
		n_threads = atoi(env);
	}
	n_threads = n_threads < 1 ? 1 : (n_threads > n_sets ? n_sets : n_threads);
	ls->author_threads = n_threads;
	std::atomic<int32_t> next(0);
	std::atomic<int> first_rc(MBIK_OK);
	std::mutex err_mu;
	auto work = [&]() {
		for (;;) {
			const int32_t s = next.fetch_add(1);
			if (s >= n_sets || first_rc.load() != MBIK_OK) {
				return;
			}
			auto set_fail = [&](int code, const std::string &msg) {
				int expected = MBIK_OK;
				if (first_rc.compare_exchange_strong(expected, code)) {
					std::lock_guard<std::mutex> lock(err_mu);
					ls->error = msg;
				}
			};
			const mbik_constraint_desc *rows = constraints + (size_t)s * n_rows;
			bool rows_ok = true;
			for (size_t r = 0; r < n_rows && rows_ok; r++) {
				const mbik_constraint_desc &mine = rig->desc.constraints[r];
				if (rows[r].bone != mine.bone || rows[r].n_cones != mine.n_cones) {
					set_fail(MBIK_ERR_INVALID_ARG, "limit set " + std::to_string(s) + ", row " + std::to_string(r) +
							": bone and n_cones must equal the rig's constraint row (only values may vary)");
					rows_ok = false;
				} else if (rows[r].n_cones > 0 && (rows[r].cone_offset < 0 || rows[r].cone_offset + rows[r].n_cones > cones_per_set || !cones)) {
					set_fail(MBIK_ERR_INVALID_ARG, "limit set " + std::to_string(s) + ", row " + std::to_string(r) + ": cone_offset out of range");
					rows_ok = false;
				}
			}
			if (!rows_ok) {
				return;
			}
			// only the limit-dependent part of the flattening runs per set (author_constraints); rows name the rig's own bones
			// and cone counts (checked above), so steps, limit flags and cone ranges are the rig's by construction
			mbik_rig_desc d = rig->desc.view();
			d.constraints = rows;
			d.cones = cones ? cones + (size_t)s * cones_per_set : nullptr;
			std::vector<mbik::BlobBone> set_bones = F.bones;
			std::vector<mbik::BlobCone> set_cones;
			set_cones.reserve(F.cones.size());
			mbik::ConstraintTables tables;
			std::string err;
			int rc = mbik::author_constraints(&d, F, set_bones, set_cones, tables, err);
			if (rc != MBIK_OK) {
				set_fail(rc, "limit set " + std::to_string(s) + ": " + err);
				return;
			}
			if (set_cones.size() != F.cones.size()) {
				set_fail(MBIK_ERR_INVALID_ARG, "limit set " + std::to_string(s) + " changes the rig's schedule (which bones are limited, or cone counts)");
				return;
			}
			unsigned char *rec = ls->table.data() + (size_t)s * ls->stride;
			if (cone_bytes) {
				memcpy(rec, set_cones.data(), cone_bytes);
			}
			memcpy(rec + cone_bytes, set_bones.data(), bone_bytes);
		}
	};
	const auto t0 = std::chrono::steady_clock::now();
	std::vector<std::thread> pool;
	for (int t = 1; t < n_threads; t++) {
		pool.emplace_back(work);
	}
	work();
	for (auto &t : pool) {
		t.join();
	}
	ls->author_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	return first_rc.load();
}

int limit_sets_new(mbik_rig *rig, int32_t n_sets, const mbik_constraint_desc *constraints, int32_t cones_per_set, mbik_limit_sets **out_sets) {
	if (!rig || !out_sets || n_sets < 1 || cones_per_set < 0) {
		return fail(MBIK_ERR_INVALID_ARG, "rig/out_sets is NULL or n_sets < 1");
	}
	*out_sets = nullptr;
	const mbik::FlatRig &F = rig->flat;
	if (!rig->desc.constraints.empty() && !constraints) {
		return fail(MBIK_ERR_INVALID_ARG, "constraints is NULL");
	}
	std::unique_ptr<mbik_limit_sets> ls(new (std::nothrow) mbik_limit_sets());
	if (!ls) {
		return fail(MBIK_ERR_ALLOC, "out of memory");
	}
	ls->rig = rig;
	ls->n_sets = n_sets;
	ls->stride = (uint32_t)(F.cones.size() * sizeof(mbik::BlobCone) + F.bones.size() * sizeof(mbik::BlobBone));
	ls->table.resize((size_t)n_sets * ls->stride);
	*out_sets = ls.release();
	return MBIK_OK;
}

} // namespace

int mbik_limit_sets_create(mbik_rig *rig, int32_t n_sets, const mbik_constraint_desc *constraints, const mbik_cone_desc *cones,
		int32_t cones_per_set, mbik_limit_sets **out_sets) {
	int rc = limit_sets_new(rig, n_sets, constraints, cones_per_set, out_sets);
	if (rc != MBIK_OK) {
		return rc;
	}
	mbik_limit_sets *ls = *out_sets;
	NvtxRange range("mbik limit-set authoring");
	ls->rc = author_limit_sets(ls, constraints, cones, cones_per_set);
	if (ls->rc != MBIK_OK) {
		rc = fail(ls->rc, ls->error);
		delete ls;
		*out_sets = nullptr;
		return rc;
	}
	return MBIK_OK;
}

int mbik_limit_sets_create_async(mbik_rig *rig, int32_t n_sets, const mbik_constraint_desc *constraints, const mbik_cone_desc *cones,
		int32_t cones_per_set, mbik_limit_sets **out_sets) {
	int rc = limit_sets_new(rig, n_sets, constraints, cones_per_set, out_sets);
	if (rc != MBIK_OK) {
		return rc;
	}
	mbik_limit_sets *ls = *out_sets;
	const size_t n_rows = rig->desc.constraints.size();
	ls->own_constraints.assign(constraints, constraints + (size_t)n_sets * n_rows);
	if (cones && cones_per_set > 0) {
		ls->own_cones.assign(cones, cones + (size_t)n_sets * cones_per_set);
	}
	ls->worker = std::thread([ls, cones_per_set]() {
		ls->rc = author_limit_sets(ls, ls->own_constraints.data(), ls->own_cones.empty() ? nullptr : ls->own_cones.data(), cones_per_set);
		ls->own_constraints = std::vector<mbik_constraint_desc>();
		ls->own_cones = std::vector<mbik_cone_desc>();
	});
	return MBIK_OK;
}

int mbik_limit_sets_wait(mbik_limit_sets *sets) {
	if (!sets) {
		return fail(MBIK_ERR_INVALID_ARG, "sets is NULL");
	}
	{
		std::lock_guard<std::mutex> lock(sets->mu);
		if (sets->worker.joinable()) {
			sets->worker.join();
		}
	}
	return sets->rc == MBIK_OK ? MBIK_OK : fail(sets->rc, sets->error);
}

int mbik_limit_sets_get_info(mbik_limit_sets *sets, mbik_limit_sets_info *out) {
	if (!sets || !out) {
		return fail(MBIK_ERR_INVALID_ARG, "NULL argument");
	}
	int rc = mbik_limit_sets_wait(sets);
	out->n_sets = sets->n_sets;
	out->bytes_per_set = sets->stride;
	out->table_bytes = (int64_t)sets->table.size();
	out->author_seconds = sets->author_seconds;
	out->author_threads = sets->author_threads;
	return rc;
}

int mbik_limit_sets_get_geometry(mbik_limit_sets *sets, int32_t set, float *out_cones, float *out_twist_basis) {
	if (!sets) {
		return fail(MBIK_ERR_INVALID_ARG, "sets is NULL");
	}
	int rc = mbik_limit_sets_wait(sets);
	if (rc != MBIK_OK) {
		return rc;
	}
	if (set < 0 || set >= sets->n_sets) {
		return fail(MBIK_ERR_INVALID_ARG, "set index out of range");
	}
	const mbik::FlatRig &F = sets->rig->flat;
	const unsigned char *rec = sets->table.data() + (size_t)set * sets->stride;
	const mbik::BlobCone *set_cones = reinterpret_cast<const mbik::BlobCone *>(rec);
	const mbik::BlobBone *set_bones = reinterpret_cast<const mbik::BlobBone *>(rec + F.cones.size() * sizeof(mbik::BlobCone));
	for (size_t i = 0; out_cones && i < F.cone_row_index.size(); i++) {
		float *o = out_cones + 9 * i;
		const int k = F.cone_row_index[i];
		if (k < 0) {
			memset(o, 0, sizeof(float) * 9);
			continue;
		}
		memcpy(o, set_cones[k].cp, sizeof(float) * 3);
		memcpy(o + 3, set_cones[k].tc1, sizeof(float) * 3);
		memcpy(o + 6, set_cones[k].tc2, sizeof(float) * 3);
	}
	for (size_t i = 0; out_twist_basis && i < F.bone_order.size(); i++) {
		memcpy(out_twist_basis + 9 * i, set_bones[F.t_of_bone[F.bone_order[i]]].twist_basis, sizeof(float) * 9);
	}
	return (int)F.cone_row_index.size();
}

int mbik_limit_sets_destroy(mbik_limit_sets *sets) {
	if (!sets) {
		return MBIK_OK;
	}
	if (sets->worker.joinable()) {
		sets->worker.join();
	}
	DeviceGuard guard;
	for (auto &kv : sets->device_tables) {
		if (cudaSetDevice(kv.first) == cudaSuccess) {
			cudaFree(kv.second);
		}
	}
	delete sets;
	return MBIK_OK;
}

int mbik_solve_batch_limits(mbik_rig *rig, mbik_limit_sets *sets, const mbik_solve_params *params, size_t n_poses, const int32_t *set_index,
		const float *targets, const float *start_pose, float *out_pose, float *out_local, uint32_t *out_status) {
	int iterations = 0;
	int rc = check_solve_args(rig, params, n_poses, targets, out_pose, &iterations);
	if (rc != MBIK_OK) {
		return rc;
	}
	if (!sets || sets->rig != rig) {
		return fail(MBIK_ERR_INVALID_ARG, "sets is NULL or was created for another rig");
	}
	if (n_poses > 0 && !set_index) {
		return fail(MBIK_ERR_INVALID_ARG, "set_index must not be NULL");
	}
	if ((rc = mbik_limit_sets_wait(sets)) != MBIK_OK) { // asynchronous authoring still running / failed
		return rc;
	}
	int device = -1;
	if ((rc = resolve_device(params, &device)) != MBIK_OK) {
		return rc;
	}
	uint32_t flags = params ? params->flags : MBIK_IO_HOST;
	cudaStream_t stream = params ? (cudaStream_t)params->stream : nullptr;
	DeviceGuard guard;
	LimitArgs lim;
	{
		// per-device copy of the table, uploaded on first use
		std::lock_guard<std::mutex> lock(sets->mu);
		auto it = sets->device_tables.find(device);
		if (it == sets->device_tables.end()) {
			unsigned char *d = nullptr;
			cudaError_t e = cudaSetDevice(device);
			if (e == cudaSuccess) {
				e = cudaMalloc((void **)&d, sets->table.size());
			}
			if (e == cudaSuccess) {
				e = cudaMemcpy(d, sets->table.data(), sets->table.size(), cudaMemcpyHostToDevice);
			}
			if (e != cudaSuccess) {
				cudaFree(d);
				return cuda_fail(e, "limit-set table upload");
			}
			it = sets->device_tables.emplace(device, d).first;
		}
		lim.table = it->second;
	}
	lim.stride = sets->stride;
	lim.n_sets = sets->n_sets;
	lim.set_index = set_index;
	return solve_on_device(rig, device, flags, stream, iterations, params ? params->newton_iters : 0, n_poses, targets, start_pose, out_pose, out_local, out_status, &lim);
}

int mbik_solve_batch_multi(mbik_rig *rig, const mbik_solve_params *params, size_t n_poses, const float *targets, const float *start_pose,
		float *out_pose, float *out_local, uint32_t *out_status, const int32_t *devices, int32_t n_devices) {
	int iterations = 0;
	int rc = check_solve_args(rig, params, n_poses, targets, out_pose, &iterations);
	if (rc != MBIK_OK) {
		return rc;
	}
	int count = mbik_device_count();
	if (count <= 0) {
		return fail(MBIK_ERR_NO_DEVICE, "no CUDA device: mbik has no CPU fallback");
	}
	if (n_devices <= 0) {
		return fail(MBIK_ERR_INVALID_ARG, "n_devices must be > 0");
	}
	if (params && (params->flags & MBIK_IO_DEVICE)) {
		return fail(MBIK_ERR_INVALID_ARG, "mbik_solve_batch_multi takes host buffers only");
	}
	std::vector<int> devs(n_devices);
	for (int i = 0; i < n_devices; i++) {
		devs[i] = devices ? devices[i] : i;
		if (devs[i] < 0 || devs[i] >= count) {
			return fail(MBIK_ERR_INVALID_ARG, "device ordinal out of range");
		}
	}
	const size_t nb = (size_t)rig->flat.n_bones, np = rig->flat.pins.size();
	const uint32_t pass_flags = params ? params->flags & (MBIK_SCHED_THROUGHPUT | MBIK_SCHED_SEGMENT_PARALLEL | MBIK_OUT_SOLVED_ONLY | MBIK_LOCAL_RECOMPOSED) : 0u;
	const size_t out_rows = (pass_flags & MBIK_OUT_SOLVED_ONLY) ? rig->flat.bones.size() : nb;
	const int newton_iters = params ? params->newton_iters : 0;
	std::vector<int> rcs(n_devices, MBIK_OK);
	std::vector<std::string> msgs(n_devices);
	std::vector<std::thread> workers;
	for (int g = 0; g < n_devices; g++) {
		// contiguous split of the pose index range: device g gets [g*n/G, (g+1)*n/G)
		size_t b = n_poses * (size_t)g / (size_t)n_devices, e = n_poses * (size_t)(g + 1) / (size_t)n_devices;
		workers.emplace_back([&, g, b, e]() {
			rcs[g] = solve_on_device(rig, devs[g], MBIK_IO_HOST | pass_flags, nullptr, iterations, newton_iters, e - b, targets + b * np * 12,
					start_pose ? start_pose + b * nb * 12 : nullptr, out_pose + b * out_rows * 10, out_local ? out_local + b * nb * 12 : nullptr,
					out_status ? out_status + b : nullptr);
			if (rcs[g] != MBIK_OK) {
				msgs[g] = g_last_error;
			}
		});
	}
	for (auto &w : workers) {
		w.join();
	}
	for (int g = 0; g < n_devices; g++) {
		if (rcs[g] != MBIK_OK) {
			return fail(rcs[g], msgs[g]);
		}
	}
	return MBIK_OK;
}

int mbik_stream_destroy(mbik_stream *st) {
	if (!st) {
		return MBIK_OK;
	}
	DeviceGuard guard;
	if (st->device >= 0) {
		cudaSetDevice(st->device);
		if (st->s_solve) {
			cudaStreamSynchronize(st->s_up);
			cudaStreamSynchronize(st->s_solve);
			cudaStreamSynchronize(st->s_down);
		}
		for (int i = 0; i < 2; i++) {
			cudaFree(st->local[i]);
			cudaFree(st->d_targets[i]);
			cudaFree(st->d_out[i]);
			cudaFree(st->d_status[i]);
			if (st->ev_up[i]) {
				cudaEventDestroy(st->ev_up[i]);
			}
			if (st->ev_solved[i]) {
				cudaEventDestroy(st->ev_solved[i]);
			}
			if (st->ev_down[i]) {
				cudaEventDestroy(st->ev_down[i]);
			}
		}
		if (st->s_up) {
			cudaStreamDestroy(st->s_up);
		}
		if (st->s_solve) {
			cudaStreamDestroy(st->s_solve);
		}
		if (st->s_down) {
			cudaStreamDestroy(st->s_down);
		}
	}
	delete st;
	return MBIK_OK;
}

int mbik_stream_reset(mbik_stream *st, const float *initial_pose) {
	if (!st) {
		return fail(MBIK_ERR_INVALID_ARG, "stream is NULL");
	}
	DeviceGuard guard;
	cudaError_t e = cudaSetDevice(st->device);
	if (e != cudaSuccess) {
		return cuda_fail(e, "cudaSetDevice");
	}
	int rc = mbik_stream_sync(st);
	if (rc != MBIK_OK) {
		return rc;
	}
	const mbik::FlatRig &F = st->rig->flat;
	const size_t row = (size_t)F.n_bones * 12, bytes = st->n_poses * row * sizeof(float);
	if (initial_pose) {
		e = cudaMemcpy(st->local[st->cur], initial_pose, bytes, cudaMemcpyHostToDevice);
	} else {
		// rest pose for every skeleton: replicate one row (rest_local is the tail section of the blob)
		std::vector<float> rest(row);
		for (int b = 0; b < F.n_bones; b++) {
			memcpy(&rest[(size_t)b * 12], &F.rest_local[b], sizeof(float) * 12);
		}
		const size_t chunk = 4096;
		std::vector<float> tile(chunk * row);
		for (size_t k = 0; k < chunk; k++) {
			memcpy(&tile[k * row], rest.data(), row * sizeof(float));
		}
		for (size_t k = 0; k < st->n_poses && e == cudaSuccess; k += chunk) {
			size_t c = st->n_poses - k < chunk ? st->n_poses - k : chunk;
			e = cudaMemcpy(st->local[st->cur] + k * row, tile.data(), c * row * sizeof(float), cudaMemcpyHostToDevice);
		}
	}
	if (e != cudaSuccess) {
		return cuda_fail(e, "stream reset");
	}
	return MBIK_OK;
}

int mbik_stream_create(mbik_rig *rig, int32_t device, size_t n_poses, const float *initial_pose, mbik_stream **out_stream) {
	return mbik_stream_create_ex(rig, device, n_poses, initial_pose, 0u, out_stream);
}

int mbik_stream_create_ex(mbik_rig *rig, int32_t device, size_t n_poses, const float *initial_pose, uint32_t flags, mbik_stream **out_stream) {
	if (!rig || !out_stream || n_poses == 0) {
		return fail(MBIK_ERR_INVALID_ARG, "rig/out_stream is NULL or n_poses is 0");
	}
	if (flags & ~MBIK_OUT_SOLVED_ONLY) {
		return fail(MBIK_ERR_INVALID_ARG, "mbik_stream_create_ex: only MBIK_OUT_SOLVED_ONLY is a stream flag");
	}
	*out_stream = nullptr;
	mbik_solve_params p = {};
	p.iterations = -1;
	p.device = device;
	int dev = -1;
	int rc = resolve_device(&p, &dev);
	if (rc != MBIK_OK) {
		return rc;
	}
	DeviceGuard guard;
	cudaError_t e = cudaSetDevice(dev);
	if (e != cudaSuccess) {
		return cuda_fail(e, "cudaSetDevice");
	}
	DeviceState *ds = nullptr;
	if ((rc = get_device_state(rig, dev, &ds)) != MBIK_OK) {
		return rc;
	}
	mbik_stream *st = new (std::nothrow) mbik_stream();
	if (!st) {
		return fail(MBIK_ERR_ALLOC, "out of memory");
	}
	st->rig = rig;
	st->device = dev;
	st->sm_count = ds->sm_count;
	st->n_poses = n_poses;
	st->blob = ds->blob;
	const mbik::FlatRig &F = rig->flat;
	const size_t nb = (size_t)F.n_bones, np = F.pins.size();
	st->flags = flags;
	st->out_rows = (flags & MBIK_OUT_SOLVED_ONLY) ? F.bones.size() : nb;
	for (int i = 0; i < 2 && e == cudaSuccess; i++) {
		e = cudaMalloc((void **)&st->local[i], n_poses * nb * 12 * sizeof(float));
		if (e == cudaSuccess) {
			e = cudaMalloc((void **)&st->d_targets[i], (np ? n_poses * np * 12 : 4) * sizeof(float));
		}
		if (e == cudaSuccess) {
			e = cudaMalloc((void **)&st->d_out[i], (st->out_rows ? n_poses * st->out_rows * 10 : 4) * sizeof(float));
		}
		if (e == cudaSuccess) {
			e = cudaMalloc((void **)&st->d_status[i], n_poses * sizeof(uint32_t));
		}
		if (e == cudaSuccess) {
			e = cudaEventCreateWithFlags(&st->ev_up[i], cudaEventDisableTiming);
		}
		if (e == cudaSuccess) {
			e = cudaEventCreateWithFlags(&st->ev_solved[i], cudaEventDisableTiming);
		}
		if (e == cudaSuccess) {
			e = cudaEventCreateWithFlags(&st->ev_down[i], cudaEventDisableTiming);
		}
	}
	if (e == cudaSuccess) {
		e = cudaStreamCreateWithFlags(&st->s_up, cudaStreamNonBlocking);
	}
	if (e == cudaSuccess) {
		e = cudaStreamCreateWithFlags(&st->s_solve, cudaStreamNonBlocking);
	}
	if (e == cudaSuccess) {
		e = cudaStreamCreateWithFlags(&st->s_down, cudaStreamNonBlocking);
	}
	if (e != cudaSuccess) {
		mbik_stream_destroy(st);
		return cuda_fail(e, "stream allocation");
	}
	rc = mbik_stream_reset(st, initial_pose);
	if (rc != MBIK_OK) {
		mbik_stream_destroy(st);
		return rc;
	}
	*out_stream = st;
	return MBIK_OK;
}

int mbik_stream_submit(mbik_stream *st, const float *targets, float *out_pose, uint32_t *out_status, int32_t iterations) {
	if (!st || (!targets && !st->rig->flat.pins.empty())) {
		return fail(MBIK_ERR_INVALID_ARG, "stream/targets is NULL");
	}
	DeviceGuard guard;
	NvtxRange range("mbik stream submit");
	cudaError_t e = cudaSetDevice(st->device);
	if (e != cudaSuccess) {
		return cuda_fail(e, "cudaSetDevice");
	}
	const mbik::FlatRig &F = st->rig->flat;
	const size_t np = F.pins.size(), n = st->n_poses;
	const int slot = (int)(st->frames & 1);
	// the slot's previous frame (f-2) must have finished downloading before its buffers are overwritten
	if (st->slot_used[slot]) {
		cudaStreamWaitEvent(st->s_up, st->ev_solved[slot], 0);   // targets[slot] no longer read
		cudaStreamWaitEvent(st->s_solve, st->ev_down[slot], 0);  // out[slot] no longer being copied
	}
	if (np) {
		cudaMemcpyAsync(st->d_targets[slot], targets, n * np * 12 * sizeof(float), cudaMemcpyHostToDevice, st->s_up);
	}
	cudaEventRecord(st->ev_up[slot], st->s_up);
	cudaStreamWaitEvent(st->s_solve, st->ev_up[slot], 0);
	mbik::SolveArgs a;
	a.blob = st->blob;
	a.blob_bytes = reinterpret_cast<const mbik::BlobHeader *>(F.blob.data())->resident_bytes; // what the kernel stages into shared memory
	a.iterations = iterations >= 0 ? iterations : F.iterations;
	{
		bool has_pins = false; // same early-out as mbik_solve_batch (reference src/many_bone_ik_3d.cpp:649-651, :671-680)
		for (const mbik_pin_desc &p : F.pins) {
			has_pins = has_pins || p.bone >= 0;
		}
		if (!has_pins) {
			a.iterations = 0;
		}
	}
	a.n_poses = n;
	a.stabilize = F.stabilization_passes > 0 ? 1 : 0;
	set_launch_hints(a, F, 0);
	a.targets = st->d_targets[slot];
	a.start_pose = st->local[st->cur];
	a.out_pose = st->d_out[slot];
	a.out_local = st->local[st->cur ^ 1];
	a.out_status = st->d_status[slot];
	// The next frame seeds from what the skeleton hands back after this frame's write-back: position / rotation / scale
	// recomposed (with the non-finite reset), not the raw IK-bone transforms -- see MBIK_LOCAL_RECOMPOSED.
	a.out_flags = mbik::OUT_LOCAL_RECOMPOSED | ((st->flags & MBIK_OUT_SOLVED_ONLY) ? mbik::OUT_COMPACT : 0u);
	e = mbik::launch_solve(a, st->rig->variant, st->sm_count, st->s_solve);
	if (e != cudaSuccess) {
		return cuda_fail(e, "kernel launch");
	}
	cudaEventRecord(st->ev_solved[slot], st->s_solve);
	st->cur ^= 1;
	cudaStreamWaitEvent(st->s_down, st->ev_solved[slot], 0);
	if (out_pose) {
		cudaMemcpyAsync(out_pose, st->d_out[slot], n * st->out_rows * 10 * sizeof(float), cudaMemcpyDeviceToHost, st->s_down);
	}
	if (out_status) {
		cudaMemcpyAsync(out_status, st->d_status[slot], n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st->s_down);
	}
	cudaEventRecord(st->ev_down[slot], st->s_down);
	st->slot_used[slot] = true;
	st->frames++;
	e = cudaGetLastError();
	if (e != cudaSuccess) {
		return cuda_fail(e, "stream submit");
	}
	return MBIK_OK;
}

int mbik_stream_sync(mbik_stream *st) {
	if (!st) {
		return fail(MBIK_ERR_INVALID_ARG, "stream is NULL");
	}
	DeviceGuard guard;
	cudaError_t e = cudaSetDevice(st->device);
	if (e == cudaSuccess) {
		e = cudaStreamSynchronize(st->s_up);
	}
	if (e == cudaSuccess) {
		e = cudaStreamSynchronize(st->s_solve);
	}
	if (e == cudaSuccess) {
		e = cudaStreamSynchronize(st->s_down);
	}
	if (e != cudaSuccess) {
		return cuda_fail(e, "stream sync");
	}
	return MBIK_OK;
}

int mbik_stream_read_local(mbik_stream *st, float *out_local) {
	if (!st || !out_local) {
		return fail(MBIK_ERR_INVALID_ARG, "NULL argument");
	}
	int rc = mbik_stream_sync(st);
	if (rc != MBIK_OK) {
		return rc;
	}
	DeviceGuard guard;
	cudaSetDevice(st->device);
	cudaError_t e = cudaMemcpy(out_local, st->local[st->cur], st->n_poses * (size_t)st->rig->flat.n_bones * 12 * sizeof(float), cudaMemcpyDeviceToHost);
	if (e != cudaSuccess) {
		return cuda_fail(e, "read_local");
	}
	return MBIK_OK;
}

int64_t mbik_stream_frames(const mbik_stream *st) { return st ? st->frames : 0; }

namespace {
// tiny RAII device buffer for the stage probes
struct DevBuf {
	void *p = nullptr;
	cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
	~DevBuf() { cudaFree(p); }
};
} // namespace

int mbik_stage_qcp(int32_t device, int32_t n, const float *moved, const float *target, const double *weight, int32_t translate, float *out7) {
	return mbik_stage_qcp_newton(device, n, moved, target, weight, translate, 0, out7);
}

int mbik_stage_qcp_newton(int32_t device, int32_t n, const float *moved, const float *target, const double *weight, int32_t translate,
		int32_t newton_iters, float *out7) {
	if (n < 0 || !out7 || (n > 0 && (!moved || !target || !weight))) {
		return fail(MBIK_ERR_INVALID_ARG, "NULL argument");
	}
	if (mbik_device_count() <= 0) {
		return fail(MBIK_ERR_NO_DEVICE, "no CUDA device: mbik has no CPU fallback");
	}
	DeviceGuard guard;
	cudaError_t e = cudaSetDevice(device);
	DevBuf dm, dt, dw, dout;
	if (e == cudaSuccess) e = dm.alloc(sizeof(float) * 3 * n);
	if (e == cudaSuccess) e = dt.alloc(sizeof(float) * 3 * n);
	if (e == cudaSuccess) e = dw.alloc(sizeof(double) * n);
	if (e == cudaSuccess) e = dout.alloc(sizeof(float) * 7);
	if (e == cudaSuccess && n) e = cudaMemcpy(dm.p, moved, sizeof(float) * 3 * n, cudaMemcpyHostToDevice);
	if (e == cudaSuccess && n) e = cudaMemcpy(dt.p, target, sizeof(float) * 3 * n, cudaMemcpyHostToDevice);
	if (e == cudaSuccess && n) e = cudaMemcpy(dw.p, weight, sizeof(double) * n, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = mbik::launch_stage_qcp(n, (const float *)dm.p, (const float *)dt.p, (const double *)dw.p, translate, newton_iters > 0 ? newton_iters : 0, (float *)dout.p);
	if (e == cudaSuccess) e = cudaMemcpy(out7, dout.p, sizeof(float) * 7, cudaMemcpyDeviceToHost);
	return e == cudaSuccess ? MBIK_OK : cuda_fail(e, "mbik_stage_qcp");
}

int mbik_stage_clamp(int32_t device, int32_t n, const float *quats, const double *cos_half, float *out) {
	if (n < 0 || (n > 0 && (!quats || !cos_half || !out))) {
		return fail(MBIK_ERR_INVALID_ARG, "NULL argument");
	}
	if (n == 0) {
		return MBIK_OK;
	}
	if (mbik_device_count() <= 0) {
		return fail(MBIK_ERR_NO_DEVICE, "no CUDA device: mbik has no CPU fallback");
	}
	DeviceGuard guard;
	cudaError_t e = cudaSetDevice(device);
	DevBuf dq, dc, dout;
	if (e == cudaSuccess) e = dq.alloc(sizeof(float) * 4 * n);
	if (e == cudaSuccess) e = dc.alloc(sizeof(double) * n);
	if (e == cudaSuccess) e = dout.alloc(sizeof(float) * 4 * n);
	if (e == cudaSuccess) e = cudaMemcpy(dq.p, quats, sizeof(float) * 4 * n, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMemcpy(dc.p, cos_half, sizeof(double) * n, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = mbik::launch_stage_clamp(n, (const float *)dq.p, (const double *)dc.p, (float *)dout.p);
	if (e == cudaSuccess) e = cudaMemcpy(out, dout.p, sizeof(float) * 4 * n, cudaMemcpyDeviceToHost);
	return e == cudaSuccess ? MBIK_OK : cuda_fail(e, "mbik_stage_clamp");
}

int mbik_stage_point_in_limits(mbik_rig *rig, int32_t device, int32_t bone, int32_t n, const float *points, float *out) {
	if (!rig || n < 0 || (n > 0 && (!points || !out))) {
		return fail(MBIK_ERR_INVALID_ARG, "NULL argument");
	}
	const mbik::FlatRig &F = rig->flat;
	if (bone < 0 || bone >= F.n_bones || F.t_of_bone[bone] < 0) {
		return fail(MBIK_ERR_INVALID_ARG, "bone is not a solved bone of this rig");
	}
	const mbik::BlobStep *S = nullptr;
	for (const mbik::BlobStep &st : F.steps) {
		if (st.bone == F.t_of_bone[bone]) {
			S = &st;
		}
	}
	if (!S) {
		return fail(MBIK_ERR_INVALID_ARG, "bone has no step");
	}
	if (n == 0) {
		return MBIK_OK;
	}
	if (mbik_device_count() <= 0) {
		return fail(MBIK_ERR_NO_DEVICE, "no CUDA device: mbik has no CPU fallback");
	}
	DeviceGuard guard;
	cudaError_t e = cudaSetDevice(device);
	DevBuf dc, dp, dout;
	if (e == cudaSuccess) e = dc.alloc(sizeof(mbik::BlobCone) * S->cone_cnt);
	if (e == cudaSuccess) e = dp.alloc(sizeof(float) * 3 * n);
	if (e == cudaSuccess) e = dout.alloc(sizeof(float) * 4 * n);
	if (e == cudaSuccess && S->cone_cnt) e = cudaMemcpy(dc.p, F.cones.data() + S->cone_off, sizeof(mbik::BlobCone) * S->cone_cnt, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMemcpy(dp.p, points, sizeof(float) * 3 * n, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = mbik::launch_stage_point_in_limits((const mbik::BlobCone *)dc.p, S->cone_cnt, n, (const float *)dp.p, (float *)dout.p);
	if (e == cudaSuccess) e = cudaMemcpy(out, dout.p, sizeof(float) * 4 * n, cudaMemcpyDeviceToHost);
	return e == cudaSuccess ? MBIK_OK : cuda_fail(e, "mbik_stage_point_in_limits");
}

void *mbik_alloc_pinned(size_t bytes) {
	void *p = nullptr;
	if (cudaMallocHost(&p, bytes) != cudaSuccess) {
		g_last_error = "cudaMallocHost failed";
		return nullptr;
	}
	return p;
}

void mbik_free_pinned(void *p) {
	if (p) {
		cudaFreeHost(p);
	}
}

int mbik_last_kernel_ms(mbik_rig *rig, int32_t device, float *out_ms) {
	if (!rig || !out_ms) {
		return fail(MBIK_ERR_INVALID_ARG, "NULL argument");
	}
	std::lock_guard<std::mutex> lock(rig->mu);
	auto it = rig->devices.find(device);
	if (it != rig->devices.end() && it->second->last_was_host && it->second->host_path_kernel_ms >= 0.f) {
		*out_ms = it->second->host_path_kernel_ms;
		return MBIK_OK;
	}
	if (it == rig->devices.end() || !it->second->timed) {
		return fail(MBIK_ERR_INVALID_ARG, "no launch recorded on that device");
	}
	DeviceGuard guard;
	cudaSetDevice(device);
	cudaError_t e = cudaEventSynchronize(it->second->ev_stop);
	if (e == cudaSuccess) {
		e = cudaEventElapsedTime(out_ms, it->second->ev_start, it->second->ev_stop);
	}
	if (e != cudaSuccess) {
		return cuda_fail(e, "event timing");
	}
	return MBIK_OK;
}

#pragma GCC visibility pop
} // extern "C"
