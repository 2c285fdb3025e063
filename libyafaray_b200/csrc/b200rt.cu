// libyafaray_b200/csrc/b200rt.cu -- the C ABI of libb200rt (include/b200rt.h): scene ownership, host-side
// flattening of the kd-tree into the HBM layout the kernels read, and the staged host-buffer queries.
#include "../../include/b200rt.h"
#include "kd_build.h"
#include "kd_kernels.cuh"
#include "kd_segments.cuh"

#include <algorithm>
#include <fcntl.h>
#include <sys/file.h>
#include <unistd.h>
#include <array>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const std::string &msg)
{
	g_last_error = msg;
	return code;
}

#define CUDA_TRY(expr)                                                                                          \
	do {                                                                                                        \
		const cudaError_t e_ = (expr);                                                                          \
		if(e_ != cudaSuccess)                                                                                   \
			return fail(B200RT_E_CUDA, std::string(#expr) + ": " + cudaGetErrorName(e_) + " (" + cudaGetErrorString(e_) + ")"); \
	} while(0)

constexpr uint32_t kTriangle = 0xFFFFFFFFu;
constexpr uint32_t kSphere = 0xFFFFFFFEu; // idx[2] of a sphere face (kd_build.h)
constexpr uint32_t kBox = 0xFFFFFFFDu;    // idx[2] of a motion-blur face: the builder only sees its bound (kd_build.h)
// staging granularity of the host-buffer queries and chunks in flight per call; the environment overrides are tuning aids
// (tools/gpu_e2e_sweep.sh), read once.  Defaults from the measured sweeps (profiles/r1m_e2e_sweep.txt, r5d_e2e_sweep.txt): 24 MiB x 6 in flight
// (round 1: 16 MiB; with the two-pass kernels 24 MiB is 4 % faster on the bench's 512 MiB batches -- 21.3 chunks, a short last one).
constexpr size_t kTwoPassRaysForward = size_t(1) << 15; // = kTwoPassRays below: tail pieces stay large enough for the two-pass path
size_t chunkBytes()
{
	static const size_t value = [] { const char *e = std::getenv("B200RT_CHUNK_MB"); const long mb = e ? std::atol(e) : 0; return size_t(mb > 0 ? mb : 24) << 20; }();
	return value;
}
int lanesPerCall()
{
	static const int value = [] { const char *e = std::getenv("B200RT_LANES"); const long n = e ? std::atol(e) : 0; return int(n > 0 && n <= 16 ? n : 6); }();
	return value;
}
#ifndef B200RT_L2_PERSIST
#define B200RT_L2_PERSIST 0
#endif
constexpr size_t kSmallBatchRays = 256;           // unpinned batches up to this size go through a pinned lane buffer, one cursor-less launch
constexpr size_t kDirectRays = size_t(1) << 16;  // pinned batches up to this size are traced in place (no staging copies)

// One staging lane: a stream with pinned host and device buffers for rays in / results out.
struct Lane
{
	cudaStream_t stream = nullptr;
	cudaEvent_t done = nullptr;
	void *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
	float *d_times = nullptr; // ray times of the chunk (timed queries), in_cap / sizeof(b200rt_ray) floats
	size_t in_cap = 0, out_cap = 0, times_cap = 0;
	// pending copy-out of the previous chunk this lane carried
	void *pending_dst = nullptr;
	size_t pending_bytes = 0;
	bool pending_direct = false;

	~Lane()
	{
		if(h_in) cudaFreeHost(h_in);
		if(h_out) cudaFreeHost(h_out);
		if(d_in) cudaFree(d_in);
		if(d_out) cudaFree(d_out);
		if(d_times) cudaFree(d_times);
		if(done) cudaEventDestroy(done);
		if(stream) cudaStreamDestroy(stream);
	}
};

} // namespace

// ---- flush combiner (kd_segments.cuh) ----------------------------------------------------------------------------------
// The in-place jobs of b200rt_trace_jobs_begin -- the flushes of the renderer's ray queues, a few hundred rays each, from sixteen
// threads -- gather in the scene's pending list and leave in ONE launch: the thread whose job brings the list to
// B200RT_COMBINE_RAYS rays launches it, and so does a thread that comes to b200rt_trace_jobs_end while its job is still waiting
// (nobody waits for a timer, and no extra thread competes with the render threads for a core: a first version with a worker
// thread and a 60 us window was starved by the sixteen busy render threads and DOUBLED the frame time, profiles/r4t_*).  A ticket
// is what the flight keeps of such a job: `batch` is set once the launch that carries the job has been issued.
namespace b200rt {
struct CombinedBatch
{
	cudaEvent_t event = nullptr;
	std::atomic<int> users{0};
	int rc = B200RT_OK;
	std::string error;
};
struct Ticket
{
	std::atomic<CombinedBatch *> batch{nullptr};
};
struct PendingJob
{
	b200rt_job job;
	std::shared_ptr<Ticket> ticket;
};
struct Combiner
{
	std::mutex mutex;
	std::vector<PendingJob> pending;      // guarded by `mutex`
	size_t pending_rays = 0;
	std::vector<cudaStream_t> streams;
	std::atomic<size_t> next_stream{0};
	std::vector<CombinedBatch *> free_batches; // guarded by `mutex`
	std::atomic<uint64_t> launches{0}, jobs{0}, rays{0};
};
} // namespace b200rt

struct b200rt_scene
{
	int device = 0;
	b200rt::BuildConfig config;
	// host mesh (all b200rt_add_mesh calls concatenated)
	std::vector<float> xyz;
	std::vector<uint32_t> idx;
	std::vector<uint8_t> flags;
	// built state
	bool built = false;
	bool has_spheres = false;               // selects the kernel variant with the branches for spheres and motion-blur faces
	// motion blur: kind per face (0 static or sphere, 1 Bezier face, 2 face of a moving instance) and, for kinds 1 and 2, the record
	// the flattener emits (b200rt_add_mesh_bezier / _moving); the builder sees such a face as a box (kBox) over its bound
	struct MotionFace { uint32_t nv; float v[3][4][3]; float t0, t1; uint32_t matrix; };
	std::vector<uint8_t> kind;
	std::vector<int32_t> motion_of_face;
	std::vector<MotionFace> motion;
	std::vector<std::array<float, 50>> matrices; // per moving instance: 3 x 16 floats row major, then time_start, time_end
	float4 *d_inst = nullptr;
	b200rt::HostTree tree;
	std::vector<uint32_t> record_of_ref; // float4 offset of every leaf reference's record (for flag updates)
	uint2 *d_nodes = nullptr;               // (points into d_scene)
	void *d_scene = nullptr;                // nodes and leaf records in ONE allocation: one L2 access-policy window can cover both (B200RT_L2_PERSIST)
	size_t scene_bytes = 0;
	cudaAccessPolicyWindow l2_window{};     // valid when l2_window.num_bytes != 0
	uint4 *d_treelets = nullptr;            // B200RT_TREELET: two-level treelets of polygon-only scenes (kd_kernels.cuh, kEmptyRef)
	uint2 *d_spill = nullptr;               // ... and the per-thread overflow area of the short stack
	float4 *d_tris = nullptr;
	size_t n_tri_vec4 = 0;
	b200rt::SceneView view{};
	b200rt_stats stats{};
	// persistent-kernel launch state
	uint32_t *d_cursors = nullptr;          // ring of ray cursors, one per launch in flight
	std::atomic<uint32_t> next_cursor{0};
	int resident_blocks[3] = {0, 0, 0};     // blocks of traceKernel<Q> that fit the whole device
	int resident_blocks_queued[3] = {0, 0, 0}; // the same for the queue-fed variant of the two-pass path
	int resident_blocks_treelet[3] = {0, 0, 0}; // ... and for its treelet variant
	int setup_blocks = 0;                   // blocks of setupKernel for one resident wave
	std::mutex lane_mutex;
	std::vector<std::unique_ptr<Lane>> free_lanes;
	std::unique_ptr<b200rt::Combiner> combiner;     // made by b200rt_build unless B200RT_COMBINE=0
	void stopCombiner();

	~b200rt_scene()
	{
		cudaSetDevice(device);
		stopCombiner();
		free_lanes.clear();
		if(d_scene) cudaFree(d_scene);
		if(d_treelets) cudaFree(d_treelets);
		if(d_spill) cudaFree(d_spill);
		if(d_inst) cudaFree(d_inst);
		if(d_cursors) cudaFree(d_cursors);
	}
};

// Jobs enqueued by b200rt_trace_jobs_begin and not yet waited for.
struct b200rt_flight
{
	struct Entry { b200rt_scene *scene; std::unique_ptr<Lane> lane; };
	std::vector<Entry> lanes;
	struct Combined { b200rt_scene *scene; std::shared_ptr<b200rt::Ticket> ticket; };
	std::vector<Combined> tickets; // jobs that travel with the scene's flush combiner
	int first_error = B200RT_OK;
	std::string error_text;
	void note(int rc) { if(first_error == B200RT_OK) { first_error = rc; error_text = g_last_error; } }
};

namespace {

int ensureLane(Lane &l, size_t in_bytes, size_t out_bytes, bool need_h_in, bool need_h_out)
{
	if(!l.stream) CUDA_TRY(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
	if(!l.done) CUDA_TRY(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
	if(l.in_cap < in_bytes)
	{
		if(l.d_in) cudaFree(l.d_in);
		if(l.h_in) cudaFreeHost(l.h_in);
		l.d_in = l.h_in = nullptr;
		l.in_cap = 0;
		CUDA_TRY(cudaMalloc(&l.d_in, in_bytes));
		l.in_cap = in_bytes;
	}
	if(need_h_in && !l.h_in) CUDA_TRY(cudaMallocHost(&l.h_in, l.in_cap));
	if(l.out_cap < out_bytes)
	{
		if(l.d_out) cudaFree(l.d_out);
		if(l.h_out) cudaFreeHost(l.h_out);
		l.d_out = l.h_out = nullptr;
		l.out_cap = 0;
		CUDA_TRY(cudaMalloc(&l.d_out, out_bytes));
		l.out_cap = out_bytes;
	}
	if(need_h_out && !l.h_out) CUDA_TRY(cudaMallocHost(&l.h_out, l.out_cap));
	return B200RT_OK;
}

// device buffer for the ray times of one chunk (timed queries only)
int ensureLaneTimes(Lane &l, size_t n_rays)
{
	if(l.times_cap >= n_rays) return B200RT_OK;
	if(l.d_times) cudaFree(l.d_times);
	l.d_times = nullptr;
	l.times_cap = 0;
	CUDA_TRY(cudaMalloc(&l.d_times, n_rays * sizeof(float)));
	l.times_cap = n_rays;
	return B200RT_OK;
}

// Staging copy between a caller's pageable buffer and a pinned lane buffer.  One thread moves ~10 GB/s, a fifth of what the
// copy engine then needs; chunks of 4 MiB and more are split over a few short-lived threads (the caller's thread takes a share).
// Measured (tools/e2e_pageable.py, profiles/r2i_pageable.jsonl): 213 Mrays/s with the caller's thread alone, 549 with 5 helpers;
// pinned buffers (b200rt_host_alloc) remain the fast path at 1570 Mrays/s.
void stagingCopy(void *dst, const void *src, size_t bytes)
{
	constexpr size_t kSplitAbove = size_t(4) << 20;
	static const unsigned helpers = [] {
		const char *e = std::getenv("B200RT_STAGING_HELPERS"); // tuning aid (tools/e2e_pageable.py), read once
		if(e) return unsigned(std::min(7L, std::max(0L, std::atol(e))));
		return std::min(5u, std::max(1u, std::thread::hardware_concurrency() / 2) - 1u);
	}();
	if(bytes < kSplitAbove || helpers == 0u) { std::memcpy(dst, src, bytes); return; }
	const size_t parts = helpers + 1u, share = ((bytes / parts) + 4095) & ~size_t(4095);
	std::thread workers[7];
	unsigned started = 0;
	for(; started < helpers; ++started)
	{
		const size_t begin = (started + 1u) * share;
		if(begin >= bytes) break;
		const size_t count = std::min(share, bytes - begin);
		try { workers[started] = std::thread([=]() { std::memcpy(static_cast<char *>(dst) + begin, static_cast<const char *>(src) + begin, count); }); }
		catch(const std::system_error &) { std::memcpy(static_cast<char *>(dst) + begin, static_cast<const char *>(src) + begin, bytes - begin); break; }
	}
	std::memcpy(dst, src, std::min(share, bytes));
	for(unsigned k = 0; k < started; ++k) if(workers[k].joinable()) workers[k].join();
}

bool isPinned(const void *p)
{
	cudaPointerAttributes a;
	if(cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
	return a.type == cudaMemoryTypeHost;
}

// `times` (optional) = one ray time per ray, in host memory like the rays; it travels to the device with them.
template <typename Out, typename LaunchFn>
int tracedStaged(b200rt_scene *s, const b200rt_ray *rays, const float *times, size_t n, Out *out, LaunchFn launch, bool known_pinned = false)
{
	if(!s || (!rays && n) || (!out && n)) return fail(B200RT_E_INVALID, "null argument");
	if(!s->built) return fail(B200RT_E_INVALID, "scene not built: call b200rt_build first");
	if(n == 0) return B200RT_OK;
	CUDA_TRY(cudaSetDevice(s->device));
	// a batch of one is the per-ray compatibility path (stack variables of the caller): not worth the pointer query
	const bool in_pinned = known_pinned || (n > 1 && isPinned(rays)), out_pinned = known_pinned || (n > 1 && isPinned(out));
	const bool times_pinned = !times || known_pinned || (n > 1 && isPinned(times));
	if(in_pinned && out_pinned && times_pinned && n <= kDirectRays)
	{
		// Batches in pinned memory (the wavefront ray queue of the renderer, integration/src/render): the kernel
		// reads the rays and writes the results across PCIe itself -- one cursor reset, one launch and one stream
		// synchronisation per call, no staging copies.
		std::unique_ptr<Lane> lane;
		{
			std::lock_guard<std::mutex> lock(s->lane_mutex);
			if(!s->free_lanes.empty()) { lane = std::move(s->free_lanes.back()); s->free_lanes.pop_back(); }
		}
		if(!lane) lane = std::make_unique<Lane>();
		int rc = B200RT_OK;
		if(!lane->stream)
		{
			const cudaError_t e = cudaStreamCreateWithFlags(&lane->stream, cudaStreamNonBlocking);
			if(e != cudaSuccess) rc = fail(B200RT_E_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
		}
		if(rc == B200RT_OK) rc = launch(rays, times, n, out, lane->stream, true);
		if(rc == B200RT_OK)
		{
			const cudaError_t e = cudaStreamSynchronize(lane->stream);
			if(e != cudaSuccess) rc = fail(B200RT_E_CUDA, std::string("direct trace: ") + cudaGetErrorString(e));
		}
		std::lock_guard<std::mutex> lock(s->lane_mutex);
		s->free_lanes.push_back(std::move(lane));
		return rc;
	}
	if(n <= kSmallBatchRays)
	{
		// Small batches (the per-ray compatibility path of the Accelerator virtuals): no staging copies at all.  The
		// kernel reads the rays from, and writes the results to, pinned host memory (device-addressable under UVA), as
		// one warp without a ray cursor: one launch and one stream synchronisation per call.
		std::unique_ptr<Lane> lane;
		{
			std::lock_guard<std::mutex> lock(s->lane_mutex);
			if(!s->free_lanes.empty()) { lane = std::move(s->free_lanes.back()); s->free_lanes.pop_back(); }
		}
		if(!lane) lane = std::make_unique<Lane>();
		int rc = ensureLane(*lane, kSmallBatchRays * sizeof(b200rt_ray), kSmallBatchRays * sizeof(Out), true, true);
		if(rc == B200RT_OK && times) rc = ensureLaneTimes(*lane, kSmallBatchRays);
		if(rc == B200RT_OK && times)
		{
			const cudaError_t e = cudaMemcpyAsync(lane->d_times, times, n * sizeof(float), cudaMemcpyHostToDevice, lane->stream);
			if(e != cudaSuccess) rc = fail(B200RT_E_CUDA, std::string("small-batch times: ") + cudaGetErrorString(e));
		}
		if(rc == B200RT_OK)
		{
			std::memcpy(lane->h_in, rays, n * sizeof(b200rt_ray));
			rc = launch(static_cast<const b200rt_ray *>(lane->h_in), times ? lane->d_times : nullptr, n, static_cast<Out *>(lane->h_out), lane->stream, /*cursorless*/ true);
			if(rc == B200RT_OK)
			{
				const cudaError_t e = cudaStreamSynchronize(lane->stream);
				if(e != cudaSuccess) rc = fail(B200RT_E_CUDA, std::string("small-batch trace: ") + cudaGetErrorString(e));
				else std::memcpy(out, lane->h_out, n * sizeof(Out));
			}
		}
		std::lock_guard<std::mutex> lock(s->lane_mutex);
		s->free_lanes.push_back(std::move(lane));
		return rc;
	}
	const size_t per_ray = std::max(sizeof(b200rt_ray), sizeof(Out));
	const size_t chunk = std::max<size_t>(1024, chunkBytes() / per_ray);
	// full chunks, then a tapered tail (half, quarter, quarter of what is left): when the last rays have arrived nothing but the last
	// piece's kernel and copy-out stands between the call and its return, so that piece should be short
	std::vector<std::pair<size_t, size_t>> pieces; // (first ray, rays)
	{
		size_t begin = 0;
		for(; n - begin > chunk; begin += chunk) pieces.push_back({begin, chunk});
		size_t rest = n - begin;
		for(int k = 0; k < 2 && rest >= 4 * kTwoPassRaysForward; ++k) { const size_t half = rest / 2; pieces.push_back({begin, half}); begin += half; rest -= half; }
		pieces.push_back({begin, rest});
	}
	const size_t n_chunks = pieces.size();
	const int n_lanes = int(std::min<size_t>(size_t(lanesPerCall()), n_chunks));

	std::vector<std::unique_ptr<Lane>> lanes;
	{
		std::lock_guard<std::mutex> lock(s->lane_mutex);
		while(int(lanes.size()) < n_lanes && !s->free_lanes.empty())
		{
			lanes.push_back(std::move(s->free_lanes.back()));
			s->free_lanes.pop_back();
		}
	}
	while(int(lanes.size()) < n_lanes) lanes.push_back(std::make_unique<Lane>());
	auto give_back = [&]() {
		std::lock_guard<std::mutex> lock(s->lane_mutex);
		for(auto &l : lanes) s->free_lanes.push_back(std::move(l));
	};
	auto drain = [&](Lane &l) -> int {
		if(!l.pending_dst) return B200RT_OK;
		CUDA_TRY(cudaEventSynchronize(l.done));
		if(!l.pending_direct) stagingCopy(l.pending_dst, l.h_out, l.pending_bytes);
		l.pending_dst = nullptr;
		return B200RT_OK;
	};
	const size_t rays_this = std::min(chunk, n);
	int rc = B200RT_OK;
	for(auto &l : lanes)
	{
		l->pending_dst = nullptr;
		rc = ensureLane(*l, rays_this * sizeof(b200rt_ray), rays_this * sizeof(Out), !in_pinned, !out_pinned);
		if(rc == B200RT_OK && times) rc = ensureLaneTimes(*l, rays_this);
		if(rc != B200RT_OK) { give_back(); return rc; }
	}
	for(size_t c = 0; c < n_chunks && rc == B200RT_OK; ++c)
	{
		Lane &l = *lanes[c % size_t(n_lanes)];
		rc = drain(l);
		if(rc != B200RT_OK) break;
		const size_t begin = pieces[c].first, count = pieces[c].second;
		const void *src = rays + begin;
		if(!in_pinned) { stagingCopy(l.h_in, src, count * sizeof(b200rt_ray)); src = l.h_in; }
		cudaError_t e = cudaMemcpyAsync(l.d_in, src, count * sizeof(b200rt_ray), cudaMemcpyHostToDevice, l.stream);
		if(e == cudaSuccess && times) e = cudaMemcpyAsync(l.d_times, times + begin, count * sizeof(float), cudaMemcpyHostToDevice, l.stream);
		if(e == cudaSuccess)
		{
			rc = launch(static_cast<const b200rt_ray *>(l.d_in), times ? l.d_times : nullptr, count, static_cast<Out *>(l.d_out), l.stream, false);
			if(rc != B200RT_OK) break;
		}
		if(e == cudaSuccess) e = cudaMemcpyAsync(out_pinned ? static_cast<void *>(out + begin) : l.h_out, l.d_out, count * sizeof(Out), cudaMemcpyDeviceToHost, l.stream);
		if(e == cudaSuccess) e = cudaEventRecord(l.done, l.stream);
		if(e != cudaSuccess) { rc = fail(B200RT_E_CUDA, std::string("staged trace: ") + cudaGetErrorString(e)); break; }
		l.pending_dst = out + begin;
		l.pending_bytes = count * sizeof(Out);
		l.pending_direct = out_pinned;
	}
	for(auto &l : lanes)
	{
		const int r = drain(*l);
		if(rc == B200RT_OK) rc = r;
	}
	if(rc != B200RT_OK) for(auto &l : lanes) cudaStreamSynchronize(l->stream);
	give_back();
	return rc;
}

// Tree cache (multi-GPU frames, tools/render_sharded.py): every rank of a tile-sharded render builds the SAME kd-tree, on host
// cores all ranks share -- at 4 ranks the preprocess went from 1.9 s to 4.5 s.  With B200RT_TREE_CACHE_DIR set (to a directory on
// /dev/shm, say) the first process to take the lock of a geometry + parameter hash builds and writes the tree, the others wait
// for the lock and read it (27 MB at 1 M triangles).  The flattening and the upload stay per rank.  Off by default.
uint64_t fnv1a(const void *data, size_t bytes, uint64_t h)
{
	const unsigned char *p = static_cast<const unsigned char *>(data);
	// 8 bytes at a time: this is a cache key, not a checksum standard
	size_t i = 0;
	for(; i + 8 <= bytes; i += 8) { uint64_t w; std::memcpy(&w, p + i, 8); h = (h ^ w) * 0x100000001b3ull; }
	for(; i < bytes; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
	return h;
}

bool loadTree(const std::string &path, b200rt::HostTree &tree)
{
	std::FILE *f = std::fopen(path.c_str(), "rb");
	if(!f) return false;
	uint64_t head[8];
	bool ok = std::fread(head, sizeof(uint64_t), 8, f) == 8 && head[0] == 0x423230304b445431ull;
	if(ok)
	{
		tree.nodes.resize(size_t(head[1]));
		tree.leaf_refs.resize(size_t(head[2]));
		tree.n_interior = head[3]; tree.n_leaves = head[4]; tree.n_empty_leaves = head[5];
		tree.depth = uint32_t(head[6]); tree.max_leaf_prims = uint32_t(head[7]);
		ok = std::fread(tree.bound, sizeof(float), 6, f) == 6
		     && std::fread(tree.nodes.data(), sizeof(b200rt::HostNode), tree.nodes.size(), f) == tree.nodes.size()
		     && std::fread(tree.leaf_refs.data(), sizeof(uint32_t), tree.leaf_refs.size(), f) == tree.leaf_refs.size();
	}
	std::fclose(f);
	return ok;
}

void saveTree(const std::string &path, const b200rt::HostTree &tree)
{
	const std::string tmp = path + ".tmp" + std::to_string(long(getpid()));
	std::FILE *f = std::fopen(tmp.c_str(), "wb");
	if(!f) return;
	const uint64_t head[8] = {0x423230304b445431ull, tree.nodes.size(), tree.leaf_refs.size(), tree.n_interior, tree.n_leaves, tree.n_empty_leaves, tree.depth, tree.max_leaf_prims};
	bool ok = std::fwrite(head, sizeof(uint64_t), 8, f) == 8 && std::fwrite(tree.bound, sizeof(float), 6, f) == 6
	          && std::fwrite(tree.nodes.data(), sizeof(b200rt::HostNode), tree.nodes.size(), f) == tree.nodes.size()
	          && std::fwrite(tree.leaf_refs.data(), sizeof(uint32_t), tree.leaf_refs.size(), f) == tree.leaf_refs.size();
	ok = (std::fclose(f) == 0) && ok;
	if(ok) std::rename(tmp.c_str(), path.c_str()); else std::remove(tmp.c_str());
}

void buildOrLoadTree(const b200rt::MeshView &mesh, const b200rt::BuildConfig &config, b200rt::HostTree &tree)
{
	const char *dir = std::getenv("B200RT_TREE_CACHE_DIR");
	if(!dir || !*dir) { b200rt::buildKdTree(mesh, config, tree); return; }
	uint64_t h = 0xcbf29ce484222325ull;
	h = fnv1a(mesh.xyz, mesh.n_verts * 3 * sizeof(float), h);
	h = fnv1a(mesh.idx, mesh.n_faces * 4 * sizeof(uint32_t), h);
	const float key[4] = {float(config.max_depth), float(config.max_leaf_size), config.cost_ratio, config.empty_bonus};
	h = fnv1a(key, sizeof key, h);
	char name[64];
	std::snprintf(name, sizeof name, "/b200rt_tree_%016llx", static_cast<unsigned long long>(h));
	const std::string path = std::string(dir) + name + ".bin", lock_path = std::string(dir) + name + ".lock";
	const int lock = ::open(lock_path.c_str(), O_CREAT | O_RDWR, 0600);
	if(lock >= 0) ::flock(lock, LOCK_EX);
	if(!loadTree(path, tree))
	{
		b200rt::buildKdTree(mesh, config, tree);
		saveTree(path, tree);
	}
	if(lock >= 0) { ::flock(lock, LOCK_UN); ::close(lock); }
}

constexpr uint32_t kCursorRing = 4096;
constexpr size_t kMaxRaysPerLaunch = size_t(1) << 30;
constexpr size_t kTwoPassRays = size_t(1) << 15;        // batches from this size on take the two-pass path (setup pass + queue-fed traversal)
static_assert(kTwoPassRays == kTwoPassRaysForward, "keep the forward copy in step");
constexpr size_t kMaxRaysPerTwoPass = size_t(1) << 26;  // 64 Mi rays = 4 GiB of queue scratch at most per launch pair

int checkDeviceCall(const b200rt_scene *s, const void *rays, size_t n, const void *out)
{
	if(!s || (!rays && n) || (!out && n)) return fail(B200RT_E_INVALID, "null argument");
	if(!s->built) return fail(B200RT_E_INVALID, "scene not built: call b200rt_build first");
	return B200RT_OK;
}

// Enqueue traceKernel<Q> over n rays on `stream`: a persistent grid (at most one resident wave) whose warps
// pull rays from a cursor that is zeroed on the same stream just before the launch.
template <int Q>
int launchTrace(b200rt_scene *s, const b200rt_ray *d_rays, size_t n, typename b200rt::OutType<Q>::type *d_out, cudaStream_t stream, int max_depth, unsigned flags = 0u, bool cursorless = false,
                const float *d_times = nullptr)
{
	if(cursorless)
	{
		// small batch (n <= kDirectRays): no ray cursor, so nothing to reset before the launch; warp w owns rays [32 w, 32 w + 32)
		const unsigned grid = unsigned((n + b200rt::kBlock - 1) / b200rt::kBlock);
		if(s->has_spheres) b200rt::traceKernel<Q, true><<<grid, b200rt::kBlock, 0, stream>>>(s->view, d_rays, uint32_t(n), d_out, nullptr, max_depth, (flags & B200RT_RAYS_TREE_SPACE) != 0u, d_times);
		else b200rt::traceKernel<Q, false><<<grid, b200rt::kBlock, 0, stream>>>(s->view, d_rays, uint32_t(n), d_out, nullptr, max_depth, (flags & B200RT_RAYS_TREE_SPACE) != 0u, d_times);
		++g_launches;
		CUDA_TRY(cudaGetLastError());
		return B200RT_OK;
	}
	const bool tree_space = (flags & B200RT_RAYS_TREE_SPACE) != 0u;
#if B200RT_TWO_PASS
	if(n >= kTwoPassRays)
	{
		// Two-pass batch (kd_kernels.cuh): setupKernel answers the rays that miss the tree bound and queues the others, set up,
		// in HBM; traceKernel<.., QUEUED> pulls the queue through shared memory with TMA bulk copies.  The queue is stream-ordered
		// scratch (cudaMallocAsync / cudaFreeAsync on the caller's stream: no synchronisation, the pool keeps the memory).
		for(size_t begin = 0; begin < n; begin += kMaxRaysPerTwoPass)
		{
			const uint32_t count = uint32_t(std::min(kMaxRaysPerTwoPass, n - begin));
			const uint32_t n_regions = (count + uint32_t(b200rt::kRegionRays) - 1u) / uint32_t(b200rt::kRegionRays);
			float *queue = nullptr;
			CUDA_TRY(cudaMallocAsync(&queue, b200rt::queueBytes(n_regions), stream));
			uint32_t *cursor = s->d_cursors + (s->next_cursor.fetch_add(1) % kCursorRing);
			cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(uint32_t), stream);
			if(e == cudaSuccess)
			{
				const unsigned setup_grid = std::max(1u, std::min((n_regions + 7u) / 8u, unsigned(s->setup_blocks)));
				b200rt::setupKernel<Q><<<setup_grid, b200rt::kSetupBlock, 0, stream>>>(s->view, d_rays + begin, count, d_out + begin, queue, tree_space, d_times ? d_times + begin : nullptr, max_depth);
				++g_launches;
				const unsigned wanted = unsigned((size_t(n_regions) + b200rt::kBlock / 32 - 1) / (b200rt::kBlock / 32));
				const unsigned grid = std::max(1u, std::min(wanted, unsigned(s->resident_blocks_queued[Q])));
				const b200rt_ray *as_rays = reinterpret_cast<const b200rt_ray *>(queue);
				if(s->has_spheres) b200rt::traceKernel<Q, true, true><<<grid, b200rt::kBlock, 0, stream>>>(s->view, as_rays, n_regions, d_out + begin, cursor, max_depth, tree_space, nullptr);
#if B200RT_TREELET
				else if(s->d_treelets && Q != b200rt::kTShadow)
				{
					const unsigned tgrid = std::max(1u, std::min(wanted, unsigned(s->resident_blocks_treelet[Q])));
					b200rt::traceKernel<Q, false, true, true><<<tgrid, b200rt::kBlock, 0, stream>>>(s->view, as_rays, n_regions, d_out + begin, cursor, max_depth, tree_space, nullptr);
				}
#endif
#if B200RT_L2_PERSIST
				else if(s->l2_window.num_bytes != 0)
				{
					cudaLaunchConfig_t cfg{};
					cfg.gridDim = dim3(grid); cfg.blockDim = dim3(b200rt::kBlock); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
					cudaLaunchAttribute attr{};
					attr.id = cudaLaunchAttributeAccessPolicyWindow;
					attr.val.accessPolicyWindow = s->l2_window;
					cfg.attrs = &attr; cfg.numAttrs = 1;
					const float *no_times = nullptr;
					e = cudaLaunchKernelEx(&cfg, b200rt::traceKernel<Q, false, true>, s->view, as_rays, n_regions, d_out + begin, cursor, max_depth, tree_space, no_times);
				}
#endif
				else b200rt::traceKernel<Q, false, true><<<grid, b200rt::kBlock, 0, stream>>>(s->view, as_rays, n_regions, d_out + begin, cursor, max_depth, tree_space, nullptr);
				++g_launches;
				e = cudaGetLastError();
			}
			const cudaError_t e_free = cudaFreeAsync(queue, stream);
			if(e != cudaSuccess) return fail(B200RT_E_CUDA, std::string("two-pass trace: ") + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
			CUDA_TRY(e_free);
		}
		return B200RT_OK;
	}
#endif
	for(size_t begin = 0; begin < n; begin += kMaxRaysPerLaunch)
	{
		const uint32_t count = uint32_t(std::min(kMaxRaysPerLaunch, n - begin));
		uint32_t *cursor = s->d_cursors + (s->next_cursor.fetch_add(1) % kCursorRing);
		CUDA_TRY(cudaMemsetAsync(cursor, 0, sizeof(uint32_t), stream));
		const unsigned wanted = unsigned((size_t(count) + b200rt::kBlock - 1) / b200rt::kBlock);
		const unsigned grid = std::max(1u, std::min(wanted, unsigned(s->resident_blocks[Q])));
		if(s->has_spheres) b200rt::traceKernel<Q, true><<<grid, b200rt::kBlock, 0, stream>>>(s->view, d_rays + begin, count, d_out + begin, cursor, max_depth, tree_space, d_times ? d_times + begin : nullptr);
		else b200rt::traceKernel<Q, false><<<grid, b200rt::kBlock, 0, stream>>>(s->view, d_rays + begin, count, d_out + begin, cursor, max_depth, tree_space, d_times ? d_times + begin : nullptr);
		++g_launches;
		CUDA_TRY(cudaGetLastError());
	}
	return B200RT_OK;
}

template <int Q, bool SPHERES, bool QUEUED, bool TREELET = false>
int queryResidencyOf(b200rt_scene *s, int &blocks)
{
	int per_sm = 0, sms = 0;
#ifdef B200RT_CARVEOUT
	CUDA_TRY(cudaFuncSetAttribute(b200rt::traceKernel<Q, SPHERES, QUEUED, TREELET>, cudaFuncAttributePreferredSharedMemoryCarveout, B200RT_CARVEOUT));
#endif
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, b200rt::traceKernel<Q, SPHERES, QUEUED, TREELET>, b200rt::kBlock, 0));
	CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
	blocks = std::max(1, per_sm) * std::max(1, sms);
	return B200RT_OK;
}

// one resident wave of the kernel variants this scene launches (with or without the sphere branch; single-kernel and queue-fed)
template <int Q>
int queryResidency(b200rt_scene *s)
{
	int rc = s->has_spheres ? queryResidencyOf<Q, true, false>(s, s->resident_blocks[Q]) : queryResidencyOf<Q, false, false>(s, s->resident_blocks[Q]);
	if(rc == B200RT_OK) rc = s->has_spheres ? queryResidencyOf<Q, true, true>(s, s->resident_blocks_queued[Q]) : queryResidencyOf<Q, false, true>(s, s->resident_blocks_queued[Q]);
	if(rc == B200RT_OK && Q == b200rt::kClosest)
	{
		int per_sm = 0, sms = 0;
		CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, b200rt::setupKernel<b200rt::kClosest>, b200rt::kSetupBlock, 0));
		CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
		s->setup_blocks = std::max(1, per_sm) * std::max(1, sms);
	}
	return rc;
}

} // namespace

// ---- flush combiner: the worker thread of a scene ---------------------------------------------------------------------
namespace {

long envLong(const char *name, long fallback)
{
	const char *v = std::getenv(name);
	return (v && *v) ? std::atol(v) : fallback;
}
// flushes are held back for at most this long, or until this many rays are pending (B200RT_COMBINE_WINDOW_US / _RAYS)
size_t combineTargetRays() { static const size_t v = size_t(std::max(1l, envLong("B200RT_COMBINE_RAYS", 2048))); return v; }

// One launch per group of up to kMaxSegments jobs with the same ray-space flag.  Called WITHOUT the combiner's mutex by the thread
// that took `jobs` off the pending list.
void launchCombined(b200rt_scene *s, std::vector<b200rt::PendingJob> &jobs)
{
	b200rt::Combiner &c = *s->combiner;
	cudaSetDevice(s->device);
	for(unsigned tree_space = 0u; tree_space < 2u; ++tree_space)
	{
		size_t j = 0;
		while(j < jobs.size())
		{
			b200rt::SegmentTable table{};
			b200rt::Ticket *riders[b200rt::kMaxSegments];
			uint32_t warps = 0u;
			size_t rays = 0;
			for(; j < jobs.size() && table.n_segments < uint32_t(b200rt::kMaxSegments); ++j)
			{
				const b200rt_job &job = jobs[j].job;
				if(((job.flags & B200RT_RAYS_TREE_SPACE) != 0u) != (tree_space != 0u)) continue;
				riders[table.n_segments] = jobs[j].ticket.get();
				b200rt::Segment &g = table.seg[table.n_segments++];
				g.rays = job.rays; g.out = job.out; g.times = job.times; g.n = uint32_t(job.n);
				g.first_warp = warps; g.query = job.query; g.max_depth = job.max_depth;
				warps += uint32_t((job.n + 31) / 32);
				rays += job.n;
			}
			if(table.n_segments == 0u) continue;
			table.n_warps = warps;
			b200rt::CombinedBatch *batch = nullptr;
			{
				std::lock_guard<std::mutex> lock(c.mutex);
				if(!c.free_batches.empty()) { batch = c.free_batches.back(); c.free_batches.pop_back(); }
			}
			if(!batch) batch = new b200rt::CombinedBatch;
			batch->rc = B200RT_OK;
			batch->error.clear();
			cudaError_t e = cudaSuccess;
			if(!batch->event) e = cudaEventCreateWithFlags(&batch->event, cudaEventDisableTiming);
			cudaStream_t stream = c.streams[c.next_stream.fetch_add(1) % c.streams.size()];
			if(e == cudaSuccess)
			{
				const unsigned grid = (warps + b200rt::kBlock / 32 - 1) / (b200rt::kBlock / 32);
				if(s->has_spheres) b200rt::traceSegmentsKernel<true><<<grid, b200rt::kBlock, 0, stream>>>(s->view, table, tree_space != 0u);
				else b200rt::traceSegmentsKernel<false><<<grid, b200rt::kBlock, 0, stream>>>(s->view, table, tree_space != 0u);
				++g_launches;
				e = cudaGetLastError();
			}
			if(e == cudaSuccess) e = cudaEventRecord(batch->event, stream);
			if(e != cudaSuccess) { batch->rc = B200RT_E_CUDA; batch->error = std::string("traceSegmentsKernel: ") + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")"; }
			c.launches.fetch_add(1);
			c.jobs.fetch_add(table.n_segments);
			c.rays.fetch_add(rays);
			batch->users.store(int(table.n_segments), std::memory_order_relaxed);
			for(uint32_t k = 0; k < table.n_segments; ++k) riders[k]->batch.store(batch, std::memory_order_release);
		}
	}
}

// Take whatever is pending and launch it (the caller's own job included, if it is still there).
void flushCombiner(b200rt_scene *s)
{
	b200rt::Combiner &c = *s->combiner;
	std::vector<b200rt::PendingJob> taken;
	{
		std::lock_guard<std::mutex> lock(c.mutex);
		taken.swap(c.pending);
		c.pending_rays = 0;
	}
	if(!taken.empty()) launchCombined(s, taken);
}

} // namespace

void b200rt_scene::stopCombiner()
{
	if(!combiner) return;
	flushCombiner(this); // (jobs whose flights were never ended)
	for(cudaStream_t st : combiner->streams) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
	for(b200rt::CombinedBatch *b : combiner->free_batches) { if(b->event) cudaEventDestroy(b->event); delete b; }
	combiner.reset();
}

// shared with the other translation units of the library (b200pm.cu): one error text per thread, one launch counter
namespace b200 {
int failWith(int code, const std::string &msg) { return fail(code, msg); }
void countLaunches(uint64_t n) { g_launches += n; }
} // namespace b200

extern "C" {

int b200rt_version(void) { return B200RT_VERSION; }
const char *b200rt_last_error(void) { return g_last_error.c_str(); }
uint64_t b200rt_launch_count(void) { return g_launches.load(); }

int b200rt_device_count(int *count)
{
	if(!count) return fail(B200RT_E_INVALID, "null argument");
	*count = 0;
	int n = 0;
	const cudaError_t e = cudaGetDeviceCount(&n);
	if(e != cudaSuccess) { cudaGetLastError(); return fail(B200RT_E_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); }
	*count = n;
	return B200RT_OK;
}

int b200rt_create(int device, const b200rt_build_params *params, b200rt_scene **out)
{
	if(!out) return fail(B200RT_E_INVALID, "null argument");
	*out = nullptr;
	int n = 0;
	const int rc = b200rt_device_count(&n);
	if(rc != B200RT_OK) return rc;
	if(n == 0) return fail(B200RT_E_NO_DEVICE, "no CUDA device: libb200rt has no CPU fallback");
	if(device < 0 || device >= n) return fail(B200RT_E_INVALID, "device index out of range");
	CUDA_TRY(cudaSetDevice(device));
	CUDA_TRY(cudaFree(nullptr));
	{
		// the two-pass path takes its ray queue from the stream-ordered allocator: let the pool keep what it has handed out
		cudaMemPool_t pool = nullptr;
		if(cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool)
		{
			uint64_t keep = ~uint64_t(0);
			cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
		}
		cudaGetLastError();
	}
	b200rt_scene *s = new(std::nothrow) b200rt_scene;
	if(!s) return fail(B200RT_E_MEMORY, "out of host memory");
	s->device = device;
	if(params)
	{
		s->config.max_depth = params->max_depth;
		s->config.max_leaf_size = params->max_leaf_size;
		s->config.cost_ratio = params->cost_ratio;
		s->config.empty_bonus = params->empty_bonus > 0.f ? params->empty_bonus : -1.f;
		s->config.threads = params->build_threads;
	}
	*out = s;
	return B200RT_OK;
}

void b200rt_destroy(b200rt_scene *scene) { delete scene; }

int b200rt_add_mesh(b200rt_scene *s, const float *xyz, size_t n_verts, const uint32_t *idx, size_t n_faces, const uint8_t *flags)
{
	if(!s || (!xyz && n_verts) || (!idx && n_faces)) return fail(B200RT_E_INVALID, "null argument");
	const size_t base = s->xyz.size() / 3;
	if(base + n_verts >= size_t(kSphere)) return fail(B200RT_E_INVALID, "too many vertices");
	if(s->idx.size() / 4 + n_faces >= (size_t(1) << 30)) return fail(B200RT_E_INVALID, "too many faces");
	for(size_t f = 0; f < n_faces; ++f)
		for(int k = 0; k < 4; ++k)
		{
			const uint32_t v = idx[4 * f + k];
			if(k == 3 && v == kTriangle) continue;
			if(v >= n_verts) return fail(B200RT_E_INVALID, "face " + std::to_string(f) + " references vertex " + std::to_string(v) + " >= n_verts");
		}
	try
	{
		s->xyz.insert(s->xyz.end(), xyz, xyz + 3 * n_verts);
		s->idx.reserve(s->idx.size() + 4 * n_faces);
		for(size_t f = 0; f < n_faces; ++f)
			for(int k = 0; k < 4; ++k)
			{
				const uint32_t v = idx[4 * f + k];
				s->idx.push_back((k == 3 && v == kTriangle) ? kTriangle : uint32_t(v + base));
			}
		if(flags) s->flags.insert(s->flags.end(), flags, flags + n_faces);
		else s->flags.insert(s->flags.end(), n_faces, uint8_t(B200RT_FACE_VISIBLE | B200RT_FACE_CASTS_SHADOWS));
		s->kind.resize(s->flags.size(), uint8_t(0));
		s->motion_of_face.resize(s->flags.size(), -1);
	}
	catch(const std::bad_alloc &) { return fail(B200RT_E_MEMORY, "out of host memory"); }
	s->built = false;
	return B200RT_OK;
}

namespace {

// SquareMatrix * Point (include/geometry/matrix.h:131-144): aux = 0; aux += m[i][j] * v[j], j = 0..2; aux += m[i][3] -- float
// operations in this order (part of a primitive's bound, hence of the tree bound and the ray bias)
void matPoint(const float *m, const float *v, float *o)
{
	for(int i = 0; i < 3; ++i)
	{
		volatile float aux = 0.f;
		for(int j = 0; j < 3; ++j) { volatile float prod = m[4 * i + j] * v[j]; aux = aux + prod; }
		aux = aux + m[4 * i + 3];
		o[i] = aux;
	}
}

// Faces of a motion-blur mesh / moving instance: the builder gets one box face per face (two corner vertices appended to the
// scene's vertex array), the flattener the MotionFace record.  steps: the vertex arrays of the three time steps (Bezier) or
// {xyz, nullptr, nullptr} with `matrix` >= 0 (moving instance).
int addMotionFaces(b200rt_scene *s, const float *const steps[3], size_t n_verts, const uint32_t *idx, size_t n_faces, const uint8_t *flags, float t0, float t1, int matrix)
{
	if(s->idx.size() / 4 + n_faces >= (size_t(1) << 30)) return fail(B200RT_E_INVALID, "too many faces");
	if(s->xyz.size() / 3 + 2 * n_faces >= size_t(kBox)) return fail(B200RT_E_INVALID, "too many vertices");
	for(size_t f = 0; f < n_faces; ++f)
		for(int k = 0; k < 4; ++k)
		{
			const uint32_t v = idx[4 * f + k];
			if(k == 3 && v == kTriangle) continue;
			if(v >= n_verts) return fail(B200RT_E_INVALID, "face " + std::to_string(f) + " references vertex " + std::to_string(v) + " >= n_verts");
		}
	try
	{
		for(size_t f = 0; f < n_faces; ++f)
		{
			b200rt_scene::MotionFace mf{};
			mf.nv = (idx[4 * f + 3] == kTriangle) ? 3u : 4u;
			mf.t0 = t0; mf.t1 = t1;
			mf.matrix = matrix >= 0 ? uint32_t(matrix) : 0u;
			float lo[3], hi[3];
			bool first = true;
			for(int step = 0; step < 3; ++step)
				for(uint32_t k = 0; k < mf.nv; ++k)
				{
					float p[3];
					if(matrix >= 0)
					{
						const float *base = steps[0] + 3 * size_t(idx[4 * f + k]);
						if(step == 0) for(int a = 0; a < 3; ++a) mf.v[0][k][a] = base[a];
						matPoint(s->matrices[size_t(matrix)].data() + 16 * step, base, p); // PrimitiveInstance::getBound, primitive_instance.h:119-128
					}
					else
					{
						const float *src = steps[step] + 3 * size_t(idx[4 * f + k]);
						for(int a = 0; a < 3; ++a) { mf.v[step][k][a] = src[a]; p[a] = src[a]; } // FacePrimitive::getBoundTimeSteps, primitive_face.h:155-170
					}
					for(int a = 0; a < 3; ++a)
					{
						if(first || p[a] < lo[a]) lo[a] = p[a];
						if(first || p[a] > hi[a]) hi[a] = p[a];
					}
					first = false;
				}
			const uint32_t corner = uint32_t(s->xyz.size() / 3);
			s->xyz.insert(s->xyz.end(), {lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]});
			s->idx.insert(s->idx.end(), {corner, corner + 1u, kBox, kTriangle});
			s->flags.push_back(flags ? flags[f] : uint8_t(B200RT_FACE_VISIBLE | B200RT_FACE_CASTS_SHADOWS));
			s->kind.push_back(matrix >= 0 ? uint8_t(2) : uint8_t(1));
			s->motion_of_face.push_back(int32_t(s->motion.size()));
			s->motion.push_back(mf);
		}
	}
	catch(const std::bad_alloc &) { return fail(B200RT_E_MEMORY, "out of host memory"); }
	s->built = false;
	return B200RT_OK;
}

} // namespace

int b200rt_add_mesh_bezier(b200rt_scene *s, const float *xyz0, const float *xyz1, const float *xyz2, size_t n_verts, const uint32_t *idx, size_t n_faces,
                           const uint8_t *flags, float time_start, float time_end)
{
	if(!s || ((!xyz0 || !xyz1 || !xyz2) && n_verts) || (!idx && n_faces)) return fail(B200RT_E_INVALID, "null argument");
	const float *const steps[3] = {xyz0, xyz1, xyz2};
	return addMotionFaces(s, steps, n_verts, idx, n_faces, flags, time_start, time_end, -1);
}

int b200rt_add_mesh_moving(b200rt_scene *s, const float *xyz, size_t n_verts, const uint32_t *idx, size_t n_faces, const uint8_t *flags,
                           const float matrices[48], float time_start, float time_end)
{
	if(!s || (!xyz && n_verts) || (!idx && n_faces) || !matrices) return fail(B200RT_E_INVALID, "null argument");
	std::array<float, 50> m{};
	std::copy(matrices, matrices + 48, m.begin());
	m[48] = time_start; m[49] = time_end;
	try { s->matrices.push_back(m); }
	catch(const std::bad_alloc &) { return fail(B200RT_E_MEMORY, "out of host memory"); }
	const float *const steps[3] = {xyz, nullptr, nullptr};
	return addMotionFaces(s, steps, n_verts, idx, n_faces, flags, time_start, time_end, int(s->matrices.size()) - 1);
}

int b200rt_add_spheres(b200rt_scene *s, const float *center_radius, size_t n_spheres, const uint8_t *flags)
{
	if(!s || (!center_radius && n_spheres)) return fail(B200RT_E_INVALID, "null argument");
	const size_t base = s->xyz.size() / 3;
	if(base + 2 * n_spheres >= size_t(kSphere)) return fail(B200RT_E_INVALID, "too many vertices");
	if(s->idx.size() / 4 + n_spheres >= (size_t(1) << 30)) return fail(B200RT_E_INVALID, "too many faces");
	try
	{
		// stored in the mesh arrays: two vertices (centre; radius in x) and one marker face per sphere
		for(size_t k = 0; k < n_spheres; ++k)
		{
			const float *c = center_radius + 4 * k;
			s->xyz.insert(s->xyz.end(), {c[0], c[1], c[2], c[3], 0.f, 0.f});
			s->idx.insert(s->idx.end(), {uint32_t(base + 2 * k), uint32_t(base + 2 * k + 1), kSphere, kTriangle});
		}
		if(flags) s->flags.insert(s->flags.end(), flags, flags + n_spheres);
		else s->flags.insert(s->flags.end(), n_spheres, uint8_t(B200RT_FACE_VISIBLE | B200RT_FACE_CASTS_SHADOWS));
		s->kind.resize(s->flags.size(), uint8_t(0));
		s->motion_of_face.resize(s->flags.size(), -1);
	}
	catch(const std::bad_alloc &) { return fail(B200RT_E_MEMORY, "out of host memory"); }
	s->built = false;
	return B200RT_OK;
}

int b200rt_build(b200rt_scene *s)
{
	if(!s) return fail(B200RT_E_INVALID, "null argument");
	CUDA_TRY(cudaSetDevice(s->device));
	s->built = false;
	const size_t n_faces = s->idx.size() / 4;
	const auto t0 = std::chrono::steady_clock::now();
	std::vector<uint2> nodes;
	std::vector<float4> tris;
	std::vector<uint4> treelets; // B200RT_TREELET
	try
	{
		const b200rt::MeshView mesh{s->xyz.data(), s->xyz.size() / 3, s->idx.data(), n_faces};
		buildOrLoadTree(mesh, s->config, s->tree);
		// flatten: one record per leaf reference, in leaf order
		const auto &tree = s->tree;
		s->record_of_ref.assign(tree.leaf_refs.size(), 0u);
		tris.reserve(tree.leaf_refs.size() * 3 + 4);
		nodes.resize(tree.nodes.size());
		uint64_t n_tri = 0, n_quad = 0, n_sphere = 0, n_bezier = 0, n_moving = 0;
		for(size_t f = 0; f < n_faces; ++f)
		{
			if(s->kind[f]) { (s->kind[f] == 1 ? n_bezier : n_moving)++; (s->motion[size_t(s->motion_of_face[f])].nv == 3u ? n_tri : n_quad)++; }
			else (s->idx[4 * f + 2] == kSphere ? n_sphere : s->idx[4 * f + 3] == kTriangle ? n_tri : n_quad)++;
		}
		for(size_t i = 0; i < tree.nodes.size(); ++i)
		{
			const b200rt::HostNode hn = tree.nodes[i];
			if((hn.b & 3u) != 3u) { nodes[i] = make_uint2(hn.a, hn.b); continue; }
			const uint32_t count = hn.b >> 2;
			if(count >= b200rt::kLeafStride4) return fail(B200RT_E_INVALID, "a leaf holds 2^29 primitives or more");
			// a leaf of static polygons with at least one quad stores all of its records in four float4 (kd_kernels.cuh, kLeafStride4)
			bool stride4 = false, plain = true;
			for(uint32_t k = 0; k < count; ++k)
			{
				const uint32_t face = tree.leaf_refs[hn.a + k];
				const uint32_t *id = s->idx.data() + 4 * size_t(face);
				if(s->kind[face] || id[2] == kSphere) plain = false;
				else if(id[3] != kTriangle) stride4 = true;
			}
			stride4 = stride4 && plain && B200RT_COOP_LEAF != 0; // only the cooperative leaf phase needs the padding
			nodes[i] = make_uint2(uint32_t(tris.size()), hn.b | (stride4 ? (b200rt::kLeafStride4 << 2) : 0u));
			for(uint32_t k = 0; k < count; ++k)
			{
				const uint32_t face = tree.leaf_refs[hn.a + k];
				const uint32_t *id = s->idx.data() + 4 * size_t(face);
				if(s->kind[face])
				{
					// motion-blur face (record layouts: kd_kernels.cuh, kFlagBezier / kFlagMoving)
					const b200rt_scene::MotionFace &mf = s->motion[size_t(s->motion_of_face[face])];
					const bool bezier = s->kind[face] == 1;
					s->record_of_ref[hn.a + k] = uint32_t(tris.size());
					uint32_t fl = (s->flags[face] & 7u) | (bezier ? b200rt::kFlagBezier : b200rt::kFlagMoving);
					if(mf.nv == 4u) fl |= b200rt::kFlagQuad;
					const uint32_t inst_offset = mf.matrix * 10u;
					for(int step = 0; step < (bezier ? 3 : 1); ++step)
						for(uint32_t v = 0; v < mf.nv; ++v)
						{
							float4 q = make_float4(mf.v[step][v][0], mf.v[step][v][1], mf.v[step][v][2], 0.f);
							if(step == 0 && v == 0) std::memcpy(&q.w, &face, 4);
							else if(step == 0 && v == 1) std::memcpy(&q.w, &fl, 4);
							else if(step == 0 && v == 2) { if(bezier) q.w = mf.t0; else std::memcpy(&q.w, &inst_offset, 4); }
							else if(step == 1 && v == 0) q.w = mf.t1;
							tris.push_back(q);
						}
					continue;
				}
				const bool quad = id[3] != kTriangle;
				const float *v0 = s->xyz.data() + 3 * size_t(id[0]);
				s->record_of_ref[hn.a + k] = uint32_t(tris.size());
				uint32_t fl = s->flags[face] & 7u;
				if(quad) fl |= b200rt::kFlagQuad;
				float4 q;
				if(id[2] == kSphere)
				{
					// sphere record, same 3-vector stride as a triangle: q0 = centre | face id, q1 = radius,0,0 | flags, q2 = 0
					fl |= b200rt::kFlagSphere;
					q.x = v0[0]; q.y = v0[1]; q.z = v0[2];
					std::memcpy(&q.w, &face, 4);
					tris.push_back(q);
					q.x = s->xyz[3 * size_t(id[1])]; q.y = 0.f; q.z = 0.f;
					std::memcpy(&q.w, &fl, 4);
					tris.push_back(q);
					tris.push_back(make_float4(0.f, 0.f, 0.f, 0.f));
					continue;
				}
				q.x = v0[0]; q.y = v0[1]; q.z = v0[2];
				std::memcpy(&q.w, &face, 4);
				tris.push_back(q);
				for(int e = 1; e <= (quad ? 3 : 2); ++e)
				{
					const float *ve = s->xyz.data() + 3 * size_t(id[e]);
					// edge_k = v_k - v_0, one float subtraction per component (shape_polygon.h:130-131,150)
					volatile float ex = ve[0] - v0[0], ey = ve[1] - v0[1], ez = ve[2] - v0[2];
					q.x = ex; q.y = ey; q.z = ez;
					const uint32_t w = (e == 1) ? fl : 0u;
					std::memcpy(&q.w, &w, 4);
					tris.push_back(q);
				}
				if(stride4 && !quad) tris.push_back(make_float4(0.f, 0.f, 0.f, 0.f));
			}
		}
		if(tris.size() >= (size_t(1) << 32)) return fail(B200RT_E_INVALID, "leaf stream exceeds 2^32 records");
#if B200RT_TREELET
		// Two-level treelets of a polygon-only scene (kd_kernels.cuh, kEmptyRef): every interior node whose parent is not in a
		// treelet already becomes a treelet root; allocation in depth-first order, like the nodes.
		treelets.clear();
		if(n_sphere == 0 && n_bezier == 0 && n_moving == 0 && tris.size() < (size_t(1) << 31) - 4)
		{
			const auto is_interior = [&](uint32_t i) { return (tree.nodes[i].b & 3u) != 3u; };
			const auto leaf_ref = [&](uint32_t i) -> uint32_t {
				const uint32_t count = tree.nodes[i].b >> 2;
				if(count == 0u) return b200rt::kEmptyRef;
				const uint32_t first = nodes[i].x;
				std::memcpy(&tris[size_t(first) + 2].w, &count, 4); // the count rides in q2.w of the leaf's first record
				return b200rt::kLeafRef | first;
			};
			const float far_plane = FLT_MAX;
			uint32_t far_bits;
			std::memcpy(&far_bits, &far_plane, 4);
			// work list of (node, slot to patch with the treelet's index); slot = index into `treelets` as uint32 words, or ~0 for the root
			std::vector<std::pair<uint32_t, size_t>> todo;
			if(is_interior(0u)) todo.push_back({0u, ~size_t(0)});
			else
			{
				// a tree of one leaf: a treelet whose planes all lie beyond every interval and whose first slot is that leaf
				treelets.push_back(make_uint4(far_bits, far_bits, far_bits, (3u << 2) | (3u << 10) | (3u << 18)));
				treelets.push_back(make_uint4(leaf_ref(0u), b200rt::kEmptyRef, b200rt::kEmptyRef, b200rt::kEmptyRef));
			}
			while(!todo.empty())
			{
				const auto [node, slot] = todo.back();
				todo.pop_back();
				const uint32_t index = uint32_t(treelets.size() / 2);
				if(slot != ~size_t(0)) reinterpret_cast<uint32_t *>(treelets.data())[slot] = index;
				const uint32_t child[2] = {node + 1u, tree.nodes[node].b >> 2};
				uint32_t split[2] = {far_bits, far_bits}, axis[2] = {3u, 3u}, ref[4] = {b200rt::kEmptyRef, b200rt::kEmptyRef, b200rt::kEmptyRef, b200rt::kEmptyRef};
				uint32_t grand[4] = {0u, 0u, 0u, 0u}; // interior grandchildren still to become treelets (0 = none: node 0 is never a grandchild)
				for(int c = 0; c < 2; ++c)
				{
					if(!is_interior(child[c])) { ref[2 * c] = leaf_ref(child[c]); continue; }
					split[c] = tree.nodes[child[c]].a;
					axis[c] = tree.nodes[child[c]].b & 3u;
					const uint32_t g[2] = {child[c] + 1u, tree.nodes[child[c]].b >> 2};
					for(int k = 0; k < 2; ++k)
					{
						if(is_interior(g[k])) grand[2 * c + k] = g[k];
						else ref[2 * c + k] = leaf_ref(g[k]);
					}
				}
				treelets.push_back(make_uint4(tree.nodes[node].a, split[0], split[1], ((tree.nodes[node].b & 3u) << 2) | (axis[0] << 10) | (axis[1] << 18)));
				treelets.push_back(make_uint4(ref[0], ref[1], ref[2], ref[3]));
				// depth-first order: the first grandchild's treelet follows its parent's directly (pushed last = popped first)
				for(int k = 3; k >= 0; --k)
					if(grand[k]) todo.push_back({grand[k], (size_t(index) * 2 + 1) * 4 + size_t(k)});
			}
			if(treelets.size() / 2 >= size_t(b200rt::kLeafRef)) treelets.clear();
		}
#endif
		for(int pad = 0; pad < 4; ++pad) tris.push_back(make_float4(0.f, 0.f, 0.f, 0.f)); // records are read up to four vectors ahead of their kind test
		s->stats = b200rt_stats{};
		s->stats.n_faces = n_faces;
		s->stats.n_triangles = n_tri;
		s->stats.n_quads = n_quad;
		s->stats.n_spheres = n_sphere;
		s->stats.n_bezier_faces = n_bezier;
		s->stats.n_moving_faces = n_moving;
		s->has_spheres = n_sphere != 0 || n_bezier != 0 || n_moving != 0; // the kernel variant with the other primitive kinds
		s->stats.n_nodes = tree.nodes.size();
		s->stats.n_interior = tree.n_interior;
		s->stats.n_leaves = tree.n_leaves;
		s->stats.n_empty_leaves = tree.n_empty_leaves;
		s->stats.n_leaf_refs = tree.leaf_refs.size();
		s->stats.max_depth = tree.depth;
		s->stats.max_leaf_prims = tree.max_leaf_prims;
	}
	catch(const std::bad_alloc &) { return fail(B200RT_E_MEMORY, "out of host memory during build"); }
	catch(const std::exception &e) { return fail(B200RT_E_INVALID, std::string("build failed: ") + e.what()); }
	const auto t1 = std::chrono::steady_clock::now();
	s->stats.build_seconds = std::chrono::duration<double>(t1 - t0).count();

	if(s->d_scene) { cudaFree(s->d_scene); s->d_scene = nullptr; s->d_nodes = nullptr; s->d_tris = nullptr; }
	{
		const size_t node_bytes = (nodes.size() * sizeof(uint2) + 255) & ~size_t(255);
		s->scene_bytes = node_bytes + tris.size() * sizeof(float4);
		CUDA_TRY(cudaMalloc(&s->d_scene, s->scene_bytes));
		s->d_nodes = static_cast<uint2 *>(s->d_scene);
		s->d_tris = reinterpret_cast<float4 *>(static_cast<char *>(s->d_scene) + node_bytes);
	}
	s->l2_window = cudaAccessPolicyWindow{};
#if B200RT_L2_PERSIST
	{
		// L2 persistence for the scene: the rays, queue entries and results that stream through L2 (1.5 GB per closest pass) keep
		// evicting it (ncu: ~1.1 GB of scene re-fetched from HBM per launch although nodes + records are 95 MB of a 126 MB L2)
		cudaDeviceProp prop{};
		CUDA_TRY(cudaGetDeviceProperties(&prop, s->device));
		const size_t carve = std::min<size_t>(size_t(prop.persistingL2CacheMaxSize), s->scene_bytes);
		if(carve > 0 && prop.accessPolicyMaxWindowSize > 0)
		{
			CUDA_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
			s->l2_window.base_ptr = s->d_scene;
			s->l2_window.num_bytes = std::min<size_t>(s->scene_bytes, size_t(prop.accessPolicyMaxWindowSize));
			s->l2_window.hitRatio = float(std::min(1.0, double(carve) / double(s->l2_window.num_bytes)));
			s->l2_window.hitProp = cudaAccessPropertyPersisting;
			s->l2_window.missProp = cudaAccessPropertyStreaming;
		}
	}
#endif
	CUDA_TRY(cudaMemcpy(s->d_nodes, nodes.data(), nodes.size() * sizeof(uint2), cudaMemcpyHostToDevice));
	CUDA_TRY(cudaMemcpy(s->d_tris, tris.data(), tris.size() * sizeof(float4), cudaMemcpyHostToDevice));
	if(s->d_treelets) { cudaFree(s->d_treelets); s->d_treelets = nullptr; }
	if(!treelets.empty())
	{
		CUDA_TRY(cudaMalloc(&s->d_treelets, treelets.size() * sizeof(uint4)));
		CUDA_TRY(cudaMemcpy(s->d_treelets, treelets.data(), treelets.size() * sizeof(uint4), cudaMemcpyHostToDevice));
	}
	if(s->d_inst) { cudaFree(s->d_inst); s->d_inst = nullptr; }
	if(!s->matrices.empty())
	{
		// per moving instance: rows 0-2 of its three matrices, then (time_start, time_end, -, -)
		std::vector<float4> inst;
		for(const auto &m : s->matrices)
		{
			for(int step = 0; step < 3; ++step)
				for(int i = 0; i < 3; ++i) inst.push_back(make_float4(m[16 * step + 4 * i], m[16 * step + 4 * i + 1], m[16 * step + 4 * i + 2], m[16 * step + 4 * i + 3]));
			inst.push_back(make_float4(m[48], m[49], 0.f, 0.f));
		}
		CUDA_TRY(cudaMalloc(&s->d_inst, inst.size() * sizeof(float4)));
		CUDA_TRY(cudaMemcpy(s->d_inst, inst.data(), inst.size() * sizeof(float4), cudaMemcpyHostToDevice));
	}
	// one resident wave of the kernel variant THIS build launches (a rebuild after b200rt_add_spheres switches variants); the
	// cursors are allocated only once the occupancy queries have succeeded, so a failed query is retried by the next build
	{
		int rc = queryResidency<b200rt::kClosest>(s);
		if(rc == B200RT_OK) rc = queryResidency<b200rt::kShadow>(s);
		if(rc == B200RT_OK) rc = queryResidency<b200rt::kTShadow>(s);
		if(rc != B200RT_OK) return rc;
	}
	if(!s->d_cursors) CUDA_TRY(cudaMalloc(&s->d_cursors, kCursorRing * sizeof(uint32_t)));
	s->n_tri_vec4 = tris.size();
	s->view.nodes = s->d_nodes;
	s->view.tris = s->d_tris;
	s->view.inst = s->d_inst;
	s->view.treelets = s->d_treelets;
	s->view.spill = nullptr;
	s->view.spill_threads = 0u;
#if B200RT_TREELET
	if(s->d_treelets)
	{
		// overflow area of the short stacks: one column per thread of the largest resident grid, kMaxTreeDepth entries deep
		int blocks_c = 0, blocks_s = 0;
		int rc2 = queryResidencyOf<b200rt::kClosest, false, true, true>(s, blocks_c);
		if(rc2 == B200RT_OK) rc2 = queryResidencyOf<b200rt::kShadow, false, true, true>(s, blocks_s);
		if(rc2 != B200RT_OK) return rc2;
		s->resident_blocks_treelet[b200rt::kClosest] = blocks_c;
		s->resident_blocks_treelet[b200rt::kShadow] = blocks_s;
		const size_t threads = size_t(std::max(blocks_c, blocks_s)) * b200rt::kBlock;
		if(s->d_spill) { cudaFree(s->d_spill); s->d_spill = nullptr; }
		CUDA_TRY(cudaMalloc(&s->d_spill, threads * 96 * sizeof(uint2))); // up to three postponed subtrees per treelet level: 1.5 x kMaxTreeDepth entries, less the 8 of the ring
		s->view.spill = s->d_spill;
		s->view.spill_threads = uint32_t(threads);
	}
#endif
	std::memcpy(s->view.bound, s->tree.bound, sizeof(s->view.bound));
	s->stats.device_bytes = nodes.size() * sizeof(uint2) + tris.size() * sizeof(float4);
	// Warm start for the renderer's ray queues (b200rt_trace_jobs_begin): their first flushes come from sixteen threads at once,
	// and what a first use costs -- a stream per lane, the lazy load of a kernel, the growth of the device's local-memory pool for
	// a kernel with a stack frame -- is serialised inside the driver while every render thread waits (measured on a one-frame SPPM
	// render: 2.1 thread-seconds inside libb200rt during the first photon pass, 0.02 in the second; profiles/r4g_*).  So: one
	// empty launch of each kernel a flush can use, and a stock of lanes with their streams, paid once here.
	{
		b200rt::MixedBatch none{};
		cudaStream_t warm = nullptr;
		CUDA_TRY(cudaStreamCreateWithFlags(&warm, cudaStreamNonBlocking));
		if(s->has_spheres)
		{
			b200rt::traceMixedKernel<true><<<1, b200rt::kBlock, 0, warm>>>(s->view, none, 0, true);
			b200rt::traceSegmentsKernel<true><<<1, b200rt::kBlock, 0, warm>>>(s->view, b200rt::SegmentTable{}, true);
			b200rt::traceKernel<b200rt::kClosest, true><<<1, b200rt::kBlock, 0, warm>>>(s->view, nullptr, 0u, nullptr, nullptr, 0, true, nullptr);
			b200rt::traceKernel<b200rt::kShadow, true><<<1, b200rt::kBlock, 0, warm>>>(s->view, nullptr, 0u, nullptr, nullptr, 0, true, nullptr);
			b200rt::traceKernel<b200rt::kTShadow, true><<<1, b200rt::kBlock, 0, warm>>>(s->view, nullptr, 0u, nullptr, nullptr, 0, true, nullptr);
		}
		else
		{
			b200rt::traceMixedKernel<false><<<1, b200rt::kBlock, 0, warm>>>(s->view, none, 0, true);
			b200rt::traceSegmentsKernel<false><<<1, b200rt::kBlock, 0, warm>>>(s->view, b200rt::SegmentTable{}, true);
			b200rt::traceKernel<b200rt::kClosest, false><<<1, b200rt::kBlock, 0, warm>>>(s->view, nullptr, 0u, nullptr, nullptr, 0, true, nullptr);
			b200rt::traceKernel<b200rt::kShadow, false><<<1, b200rt::kBlock, 0, warm>>>(s->view, nullptr, 0u, nullptr, nullptr, 0, true, nullptr);
			b200rt::traceKernel<b200rt::kTShadow, false><<<1, b200rt::kBlock, 0, warm>>>(s->view, nullptr, 0u, nullptr, nullptr, 0, true, nullptr);
		}
		cudaError_t e = cudaStreamSynchronize(warm);
		cudaStreamDestroy(warm);
		if(e != cudaSuccess) return fail(B200RT_E_CUDA, std::string("warm-up launches: ") + cudaGetErrorString(e));
		std::lock_guard<std::mutex> lock(s->lane_mutex);
		const size_t stock = std::min<size_t>(64, 2 * std::max(1u, std::thread::hardware_concurrency()));
		while(s->free_lanes.size() < stock)
		{
			auto lane = std::make_unique<Lane>();
			if(cudaStreamCreateWithFlags(&lane->stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); break; }
			s->free_lanes.push_back(std::move(lane));
		}
	}
	// the flush combiner (kd_segments.cuh): the in-place jobs of b200rt_trace_jobs_begin -- the flushes of the renderer's sixteen ray
	// queues -- leave in shared launches
	// OFF unless B200RT_COMBINE=1: measured on path-traced and direct-lighting frames it reaches 2-4 k rays per launch and 8-16 x fewer
	// launches, and the frames get SLOWER (2.9 s -> 3.0 / 3.9 / 4.8 s at 1024 / 2048 / 4096 rays per launch): a render thread wants a
	// flush back within the ~180 us it takes to shade its other fiber group, and waiting for company spends that (profiles/r4u_*).
	if(!s->combiner && envLong("B200RT_COMBINE", 0) != 0)
	{
		auto combiner = std::make_unique<b200rt::Combiner>();
		for(int k = 0; k < 4; ++k)
		{
			cudaStream_t st = nullptr;
			CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
			combiner->streams.push_back(st);
		}
		s->combiner = std::move(combiner);
	}
	s->stats.upload_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
	s->built = true;
	return B200RT_OK;
}

int b200rt_get_bound(const b200rt_scene *s, float out6[6])
{
	if(!s || !out6) return fail(B200RT_E_INVALID, "null argument");
	if(!s->built) return fail(B200RT_E_INVALID, "scene not built");
	std::memcpy(out6, s->tree.bound, 6 * sizeof(float));
	return B200RT_OK;
}

int b200rt_get_stats(const b200rt_scene *s, b200rt_stats *out)
{
	if(!s || !out) return fail(B200RT_E_INVALID, "null argument");
	if(!s->built) return fail(B200RT_E_INVALID, "scene not built");
	*out = s->stats;
	return B200RT_OK;
}

int b200rt_update_face_flags(b200rt_scene *s, const uint8_t *flags, size_t n_faces)
{
	if(!s || !flags) return fail(B200RT_E_INVALID, "null argument");
	if(!s->built) return fail(B200RT_E_INVALID, "scene not built");
	if(n_faces != s->idx.size() / 4) return fail(B200RT_E_INVALID, "n_faces does not match the scene");
	CUDA_TRY(cudaSetDevice(s->device));
	s->flags.assign(flags, flags + n_faces);
	// patch the flag word (q1.w) of every record; small strided writes, done once per material change
	std::vector<float4> tris(s->n_tri_vec4);
	CUDA_TRY(cudaMemcpy(tris.data(), s->d_tris, tris.size() * sizeof(float4), cudaMemcpyDeviceToHost));
	for(size_t r = 0; r < s->record_of_ref.size(); ++r)
	{
		const uint32_t face = s->tree.leaf_refs[r];
		uint32_t fl = s->flags[face] & 7u;
		if(s->kind[face])
		{
			fl |= (s->kind[face] == 1) ? b200rt::kFlagBezier : b200rt::kFlagMoving;
			if(s->motion[size_t(s->motion_of_face[face])].nv == 4u) fl |= b200rt::kFlagQuad;
		}
		else
		{
			if(s->idx[4 * size_t(face) + 3] != kTriangle) fl |= b200rt::kFlagQuad;
			if(s->idx[4 * size_t(face) + 2] == kSphere) fl |= b200rt::kFlagSphere;
		}
		std::memcpy(&tris[size_t(s->record_of_ref[r]) + 1].w, &fl, 4);
	}
	CUDA_TRY(cudaMemcpy(s->d_tris, tris.data(), tris.size() * sizeof(float4), cudaMemcpyHostToDevice));
	return B200RT_OK;
}

// ---- device-buffer queries --------------------------------------------------------------------
int b200rt_trace_closest_device(b200rt_scene *s, const b200rt_ray *d_rays, size_t n, b200rt_hit *d_out, void *stream)
{
	const int rc = checkDeviceCall(s, d_rays, n, d_out);
	if(rc != B200RT_OK || n == 0) return rc;
	CUDA_TRY(cudaSetDevice(s->device));
	return launchTrace<b200rt::kClosest>(s, d_rays, n, d_out, static_cast<cudaStream_t>(stream), 0);
}

int b200rt_trace_shadow_device(b200rt_scene *s, const b200rt_ray *d_rays, size_t n, uint32_t *d_out, void *stream)
{
	const int rc = checkDeviceCall(s, d_rays, n, d_out);
	if(rc != B200RT_OK || n == 0) return rc;
	CUDA_TRY(cudaSetDevice(s->device));
	return launchTrace<b200rt::kShadow>(s, d_rays, n, d_out, static_cast<cudaStream_t>(stream), 0);
}

int b200rt_trace_tshadow_device(b200rt_scene *s, const b200rt_ray *d_rays, size_t n, int max_depth, b200rt_tshadow *d_out, void *stream)
{
	const int rc = checkDeviceCall(s, d_rays, n, d_out);
	if(rc != B200RT_OK) return rc;
	if(max_depth < 0 || max_depth > B200RT_TSHADOW_MAX) return fail(B200RT_E_INVALID, "max_depth must be in [0, B200RT_TSHADOW_MAX]");
	if(n == 0) return B200RT_OK;
	CUDA_TRY(cudaSetDevice(s->device));
	return launchTrace<b200rt::kTShadow>(s, d_rays, n, d_out, static_cast<cudaStream_t>(stream), max_depth);
}

// ---- generic entry points (query kind + flags [+ ray times]); everything else forwards here ------------------------
int b200rt_trace_timed_device(b200rt_scene *s, int query, unsigned flags, const b200rt_ray *d_rays, const float *d_times, size_t n, void *d_out, int max_depth, void *stream)
{
	const int rc = checkDeviceCall(s, d_rays, n, d_out);
	if(rc != B200RT_OK) return rc;
	if(query == B200RT_QUERY_TSHADOW && (max_depth < 0 || max_depth > B200RT_TSHADOW_MAX)) return fail(B200RT_E_INVALID, "max_depth must be in [0, B200RT_TSHADOW_MAX]");
	if(n == 0) return B200RT_OK;
	CUDA_TRY(cudaSetDevice(s->device));
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	switch(query)
	{
		case B200RT_QUERY_CLOSEST: return launchTrace<b200rt::kClosest>(s, d_rays, n, static_cast<b200rt_hit *>(d_out), st, 0, flags, false, d_times);
		case B200RT_QUERY_SHADOW: return launchTrace<b200rt::kShadow>(s, d_rays, n, static_cast<uint32_t *>(d_out), st, 0, flags, false, d_times);
		case B200RT_QUERY_TSHADOW: return launchTrace<b200rt::kTShadow>(s, d_rays, n, static_cast<b200rt_tshadow *>(d_out), st, max_depth, flags, false, d_times);
		default: return fail(B200RT_E_INVALID, "unknown query kind");
	}
}

int b200rt_trace_device(b200rt_scene *s, int query, unsigned flags, const b200rt_ray *d_rays, size_t n, void *d_out, int max_depth, void *stream)
{
	return b200rt_trace_timed_device(s, query, flags, d_rays, nullptr, n, d_out, max_depth, stream);
}

int b200rt_trace_timed(b200rt_scene *s, int query, unsigned flags, const b200rt_ray *rays, const float *times, size_t n, void *out, int max_depth)
{
	const bool known_pinned = (flags & B200RT_BUFFERS_PINNED) != 0u;
	switch(query)
	{
		case B200RT_QUERY_CLOSEST:
			return tracedStaged(s, rays, times, n, static_cast<b200rt_hit *>(out), [&](const b200rt_ray *d_rays, const float *d_times, size_t count, b200rt_hit *d_out, cudaStream_t st, bool small) {
				return launchTrace<b200rt::kClosest>(s, d_rays, count, d_out, st, 0, flags, small, d_times);
			}, known_pinned);
		case B200RT_QUERY_SHADOW:
			return tracedStaged(s, rays, times, n, static_cast<uint32_t *>(out), [&](const b200rt_ray *d_rays, const float *d_times, size_t count, uint32_t *d_out, cudaStream_t st, bool small) {
				return launchTrace<b200rt::kShadow>(s, d_rays, count, d_out, st, 0, flags, small, d_times);
			}, known_pinned);
		case B200RT_QUERY_TSHADOW:
			if(max_depth < 0 || max_depth > B200RT_TSHADOW_MAX) return fail(B200RT_E_INVALID, "max_depth must be in [0, B200RT_TSHADOW_MAX]");
			return tracedStaged(s, rays, times, n, static_cast<b200rt_tshadow *>(out), [&](const b200rt_ray *d_rays, const float *d_times, size_t count, b200rt_tshadow *d_out, cudaStream_t st, bool small) {
				return launchTrace<b200rt::kTShadow>(s, d_rays, count, d_out, st, max_depth, flags, small, d_times);
			}, known_pinned);
		default: return fail(B200RT_E_INVALID, "unknown query kind");
	}
}

int b200rt_trace(b200rt_scene *s, int query, unsigned flags, const b200rt_ray *rays, size_t n, void *out, int max_depth)
{
	return b200rt_trace_timed(s, query, flags, rays, nullptr, n, out, max_depth);
}

// ---- host-buffer queries ----------------------------------------------------------------------
int b200rt_trace_closest(b200rt_scene *s, const b200rt_ray *rays, size_t n, b200rt_hit *out) { return b200rt_trace_timed(s, B200RT_QUERY_CLOSEST, 0u, rays, nullptr, n, out, 0); }
int b200rt_trace_shadow(b200rt_scene *s, const b200rt_ray *rays, size_t n, uint32_t *out) { return b200rt_trace_timed(s, B200RT_QUERY_SHADOW, 0u, rays, nullptr, n, out, 0); }
int b200rt_trace_tshadow(b200rt_scene *s, const b200rt_ray *rays, size_t n, int max_depth, b200rt_tshadow *out) { return b200rt_trace_timed(s, B200RT_QUERY_TSHADOW, 0u, rays, nullptr, n, out, max_depth); }

int b200rt_trace_tshadow_deep_device(b200rt_scene *s, unsigned flags, const b200rt_ray *d_rays, const float *d_times, size_t n, int max_depth, int capacity, void *d_out, void *stream)
{
	const int rc = checkDeviceCall(s, d_rays, n, d_out);
	if(rc != B200RT_OK) return rc;
	if(max_depth < 0 || capacity < 1 || capacity > 4096 || max_depth > capacity) return fail(B200RT_E_INVALID, "need 0 <= max_depth <= capacity <= 4096");
	if(n > kMaxRaysPerTwoPass) return fail(B200RT_E_INVALID, "at most 2^26 rays per deep transparent-shadow call");
	if(n == 0) return B200RT_OK;
	CUDA_TRY(cudaSetDevice(s->device));
	// the kernels take the record capacity in the upper half of max_depth (kd_kernels.cuh, TShadowState); one launch (pair) covers n
	return launchTrace<b200rt::kTShadow>(s, d_rays, n, static_cast<b200rt_tshadow *>(d_out), static_cast<cudaStream_t>(stream), max_depth | (capacity << 16), flags, false, d_times);
}

int b200rt_trace_tshadow_deep(b200rt_scene *s, unsigned flags, const b200rt_ray *rays, const float *times, size_t n, int max_depth, int capacity, void *out)
{
	if(!s || (!rays && n) || (!out && n)) return fail(B200RT_E_INVALID, "null argument");
	if(!s->built) return fail(B200RT_E_INVALID, "scene not built: call b200rt_build first");
	if(max_depth < 0 || capacity < 1 || capacity > 4096 || max_depth > capacity) return fail(B200RT_E_INVALID, "need 0 <= max_depth <= capacity <= 4096");
	if(n > kMaxRaysPerTwoPass) return fail(B200RT_E_INVALID, "at most 2^26 rays per deep transparent-shadow call");
	if(n == 0) return B200RT_OK;
	CUDA_TRY(cudaSetDevice(s->device));
	// a rare query (shadow_depth > 8): plain copies around one launch, no staging pipeline
	const size_t record = 16u + 16u * size_t(capacity);
	b200rt_ray *d_rays = nullptr;
	float *d_times = nullptr;
	void *d_out = nullptr;
	cudaError_t e = cudaMalloc(&d_rays, n * sizeof(b200rt_ray));
	if(e == cudaSuccess) e = cudaMalloc(&d_out, n * record);
	if(e == cudaSuccess && times) e = cudaMalloc(&d_times, n * sizeof(float));
	if(e == cudaSuccess) e = cudaMemcpy(d_rays, rays, n * sizeof(b200rt_ray), cudaMemcpyHostToDevice);
	if(e == cudaSuccess && times) e = cudaMemcpy(d_times, times, n * sizeof(float), cudaMemcpyHostToDevice);
	int rc = B200RT_OK;
	if(e == cudaSuccess) rc = b200rt_trace_tshadow_deep_device(s, flags, d_rays, d_times, n, max_depth, capacity, d_out, nullptr);
	if(e == cudaSuccess && rc == B200RT_OK) e = cudaMemcpy(out, d_out, n * record, cudaMemcpyDeviceToHost);
	cudaFree(d_rays); cudaFree(d_times); cudaFree(d_out);
	if(e != cudaSuccess) return fail(B200RT_E_CUDA, std::string("b200rt_trace_tshadow_deep: ") + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
	return rc;
}

int b200rt_trace_jobs_begin(const b200rt_job *jobs, size_t n_jobs, b200rt_flight **out_flight)
{
	if((!jobs && n_jobs) || !out_flight) return fail(B200RT_E_INVALID, "null argument");
	*out_flight = nullptr;
	b200rt_flight *flight = new(std::nothrow) b200rt_flight;
	if(!flight) return fail(B200RT_E_MEMORY, "out of host memory");
	// in-place jobs of one scene with different query kinds share ONE launch (traceMixedKernel): with 16 render threads
	// flushing small batches the driver's launch path is the contended resource, not the GPU
	struct Bundle { b200rt_scene *scene; unsigned tree_space; const b200rt_job *job[3]; };
	std::vector<Bundle> bundles;
	std::vector<size_t> staged;
	auto launchBundle = [&](const Bundle &bundle) {
		b200rt_scene *s = bundle.scene;
		std::unique_ptr<Lane> lane;
		{
			std::lock_guard<std::mutex> lock(s->lane_mutex);
			if(!s->free_lanes.empty()) { lane = std::move(s->free_lanes.back()); s->free_lanes.pop_back(); }
		}
		if(!lane) lane = std::make_unique<Lane>();
		int rc = B200RT_OK;
		cudaError_t e = cudaSetDevice(s->device);
		if(e == cudaSuccess && !lane->stream) e = cudaStreamCreateWithFlags(&lane->stream, cudaStreamNonBlocking);
		if(e != cudaSuccess) rc = fail(B200RT_E_CUDA, std::string("b200rt_trace_jobs: ") + cudaGetErrorString(e));
		const int kinds = (bundle.job[0] ? 1 : 0) + (bundle.job[1] ? 1 : 0) + (bundle.job[2] ? 1 : 0);
		if(rc == B200RT_OK && kinds == 1)
		{
			const b200rt_job &job = *(bundle.job[0] ? bundle.job[0] : bundle.job[1] ? bundle.job[1] : bundle.job[2]);
			switch(job.query)
			{
				case B200RT_QUERY_CLOSEST: rc = launchTrace<b200rt::kClosest>(s, job.rays, job.n, static_cast<b200rt_hit *>(job.out), lane->stream, 0, job.flags, true, job.times); break;
				case B200RT_QUERY_SHADOW: rc = launchTrace<b200rt::kShadow>(s, job.rays, job.n, static_cast<uint32_t *>(job.out), lane->stream, 0, job.flags, true, job.times); break;
				default: rc = launchTrace<b200rt::kTShadow>(s, job.rays, job.n, static_cast<b200rt_tshadow *>(job.out), lane->stream, job.max_depth, job.flags, true, job.times); break;
			}
		}
		else if(rc == B200RT_OK)
		{
			b200rt::MixedBatch batch{};
			size_t warps = 0;
			for(int k = 0; k < 3; ++k)
			{
				if(!bundle.job[k]) continue;
				batch.rays[k] = bundle.job[k]->rays;
				batch.out[k] = bundle.job[k]->out;
				batch.n[k] = uint32_t(bundle.job[k]->n);
				batch.times[k] = bundle.job[k]->times;
				warps += (bundle.job[k]->n + 31) / 32;
			}
			const unsigned grid = unsigned((warps + b200rt::kBlock / 32 - 1) / (b200rt::kBlock / 32));
			if(s->has_spheres) b200rt::traceMixedKernel<true><<<grid, b200rt::kBlock, 0, lane->stream>>>(s->view, batch, bundle.job[2] ? bundle.job[2]->max_depth : 0, bundle.tree_space != 0u);
			else b200rt::traceMixedKernel<false><<<grid, b200rt::kBlock, 0, lane->stream>>>(s->view, batch, bundle.job[2] ? bundle.job[2]->max_depth : 0, bundle.tree_space != 0u);
			++g_launches;
			e = cudaGetLastError();
			if(e != cudaSuccess) rc = fail(B200RT_E_CUDA, std::string("traceMixedKernel: ") + cudaGetErrorString(e));
		}
		if(rc != B200RT_OK) flight->note(rc);
		flight->lanes.push_back({s, std::move(lane)});
	};
	for(size_t j = 0; j < n_jobs; ++j)
	{
		const b200rt_job &job = jobs[j];
		if(job.n == 0) continue;
		int rc = checkDeviceCall(job.scene, job.rays, job.n, job.out);
		if(rc == B200RT_OK && (job.query < B200RT_QUERY_CLOSEST || job.query > B200RT_QUERY_TSHADOW)) rc = fail(B200RT_E_INVALID, "unknown query kind");
		if(rc == B200RT_OK && job.query == B200RT_QUERY_TSHADOW && (job.max_depth < 0 || job.max_depth > B200RT_TSHADOW_MAX)) rc = fail(B200RT_E_INVALID, "max_depth must be in [0, B200RT_TSHADOW_MAX]");
		if(rc != B200RT_OK) { flight->note(rc); continue; }
		const bool pinned = (job.flags & B200RT_BUFFERS_PINNED) != 0u || (isPinned(job.rays) && isPinned(job.out) && (!job.times || isPinned(job.times)));
		if(!pinned || job.n > kDirectRays) { staged.push_back(j); continue; }
		if(job.scene->combiner)
		{
			// rides with the flushes of the other render threads (flush combiner); b200rt_trace_jobs_end waits for its launch
			b200rt::Combiner &c = *job.scene->combiner;
			auto ticket = std::make_shared<b200rt::Ticket>();
			bool full = false;
			{
				std::lock_guard<std::mutex> lock(c.mutex);
				c.pending.push_back(b200rt::PendingJob{job, ticket});
				c.pending_rays += job.n;
				full = c.pending_rays >= combineTargetRays() || c.pending.size() >= size_t(b200rt::kMaxSegments);
			}
			flight->tickets.push_back({job.scene, std::move(ticket)});
			if(full) flushCombiner(job.scene);
			continue;
		}
		const unsigned tree_space = job.flags & B200RT_RAYS_TREE_SPACE;
		Bundle *home = nullptr;
		for(Bundle &bundle : bundles)
			if(bundle.scene == job.scene && bundle.tree_space == tree_space && !bundle.job[job.query]) { home = &bundle; break; }
		if(!home)
		{
			bundles.push_back(Bundle{job.scene, tree_space, {nullptr, nullptr, nullptr}});
			home = &bundles.back();
		}
		home->job[job.query] = &job;
	}
	for(const Bundle &bundle : bundles) launchBundle(bundle);
	for(size_t j : staged)
	{
		const int rc = b200rt_trace_timed(jobs[j].scene, jobs[j].query, jobs[j].flags & ~unsigned(B200RT_BUFFERS_PINNED), jobs[j].rays, jobs[j].times, jobs[j].n, jobs[j].out, jobs[j].max_depth);
		if(rc != B200RT_OK) flight->note(rc);
	}
	*out_flight = flight;
	return flight->first_error;
}

int b200rt_trace_jobs_end(b200rt_flight *flight)
{
	if(!flight) return fail(B200RT_E_INVALID, "null argument");
	for(auto &f : flight->lanes)
	{
		if(f.lane->stream)
		{
			const cudaError_t e = cudaStreamSynchronize(f.lane->stream);
			if(e != cudaSuccess && flight->first_error == B200RT_OK) { flight->first_error = B200RT_E_CUDA; flight->error_text = std::string("b200rt_trace_jobs: ") + cudaGetErrorString(e); }
		}
		std::lock_guard<std::mutex> lock(f.scene->lane_mutex);
		f.scene->free_lanes.push_back(std::move(f.lane));
	}
	for(auto &t : flight->tickets)
	{
		b200rt::CombinedBatch *batch = t.ticket->batch.load(std::memory_order_acquire);
		if(!batch) flushCombiner(t.scene); // still waiting for company: launch what is there, this job included
		for(unsigned spins = 0; !(batch = t.ticket->batch.load(std::memory_order_acquire)); ++spins)
		{
			// another thread has taken the list and is launching it this moment
			if(spins < 2000u) {
#if defined(__x86_64__)
				__builtin_ia32_pause();
#endif
			}
			else std::this_thread::yield();
		}
		if(batch->rc != B200RT_OK)
		{
			if(flight->first_error == B200RT_OK) { flight->first_error = batch->rc; flight->error_text = batch->error; }
		}
		else
		{
			const cudaError_t e = cudaEventSynchronize(batch->event);
			if(e != cudaSuccess && flight->first_error == B200RT_OK) { flight->first_error = B200RT_E_CUDA; flight->error_text = std::string("b200rt_trace_jobs (combined launch): ") + cudaGetErrorString(e); }
		}
		if(batch->users.fetch_sub(1, std::memory_order_acq_rel) == 1)
		{
			b200rt::Combiner &c = *t.scene->combiner;
			std::lock_guard<std::mutex> lock(c.mutex);
			c.free_batches.push_back(batch);
		}
	}
	const int rc = flight->first_error;
	if(rc != B200RT_OK) g_last_error = flight->error_text;
	delete flight;
	return rc;
}

int b200rt_trace_jobs(const b200rt_job *jobs, size_t n_jobs)
{
	b200rt_flight *flight = nullptr;
	const int rc = b200rt_trace_jobs_begin(jobs, n_jobs, &flight);
	if(!flight) return rc;
	return b200rt_trace_jobs_end(flight);
}

int b200rt_host_alloc(void **ptr, size_t bytes)
{
	if(!ptr) return fail(B200RT_E_INVALID, "null argument");
	*ptr = nullptr;
	CUDA_TRY(cudaMallocHost(ptr, bytes ? bytes : 1));
	return B200RT_OK;
}

int b200rt_host_free(void *ptr)
{
	if(!ptr) return B200RT_OK;
	CUDA_TRY(cudaFreeHost(ptr));
	return B200RT_OK;
}

} // extern "C"

// ---- diagnostics: host builder only -----------------------------------------------------------
struct b200rt_host_tree
{
	b200rt::HostTree tree;
};

namespace {
b200rt::BuildConfig configFrom(const b200rt_build_params *params)
{
	b200rt::BuildConfig c;
	if(params)
	{
		c.max_depth = params->max_depth;
		c.max_leaf_size = params->max_leaf_size;
		c.cost_ratio = params->cost_ratio;
		c.empty_bonus = params->empty_bonus > 0.f ? params->empty_bonus : -1.f;
		c.threads = params->build_threads;
	}
	return c;
}
} // namespace

extern "C" {

int b200rt_host_tree_build(const float *xyz, size_t n_verts, const uint32_t *idx, size_t n_faces, const b200rt_build_params *params, b200rt_host_tree **out)
{
	if(!out || (!xyz && n_verts) || (!idx && n_faces)) return fail(B200RT_E_INVALID, "null argument");
	*out = nullptr;
	for(size_t f = 0; f < n_faces; ++f)
		for(int k = 0; k < 4; ++k)
		{
			const uint32_t v = idx[4 * f + k];
			if(k == 3 && v == kTriangle) continue;
			if(k == 2 && (v == kSphere || v == kBox) && idx[4 * f + 3] == kTriangle) continue; // sphere / box face (kd_build.h)
			if(v >= n_verts) return fail(B200RT_E_INVALID, "face references a vertex >= n_verts");
		}
	try
	{
		auto t = std::make_unique<b200rt_host_tree>();
		b200rt::buildKdTree(b200rt::MeshView{xyz, n_verts, idx, n_faces}, configFrom(params), t->tree);
		*out = t.release();
	}
	catch(const std::exception &e) { return fail(B200RT_E_MEMORY, std::string("host build failed: ") + e.what()); }
	return B200RT_OK;
}

int b200rt_host_tree_sizes(const b200rt_host_tree *t, size_t *n_nodes, size_t *n_refs)
{
	if(!t) return fail(B200RT_E_INVALID, "null argument");
	if(n_nodes) *n_nodes = t->tree.nodes.size();
	if(n_refs) *n_refs = t->tree.leaf_refs.size();
	return B200RT_OK;
}

int b200rt_host_tree_export(const b200rt_host_tree *t, uint32_t *node_a, uint32_t *node_b, uint32_t *refs, float bound6[6])
{
	if(!t) return fail(B200RT_E_INVALID, "null argument");
	for(size_t i = 0; i < t->tree.nodes.size(); ++i)
	{
		if(node_a) node_a[i] = t->tree.nodes[i].a;
		if(node_b) node_b[i] = t->tree.nodes[i].b;
	}
	if(refs && !t->tree.leaf_refs.empty()) std::memcpy(refs, t->tree.leaf_refs.data(), t->tree.leaf_refs.size() * sizeof(uint32_t));
	if(bound6) std::memcpy(bound6, t->tree.bound, 6 * sizeof(float));
	return B200RT_OK;
}

void b200rt_host_tree_destroy(b200rt_host_tree *t) { delete t; }

} // extern "C"
