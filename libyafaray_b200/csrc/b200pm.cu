// libyafaray_b200/csrc/b200pm.cu -- the C ABI of the photon-map queries (include/b200pm.h): map ownership, packing of the
// reference's point kd-tree into the 16-byte node layout the kernels read, host-buffer and device-buffer lookups.
#include "../../include/b200pm.h"
#include "../../include/b200rt.h"
#include "pm_build.h"
#include "pm_kernels.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

namespace b200 {
int failWith(int code, const std::string &msg); // b200rt.cu: text for b200rt_last_error()
void countLaunches(uint64_t n);                  // b200rt.cu: b200rt_launch_count()
} // namespace b200

namespace {

using b200::failWith;

#define PM_CUDA_TRY(expr)                                                                                                   \
	do {                                                                                                                    \
		const cudaError_t e_ = (expr);                                                                                      \
		if(e_ != cudaSuccess)                                                                                               \
			return failWith(B200RT_E_CUDA, std::string(#expr) + ": " + cudaGetErrorName(e_) + " (" + cudaGetErrorString(e_) + ")"); \
	} while(0)

double now()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// device scratch of the host-buffer calls, grown on demand and kept with the map
struct Scratch
{
	void *ptr = nullptr;
	size_t cap = 0;
	int reserve(size_t bytes)
	{
		if(bytes <= cap) return B200RT_OK;
		if(ptr) cudaFree(ptr);
		ptr = nullptr;
		cap = 0;
		PM_CUDA_TRY(cudaMalloc(&ptr, bytes));
		cap = bytes;
		return B200RT_OK;
	}
	~Scratch() { if(ptr) cudaFree(ptr); }
};

} // namespace

struct b200pm_map
{
	int device = 0;
	size_t n_photons = 0, n_nodes = 0;
	bool has_dirs = false;
	uint4 *d_nodes = nullptr;
	float4 *d_dirs = nullptr;
	b200pm_stats stats{};
	// host-buffer calls: two lanes (stream + device scratch) so that the copies of one chunk overlap the kernel of the next
	struct HostLane
	{
		cudaStream_t stream = nullptr;
		Scratch in, out;
	} lanes[2];
	std::mutex host_call; // host-buffer calls on one map take turns (they share the lanes)
	bool smem_opt_in = false;

	~b200pm_map()
	{
		cudaSetDevice(device);
		if(d_nodes) cudaFree(d_nodes);
		if(d_dirs) cudaFree(d_dirs);
		for(auto &lane : lanes)
			if(lane.stream) cudaStreamDestroy(lane.stream);
	}
};

namespace {

// Tuning aids (tools/pm_sweep.py sweeps them through b200pm_debug_set_tuning; the defaults are the measured best,
// profiles/r6*_pm_*): environment, read once --
//   B200PM_KERNEL=plain|phased|phased1   gather above smem_k: the plain per-thread loop (pmLookupKernel), the phased state machine
//                         with one stack pop per step (pmLookupPhasedKernel<.,true>), or with the pop loop inside a step
//   B200PM_ROUND=<n>      node visits per round of the phased kernel
//   B200PM_PATIENCE=<n>   lanes that must have a make_heap pending before the warp does it (1 = no waiting)
//   B200PM_SMEM_K=<k>     largest k whose heaps live in shared memory (0 = always in `found`)
struct Tuning
{
	// defaults from profiles/r6c_pm_sweep.jsonl (1 M photons, 1 M points): k = 100 plain 10.1 ms, phased 9.7, phased + single pop 8.9,
	// + patience 16 7.0; k = 8 with heaps in shared memory: plain 2.4 ms, phased 2.9 -- small k keeps the plain loop
	int kernel = 2; // 0 plain, 1 phased, 2 phased + single pop
	bool small_k_plain = true; // k <= smem_k: the plain loop with shared-memory heaps, unless a kernel was asked for explicitly
	int round_steps = 8;
	int patience = 16;
	uint32_t smem_k = 16;
	Tuning()
	{
		if(const char *e = std::getenv("B200PM_KERNEL"))
		{
			kernel = std::string(e) == "plain" ? 0 : (std::string(e) == "phased1" ? 1 : 2);
			small_k_plain = false;
		}
		if(const char *e = std::getenv("B200PM_ROUND")) { const long v = std::atol(e); if(v >= 1 && v <= 4096) round_steps = int(v); }
		if(const char *e = std::getenv("B200PM_PATIENCE")) { const long v = std::atol(e); if(v >= 1 && v <= 32) patience = int(v); }
		if(const char *e = std::getenv("B200PM_SMEM_K")) { const long v = std::atol(e); if(v >= 0 && uint32_t(v) <= b200pm::kPmSmemK) smem_k = uint32_t(v); }
	}
};
Tuning &tuning()
{
	static Tuning t;
	return t;
}

int launchGather(b200pm_map *map, const float *d_points, size_t n_points, uint32_t k, float sq_radius, const float *d_sq_radii, b200pm_found *d_found,
                 uint32_t *d_n_found, float *d_sq_radius_out, cudaStream_t stream)
{
	if(!n_points) return B200RT_OK;
	const Tuning t = tuning();
	const unsigned blocks = unsigned((n_points + b200pm::kPmThreads - 1) / b200pm::kPmThreads);
	uint2 *found2 = reinterpret_cast<uint2 *>(d_found);
	const bool in_smem = k <= t.smem_k;
	const size_t smem = in_smem ? size_t(k) * b200pm::kPmThreads * sizeof(uint2) : 0;
	if(in_smem && !map->smem_opt_in)
	{
		const int most = int(size_t(b200pm::kPmSmemK) * b200pm::kPmThreads * sizeof(uint2));
		PM_CUDA_TRY(cudaFuncSetAttribute(b200pm::pmLookupKernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, most));
		PM_CUDA_TRY(cudaFuncSetAttribute(b200pm::pmLookupPhasedKernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, most));
		PM_CUDA_TRY(cudaFuncSetAttribute(b200pm::pmLookupPhasedKernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, most));
		map->smem_opt_in = true;
	}
#define PM_ARGS map->d_nodes, map->d_dirs, d_points, nullptr, uint32_t(n_points), k, sq_radius, d_sq_radii, found2, d_n_found, d_sq_radius_out, nullptr
	if(t.kernel == 0 || (in_smem && t.small_k_plain))
	{
		if(in_smem) b200pm::pmLookupKernel<0><<<blocks, b200pm::kPmThreads, smem, stream>>>(PM_ARGS);
		else b200pm::pmLookupKernel<1><<<blocks, b200pm::kPmThreads, 0, stream>>>(PM_ARGS);
	}
	else if(t.kernel == 1)
	{
		if(in_smem) b200pm::pmLookupPhasedKernel<0, false><<<blocks, b200pm::kPmThreads, smem, stream>>>(PM_ARGS, t.round_steps, t.patience);
		else b200pm::pmLookupPhasedKernel<1, false><<<blocks, b200pm::kPmThreads, 0, stream>>>(PM_ARGS, t.round_steps, t.patience);
	}
	else
	{
		if(in_smem) b200pm::pmLookupPhasedKernel<0, true><<<blocks, b200pm::kPmThreads, smem, stream>>>(PM_ARGS, t.round_steps, t.patience);
		else b200pm::pmLookupPhasedKernel<1, true><<<blocks, b200pm::kPmThreads, 0, stream>>>(PM_ARGS, t.round_steps, t.patience);
	}
#undef PM_ARGS
	PM_CUDA_TRY(cudaGetLastError());
	b200::countLaunches(1);
	return B200RT_OK;
}

int launchNearest(b200pm_map *map, const float *d_points, const float *d_normals, size_t n_points, float dist, uint32_t *d_out, cudaStream_t stream)
{
	if(!n_points) return B200RT_OK;
	const Tuning t = tuning();
	const unsigned blocks = unsigned((n_points + b200pm::kPmThreads - 1) / b200pm::kPmThreads);
#define PM_ARGS map->d_nodes, map->d_dirs, d_points, d_normals, uint32_t(n_points), 1u, dist, nullptr, nullptr, nullptr, nullptr, d_out
	// findNearest has no heap work to gather: the plain loop is the fastest (profiles/r6b_pm_sweep.jsonl); the phased kernels only on request
	if(t.kernel == 0 || !std::getenv("B200PM_NEAREST_PHASED")) b200pm::pmLookupKernel<2><<<blocks, b200pm::kPmThreads, 0, stream>>>(PM_ARGS);
	else if(t.kernel == 1) b200pm::pmLookupPhasedKernel<2, false><<<blocks, b200pm::kPmThreads, 0, stream>>>(PM_ARGS, t.round_steps, t.patience);
	else b200pm::pmLookupPhasedKernel<2, true><<<blocks, b200pm::kPmThreads, 0, stream>>>(PM_ARGS, t.round_steps, t.patience);
#undef PM_ARGS
	PM_CUDA_TRY(cudaGetLastError());
	b200::countLaunches(1);
	return B200RT_OK;
}

constexpr size_t kMaxPoints = size_t(1) << 31; // blocks * 64 threads must fit the x grid dimension and uint32 point ids

size_t alignUp(size_t v) { return (v + 255) & ~size_t(255); }

// the median splits order photons by coordinate: a NaN has no place in that order (std::nth_element would be undefined)
bool allFinite(const float *pos, size_t n)
{
	for(size_t i = 0; i < 3 * n; ++i)
		if(!std::isfinite(pos[i])) return false;
	return true;
}

} // namespace

extern "C" {

int b200pm_host_tree_build(const float *pos, size_t n, int build_threads, uint32_t *a, uint32_t *b)
{
	if(!pos || !a || !b || !n || n >= (size_t(1) << 29)) return failWith(B200RT_E_INVALID, "b200pm_host_tree_build: need 1 <= n < 2^29 photons and non-null arrays");
	if(!allFinite(pos, n)) return failWith(B200RT_E_INVALID, "b200pm_host_tree_build: photon positions must be finite");
	try
	{
		b200pm::HostTree tree;
		b200pm::buildTree(pos, n, build_threads, tree);
		std::memcpy(a, tree.a.data(), 4 * tree.a.size());
		std::memcpy(b, tree.b.data(), 4 * tree.b.size());
	}
	catch(const std::bad_alloc &) { return failWith(B200RT_E_MEMORY, "out of host memory"); }
	catch(const std::exception &e) { return failWith(B200RT_E_INVALID, e.what()); }
	return B200RT_OK;
}

int b200pm_debug_set_tuning(int kernel, int round_steps, int smem_k, int patience)
{
	Tuning &t = tuning();
	if(kernel >= 0 && kernel <= 2) { t.kernel = kernel; t.small_k_plain = false; }
	if(kernel == 3) { t.kernel = 2; t.small_k_plain = true; } // the library default: phased + single pop above smem_k, plain below
	if(round_steps >= 1) t.round_steps = round_steps;
	if(smem_k >= 0 && uint32_t(smem_k) <= b200pm::kPmSmemK) t.smem_k = uint32_t(smem_k);
	if(patience >= 1 && patience <= 32) t.patience = patience;
	return B200RT_OK;
}

int b200pm_create(int device, const float *pos, const float *dir, size_t n, int build_threads, b200pm_map **out)
{
	if(!out) return failWith(B200RT_E_INVALID, "null argument");
	*out = nullptr;
	if(!pos || !n || n >= (size_t(1) << 29)) return failWith(B200RT_E_INVALID, "b200pm_create: need 1 <= n < 2^29 photons");
	if(!allFinite(pos, n)) return failWith(B200RT_E_INVALID, "b200pm_create: photon positions must be finite");
	int count = 0;
	const int rc = b200rt_device_count(&count);
	if(rc != B200RT_OK) return rc;
	if(device < 0 || device >= count) return failWith(B200RT_E_NO_DEVICE, "b200pm_create: no such CUDA device");
	PM_CUDA_TRY(cudaSetDevice(device));
	b200pm_map *map = nullptr;
	try
	{
		map = new b200pm_map;
		map->device = device;
		map->n_photons = n;
		map->n_nodes = 2 * n - 1;
		map->has_dirs = dir != nullptr;
		const double t0 = now();
		b200pm::HostTree tree;
		b200pm::buildTree(pos, n, build_threads, tree);
		// pack: the leaf carries its photon's position
		std::vector<uint4> nodes(map->n_nodes);
		for(size_t i = 0; i < map->n_nodes; ++i)
		{
			if((tree.b[i] & 3u) == 3u)
			{
				const uint32_t photon = tree.a[i];
				uint32_t xyz[3];
				std::memcpy(xyz, pos + 3 * size_t(photon), 12);
				nodes[i] = make_uint4(xyz[0], xyz[1], xyz[2], (photon << 2) | 3u);
			}
			else
				nodes[i] = make_uint4(tree.a[i], 0u, 0u, tree.b[i]);
		}
		const double t1 = now();
		map->stats.n_photons = n;
		map->stats.n_nodes = map->n_nodes;
		map->stats.depth = tree.depth;
		map->stats.build_seconds = t1 - t0;
		cudaError_t e = cudaMalloc(&map->d_nodes, sizeof(uint4) * map->n_nodes);
		if(e == cudaSuccess) e = cudaMemcpy(map->d_nodes, nodes.data(), sizeof(uint4) * map->n_nodes, cudaMemcpyHostToDevice);
		map->stats.device_bytes = sizeof(uint4) * map->n_nodes;
		if(e == cudaSuccess && dir)
		{
			std::vector<float4> dirs(n);
			for(size_t i = 0; i < n; ++i) dirs[i] = make_float4(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2], 0.f);
			e = cudaMalloc(&map->d_dirs, sizeof(float4) * n);
			if(e == cudaSuccess) e = cudaMemcpy(map->d_dirs, dirs.data(), sizeof(float4) * n, cudaMemcpyHostToDevice);
			map->stats.device_bytes += sizeof(float4) * n;
		}
		for(auto &lane : map->lanes)
			if(e == cudaSuccess) e = cudaStreamCreateWithFlags(&lane.stream, cudaStreamNonBlocking);
		if(e != cudaSuccess)
		{
			delete map;
			return failWith(B200RT_E_CUDA, std::string("b200pm_create: ") + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
		}
		map->stats.upload_seconds = now() - t1;
	}
	catch(const std::bad_alloc &) { delete map; return failWith(B200RT_E_MEMORY, "out of host memory"); }
	catch(const std::exception &e) { delete map; return failWith(B200RT_E_INVALID, e.what()); }
	*out = map;
	return B200RT_OK;
}

void b200pm_destroy(b200pm_map *map) { delete map; }

int b200pm_get_stats(const b200pm_map *map, b200pm_stats *out)
{
	if(!map || !out) return failWith(B200RT_E_INVALID, "null argument");
	*out = map->stats;
	return B200RT_OK;
}

int b200pm_gather_device(b200pm_map *map, const float *d_points, size_t n_points, uint32_t k, float sq_radius, const float *d_sq_radii, b200pm_found *d_found,
                         uint32_t *d_n_found, float *d_sq_radius_out, void *stream)
{
	if(!map || (n_points && (!d_points || !d_found || !d_n_found))) return failWith(B200RT_E_INVALID, "b200pm_gather_device: null argument");
	if(!k) return failWith(B200RT_E_INVALID, "b200pm_gather: k must be at least 1");
	if(n_points >= kMaxPoints) return failWith(B200RT_E_INVALID, "b200pm_gather: at most 2^31 - 1 points per call");
	PM_CUDA_TRY(cudaSetDevice(map->device));
	return launchGather(map, d_points, n_points, k, sq_radius, d_sq_radii, d_found, d_n_found, d_sq_radius_out, static_cast<cudaStream_t>(stream));
}

int b200pm_gather(b200pm_map *map, const float *points, size_t n_points, uint32_t k, float sq_radius, const float *sq_radii, b200pm_found *found, uint32_t *n_found,
                  float *sq_radius_out)
{
	if(!map || (n_points && (!points || !found || !n_found))) return failWith(B200RT_E_INVALID, "b200pm_gather: null argument");
	if(!k) return failWith(B200RT_E_INVALID, "b200pm_gather: k must be at least 1");
	if(n_points >= kMaxPoints) return failWith(B200RT_E_INVALID, "b200pm_gather: at most 2^31 - 1 points per call");
	if(!n_points) return B200RT_OK;
	std::lock_guard<std::mutex> lock(map->host_call);
	PM_CUDA_TRY(cudaSetDevice(map->device));
	// chunks of about 32 MiB of results alternate between the two lanes: with page-locked caller buffers the copy back of one
	// chunk and the copy in of the next overlap the kernel in between (pageable buffers take the same path, the driver stages them)
	const size_t per_point = sizeof(b200pm_found) * size_t(k) + 8;
	size_t chunk = (size_t(32) << 20) / per_point;
	chunk = std::max<size_t>(4096, std::min<size_t>(chunk, size_t(1) << 20));
	chunk = std::min(chunk, n_points);
	const size_t in_points = alignUp(12 * chunk), in_radii = alignUp(4 * chunk);
	const size_t out_found = alignUp(sizeof(b200pm_found) * chunk * k), out_count = alignUp(4 * chunk), out_radius = alignUp(4 * chunk);
	for(auto &lane : map->lanes)
	{
		int rc = lane.in.reserve(in_points + in_radii);
		if(rc == B200RT_OK) rc = lane.out.reserve(out_found + out_count + out_radius);
		if(rc != B200RT_OK) return rc;
		if(n_points <= chunk) break; // one chunk: one lane
	}
	size_t index = 0;
	for(size_t off = 0; off < n_points; off += chunk, ++index)
	{
		auto &lane = map->lanes[index & 1];
		const size_t m = std::min(chunk, n_points - off);
		char *d_in = static_cast<char *>(lane.in.ptr), *d_out = static_cast<char *>(lane.out.ptr);
		float *d_points = reinterpret_cast<float *>(d_in), *d_radii = reinterpret_cast<float *>(d_in + in_points);
		b200pm_found *d_found = reinterpret_cast<b200pm_found *>(d_out);
		uint32_t *d_count = reinterpret_cast<uint32_t *>(d_out + out_found);
		float *d_radius = reinterpret_cast<float *>(d_out + out_found + out_count);
		PM_CUDA_TRY(cudaMemcpyAsync(d_points, points + 3 * off, 12 * m, cudaMemcpyHostToDevice, lane.stream));
		if(sq_radii) PM_CUDA_TRY(cudaMemcpyAsync(d_radii, sq_radii + off, 4 * m, cudaMemcpyHostToDevice, lane.stream));
		const int rc = launchGather(map, d_points, m, k, sq_radius, sq_radii ? d_radii : nullptr, d_found, d_count, d_radius, lane.stream);
		if(rc != B200RT_OK) return rc;
		PM_CUDA_TRY(cudaMemcpyAsync(found + off * k, d_found, sizeof(b200pm_found) * m * k, cudaMemcpyDeviceToHost, lane.stream));
		PM_CUDA_TRY(cudaMemcpyAsync(n_found + off, d_count, 4 * m, cudaMemcpyDeviceToHost, lane.stream));
		if(sq_radius_out) PM_CUDA_TRY(cudaMemcpyAsync(sq_radius_out + off, d_radius, 4 * m, cudaMemcpyDeviceToHost, lane.stream));
	}
	for(auto &lane : map->lanes) PM_CUDA_TRY(cudaStreamSynchronize(lane.stream));
	return B200RT_OK;
}

int b200pm_find_nearest_device(b200pm_map *map, const float *d_points, const float *d_normals, size_t n_points, float dist, uint32_t *d_out_photon, void *stream)
{
	if(!map || (n_points && (!d_points || !d_normals || !d_out_photon))) return failWith(B200RT_E_INVALID, "b200pm_find_nearest_device: null argument");
	if(!map->has_dirs) return failWith(B200RT_E_INVALID, "b200pm_find_nearest: the map was created without photon directions");
	if(n_points >= kMaxPoints) return failWith(B200RT_E_INVALID, "b200pm_find_nearest: at most 2^31 - 1 points per call");
	PM_CUDA_TRY(cudaSetDevice(map->device));
	return launchNearest(map, d_points, d_normals, n_points, dist, d_out_photon, static_cast<cudaStream_t>(stream));
}

int b200pm_find_nearest(b200pm_map *map, const float *points, const float *normals, size_t n_points, float dist, uint32_t *out_photon)
{
	if(!map || (n_points && (!points || !normals || !out_photon))) return failWith(B200RT_E_INVALID, "b200pm_find_nearest: null argument");
	if(!map->has_dirs) return failWith(B200RT_E_INVALID, "b200pm_find_nearest: the map was created without photon directions");
	if(n_points >= kMaxPoints) return failWith(B200RT_E_INVALID, "b200pm_find_nearest: at most 2^31 - 1 points per call");
	if(!n_points) return B200RT_OK;
	std::lock_guard<std::mutex> lock(map->host_call);
	PM_CUDA_TRY(cudaSetDevice(map->device));
	const size_t in_points = alignUp(12 * n_points);
	auto &lane = map->lanes[0];
	int rc = lane.in.reserve(2 * in_points);
	if(rc == B200RT_OK) rc = lane.out.reserve(alignUp(4 * n_points));
	if(rc != B200RT_OK) return rc;
	char *d_in = static_cast<char *>(lane.in.ptr);
	float *d_points = reinterpret_cast<float *>(d_in), *d_normals = reinterpret_cast<float *>(d_in + in_points);
	uint32_t *d_out = static_cast<uint32_t *>(lane.out.ptr);
	PM_CUDA_TRY(cudaMemcpyAsync(d_points, points, 12 * n_points, cudaMemcpyHostToDevice, lane.stream));
	PM_CUDA_TRY(cudaMemcpyAsync(d_normals, normals, 12 * n_points, cudaMemcpyHostToDevice, lane.stream));
	rc = launchNearest(map, d_points, d_normals, n_points, dist, d_out, lane.stream);
	if(rc != B200RT_OK) return rc;
	PM_CUDA_TRY(cudaMemcpyAsync(out_photon, d_out, 4 * n_points, cudaMemcpyDeviceToHost, lane.stream));
	PM_CUDA_TRY(cudaStreamSynchronize(lane.stream));
	return B200RT_OK;
}

} // extern "C"
