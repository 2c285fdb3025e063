// libyafaray_b200/csrc/pm_kernels.cuh -- photon-map lookups on sm_100a (include/b200pm.h).
//
// One thread answers one query point: it walks the reference's point kd-tree in the reference's order
// (PointKdTree::lookup, include/photon/pkdtree.h:221-279) and feeds the photons it meets to the reference's lookup procedure
// (PhotonGather src/photon/photon.cc:26-44, NearestPhoton include/photon/photon.h:101-109).  The arithmetic is the reference's,
// operation for operation (__f*_rn: no FMA contraction), so counts, orders, distances and radii come out bit-identical.
//
// Data layout (HBM): one 16-byte node per tree node, read with a single 128-bit load --
//   interior  x = split (float bits)   w = (right child << 2) | axis         left child = this + 1
//   leaf      x y z = photon position  w = (photon index << 2) | 3
// -- the leaf carries the position, so testing a photon costs no second dependent load; dirs (float4 per photon) are read only
// by findNearest, and only for photons inside the radius.
//
// Per-thread state: the traversal stack (far child, split, axis: 8 bytes per level) lives in local memory, interleaved per lane
// by the hardware (one 128-byte line per warp and level); the k-entry max-heap of a gather lives directly in the caller's `found`
// array, or -- for small k (<= 16 by default, b200pm.cu Tuning; at most kPmSmemK) -- in shared memory, interleaved per thread
// ([entry][thread], 8-byte entries) and copied out to `found` at the end: at k = 100 shared-memory heaps cost 3.7x in occupancy.
#pragma once
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include "../../include/b200pm.h"

namespace b200pm {

constexpr int kPmThreads = 64;      // threads per block of the lookup kernels
constexpr uint32_t kPmSmemK = 256;  // largest k whose heaps live in shared memory (256 * 8 * 64 = 128 KiB per block)
constexpr int kPmStack = 40;        // tree depth <= 30 for n < 2^29 photons; the lookup pushes one entry per level + sentinel
constexpr uint32_t kPmEmpty = 0xFFFFFFFFu;

// The lookup itself is __host__ __device__: tests/native/pm_host_model.cu runs the SAME code on the CPU (one "thread" at a time)
// so that the CPU test-suite checks the kernel's logic against the reference without a GPU.  On the device the arithmetic uses
// the round-to-nearest intrinsics (never contracted into FMAs); the host build has no FMA target.
#ifdef __CUDA_ARCH__
#define PM_HD __host__ __device__ __forceinline__
__device__ __forceinline__ float pmSub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float pmAdd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float pmMul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float pmBitsToFloat(uint32_t u) { return __uint_as_float(u); }
__device__ __forceinline__ uint32_t pmFloatToBits(float f) { return __float_as_uint(f); }
__device__ __forceinline__ uint4 pmLoadNode(const uint4 *p) { return __ldg(p); }
__device__ __forceinline__ float4 pmLoadDir(const float4 *p) { return __ldg(p); }
#else
#define PM_HD __host__ __device__ inline
inline float pmSub(float a, float b) { return a - b; }
inline float pmAdd(float a, float b) { return a + b; }
inline float pmMul(float a, float b) { return a * b; }
inline float pmBitsToFloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t pmFloatToBits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline uint4 pmLoadNode(const uint4 *p) { return *p; }
inline float4 pmLoadDir(const float4 *p) { return *p; }
#endif

struct HeapSmem
{
	uint2 *base; // &smem[threadIdx.x]; entry j at base[j * kPmThreads]
	PM_HD uint2 get(int j) const { return base[j * kPmThreads]; }
	PM_HD void set(int j, uint2 v) const { base[j * kPmThreads] = v; }
};
struct HeapGlobal
{
	uint2 *base; // &found[point * k]
	PM_HD uint2 get(int j) const { return base[j]; }
	PM_HD void set(int j, uint2 v) const { base[j] = v; }
};

PM_HD float heapKey(uint2 v) { return pmBitsToFloat(v.y); }

// libstdc++ std::__push_heap (bits/stl_heap.h) for FoundPhoton::operator< (dist_square_ <)
template <typename Heap>
PM_HD void heapPush(const Heap &h, int hole, int top, uint2 value)
{
	int parent = (hole - 1) / 2;
	while(hole > top)
	{
		const uint2 pv = h.get(parent);
		if(!(heapKey(pv) < heapKey(value))) break;
		h.set(hole, pv);
		hole = parent;
		parent = (hole - 1) / 2;
	}
	h.set(hole, value);
}

// libstdc++ std::__adjust_heap
template <typename Heap>
PM_HD void heapAdjust(const Heap &h, int hole, int len, uint2 value)
{
	const int top = hole;
	int second = hole;
	while(second < (len - 1) / 2)
	{
		second = 2 * (second + 1);
		uint2 sv = h.get(second);
		const uint2 lv = h.get(second - 1);
		if(heapKey(sv) < heapKey(lv)) { --second; sv = lv; }
		h.set(hole, sv);
		hole = second;
	}
	if((len & 1) == 0 && second == (len - 2) / 2)
	{
		second = 2 * (second + 1);
		h.set(hole, h.get(second - 1));
		hole = second - 1;
	}
	heapPush(h, hole, top, value);
}

// libstdc++ std::__make_heap
template <typename Heap>
PM_HD void heapMake(const Heap &h, int len)
{
	if(len < 2) return;
	int parent = (len - 2) / 2;
	for(;;)
	{
		const uint2 value = h.get(parent);
		heapAdjust(h, parent, len, value);
		if(parent == 0) return;
		--parent;
	}
}

// PhotonGather::operator() (photon.cc:26-44)
template <typename Heap>
PM_HD void gatherProc(const Heap &h, uint32_t n_lookup, uint32_t &n_found, uint32_t photon, float dist_2, float &max_dist_squared)
{
	const uint2 entry = make_uint2(photon, pmFloatToBits(dist_2));
	if(n_found < n_lookup)
	{
		h.set(int(n_found++), entry);
		if(n_found == n_lookup)
		{
			heapMake(h, int(n_lookup));
			max_dist_squared = heapKey(h.get(0));
		}
	}
	else
	{
		const int n = int(n_lookup);
		if(n > 1)
		{
			// std::pop_heap: the top moves to [n - 1] (overwritten just below), the old last entry is sifted in from the root
			const uint2 value = h.get(n - 1);
			heapAdjust(h, 0, n - 1, value);
		}
		// found[n - 1] = {photon, dist_2}; std::push_heap
		heapPush(h, n - 1, 0, entry);
		max_dist_squared = heapKey(h.get(0));
	}
}

// MODE 0: gather with the heap in shared memory, 1: gather with the heap in `found`, 2: findNearest
// One query point, start to finish.  heap_s: this thread's shared-memory heap (MODE 0 only).
template <int MODE>
PM_HD void pmLookupOne(const uint4 *__restrict__ nodes, const float4 *__restrict__ dirs, const float *__restrict__ points, const float *__restrict__ normals,
                       uint32_t point, uint32_t k, float sq_radius, const float *__restrict__ sq_radii, uint2 *__restrict__ found, uint32_t *__restrict__ n_found_out,
                       float *__restrict__ sq_radius_out, uint32_t *__restrict__ nearest_out, const HeapSmem heap_s)
{
	const float px = points[3 * size_t(point)], py = points[3 * size_t(point) + 1], pz = points[3 * size_t(point) + 2];
	float nx = 0.f, ny = 0.f, nz = 0.f;
	if(MODE == 2)
	{
		nx = normals[3 * size_t(point)];
		ny = normals[3 * size_t(point) + 1];
		nz = normals[3 * size_t(point) + 2];
	}
	float max_dist_squared = sq_radii ? sq_radii[point] : sq_radius;
	uint32_t n_found = 0, nearest = kPmEmpty;
	const HeapGlobal heap_g{MODE == 2 ? nullptr : found + size_t(point) * k};

	uint32_t st_node[kPmStack]; // (far child << 2) | axis of the parent's split
	float st_split[kPmStack];
	int sp = 1;
	st_node[1] = kPmEmpty; // "nowhere", the reference's termination flag (pkdtree.h:230)
	uint32_t curr = 0;
	for(;;)
	{
		uint4 node = pmLoadNode(nodes + curr);
		while((node.w & 3u) != 3u)
		{
			const uint32_t axis = node.w & 3u;
			const float split = pmBitsToFloat(node.x);
			const float pa = axis == 0 ? px : (axis == 1 ? py : pz);
			const uint32_t right = node.w >> 2;
			uint32_t far_child;
			if(pa <= split) { far_child = right; curr = curr + 1; }
			else { far_child = curr + 1; curr = right; }
			++sp;
			st_node[sp] = (far_child << 2) | axis;
			st_split[sp] = split;
			node = pmLoadNode(nodes + curr);
		}
		{
			const float vx = pmSub(pmBitsToFloat(node.x), px), vy = pmSub(pmBitsToFloat(node.y), py), vz = pmSub(pmBitsToFloat(node.z), pz);
			const float dist_2 = pmAdd(pmAdd(pmMul(vx, vx), pmMul(vy, vy)), pmMul(vz, vz));
			if(dist_2 < max_dist_squared)
			{
				const uint32_t photon = node.w >> 2;
				if(MODE == 0) gatherProc(heap_s, k, n_found, photon, dist_2, max_dist_squared);
				else if(MODE == 1) gatherProc(heap_g, k, n_found, photon, dist_2, max_dist_squared);
				else
				{
					const float4 d = pmLoadDir(dirs + photon);
					const float dot = pmAdd(pmAdd(pmMul(d.x, nx), pmMul(d.y, ny)), pmMul(d.z, nz));
					if(dot > 0.f) { nearest = photon; max_dist_squared = dist_2; }
				}
			}
		}
		if(st_node[sp] == kPmEmpty) break;
		bool done = false;
		for(;;)
		{
			const uint32_t axis = st_node[sp] & 3u;
			const float pa = axis == 0 ? px : (axis == 1 ? py : pz);
			float d = pmSub(pa, st_split[sp]);
			d = pmMul(d, d);
			if(!(d > max_dist_squared)) break;
			--sp;
			if(st_node[sp] == kPmEmpty) { done = true; break; }
		}
		if(done) break;
		curr = st_node[sp] >> 2;
		--sp;
	}
	if(MODE == 2)
	{
		nearest_out[point] = nearest;
		return;
	}
	if(MODE == 0)
		for(uint32_t j = 0; j < n_found; ++j) heap_g.set(int(j), heap_s.get(int(j)));
	n_found_out[point] = n_found;
	if(sq_radius_out) sq_radius_out[point] = max_dist_squared;
}

template <int MODE>
__global__ void __launch_bounds__(kPmThreads) pmLookupKernel(const uint4 *__restrict__ nodes, const float4 *__restrict__ dirs, const float *__restrict__ points,
                                                             const float *__restrict__ normals, uint32_t n_points, uint32_t k, float sq_radius,
                                                             const float *__restrict__ sq_radii, uint2 *__restrict__ found, uint32_t *__restrict__ n_found_out,
                                                             float *__restrict__ sq_radius_out, uint32_t *__restrict__ nearest_out)
{
	extern __shared__ uint2 pm_heap_smem[];
	const uint32_t point = blockIdx.x * uint32_t(kPmThreads) + threadIdx.x;
	if(point >= n_points) return;
	pmLookupOne<MODE>(nodes, dirs, points, normals, point, k, sq_radius, sq_radii, found, n_found_out, sq_radius_out, nearest_out, HeapSmem{pm_heap_smem + threadIdx.x});
}

// ---- phased lookup (the default) --------------------------------------------------------------------------------------------
// The ncu source view of pmLookupKernel (profiles/r6a_pm_gather_smem_heap.txt) shows where the plain loop above loses the warp:
// the descent loop runs at 5.3 of 32 lanes (lanes that reached their leaf wait for the longest descent), the stack pops at 6.2,
// and the heap replacement -- 40 % of all warp instructions -- at 1.1 lanes, because each lane accepts a photon at a different
// moment.  The same lookup as a per-lane state machine:
//   pmStep     one node visit: (pop the stack if the last visit was a leaf) -> load the node -> interior: push the far child and
//              step down | leaf: test the photon; a cheap accept (append below k - 1, findNearest) happens on the spot, an accept
//              that needs heap work (the k-th photon: make_heap; later ones: pop_heap / push_heap) is only RECORDED (pending);
//   pmResolve  the pending heap work.
// The warp alternates a round of up to `round_steps` pmStep calls (a lane with pending work or without nodes left sits the round
// out) with one pmResolve in which every pending lane does its heap work at the same time.  Visit order, accept tests and heap
// calls of a lane are exactly those of the plain loop, so the results stay bit-identical; pmLookupPhasedOne (used by the host
// model) runs the same two functions for one lane.
struct PmLane
{
	float px, py, pz, nx, ny, nz;
	float max_dist_squared;
	uint32_t curr, n_found, nearest;
	uint32_t cand_photon;
	float cand_dist_2;
	int sp;
	bool alive, need_pop, pending;
};

// SINGLE_POP: a step pops ONE stack entry; a lane whose entry was beyond the radius pops the next one in its next step instead of
// looping here while the rest of the warp waits (the pop loop ran at 4.6 of 32 lanes and was 36 % of all warp instructions,
// profiles/r6b_pm_gather_phased.txt).
template <int MODE, bool SINGLE_POP, typename Heap>
PM_HD void pmStep(PmLane &s, uint2 *stack, const uint4 *__restrict__ nodes, const float4 *__restrict__ dirs, uint32_t k, const Heap &heap)
{
	if(s.need_pop)
	{
		// pkdtree.h:263-278: leave the leaf -- drop far children whose split plane is beyond the search radius
		for(;;)
		{
			const uint2 top = stack[s.sp];
			if(top.x == kPmEmpty) { s.alive = false; return; }
			const uint32_t axis = top.x & 3u;
			const float pa = axis == 0 ? s.px : (axis == 1 ? s.py : s.pz);
			float d = pmSub(pa, pmBitsToFloat(top.y));
			d = pmMul(d, d);
			--s.sp;
			if(!(d > s.max_dist_squared)) { s.curr = top.x >> 2; break; }
			if(SINGLE_POP) return;
		}
		s.need_pop = false;
	}
	const uint4 node = pmLoadNode(nodes + s.curr);
	if((node.w & 3u) != 3u)
	{
		// pkdtree.h:234-252
		const uint32_t axis = node.w & 3u;
		const float split = pmBitsToFloat(node.x);
		const float pa = axis == 0 ? s.px : (axis == 1 ? s.py : s.pz);
		const uint32_t right = node.w >> 2;
		uint32_t far_child;
		if(pa <= split) { far_child = right; s.curr = s.curr + 1; }
		else { far_child = s.curr + 1; s.curr = right; }
		++s.sp;
		stack[s.sp] = make_uint2((far_child << 2) | axis, node.x);
		return;
	}
	// pkdtree.h:254-261
	s.need_pop = true;
	const float vx = pmSub(pmBitsToFloat(node.x), s.px), vy = pmSub(pmBitsToFloat(node.y), s.py), vz = pmSub(pmBitsToFloat(node.z), s.pz);
	const float dist_2 = pmAdd(pmAdd(pmMul(vx, vx), pmMul(vy, vy)), pmMul(vz, vz));
	if(!(dist_2 < s.max_dist_squared)) return;
	const uint32_t photon = node.w >> 2;
	if(MODE == 2)
	{
		const float4 d = pmLoadDir(dirs + photon);
		const float dot = pmAdd(pmAdd(pmMul(d.x, s.nx), pmMul(d.y, s.ny)), pmMul(d.z, s.nz));
		if(dot > 0.f) { s.nearest = photon; s.max_dist_squared = dist_2; }
	}
	else if(s.n_found + 1u < k)
		heap.set(int(s.n_found++), make_uint2(photon, pmFloatToBits(dist_2))); // photon.cc:29-32, below the k-th photon
	else
	{
		s.pending = true;
		s.cand_photon = photon;
		s.cand_dist_2 = dist_2;
	}
}

template <typename Heap>
PM_HD void pmResolve(PmLane &s, uint32_t k, const Heap &heap)
{
	gatherProc(heap, k, s.n_found, s.cand_photon, s.cand_dist_2, s.max_dist_squared);
	s.pending = false;
}

PM_HD void pmLaneInit(PmLane &s, uint2 *stack, const float *__restrict__ points, const float *__restrict__ normals, bool with_normal, uint32_t point, float sq_radius,
                      const float *__restrict__ sq_radii)
{
	s.px = points[3 * size_t(point)];
	s.py = points[3 * size_t(point) + 1];
	s.pz = points[3 * size_t(point) + 2];
	s.nx = s.ny = s.nz = 0.f;
	if(with_normal)
	{
		s.nx = normals[3 * size_t(point)];
		s.ny = normals[3 * size_t(point) + 1];
		s.nz = normals[3 * size_t(point) + 2];
	}
	s.max_dist_squared = sq_radii ? sq_radii[point] : sq_radius;
	s.curr = 0;
	s.n_found = 0;
	s.nearest = kPmEmpty;
	s.cand_photon = 0;
	s.cand_dist_2 = 0.f;
	s.sp = 1;
	stack[1] = make_uint2(kPmEmpty, 0u); // "nowhere", the reference's termination flag (pkdtree.h:230)
	s.alive = true;
	s.need_pop = false;
	s.pending = false;
}

template <int MODE>
PM_HD void pmLaneFinish(const PmLane &s, uint32_t point, uint32_t k, uint2 *__restrict__ found, uint32_t *__restrict__ n_found_out, float *__restrict__ sq_radius_out,
                        uint32_t *__restrict__ nearest_out, const HeapSmem heap_s)
{
	if(MODE == 2)
	{
		nearest_out[point] = s.nearest;
		return;
	}
	if(MODE == 0)
	{
		const HeapGlobal heap_g{found + size_t(point) * k};
		for(uint32_t j = 0; j < s.n_found; ++j) heap_g.set(int(j), heap_s.get(int(j)));
	}
	n_found_out[point] = s.n_found;
	if(sq_radius_out) sq_radius_out[point] = s.max_dist_squared;
}

// one lane, start to finish, through the same state machine (host model; also what a warp of one lane would do)
template <int MODE, bool SINGLE_POP>
PM_HD void pmLookupPhasedOne(const uint4 *__restrict__ nodes, const float4 *__restrict__ dirs, const float *__restrict__ points, const float *__restrict__ normals,
                             uint32_t point, uint32_t k, float sq_radius, const float *__restrict__ sq_radii, uint2 *__restrict__ found,
                             uint32_t *__restrict__ n_found_out, float *__restrict__ sq_radius_out, uint32_t *__restrict__ nearest_out, const HeapSmem heap_s, int round_steps)
{
	PmLane s;
	uint2 stack[kPmStack];
	pmLaneInit(s, stack, points, normals, MODE == 2, point, sq_radius, sq_radii);
	const HeapGlobal heap_g{MODE == 2 ? nullptr : found + size_t(point) * k};
	while(s.alive)
	{
		for(int step = 0; step < round_steps && s.alive && !s.pending; ++step)
		{
			if(MODE == 0) pmStep<MODE, SINGLE_POP>(s, stack, nodes, dirs, k, heap_s);
			else pmStep<MODE, SINGLE_POP>(s, stack, nodes, dirs, k, heap_g);
		}
		if(s.pending)
		{
			if(MODE == 0) pmResolve(s, k, heap_s);
			else pmResolve(s, k, heap_g);
		}
	}
	pmLaneFinish<MODE>(s, point, k, found, n_found_out, sq_radius_out, nearest_out, heap_s);
}

#ifdef __CUDACC__
// patience: a lane whose pending work is the expensive one -- the k-th photon, i.e. std::make_heap over k entries, ~3000
// instructions for k = 100, which ran at 1.2 lanes when every lane did it on its own -- waits until `patience` lanes of the warp
// have the same work pending, or no lane of the warp can take another step; replacements (pop_heap / push_heap) are resolved at
// the end of every round.  1 = no waiting.
template <int MODE, bool SINGLE_POP>
__global__ void __launch_bounds__(kPmThreads) pmLookupPhasedKernel(const uint4 *__restrict__ nodes, const float4 *__restrict__ dirs, const float *__restrict__ points,
                                                                   const float *__restrict__ normals, uint32_t n_points, uint32_t k, float sq_radius,
                                                                   const float *__restrict__ sq_radii, uint2 *__restrict__ found, uint32_t *__restrict__ n_found_out,
                                                                   float *__restrict__ sq_radius_out, uint32_t *__restrict__ nearest_out, int round_steps, int patience)
{
	extern __shared__ uint2 pm_heap_smem[];
	const uint32_t point = blockIdx.x * uint32_t(kPmThreads) + threadIdx.x;
	const bool in_range = point < n_points;
	PmLane s;
	uint2 stack[kPmStack];
	pmLaneInit(s, stack, points, normals, MODE == 2, in_range ? point : 0u, sq_radius, sq_radii);
	s.alive = in_range;
	const HeapSmem heap_s{pm_heap_smem + threadIdx.x};
	const HeapGlobal heap_g{MODE == 2 ? nullptr : found + size_t(in_range ? point : 0u) * k};
	// all 32 lanes stay in the loop until the whole warp is done, so that the votes and __syncwarp below are always complete
	while(__any_sync(0xFFFFFFFFu, s.alive))
	{
#pragma unroll 1
		for(int step = 0; step < round_steps; ++step)
		{
			if(s.alive && !s.pending)
			{
				if(MODE == 0) pmStep<MODE, SINGLE_POP>(s, stack, nodes, dirs, k, heap_s);
				else pmStep<MODE, SINGLE_POP>(s, stack, nodes, dirs, k, heap_g);
			}
			// no lane left that could use another step of this round
			if(!__any_sync(0xFFFFFFFFu, s.alive && !s.pending)) break;
		}
		if(MODE != 2)
		{
			const bool heavy = s.pending && s.n_found + 1u == k && k > 2u;
			const int n_heavy = __popc(__ballot_sync(0xFFFFFFFFu, heavy));
			const bool someone_can_step = __any_sync(0xFFFFFFFFu, s.alive && !s.pending);
			const bool heavy_now = n_heavy >= patience || !someone_can_step;
			if(s.pending && (!heavy || heavy_now))
			{
				if(MODE == 0) pmResolve(s, k, heap_s);
				else pmResolve(s, k, heap_g);
			}
			__syncwarp();
		}
	}
	if(in_range) pmLaneFinish<MODE>(s, point, k, found, n_found_out, sq_radius_out, nearest_out, heap_s);
}
#endif

} // namespace b200pm
