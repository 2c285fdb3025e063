// libyafaray_b200/csrc/pm_build.cc -- the reference's balanced point kd-tree, built with every host thread.
//
// What fixes the tree (include/photon/pkdtree.h:144-218): a node over photons [start, end) splits on the largest axis of its
// CLIPPED bound (the map's bound cut by the split planes above it, Bound::largestAxis include/geometry/bound.h:100-104), at
// the photon of rank (start + end) / 2 under the order (coordinate, then address) -- std::nth_element with CompareNode,
// pkdtree.h:71-79 -- which itself goes to the right child; one photon per leaf; nodes in preorder.  Under a strict total order
// the two sides of an nth_element are fixed SETS, so any selection algorithm gives the reference's tree.
//
// Own design: a subtree over m photons always has 2 m - 1 nodes, so every subtree's place in the node array is known before it
// is built -- the reference builds subtrees in scratch arrays and splices them (pkdtree.h:168-206); here every task writes its
// nodes in place and the top levels fan out over a pool of threads.
#include "pm_build.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

namespace b200pm {
namespace {

struct Builder
{
	const float *pos;
	uint32_t *prims;
	uint32_t *a, *b;
	std::atomic<uint32_t> depth{0};
	size_t spawn_above; // subtrees larger than this hand their left half to another thread while one is free
	std::atomic<int> free_threads{0};

	void noteDepth(uint32_t d)
	{
		uint32_t seen = depth.load(std::memory_order_relaxed);
		while(d > seen && !depth.compare_exchange_weak(seen, d, std::memory_order_relaxed)) {}
	}

	void build(size_t start, size_t end, size_t node, float lo_x, float lo_y, float lo_z, float hi_x, float hi_y, float hi_z, uint32_t level)
	{
		for(;;)
		{
			if(end - start == 1)
			{
				a[node] = prims[start];
				b[node] = 3u;
				noteDepth(level);
				return;
			}
			const float dx = hi_x - lo_x, dy = hi_y - lo_y, dz = hi_z - lo_z;
			const int axis = (dx > dy) ? ((dx > dz) ? 0 : 2) : ((dy > dz) ? 1 : 2);
			const size_t split_el = (start + end) / 2;
			const float *p = pos;
			std::nth_element(prims + start, prims + split_el, prims + end, [p, axis](uint32_t i, uint32_t j) {
				const float ci = p[3 * size_t(i) + axis], cj = p[3 * size_t(j) + axis];
				return ci == cj ? (i < j) : (ci < cj);
			});
			const float split = pos[3 * size_t(prims[split_el]) + axis];
			const size_t right = node + 2 * (split_el - start); // 1 + (2 m_left - 1)
			std::memcpy(&a[node], &split, 4);
			b[node] = (uint32_t(right) << 2) | uint32_t(axis);
			float l_hi_x = hi_x, l_hi_y = hi_y, l_hi_z = hi_z;
			(axis == 0 ? l_hi_x : axis == 1 ? l_hi_y : l_hi_z) = split;
			std::thread helper;
			if(end - start > spawn_above && free_threads.fetch_sub(1) > 0)
				helper = std::thread([=] { build(start, split_el, node + 1, lo_x, lo_y, lo_z, l_hi_x, l_hi_y, l_hi_z, level + 1); free_threads.fetch_add(1); });
			else
			{
				if(end - start > spawn_above) free_threads.fetch_add(1);
				build(start, split_el, node + 1, lo_x, lo_y, lo_z, l_hi_x, l_hi_y, l_hi_z, level + 1);
			}
			// right half: iterate
			if(helper.joinable())
			{
				// this thread goes on with the right half; join when it is done
				float r_lo_x = lo_x, r_lo_y = lo_y, r_lo_z = lo_z;
				(axis == 0 ? r_lo_x : axis == 1 ? r_lo_y : r_lo_z) = split;
				build(split_el, end, right, r_lo_x, r_lo_y, r_lo_z, hi_x, hi_y, hi_z, level + 1);
				helper.join();
				return;
			}
			(axis == 0 ? lo_x : axis == 1 ? lo_y : lo_z) = split;
			start = split_el;
			node = right;
			++level;
		}
	}
};

} // namespace

void buildTree(const float *pos, size_t n, int threads, HostTree &out)
{
	out.a.assign(2 * n - 1, 0u);
	out.b.assign(2 * n - 1, 0u);
	std::vector<uint32_t> prims(n);
	float lo[3] = {pos[0], pos[1], pos[2]}, hi[3] = {pos[0], pos[1], pos[2]};
	for(size_t i = 0; i < n; ++i)
	{
		prims[i] = uint32_t(i);
		for(int c = 0; c < 3; ++c)
		{
			const float v = pos[3 * i + c];
			if(v < lo[c]) lo[c] = v;
			if(v > hi[c]) hi[c] = v;
		}
	}
	if(threads <= 0) threads = int(std::max(1u, std::thread::hardware_concurrency()));
	Builder builder;
	builder.pos = pos;
	builder.prims = prims.data();
	builder.a = out.a.data();
	builder.b = out.b.data();
	builder.spawn_above = std::max<size_t>(8192, n / (size_t(threads) * 8));
	builder.free_threads.store(threads - 1);
	builder.build(0, n, 0, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], 1);
	out.depth = builder.depth.load();
}

} // namespace b200pm
