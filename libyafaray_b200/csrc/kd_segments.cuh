// libyafaray_b200/csrc/kd_segments.cuh -- one launch for the flushes of SEVERAL render threads.
//
// The renderer's wavefront ray queues (integration/src/render/wavefront_b200.cc) flush a few hundred rays per render thread at a
// time.  The flush combiner of b200rt.cu collects the flushes that arrive within a short window from all threads and traces them
// with ONE launch of this kernel: a table of segments (rays / results / times of one job, in the pinned host memory the job
// named, read and written in place across PCIe) travels as a kernel parameter, warp w of the grid finds the segment its index
// falls into and runs the traversal of kd_kernels.cuh on 32 rays of it -- cursor-less, like traceMixedKernel.
#ifndef B200RT_KD_SEGMENTS_CUH
#define B200RT_KD_SEGMENTS_CUH

#include "kd_kernels.cuh"

namespace b200rt {

struct Segment
{
	const b200rt_ray *rays;
	void *out;
	const float *times; // nullptr = 0
	uint32_t n;
	uint32_t first_warp; // index of the segment's first warp in the grid
	int32_t query;       // Query
	int32_t max_depth;   // transparent shadows
};

static constexpr int kMaxSegments = 72; // 72 x 40 bytes + header: below the 4 KB of kernel parameters
struct SegmentTable
{
	Segment seg[kMaxSegments];
	uint32_t n_segments;
	uint32_t n_warps;
};

template <bool SPHERES>
__global__ void __launch_bounds__(kBlock, 4) traceSegmentsKernel(const __grid_constant__ SceneView s, const __grid_constant__ SegmentTable table, bool tree_space)
{
	__shared__ __align__(kShortStack * kBlock * 8) uint2 sh_stack[kShortStack][kBlock];
	__shared__ float2 sh_axis[4][kBlock];
	__shared__ uint32_t sh_task[kBlock / 32][32];
#if B200RT_LEAF_PREFETCH == 2
	__shared__ float4 sh_leaf[3][kBlock];
#else
	float4 (*sh_leaf)[kBlock] = nullptr;
#endif
	const uint32_t warp = blockIdx.x * uint32_t(kBlock / 32) + (threadIdx.x >> 5);
	if(warp >= table.n_warps) return;
	// the last segment that starts at or before this warp (warp-uniform; the table sits in the constant bank)
	uint32_t lo = 0u, hi = table.n_segments;
	while(hi - lo > 1u)
	{
		const uint32_t mid = (lo + hi) >> 1;
		if(table.seg[mid].first_warp <= warp) lo = mid;
		else hi = mid;
	}
	const Segment &g = table.seg[lo];
	const uint32_t base = (warp - g.first_warp) * 32u;
	if(g.query == kClosest) traceWarps<kClosest, SPHERES>(s, g.rays, g.n, static_cast<b200rt_hit *>(g.out), nullptr, 0, tree_space, sh_stack, sh_axis, base, sh_leaf, g.times, sh_task);
	else if(g.query == kShadow) traceWarps<kShadow, SPHERES>(s, g.rays, g.n, static_cast<uint32_t *>(g.out), nullptr, 0, tree_space, sh_stack, sh_axis, base, sh_leaf, g.times, sh_task);
	else traceWarps<kTShadow, SPHERES>(s, g.rays, g.n, static_cast<b200rt_tshadow *>(g.out), nullptr, g.max_depth, tree_space, sh_stack, sh_axis, base, sh_leaf, g.times, sh_task);
}

} // namespace b200rt
#endif
