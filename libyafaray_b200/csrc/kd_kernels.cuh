// libyafaray_b200/csrc/kd_kernels.cuh -- sm_100a traversal kernels of libb200rt.
//
// What is reproduced from the reference, operation for operation (paths relative to the reference tree):
//   * root slab test            include/geometry/bound.h:156-198          (enter/leave feed the ray bias)
//   * ray bias t_min            include/accelerator/accelerator.h:64, accelerator_kdtree_common.h:139
//   * wrapper t_max / origin    include/accelerator/accelerator.h:89-120
//   * polygon test              include/geometry/shape/shape_polygon.h:126-176 (Moeller-Trumbore, quads)
//   * accept rules              include/accelerator/accelerator.h:122-169
// every float operation of those is issued through __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn so that nvcc
// can neither contract it into an FMA nor reassociate it (the reference build has no FMA).
//
// What is NOT taken from the reference: the tree (kd_build.cc), its memory layout and the traversal
// order (t-interval based, below).  Those only decide WHICH leaves are opened, never the value of an
// accepted hit; exact-t ties between primitives may resolve differently (documented).
#ifndef B200RT_KD_KERNELS_CUH
#define B200RT_KD_KERNELS_CUH

#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/b200rt.h"

namespace b200rt {

// Flattened scene in HBM (DESIGN.md "Data layout").
//   nodes: uint2 per node, depth-first, left child = i + 1
//          interior: x = float bits of split, y = (right child << 2) | axis
//          leaf:     x = first float4 of its records in `tris`, y = (count << 2) | 3
//   tris:  per leaf reference 3 float4 (triangle) or 4 float4 (quad):
//          q0 = v0.xyz | face id      q1 = e1.xyz | flags (bit3 = quad)      q2 = e2.xyz | 0     [q3 = e3.xyz | 0]
//          with e_k = v_k - v0 computed in float on the host exactly as shape_polygon.h:130-131,150 does.
struct SceneView
{
	const uint2 *nodes;
	const float4 *tris;
	float bound[6];
};

enum Query { kClosest = 0, kShadow = 1, kTShadow = 2 };

static constexpr int kStackSize = 64;
static constexpr uint32_t kFlagQuad = 8u;

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz)
{
	// vector.h:163-164: a0*b0 + a1*b1 + a2*b2, left to right
	return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

#define B200RT_CROSS(ox, oy, oz, ax, ay, az, bx, by, bz)                 \
	const float ox = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by)); \
	const float oy = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz)); \
	const float oz = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));

// bound.h:156-198.  Returns false on a miss.
__device__ __forceinline__ bool boundCross(const float *b, float ox, float oy, float oz, float dx, float dy, float dz, float t_max, float &enter, float &leave)
{
	float lmin = -FLT_MAX, lmax = FLT_MAX;
	const float o[3] = {ox, oy, oz}, d[3] = {dx, dy, dz};
#pragma unroll
	for(int axis = 0; axis < 3; ++axis)
	{
		if(d[axis] != 0.f)
		{
			const float p = __fsub_rn(o[axis], b[axis]);
			const float inv_dir = __fdiv_rn(1.f, d[axis]);
			const float near_t = __fmul_rn(-p, inv_dir);
			const float far_t = __fmul_rn(__fsub_rn(__fsub_rn(b[3 + axis], b[axis]), p), inv_dir);
			const float ltmin = (inv_dir > 0.f) ? near_t : far_t;
			const float ltmax = (inv_dir > 0.f) ? far_t : near_t;
			if(axis == 0) { lmin = ltmin; lmax = ltmax; }
			else
			{
				lmin = (ltmin < lmin) ? lmin : ltmin; // std::max(ltmin, lmin)
				lmax = (lmax < ltmax) ? lmax : ltmax; // std::min(ltmax, lmax)
			}
			if((lmax < 0.f) || (lmin > t_max)) return false;
		}
	}
	if((lmin <= lmax) && (lmax >= 0.f) && (lmin <= t_max))
	{
		enter = lmin;
		leave = lmax;
		return true;
	}
	return false;
}

// shape_polygon.h:126-176 on a record of the leaf stream.  Returns t (0 = miss) and uv.
__device__ __forceinline__ float polyIntersect(const float4 q0, const float4 q1, const float4 q2, const float4 *q3_ptr, bool quad,
                                               float ox, float oy, float oz, float dx, float dy, float dz, float &out_u, float &out_v)
{
	B200RT_CROSS(px, py, pz, dx, dy, dz, q2.x, q2.y, q2.z)            // pvec_2 = dir ^ edge_2
	const float det = dot3(q1.x, q1.y, q1.z, px, py, pz);             // edge_1 * pvec_2
	if(det != 0.f)
	{
		const float inv_det = __fdiv_rn(1.f, det);
		const float tx = __fsub_rn(ox, q0.x), ty = __fsub_rn(oy, q0.y), tz = __fsub_rn(oz, q0.z); // tvec = from - v0
		float u = __fmul_rn(dot3(tx, ty, tz, px, py, pz), inv_det);
		if(u >= 0.f && u <= 1.f)
		{
			B200RT_CROSS(qx, qy, qz, tx, ty, tz, q1.x, q1.y, q1.z)     // qvec_1 = tvec ^ edge_1
			const float v = __fmul_rn(dot3(dx, dy, dz, qx, qy, qz), inv_det);
			if(v >= 0.f && __fadd_rn(u, v) <= 1.f)
			{
				const float t = __fmul_rn(dot3(q2.x, q2.y, q2.z, qx, qy, qz), inv_det);
				if(t > 0.f)
				{
					out_u = quad ? __fadd_rn(u, v) : u;
					out_v = v;
					return t;
				}
			}
		}
		else if(quad)
		{
			const float4 q3 = __ldg(q3_ptr);
			B200RT_CROSS(p3x, p3y, p3z, dx, dy, dz, q3.x, q3.y, q3.z)  // pvec_3 = dir ^ edge_3
			const float det2 = dot3(q2.x, q2.y, q2.z, p3x, p3y, p3z);  // edge_2 * pvec_3
			if(det2 != 0.f)
			{
				const float inv_det2 = __fdiv_rn(1.f, det2);
				u = __fmul_rn(dot3(tx, ty, tz, p3x, p3y, p3z), inv_det2);
				if(u >= 0.f && u <= 1.f)
				{
					B200RT_CROSS(q2x, q2y, q2z, tx, ty, tz, q2.x, q2.y, q2.z) // qvec_2 = tvec ^ edge_2
					const float v = __fmul_rn(dot3(dx, dy, dz, q2x, q2y, q2z), inv_det2);
					if(v >= 0.f && __fadd_rn(u, v) <= 1.f)
					{
						const float t = __fmul_rn(dot3(q3.x, q3.y, q3.z, q2x, q2y, q2z), inv_det2);
						if(t > 0.f)
						{
							out_u = u;
							out_v = __fadd_rn(u, v);
							return t;
						}
					}
				}
			}
		}
	}
	out_u = 0.f;
	out_v = 0.f;
	return 0.f;
}

struct TShadowState
{
	int depth, max_depth;
	b200rt_hit list[B200RT_TSHADOW_MAX];
};

// One ray through the tree.  (ox..dz) is the ray the TREE sees (shadow wrappers have already moved the
// origin); ray_tmin is Ray::tmin_.  Closest: returns hit in best_*.  Shadow/TShadow: returns true when
// "shadowed" with the occluder in best_prim.
template <int QUERY>
__device__ __forceinline__ bool traverse(const SceneView &s, float ox, float oy, float oz, float dx, float dy, float dz,
                                         float ray_tmin, const float t_max,
                                         float &best_t, float &best_u, float &best_v, uint32_t &best_prim, TShadowState *ts)
{
	best_t = 0.f;
	best_u = 0.f;
	best_v = 0.f;
	best_prim = B200RT_MISS;
	float enter, leave;
	if(!boundCross(s.bound, ox, oy, oz, dx, dy, dz, t_max, enter, leave)) return false;
	// math::inverse (math.h:71-77): FLT_MAX for a zero component, so that no NaN can appear below
	const float ix = (dx == 0.f) ? FLT_MAX : __fdiv_rn(1.f, dx);
	const float iy = (dy == 0.f) ? FLT_MAX : __fdiv_rn(1.f, dy);
	const float iz = (dz == 0.f) ? FLT_MAX : __fdiv_rn(1.f, dz);
	const float bias = __fmul_rn(__fmul_rn(0.1f, 0.00005f), fabsf(__fsub_rn(leave, enter)));
	const float t_min = (QUERY == kShadow) ? bias : ((ray_tmin < bias) ? bias : ray_tmin);
	float cur_t_max = t_max; // closest: shrinks with every accepted hit

	uint32_t st_node[kStackSize];
	float st_far[kStackSize];
	int sp = 0;
	uint32_t node = 0;
	float seg_lo = enter, seg_hi = fminf(leave, t_max);
	for(;;)
	{
		uint2 nd = __ldg(&s.nodes[node]);
		while((nd.y & 3u) != 3u)
		{
			const uint32_t axis = nd.y & 3u;
			const float split = __uint_as_float(nd.x);
			const float o = (axis == 0u) ? ox : ((axis == 1u) ? oy : oz);
			const float d = (axis == 0u) ? dx : ((axis == 1u) ? dy : dz);
			const float inv = (axis == 0u) ? ix : ((axis == 1u) ? iy : iz);
			const float t_plane = (split - o) * inv;
			const bool left_first = (o < split) || (o == split && d <= 0.f);
			const uint32_t left = node + 1u, right = nd.y >> 2;
			const uint32_t first = left_first ? left : right, second = left_first ? right : left;
			if(t_plane > seg_hi || t_plane <= 0.f) node = first;
			else if(t_plane < seg_lo) node = second;
			else
			{
				st_node[sp] = second;
				st_far[sp] = seg_hi;
				++sp;
				node = first;
				seg_hi = t_plane;
			}
			nd = __ldg(&s.nodes[node]);
		}
		// leaf
		uint32_t count = nd.y >> 2;
		const float4 *rec = s.tris + nd.x;
		for(; count != 0u; --count)
		{
			const float4 q0 = __ldg(rec), q1 = __ldg(rec + 1), q2 = __ldg(rec + 2);
			const uint32_t flags = __float_as_uint(q1.w);
			const bool quad = (flags & kFlagQuad) != 0u;
			float u, v;
			const float t = polyIntersect(q0, q1, q2, rec + 3, quad, ox, oy, oz, dx, dy, dz, u, v);
			rec += quad ? 4 : 3;
			// accelerator.h:125 / :137 / :150
			if(t <= 0.f || t < t_min || t >= cur_t_max) continue;
			if(QUERY == kClosest)
			{
				if(!(flags & B200RT_FACE_VISIBLE)) continue;
				best_t = t; best_u = u; best_v = v; best_prim = __float_as_uint(q0.w);
				cur_t_max = t;
			}
			else
			{
				if(!(flags & B200RT_FACE_CASTS_SHADOWS)) continue;
				best_t = t; best_u = u; best_v = v; best_prim = __float_as_uint(q0.w);
				if(QUERY == kShadow) return true;
				if(!(flags & B200RT_FACE_TRANSPARENT)) return true; // opaque caster
				bool seen = false;
				for(int k = 0; k < ts->depth; ++k) seen = seen || (ts->list[k].prim == best_prim);
				if(!seen)
				{
					if(ts->depth >= ts->max_depth) return true;
					ts->list[ts->depth].t = t; ts->list[ts->depth].u = u; ts->list[ts->depth].v = v; ts->list[ts->depth].prim = best_prim;
					++ts->depth;
				}
			}
		}
		if(QUERY == kClosest && best_prim != B200RT_MISS && best_t <= seg_hi) return true; // accelerator_kdtree_common.h:232
		if(sp == 0) break;
		--sp;
		node = st_node[sp];
		seg_lo = seg_hi;
		seg_hi = st_far[sp];
		if(QUERY == kClosest && best_prim != B200RT_MISS && best_t <= seg_lo) return true;
	}
	return QUERY == kClosest ? (best_prim != B200RT_MISS) : false;
}

// ---- kernels: one ray per thread ----------------------------------------------------------------
static constexpr int kBlock = 128;

__global__ void __launch_bounds__(kBlock) traceClosestKernel(SceneView s, const b200rt_ray *__restrict__ rays, size_t n, b200rt_hit *__restrict__ out)
{
	const size_t i = size_t(blockIdx.x) * kBlock + threadIdx.x;
	if(i >= n) return;
	const float4 a = __ldg(reinterpret_cast<const float4 *>(rays) + 2 * i);
	const float4 b = __ldg(reinterpret_cast<const float4 *>(rays) + 2 * i + 1);
	const float t_max = (b.w >= 0.f) ? b.w : FLT_MAX; // accelerator.h:91
	float t, u, v;
	uint32_t prim;
	traverse<kClosest>(s, a.x, a.y, a.z, b.x, b.y, b.z, a.w, t_max, t, u, v, prim, nullptr);
	float4 r;
	r.x = t; r.y = u; r.z = v; r.w = __uint_as_float(prim);
	reinterpret_cast<float4 *>(out)[i] = r;
}

// accelerator.h:103-111: origin moved by dir * tmin, t_max = tmax - 2 tmin (unbounded if tmax < 0)
__device__ __forceinline__ void shadowRay(const float4 a, const float4 b, float &ox, float &oy, float &oz, float &t_max)
{
	ox = __fadd_rn(a.x, __fmul_rn(b.x, a.w));
	oy = __fadd_rn(a.y, __fmul_rn(b.y, a.w));
	oz = __fadd_rn(a.z, __fmul_rn(b.z, a.w));
	t_max = (b.w >= 0.f) ? __fsub_rn(b.w, __fmul_rn(2.f, a.w)) : FLT_MAX;
}

__global__ void __launch_bounds__(kBlock) traceShadowKernel(SceneView s, const b200rt_ray *__restrict__ rays, size_t n, uint32_t *__restrict__ out)
{
	const size_t i = size_t(blockIdx.x) * kBlock + threadIdx.x;
	if(i >= n) return;
	const float4 a = __ldg(reinterpret_cast<const float4 *>(rays) + 2 * i);
	const float4 b = __ldg(reinterpret_cast<const float4 *>(rays) + 2 * i + 1);
	float ox, oy, oz, t_max, t, u, v;
	shadowRay(a, b, ox, oy, oz, t_max);
	uint32_t prim;
	const bool shadowed = traverse<kShadow>(s, ox, oy, oz, b.x, b.y, b.z, a.w, t_max, t, u, v, prim, nullptr);
	out[i] = shadowed ? prim : B200RT_MISS;
}

__global__ void __launch_bounds__(kBlock) traceTShadowKernel(SceneView s, const b200rt_ray *__restrict__ rays, size_t n, int max_depth, b200rt_tshadow *__restrict__ out)
{
	const size_t i = size_t(blockIdx.x) * kBlock + threadIdx.x;
	if(i >= n) return;
	const float4 a = __ldg(reinterpret_cast<const float4 *>(rays) + 2 * i);
	const float4 b = __ldg(reinterpret_cast<const float4 *>(rays) + 2 * i + 1);
	float ox, oy, oz, t_max, t, u, v;
	shadowRay(a, b, ox, oy, oz, t_max);
	TShadowState ts;
	ts.depth = 0;
	ts.max_depth = max_depth;
	uint32_t prim;
	const bool shadowed = traverse<kTShadow>(s, ox, oy, oz, b.x, b.y, b.z, a.w, t_max, t, u, v, prim, &ts);
	uint4 *o = reinterpret_cast<uint4 *>(out + i);
	o[0] = make_uint4(shadowed ? 1u : 0u, uint32_t(ts.depth), shadowed ? prim : B200RT_MISS, 0u); // setNoHit() clears primitive_
#pragma unroll
	for(int k = 0; k < B200RT_TSHADOW_MAX; ++k)
	{
		uint4 e = make_uint4(0u, 0u, 0u, B200RT_MISS);
		if(k < ts.depth) e = make_uint4(__float_as_uint(ts.list[k].t), __float_as_uint(ts.list[k].u), __float_as_uint(ts.list[k].v), ts.list[k].prim);
		o[1 + k] = e;
	}
}

} // namespace b200rt
#endif
