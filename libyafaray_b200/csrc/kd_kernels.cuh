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
	const float4 *inst; // moving instances: per instance 10 float4 = three obj_to_world matrices (rows 0-2 of each) + (time_start, time_end, -, -)
	// two-level treelets (TREELET kernel variants, polygon-only scenes): 2 uint4 per treelet, see kEmptyRef below
	const uint4 *treelets;
	uint2 *spill;           // per-thread overflow of the short stack: entry e of global thread g at spill[e * spill_threads + g]
	uint32_t spill_threads;
};

enum Query { kClosest = 0, kShadow = 1, kTShadow = 2 };

static constexpr uint32_t kFlagQuad = 8u;
static constexpr uint32_t kFlagSphere = 16u;
// Motion blur (scenes built with b200rt_add_mesh_bezier / _moving; kernels of the SPHERES = "other primitive kinds" variant):
//   Bezier face, nv vertices: 3 nv float4 = the vertices at time steps 0, 1, 2; w of the first four: face id, flags, time_start,
//     - (quads); w of the first vertex of step 1: time_end.  (Static records hold v0 and EDGES; these hold vertices.)
//   face of a moving instance: nv float4 = the base vertices; w: face id, flags, float4 index of the instance in SceneView::inst, -
static constexpr uint32_t kFlagBezier = 32u;
static constexpr uint32_t kFlagMoving = 64u;
// Leaf nodes: y = (count << 2) | 3 with bit 31 (bit 29 of the count field) set when every record of the leaf is a static polygon
// stored in FOUR float4 (a leaf with at least one quad pads its triangles), so that record k starts at first + 4 k; without the
// bit the records of a polygon-only scene are triangles of three float4.  (Scenes with spheres / motion-blur faces are walked
// record by record.)
// Treelets (B200RT_TREELET): an interior node together with its two children in one 32-byte sector, so that one dependent load
// advances the ray TWO levels.  uint4 A = (split of the node, split of its left child, split of its right child, meta),
// uint4 B = references to the four grandchildren (left-left, left-right, right-left, right-right).  meta: byte 0 / 1 / 2 = axis << 2
// of the node / left child / right child; axis 3 = "this child is a leaf": its split is FLT_MAX (with row 3 of the axis table =
// (0, 1) the plane then lies beyond every interval: near side only), its reference sits in the child's FIRST slot and the second
// slot is kEmptyRef.  A reference is a treelet index, or kLeafRef | first float4 of a non-empty leaf's records (the count rides in
// q2.w of the leaf's first record), or kEmptyRef for an empty leaf -- never visited, never pushed.
static constexpr uint32_t kEmptyRef = 0xFFFFFFFFu;
static constexpr uint32_t kLeafRef = 0x80000000u;
static constexpr uint32_t kLeafStride4 = 1u << 29;
static constexpr uint32_t kLeafCountMask = kLeafStride4 - 1u;

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
// `edge3` returns the quad's third edge (only called when the second triangle of a quad has to be tested).
template <typename Edge3>
__device__ __forceinline__ float polyIntersectWith(const float4 q0, const float4 q1, const float4 q2, Edge3 edge3, bool quad,
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
			const float4 q3 = edge3();
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

__device__ __forceinline__ float polyIntersect(const float4 q0, const float4 q1, const float4 q2, const float4 *q3_ptr, bool quad,
                                               float ox, float oy, float oz, float dx, float dy, float dz, float &out_u, float &out_v)
{
	return polyIntersectWith(q0, q1, q2, [&]() { return __ldg(q3_ptr); }, quad, ox, oy, oz, dx, dy, dz, out_u, out_v);
}

// Quadratic Bezier factors of a ray time inside (start, end): math::lerpSegment(time, 0, start, 1, end) then
// math::bezierCalculateFactors (include/math/interpolation.h:50-93).  Returns 0 / 2 when the time is at or outside the start /
// end of the range -- time step 0 / 2 is then used as it is (primitive_polygon.h:245-248, instance.h:78-80) -- else 1.
__device__ __forceinline__ int bezierFactors(float time, float start, float end, float &f0, float &f1, float &f2)
{
	if(time <= start) return 0;
	if(time >= end) return 2;
	const float x = __fadd_rn(0.f, __fmul_rn(__fdiv_rn(__fsub_rn(time, start), __fsub_rn(end, start)), __fsub_rn(1.f, 0.f)));
	const float xr = __fsub_rn(1.f, x);
	f0 = __fmul_rn(xr, xr);
	f1 = __fmul_rn(__fmul_rn(2.f, x), xr);
	f2 = __fmul_rn(x, x);
	return 1;
}

// math::bezierInterpolate (interpolation.h:83-86): y0 * f0 + y1 * f1 + y2 * f2, left to right, every product and sum rounded
__device__ __forceinline__ float bezier3(int where, float y0, float y1, float y2, float f0, float f1, float f2)
{
	return where == 0 ? y0 : where == 2 ? y2 : __fadd_rn(__fadd_rn(__fmul_rn(y0, f0), __fmul_rn(y1, f1)), __fmul_rn(y2, f2));
}

// A motion-blur face at the ray's time (records: kFlagBezier / kFlagMoving above).  The vertices are rebuilt with the reference's
// operations (FacePrimitive::getVertex with Bezier factors, primitive_face.h:86-90; SquareMatrix * Point, matrix.h:131-144, on the
// interpolated instance matrix, instance.h:82-85), the edges are the test's own subtractions (shape_polygon.h:130-131,150).
// `rec` is advanced past the record.
__device__ __forceinline__ float motionIntersect(const SceneView &s, const float4 *&rec, const float4 q0, const float4 q1, const float4 q2, uint32_t flags, bool quad, float time,
                                                 float ox, float oy, float oz, float dx, float dy, float dz, float &u, float &v)
{
	const int nv = quad ? 4 : 3;
	float4 p[4];
	p[0] = q0; p[1] = q1; p[2] = q2;
	if(quad) p[3] = __ldg(rec + 3);
	float f0 = 0.f, f1 = 0.f, f2 = 0.f;
	if(flags & kFlagBezier)
	{
		const float4 first1 = __ldg(rec + nv);
		const int where = bezierFactors(time, q2.w, first1.w, f0, f1, f2);
		for(int k = 0; k < nv; ++k)
		{
			const float4 a = (k == 0) ? first1 : __ldg(rec + nv + k), b = __ldg(rec + 2 * nv + k);
			p[k].x = bezier3(where, p[k].x, a.x, b.x, f0, f1, f2);
			p[k].y = bezier3(where, p[k].y, a.y, b.y, f0, f1, f2);
			p[k].z = bezier3(where, p[k].z, a.z, b.z, f0, f1, f2);
		}
		rec += 3 * nv;
	}
	else
	{
		const float4 *m = s.inst + __float_as_uint(q2.w);
		const float4 range = __ldg(m + 9);
		const int where = bezierFactors(time, range.x, range.y, f0, f1, f2);
		float4 row[3];
		for(int i = 0; i < 3; ++i)
		{
			const float4 a = __ldg(m + i), b = __ldg(m + 3 + i), c = __ldg(m + 6 + i);
			row[i].x = bezier3(where, a.x, b.x, c.x, f0, f1, f2);
			row[i].y = bezier3(where, a.y, b.y, c.y, f0, f1, f2);
			row[i].z = bezier3(where, a.z, b.z, c.z, f0, f1, f2);
			row[i].w = bezier3(where, a.w, b.w, c.w, f0, f1, f2);
		}
		for(int k = 0; k < nv; ++k)
		{
			const float vx = p[k].x, vy = p[k].y, vz = p[k].z;
			float o[3];
			for(int i = 0; i < 3; ++i)
			{
				float aux = __fadd_rn(0.f, __fmul_rn(row[i].x, vx));
				aux = __fadd_rn(aux, __fmul_rn(row[i].y, vy));
				aux = __fadd_rn(aux, __fmul_rn(row[i].z, vz));
				o[i] = __fadd_rn(aux, row[i].w);
			}
			p[k].x = o[0]; p[k].y = o[1]; p[k].z = o[2];
		}
		rec += nv;
	}
	const float4 e1 = make_float4(__fsub_rn(p[1].x, p[0].x), __fsub_rn(p[1].y, p[0].y), __fsub_rn(p[1].z, p[0].z), 0.f);
	const float4 e2 = make_float4(__fsub_rn(p[2].x, p[0].x), __fsub_rn(p[2].y, p[0].y), __fsub_rn(p[2].z, p[0].z), 0.f);
	const float4 e3 = quad ? make_float4(__fsub_rn(p[3].x, p[0].x), __fsub_rn(p[3].y, p[0].y), __fsub_rn(p[3].z, p[0].z), 0.f) : e2;
	return polyIntersectWith(p[0], e1, e2, [&]() { return e3; }, quad, ox, oy, oz, dx, dy, dz, u, v);
}

// SpherePrimitive::intersect, src/geometry/primitive/primitive_sphere.cc:83-102, on a sphere record (q0 = centre, q1.x =
// radius).  math::sqrt is std::sqrt in the reference build (include/math/math.h:148-173): IEEE, __fsqrt_rn.  Returns t (0 = miss).
__device__ __forceinline__ float sphereIntersect(const float4 q0, float radius, float ox, float oy, float oz, float dx, float dy, float dz)
{
	const float vx = __fsub_rn(ox, q0.x), vy = __fsub_rn(oy, q0.y), vz = __fsub_rn(oz, q0.z); // vf = from - center
	const float ea = dot3(dx, dy, dz, dx, dy, dz);
	const float eb = __fmul_rn(2.f, dot3(vx, vy, vz, dx, dy, dz));
	const float ec = __fsub_rn(dot3(vx, vy, vz, vx, vy, vz), __fmul_rn(radius, radius));
	float osc = __fsub_rn(__fmul_rn(eb, eb), __fmul_rn(__fmul_rn(4.f, ea), ec));
	if(osc < 0.f) return 0.f;
	osc = __fsqrt_rn(osc);
	const float two_ea = __fmul_rn(2.f, ea);
	float sol = __fdiv_rn(__fsub_rn(-eb, osc), two_ea);
	if(sol < 0.f)
	{
		sol = __fdiv_rn(__fadd_rn(-eb, osc), two_ea);
		if(sol < 0.f) return 0.f;
	}
	return sol;
}

// -------------------------------------------------------------------------------------------------
// Persistent-warp traversal ("while-while" with ray replacement).
//
// A warp owns 32 ray slots.  Whenever at least kRefill slots are idle, the idle lanes take the next rays of
// the warp's pool (the pool is refilled kPoolRays at a time with one atomicAdd on a global cursor; small batches are
// launched without a cursor, warp w of the grid owning rays [32 w, 32 w + 32)), so a warp's
// lifetime is no longer the maximum over its 32 first rays: measured with the one-thread-per-ray kernel, only
// 3.3 of 32 lanes were active per issued instruction (profiles/r1a_*), here they are kept busy.
// Every lane then alternates between
//   (1) descending the tree -- interior nodes AND empty leaves (two thirds of all leaves) are consumed here,
//   (2) testing the primitives of one non-empty leaf,
// and the two phases are warp-converged: the control flow below has one exit per loop (flags, no return
// from inside), so that the compiler's reconvergence points sit right after each phase.
//
// The traversal is t-interval based: [seg_lo, seg_hi] is the ray parameter range inside the current node.
//   t_plane >  min(seg_hi, closest so far)  -> near child only
//   t_plane <  seg_lo                       -> far child only
//   otherwise near first, far child pushed with the current seg_hi
// The comparisons are STRICT on purpose: a slab thinner than the resolution of t along this ray (a 1-ulp slab around an
// axis-aligned face, say) has t_plane == seg_lo on entry, and the faces inside it must still be met -- with "<=" whole cube
// faces of the reference's tests/test02 scene were skipped.  (split - o) * inv is monotonic in split, so entry and exit of
// a slab can coincide but never swap; a zero-length interval is traversed like any other.
// "near" is decided by the sign of the direction component (of its inverse, so that a zero component, whose
// inverse is +FLT_MAX as in math::inverse, math.h:71-77, behaves as "positive").
// -------------------------------------------------------------------------------------------------
// tuning knobs (overridable with -D for the sweeps under tools/; the defaults are the measured best)
#ifndef B200RT_BLOCK
#define B200RT_BLOCK 128
#endif
#ifndef B200RT_MIN_BLOCKS
#define B200RT_MIN_BLOCKS 9
#endif
#ifndef B200RT_REFILL
#define B200RT_REFILL 8
#endif
#ifndef B200RT_LEAF_BATCH
#define B200RT_LEAF_BATCH 8
#endif
#ifndef B200RT_POP_IN_LEAF_PHASE
#define B200RT_POP_IN_LEAF_PHASE 0
#endif
#ifndef B200RT_STREAM_IO
#define B200RT_STREAM_IO 1
#endif
#ifndef B200RT_STEPS
#define B200RT_STEPS 16
#endif
#ifndef B200RT_UNROLL
#define B200RT_UNROLL 4
#endif
#ifndef B200RT_BREAK
#define B200RT_BREAK 1
#endif
// Prefetch experiments (north_star (b), DESIGN.md "Prefetch"):
//   LEAF_PREFETCH 1: prefetch.global.L1 of the leaf's first record when a lane stops at a non-empty leaf (it is tested one vote later)
//   LEAF_PREFETCH 2: cp.async (LDGSTS) of that record into the lane's shared-memory staging slot; the leaf phase reads it from there
//   FAR_PREFETCH  1: prefetch.global.L1 of the postponed far child when it is pushed
//   POOL_PREFETCH 1: prefetch.global.L2 of the whole 256-ray pool when the warp takes it from the cursor
//   POOL_PREFETCH 2: (two-pass batches) prefetch.global.L1 of the next 32 queue entries right after a hand-out
// Two-pass batches (setupKernel feeding traceKernel<.., QUEUED>).  Ray setup -- ~170 instructions with three IEEE divisions,
// executed by ~10 of 32 lanes when idle lanes do it inside the traversal loop -- runs as its own fully converged pass that
// streams the ray records through shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier; SASS UBLKCP / SYNCS, the next
// 32 rays in flight while the current 32 are set up), answers the rays that miss the tree bound (41 % of the BASELINE closest
// rays) on the spot and writes the others, set up, to a queue in HBM: 64-byte entries, compacted per 256-ray region.  In the
// traversal pass idle lanes then take a ready entry (four 16-byte loads) instead of setting a ray up.  History of the idea
// (profiles/r3d_*, r3e_*, r3f_*; tools/simt_model.cc is the cost model): a converged setup pass INSIDE the traversal loop
// executed 10 % fewer instructions at 18.8 instead of 16.0 lanes but needed 72 registers -- held to 56 its spills made it 13 %
// slower; staging the queue in shared memory by bulk copies (7 KB per block) moved the SM to the 196 KB carve-out and left 32 KB
// of L1 -- 33 % instead of 54 % hits on node loads; reading a [field][entry] queue directly re-fetched every sector several times
// (4.5 GB of DRAM reads per launch instead of 1.6).  TWO_PASS 0 keeps large batches on the single-kernel path, which small
// batches and the renderer's mixed-kind flushes always use.
#ifndef B200RT_TWO_PASS
#define B200RT_TWO_PASS 1
#endif
#ifndef B200RT_TDONE
#define B200RT_TDONE 0
#endif
// AXIS_MAD 1: the address of the ray's (origin, inverse direction) row of the node's axis is one mad.lo + ld.shared instead of the
// shift / mask / add the compiler makes of sh_axis[axis][tid]; with LEAF_PREFETCH 1 (below) +1.7-2 % on both passes, byte-identical
// (profiles/r4j_knob_sweep.txt)
#ifndef B200RT_AXIS_MAD
#define B200RT_AXIS_MAD 1
#endif
// RING_OR 1: the ring is aligned to its own size, so a slot's address is (column address | slot bits): one LOP3 instead of
// LOP3 + IADD for the speculative store and for the read of the top entry.
#ifndef B200RT_RING_OR
#define B200RT_RING_OR 0
#endif
#ifndef B200RT_LEAF_PREFETCH
#define B200RT_LEAF_PREFETCH 1
#endif
// COOP_LEAF 1: warp-cooperative leaf phase (closest and shadow queries of polygon-only scenes).  The (ray, record) pairs of all
// lanes holding a leaf are laid out over the warp's 32 lanes -- an exclusive scan of the leaf sizes gives every pair a slot, the
// owners publish (owner lane, record number) per slot through 128 bytes of shared memory, the slot's lane fetches the owner's
// ray by shuffles and tests ONE record -- and each owner then takes, in leaf order, the nearest accepted candidate of its slots
// (the same result as the sequential accept rule: the first record among those with the smallest t).  Needs a uniform record
// stride per leaf (kLeafStride4, b200rt.cu flattening).  MEASURED DEAD END (profiles/r4a_knob_sweep.txt, r4b_coop_leaf_kernels.txt):
// byte-identical results, 21.9 instead of 18.5 active lanes per instruction and the wait for leaf records down from 13.8 % to
// 5.3 % of the warp time -- but the scan, the slot table, 13 shuffles and the gather loop cost ~240 instructions per leaf phase
// (344 per phase against 312 for the 2.7 sequential passes they replace), and the longer live ranges spill ~15 registers across
// every round of the outer loop (39 M local loads per launch instead of 1.8 M): closest 4655 instead of 5930 Mrays/s.  Off.
#ifndef B200RT_COOP_LEAF
#define B200RT_COOP_LEAF 0
#endif
#ifndef B200RT_FAR_PREFETCH
#define B200RT_FAR_PREFETCH 0
#endif
#ifndef B200RT_POOL_PREFETCH
#define B200RT_POOL_PREFETCH 0
#endif

static constexpr int kBlock = B200RT_BLOCK;
#ifndef B200RT_POOL
#define B200RT_POOL 256
#endif
static constexpr int kPoolRays = B200RT_POOL;       // rays taken from the global cursor per atomicAdd
static constexpr int kRefill = B200RT_REFILL;       // idle lanes that trigger a refill (single-kernel path)
#ifndef B200RT_TAKE
#define B200RT_TAKE 4
#endif
static constexpr int kTake = B200RT_TAKE;           // idle lanes that trigger a hand-out from the ray queue (two-pass batches)
// Ray queue between the two passes: per REGION (the rays of 256 consecutive batch indices) up to 256 ENTRIES of 16 floats, the rays
// that cross the tree bound in batch order:  ox oy oz dx | dy dz 1/dx 1/dy | 1/dz t_min t_max seg_lo | seg_hi index - -
// (traversal inverse direction; [seg_lo, seg_hi] = the ray's interval inside the tree bound; the 15th float is the ray time),
// followed by one uint32 per region,
// its number of entries.  Scratch bytes for a batch: see queueBytes().
static constexpr int kRegionRays = 256;
static constexpr int kEntryFloats = 16;
static constexpr int kRegionFloats = kRegionRays * kEntryFloats;
__host__ __device__ inline size_t queueBytes(uint32_t n_regions) { return size_t(n_regions) * (kRegionFloats * sizeof(float) + sizeof(uint32_t)); }
// TREELET 1: batches of polygon-only scenes on the two-pass path are traversed over two-level treelets (kEmptyRef above)
#ifndef B200RT_TREELET
#define B200RT_TREELET 0
#endif
#ifndef B200RT_TREELET_STEPS
#define B200RT_TREELET_STEPS 8
#endif
#ifndef B200RT_TREELET_UNROLL
#define B200RT_TREELET_UNROLL 2
#endif
static constexpr int kTreeletSteps = B200RT_TREELET_STEPS;   // treelet steps (two levels each) per lane between two warp votes
static constexpr int kTreeletUnroll = B200RT_TREELET_UNROLL;
static constexpr int kLeafBatch = B200RT_LEAF_BATCH; // lanes holding a leaf that trigger the leaf phase
static constexpr int kUnroll = B200RT_UNROLL;       // unroll factor of the descent loop
static constexpr int kSteps = B200RT_STEPS;         // node steps per lane between two warp votes
static constexpr int kMinBlocks = B200RT_MIN_BLOCKS; // blocks per SM the register allocation is held to
static constexpr unsigned kFullMask = 0xFFFFFFFFu;

struct RayState
{
	float ox, oy, oz, dx, dy, dz; // the ray the tree sees
	float ix, iy, iz;             // traversal inverse direction
	float t_min;                  // ray bias
	float t_max;                  // closest: shrinking; shadow: fixed
	float t_done;                 // closest (TDONE): +inf until a hit is accepted, then its t
	float seg_lo, seg_hi;
	float best_u, best_v;
	uint32_t best_prim;
	uint32_t node;
	uint32_t index;               // ray index in the batch
	float time;                   // Ray::time_ (motion blur; read by the SPHERES kernel variants only)
	int sp;                       // ring depth in bytes (kRingStride per entry)
};

// closest: once the best hit is not beyond the end of the current leaf nothing nearer can follow (accelerator_kdtree_common.h:232)
__device__ __forceinline__ bool closestDone(const RayState &r)
{
#if B200RT_TDONE
	return r.t_done <= r.seg_hi;
#else
	return r.best_prim != B200RT_MISS && r.t_max <= r.seg_hi;
#endif
}

// Ray setup: root slab test (bound.h:156-198), bias (accelerator.h:64), traversal interval.  Returns false when
// the ray misses the tree bound.
template <int QUERY>
__device__ __forceinline__ bool setupRay(const SceneView &s, const float4 a, const float4 b, RayState &r, bool tree_space)
{
	float t_max;
	if(QUERY == kClosest || tree_space)
	{
		// closest (accelerator.h:91), or a shadow ray its caller has already wrapped (B200RT_RAYS_TREE_SPACE)
		r.ox = a.x; r.oy = a.y; r.oz = a.z;
		t_max = (b.w >= 0.f) ? b.w : FLT_MAX;
	}
	else
	{
		// accelerator.h:103-111: origin moved by dir * tmin, t_max = tmax - 2 tmin (unbounded if tmax < 0)
		r.ox = __fadd_rn(a.x, __fmul_rn(b.x, a.w));
		r.oy = __fadd_rn(a.y, __fmul_rn(b.y, a.w));
		r.oz = __fadd_rn(a.z, __fmul_rn(b.z, a.w));
		t_max = (b.w >= 0.f) ? __fsub_rn(b.w, __fmul_rn(2.f, a.w)) : FLT_MAX;
	}
	r.dx = b.x; r.dy = b.y; r.dz = b.z;
	// math::inverse (math.h:71-77); the same quotient is what Bound::cross computes for a non-zero component
	const float ix = (r.dx == 0.f) ? FLT_MAX : __fdiv_rn(1.f, r.dx);
	const float iy = (r.dy == 0.f) ? FLT_MAX : __fdiv_rn(1.f, r.dy);
	const float iz = (r.dz == 0.f) ? FLT_MAX : __fdiv_rn(1.f, r.dz);
	float lmin = -FLT_MAX, lmax = FLT_MAX;
	bool crossed = true;
	const float o[3] = {r.ox, r.oy, r.oz}, d[3] = {r.dx, r.dy, r.dz}, inv[3] = {ix, iy, iz};
#pragma unroll
	for(int axis = 0; axis < 3; ++axis)
	{
		if(d[axis] != 0.f && crossed)
		{
			const float p = __fsub_rn(o[axis], s.bound[axis]);
			const float near_t = __fmul_rn(-p, inv[axis]);
			const float far_t = __fmul_rn(__fsub_rn(__fsub_rn(s.bound[3 + axis], s.bound[axis]), p), inv[axis]);
			const float ltmin = (inv[axis] > 0.f) ? near_t : far_t;
			const float ltmax = (inv[axis] > 0.f) ? far_t : near_t;
			if(axis == 0) { lmin = ltmin; lmax = ltmax; }
			else
			{
				lmin = (ltmin < lmin) ? lmin : ltmin; // std::max(ltmin, lmin)
				lmax = (lmax < ltmax) ? lmax : ltmax; // std::min(ltmax, lmax)
			}
			if((lmax < 0.f) || (lmin > t_max)) crossed = false;
		}
	}
	crossed = crossed && (lmin <= lmax) && (lmax >= 0.f) && (lmin <= t_max);
	const float bias = __fmul_rn(__fmul_rn(0.1f, 0.00005f), fabsf(__fsub_rn(lmax, lmin)));
	r.t_min = (QUERY == kShadow) ? bias : ((a.w < bias) ? bias : a.w); // accelerator_kdtree_common.h:139
	r.t_max = t_max;
	// traversal-only quantities (not part of the reference arithmetic)
	r.ix = isinf(ix) ? copysignf(FLT_MAX, ix) : ix;
	r.iy = isinf(iy) ? copysignf(FLT_MAX, iy) : iy;
	r.iz = isinf(iz) ? copysignf(FLT_MAX, iz) : iz;
	r.seg_lo = fmaxf(lmin, 0.f);
	r.seg_hi = fminf(lmax, t_max);
	r.best_u = 0.f; r.best_v = 0.f; r.best_prim = B200RT_MISS;
	r.t_done = __int_as_float(0x7f800000);
	r.node = 0u;
	r.sp = 0;
	return crossed;
}

// Transparent shadows: the distinct transparent casters a ray has met are written straight into its result record as they are
// found (and read back from there for the "seen before?" test of accelerator.h:160), so the kernel keeps only their number --
// no per-thread list in local memory, and no limit on the number but the record's capacity.  The kernels' max_depth parameter
// carries that capacity in its upper half: max_depth | capacity << 16 (capacity 0 = the 8 entries of b200rt_tshadow).
struct TShadowState
{
	int depth;
};
__device__ __forceinline__ int tsDepthLimit(int packed) { return packed & 0xFFFF; }
__device__ __forceinline__ int tsCapacity(int packed) { return (packed >> 16) ? (packed >> 16) : B200RT_TSHADOW_MAX; }

template <int QUERY> struct OutType;
template <> struct OutType<kClosest> { using type = b200rt_hit; };
template <> struct OutType<kShadow> { using type = uint32_t; };
template <> struct OutType<kTShadow> { using type = b200rt_tshadow; };

template <int QUERY>
__device__ __forceinline__ void writeResult(typename OutType<QUERY>::type *out, const RayState &r, bool hit, const TShadowState &ts, int ts_capacity = B200RT_TSHADOW_MAX)
{
	if(QUERY == kClosest)
	{
		float4 v;
		v.x = hit ? r.t_max : 0.f; v.y = r.best_u; v.z = r.best_v; v.w = __uint_as_float(r.best_prim);
#if B200RT_STREAM_IO
		__stcs(reinterpret_cast<float4 *>(out) + r.index, v);
#else
		reinterpret_cast<float4 *>(out)[r.index] = v;
#endif
	}
	else if(QUERY == kShadow)
	{
#if B200RT_STREAM_IO
		__stcs(reinterpret_cast<uint32_t *>(out) + r.index, hit ? r.best_prim : B200RT_MISS);
#else
		reinterpret_cast<uint32_t *>(out)[r.index] = hit ? r.best_prim : B200RT_MISS;
#endif
	}
	else
	{
		// record = header + ts_capacity entries of 16 bytes; entries [0, depth) are in place already
		uint4 *o = reinterpret_cast<uint4 *>(out) + size_t(r.index) * size_t(1 + ts_capacity);
		o[0] = make_uint4(hit ? 1u : 0u, uint32_t(ts.depth), hit ? r.best_prim : B200RT_MISS, 0u); // setNoHit() clears primitive_
		if(ts_capacity == B200RT_TSHADOW_MAX) // the fixed-size record of b200rt_tshadow is defined to its last byte; larger ones only up to n_transparent
			for(int k = ts.depth; k < B200RT_TSHADOW_MAX; ++k) o[1 + k] = make_uint4(0u, 0u, 0u, B200RT_MISS);
	}
}

// ---- per-thread short stack in shared memory -----------------------------------------------------
// kShortStack entries per thread, laid out [entry][thread] so that a warp's accesses are conflict-free
// whatever the lanes' stack depths are (bank = thread % 32).  It is a ring: a push onto a full ring
// overwrites the OLDEST entry and raises `floor`; popping down to a raised floor means entries were lost,
// and the ray then REPLAYS its descent from the root along the path to the leaf it just left, postponing the
// far children of that path again (the ring keeps the deepest ones, the ones needed next), and pops (an exact
// kd-restart: progress does not depend on t, so zero-length intervals cannot make it loop).  Pushes are unconditional stores (the slot is simply not committed when the ray visits
// one child only), which keeps the node step free of divergent branches; the price is that the ring
// effectively holds kShortStack - 1 entries.
#ifndef B200RT_SHORT_STACK
#define B200RT_SHORT_STACK 8
#endif
static constexpr int kShortStack = B200RT_SHORT_STACK;
static_assert((kShortStack & (kShortStack - 1)) == 0, "ring size must be a power of two");
static constexpr int kRingStride = kBlock * int(sizeof(uint2));        // bytes between two entries of one thread's ring
static constexpr int kRingMask = (kShortStack - 1) * kRingStride;

// The largest float below x (x finite).  Used for the traversal copy of the ray origin on axes where the direction is exactly
// zero: such a ray never crosses a plane of that axis, it lies on one side of it or IN it, and for "in it" the reference takes
// the left child only (entry[axis] <= split and exit[axis] <= split, accelerator_kdtree_common.h:148-174).  With the origin
// one ulp lower, (split - o) * FLT_MAX is +huge for o <= split (near = left only) and <= 0 for o > split (right), the same
// choice -- measured on the thin-slab cube grid, 18 % of the reference's hits on in-plane rays lay in leaves the unshifted
// rule did not open.  The leaf tests use the true origin.
__device__ __forceinline__ float floatBelow(float x)
{
	const int bits = __float_as_int(x);
	return (x > 0.f) ? __int_as_float(bits - 1) : ((x < 0.f) ? __int_as_float(bits + 1) : -1.401298464e-45f);
}

__device__ __forceinline__ float selectf(bool p, float a, float b)
{
	float r;
	asm("{ .reg .pred q; setp.ne.s32 q, %3, 0; selp.f32 %0, %1, %2, q; }" : "=f"(r) : "f"(a), "f"(b), "r"(int(p)));
	return r;
}
__device__ __forceinline__ uint32_t selectu(bool p, uint32_t a, uint32_t b)
{
	uint32_t r;
	asm("{ .reg .pred q; setp.ne.s32 q, %3, 0; selp.u32 %0, %1, %2, q; }" : "=r"(r) : "r"(a), "r"(b), "r"(int(p)));
	return r;
}

// LEAF_NOALLOC 1: leaf records (read once per test, 73 MB of them on the BASELINE scene) are loaded without allocating in L1, so
// that they do not displace tree nodes there
#ifndef B200RT_LEAF_NOALLOC
#define B200RT_LEAF_NOALLOC 0
#endif
__device__ __forceinline__ float4 loadRecord(const float4 *p)
{
#if B200RT_LEAF_NOALLOC
	float4 v;
	asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
	return v;
#else
	return __ldg(p);
#endif
}
__device__ __forceinline__ void prefetchL1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetchL2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cpAsync16(void *smem, const void *gmem)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(uint32_t(__cvta_generic_to_shared(smem))), "l"(gmem));
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cpAsyncWaitAll() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- 1-D TMA bulk copy global -> shared, completion on an mbarrier (one warp = one consumer, lane 0 = producer) ---------------
__device__ __forceinline__ void mbarInit(uint32_t mbar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulkCopyToShared(void *dst, const void *src, uint32_t bytes, uint32_t mbar)
{
	// the buffer was last touched by ordinary shared-memory loads: order them before the async-proxy write
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(uint32_t(__cvta_generic_to_shared(dst))), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
// shared -> global: the block of entries a warp has staged in shared memory leaves as ONE bulk store
__device__ __forceinline__ void bulkStoreFromShared(void *dst, const void *src, uint32_t bytes)
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the staging writes were ordinary shared-memory stores
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(uint32_t(__cvta_generic_to_shared(src))), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulkStoreWaitRead() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulkStoreWaitAll() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ bool mbarTryWait(uint32_t mbar, uint32_t parity)
{
	uint32_t done;
	asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
	return done != 0u;
}

// The traversal proper.  Called by every thread of the block; warps are independent of each other (no block-level
// synchronisation).  sh_stack / sh_axis are the block's shared arrays; static_base is the first ray of the calling warp's
// static pool in a cursor-less launch (ignored when `cursor` is given).
// SPHERES: the scene holds sphere records (b200rt_add_spheres).  A compile-time switch, so that scenes of polygons only -- the
// BASELINE workloads -- do not pay for the flag test and the extra code in the leaf loop (measured: 3 % on S1M-hf).
// QUEUED: `rays` is the queue setupKernel wrote and n its number of regions; otherwise `rays` are the batch's ray records and idle
// lanes set their rays up themselves.
template <int QUERY, bool SPHERES, bool QUEUED = false, bool TREELET = false>
__device__ __forceinline__ void traceWarps(const SceneView &s, const b200rt_ray *__restrict__ rays, uint32_t n, typename OutType<QUERY>::type *__restrict__ out,
                                           uint32_t *__restrict__ cursor, int max_depth, bool tree_space, uint2 (*sh_stack)[kBlock], float2 (*sh_axis)[kBlock], uint32_t static_base, float4 (*sh_leaf)[kBlock] = nullptr,
                                           const float *__restrict__ times = nullptr, uint32_t (*sh_task)[32] = nullptr)
{
	constexpr bool kCoopLeaf = B200RT_COOP_LEAF && !SPHERES && QUERY != kTShadow;
	const unsigned tid = threadIdx.x;
	const unsigned lane = tid & 31u;
	const unsigned lanes_below = (1u << lane) - 1u;
	// the thread's column of the ring; r.sp and floor count in bytes of it (kRingStride per entry), which saves the shifts
	char *const ring = reinterpret_cast<char *>(&sh_stack[0][tid]);
	const uint32_t axis_base = uint32_t(__cvta_generic_to_shared(&sh_axis[0][tid])); (void)axis_base;
	const uint32_t ring_addr = uint32_t(__cvta_generic_to_shared(ring)); (void)ring_addr; // RING_OR: the slot bits of this address are zero
	RayState r;
	TShadowState ts;
	ts.depth = 0;
	// lane state: !alive = idle slot; alive && !pending = descending; alive && pending = holds a non-empty leaf
	bool alive = false, pending = false;
	uint32_t leaf_count = 0u, leaf_first = 0u;
	int floor = 0;          // ring entries below this index were overwritten
	uint32_t pool_next = 0u, pool_end = 0u; // warp-uniform
	bool exhausted = false;                 // warp-uniform
	bool first_pool = true;                 // warp-uniform
	// two-pass batches (QUEUED): pool_next / pool_end count queue entries of the region being consumed; q_next is the region the
	// cursor has already handed to this warp for later and q_next_count its size, both valid in lane 0 and fetched one region
	// ahead so that neither the atomic nor the load is waited for
	uint32_t q_next = 0u, q_next_count = 0u;
	if(QUEUED && lane == 0u)
	{
		q_next = atomicAdd(cursor, 1u);
		if(q_next < n) q_next_count = __ldg(reinterpret_cast<const uint32_t *>(reinterpret_cast<const float *>(rays) + size_t(n) * kRegionFloats) + q_next);
	}

	// Exact kd-restart.  `target` is the leaf the ray has just left (ring empty, older entries lost).  The tree is stored
	// depth first (left child = node + 1, the right subtree starts at `right`), so "target < right" tells which child holds
	// it.  On the way down the decisions of the first descent are taken again -- the child without the target is
	// postponed when the ray enters it AFTER the target's side, and skipped when it lies before (already traversed) --
	// then the next postponed subtree is popped.  Returns true when nothing is left.  Rare path (the ring holds the 7
	// deepest entries), divergent on purpose.
	bool need_replay = false;
	// TREELET: the ring holds the entries [floor, sp) of the stack, the older ones [0, floor) were moved to global memory (spill)
	// before they could be overwritten; an empty ring with floor > 0 continues there.  No replay in this variant.
	const uint32_t gtid = blockIdx.x * uint32_t(kBlock) + tid; (void)gtid;
	auto globalPop = [&]() -> bool {
		floor -= kRingStride;
		r.sp = floor;
		const uint2 e = s.spill[size_t(floor / kRingStride) * s.spill_threads + gtid];
		r.node = e.x;
		r.seg_lo = r.seg_hi;
		r.seg_hi = __uint_as_float(e.y);
		return false;
	};
	auto replayTo = [&](const uint32_t target) -> bool {
		r.sp = 0;
		floor = 0;
		uint32_t node = 0u;
		const float2 whole = sh_axis[3][tid]; // the interval the ray was set up with
		// the leaf just left reaches the end of the ray's stay in the (inflated, hence empty at its faces) tree bound: nothing
		// can follow.  Most rays whose ring ever overflowed end this way, and are spared the descent.
		if(!(r.seg_hi < whole.y)) return true;
		float hi = whole.y;
		while(node != target)
		{
			const uint2 nd = __ldg(&s.nodes[node]); // interior: the target lies below it
			const uint32_t right = nd.y >> 2;
			const float2 oi = sh_axis[nd.y & 3u][tid];
			const float t_plane = (__uint_as_float(nd.x) - oi.x) * oi.y;
			const bool negative = __float_as_int(oi.y) < 0;
			const uint32_t left = node + 1u;
			const uint32_t near = negative ? right : left, far = negative ? left : right;
			const bool target_near = ((target < right) != negative);
			if(target_near)
			{
				const float limit = (QUERY == kClosest) ? fminf(hi, B200RT_TDONE ? r.t_done : r.t_max) : hi;
				if(!(t_plane > limit)) // (t_plane < interval start cannot be: the first descent would have taken the far child only)
				{
					*reinterpret_cast<uint2 *>(ring + (r.sp & kRingMask)) = make_uint2(far, __float_as_uint(hi));
					floor = max(floor, r.sp + (1 - kShortStack) * kRingStride);
					r.sp += kRingStride;
					hi = t_plane;
				}
				node = near;
			}
			else node = far; // the near side lies before the target: done
		}
		if(r.sp <= floor) return true; // nothing was postponed behind the target
		r.sp -= kRingStride;
		const uint2 e = *reinterpret_cast<const uint2 *>(ring + (r.sp & kRingMask));
		r.node = e.x;
		r.seg_lo = hi;
		r.seg_hi = __uint_as_float(e.y);
		return false;
	};

	// Leave the current leaf.  Closest queries stop once the best hit is not beyond the end of this leaf
	// (accelerator_kdtree_common.h:232); otherwise continue with the nearest postponed subtree, or replay the descent
	// when ring entries were lost.  Returns true when the ray has ended.
	auto popNode = [&]() -> bool {
		if(QUERY == kClosest && closestDone(r)) return true;
		if(r.sp > floor)
		{
			r.sp -= kRingStride;
			const uint2 e = *reinterpret_cast<const uint2 *>(ring + (r.sp & kRingMask));
			r.node = e.x;
			r.seg_lo = r.seg_hi;
			r.seg_hi = __uint_as_float(e.y);
			return false;
		}
		if(floor == 0) return true;
		need_replay = true; // done once, after the phase (one copy of the replay code)
		return false;
	};

	for(;;)
	{
		const unsigned idle = __ballot_sync(kFullMask, !alive);
		if(QUEUED)
		{
			// ---------------- hand-out of set-up rays from the queue ----------------
			if(!exhausted && __popc(idle) >= kTake)
			{
				if(pool_next == pool_end)
				{
					// next region (possibly empty: all of its rays missed the bound -- the next round moves on)
					const uint32_t region = __shfl_sync(kFullMask, q_next, 0);
					const uint32_t count = __shfl_sync(kFullMask, q_next_count, 0);
					if(region >= n) exhausted = true;
					else
					{
						pool_next = region * uint32_t(kRegionRays);
						pool_end = pool_next + count;
#if B200RT_POOL_PREFETCH == 1
						// the region's entries (64 bytes each) will be read a few at a time over the next rounds: pull them into L2 now
						for(uint32_t line = lane; line * 2u < count; line += 32u) prefetchL2(reinterpret_cast<const float *>(rays) + (size_t(pool_next) + line * 2u) * kEntryFloats);
#endif
						if(lane == 0u)
						{
							q_next = atomicAdd(cursor, 1u);
							if(q_next < n) q_next_count = __ldg(reinterpret_cast<const uint32_t *>(reinterpret_cast<const float *>(rays) + size_t(n) * kRegionFloats) + q_next);
						}
					}
				}
				if(pool_next != pool_end)
				{
					const uint32_t avail = pool_end - pool_next;
					const uint32_t rank = __popc(idle & lanes_below);
					if(!alive && rank < avail)
					{
						const float4 *e = reinterpret_cast<const float4 *>(rays) + size_t(pool_next + rank) * (kEntryFloats / 4);
						const float4 e0 = __ldcs(e), e1 = __ldcs(e + 1), e2 = __ldcs(e + 2), e3 = __ldcs(e + 3);
						r.ox = e0.x; r.oy = e0.y; r.oz = e0.z; r.dx = e0.w; r.dy = e1.x; r.dz = e1.y;
						r.t_min = e2.y; r.t_max = e2.z; r.seg_lo = e2.w; r.seg_hi = e3.x;
						r.index = __float_as_uint(e3.y);
						if(SPHERES) r.time = e3.z;
						r.best_u = 0.f; r.best_v = 0.f; r.best_prim = B200RT_MISS;
						r.t_done = __int_as_float(0x7f800000);
						r.node = 0u;
						r.sp = 0;
						ts.depth = 0;
						floor = 0;
						alive = true;
						sh_axis[0][tid] = make_float2(r.ox, e1.z);
						sh_axis[1][tid] = make_float2(r.oy, e1.w);
						sh_axis[2][tid] = make_float2(r.oz, e2.x);
						if(__builtin_expect(r.dx == 0.f || r.dy == 0.f || r.dz == 0.f, 0))
						{
							// axis-parallel ray: traversal copy of the origin one ulp lower on the zero-direction axes (floatBelow)
							if(r.dx == 0.f) sh_axis[0][tid].x = floatBelow(r.ox);
							if(r.dy == 0.f) sh_axis[1][tid].x = floatBelow(r.oy);
							if(r.dz == 0.f) sh_axis[2][tid].x = floatBelow(r.oz);
						}
						sh_axis[3][tid] = TREELET ? make_float2(0.f, 1.f) : make_float2(r.seg_lo, r.seg_hi); // where the ray enters and leaves the tree bound (read by replayTo); a leaf's "axis" 3 also reads this row, value unused.  TREELET: the "axis" of a child that is a leaf, (FLT_MAX - 0) * 1 puts its plane beyond every interval
					}
					pool_next += min(avail, uint32_t(__popc(idle)));
#if B200RT_POOL_PREFETCH == 2
					// the entries the next hand-outs will take: on their way into L1 while this round's rays descend
					if(pool_next + lane < pool_end)
					{
						const float4 *e = reinterpret_cast<const float4 *>(rays) + size_t(pool_next + lane) * (kEntryFloats / 4);
						prefetchL1(e);
						prefetchL1(e + 2);
					}
#endif
				}
			}
		}
		else
		{
			// ---------------- refill idle lanes ----------------
			// One pass per round: every idle lane takes the next ray of the warp's pool.  Rays that miss the tree bound
			// are answered on the spot and leave their lane idle until the next round (looping here until every lane
			// holds a live ray ran the ~160-instruction setup with only a few lanes active).
			if(!exhausted && (__popc(idle) >= kRefill))
			{
				if(pool_next == pool_end)
				{
					uint32_t base = n;
					if(cursor != nullptr)
					{
						if(lane == 0u) base = atomicAdd(cursor, uint32_t(kPoolRays));
						base = __shfl_sync(kFullMask, base, 0);
					}
					else if(first_pool) base = static_base; // cursor-less launch: the warp owns rays [static_base, static_base + 32)
					first_pool = false;
					if(base >= n) exhausted = true;
					else
					{
						pool_next = base;
						const uint32_t pool = (cursor != nullptr) ? uint32_t(kPoolRays) : 32u;
						pool_end = (n - base < pool) ? n : base + pool;
#if B200RT_POOL_PREFETCH == 1
						// the pool is 8 KB of rays that will be read 8..32 rays at a time over the next few thousand cycles: pull it from HBM into L2 now
						for(uint32_t line = lane; line * 4u < pool_end - base; line += 32u) prefetchL2(rays + base + line * 4u);
#endif
					}
				}
				if(!exhausted)
				{
					const uint32_t avail = pool_end - pool_next;
					const uint32_t rank = __popc(idle & lanes_below);
					if(!alive && rank < avail)
					{
						r.index = pool_next + rank;
						if(SPHERES) r.time = times ? __ldg(times + r.index) : 0.f;
#if B200RT_STREAM_IO
						// rays are read once: do not let them displace tree nodes from L1/L2
						const float4 a = __ldcs(reinterpret_cast<const float4 *>(rays) + 2 * size_t(r.index));
						const float4 b = __ldcs(reinterpret_cast<const float4 *>(rays) + 2 * size_t(r.index) + 1);
#else
						const float4 a = __ldg(reinterpret_cast<const float4 *>(rays) + 2 * size_t(r.index));
						const float4 b = __ldg(reinterpret_cast<const float4 *>(rays) + 2 * size_t(r.index) + 1);
#endif
						ts.depth = 0;
						floor = 0;
						alive = setupRay<QUERY>(s, a, b, r, tree_space);
						sh_axis[0][tid] = make_float2(r.ox, r.ix);
						sh_axis[1][tid] = make_float2(r.oy, r.iy);
						sh_axis[2][tid] = make_float2(r.oz, r.iz);
						if(__builtin_expect(r.dx == 0.f || r.dy == 0.f || r.dz == 0.f, 0))
						{
							// axis-parallel ray: traversal copy of the origin one ulp lower on the zero-direction axes (floatBelow)
							if(r.dx == 0.f) sh_axis[0][tid].x = floatBelow(r.ox);
							if(r.dy == 0.f) sh_axis[1][tid].x = floatBelow(r.oy);
							if(r.dz == 0.f) sh_axis[2][tid].x = floatBelow(r.oz);
						}
						sh_axis[3][tid] = make_float2(r.seg_lo, r.seg_hi); // where the ray enters and leaves the tree bound (read by replayTo); a leaf's "axis" 3 also reads this row, value unused
						if(!alive) writeResult<QUERY>(out, r, false, ts, tsCapacity(max_depth)); // missed the tree bound
					}
					pool_next += min(avail, uint32_t(__popc(idle)));
				}
			}
		}
		const unsigned m_alive = __ballot_sync(kFullMask, alive);
		if(m_alive == 0u)
		{
			if(exhausted) break;
			continue;
		}
		const unsigned m_pending = __ballot_sync(kFullMask, pending);
		bool finished = false; // ray ended in this round (result to be written)
		bool hit = false;      // shadow queries: occluded

		if(__popc(m_pending) >= kLeafBatch || m_pending == m_alive)
		{
			// ---------------- leaf phase: every lane holding a leaf tests its primitives ----------------
			if(kCoopLeaf)
			{
				// warp-cooperative (B200RT_COOP_LEAF): all 32 lanes take part, whatever their own state
				const uint32_t w = tid >> 5;
				const uint32_t c = pending ? (leaf_count & kLeafCountMask) : 0u;
				uint32_t incl = c; // inclusive scan of the leaf sizes over the lanes
#pragma unroll
				for(int d = 1; d < 32; d <<= 1)
				{
					const uint32_t up = __shfl_up_sync(kFullMask, incl, d);
					if(lane >= uint32_t(d)) incl += up;
				}
				const uint32_t start = incl - c, total = __shfl_sync(kFullMask, incl, 31);
				const uint32_t my_stride = (leaf_count & kLeafStride4) ? 4u : 3u;
				const uint32_t need = (QUERY == kClosest) ? uint32_t(B200RT_FACE_VISIBLE) : uint32_t(B200RT_FACE_CASTS_SHADOWS);
				for(uint32_t base = 0u; base < total; base += 32u)
				{
					// the owners publish their pairs of this chunk: slot -> (owner lane, record number within the leaf)
					const uint32_t own_lo = max(start, base), own_hi = min(start + c, base + 32u); // own slots of this chunk: [own_lo, own_hi), empty when own_lo >= own_hi
					for(uint32_t slot = own_lo; slot < own_hi; ++slot) sh_task[w][slot - base] = lane | ((slot - start) << 5);
					__syncwarp();
					const uint32_t task = sh_task[w][lane];
					const uint32_t owner = task & 31u, k = task >> 5;
					const bool active = base + lane < total;
					// the owner's ray and leaf (stale slots of the last chunk name some lane: harmless, their result is not used)
					const float tox = __shfl_sync(kFullMask, r.ox, owner), toy = __shfl_sync(kFullMask, r.oy, owner), toz = __shfl_sync(kFullMask, r.oz, owner);
					const float tdx = __shfl_sync(kFullMask, r.dx, owner), tdy = __shfl_sync(kFullMask, r.dy, owner), tdz = __shfl_sync(kFullMask, r.dz, owner);
					const float t_lo = __shfl_sync(kFullMask, r.t_min, owner), t_hi = __shfl_sync(kFullMask, r.t_max, owner);
					const uint32_t first = __shfl_sync(kFullMask, leaf_first, owner), stride = __shfl_sync(kFullMask, my_stride, owner);
					float cand = __int_as_float(0x7f800000), cu = 0.f, cv = 0.f; // +inf = no acceptable hit in this slot
					uint32_t cprim = B200RT_MISS;
					if(active)
					{
						const float4 *rec = s.tris + first + k * stride;
						const float4 q0 = __ldg(rec), q1 = __ldg(rec + 1), q2 = __ldg(rec + 2);
						const uint32_t flags = __float_as_uint(q1.w);
						float u, v;
						const float t = polyIntersect(q0, q1, q2, rec + 3, (flags & kFlagQuad) != 0u, tox, toy, toz, tdx, tdy, tdz, u, v);
						// accept rules, accelerator.h:125-127 / :137-139
						if(!(t <= 0.f || t < t_lo || t >= t_hi) && (flags & need)) { cand = t; cu = u; cv = v; cprim = __float_as_uint(q0.w); }
					}
					// every owner goes through its slots in leaf order: closest keeps the first of the nearest candidates (what the
					// sequential rule "accept when t < t_max, then t_max = t" ends with), shadow the first candidate
					const uint32_t n_own = (own_hi > own_lo) ? own_hi - own_lo : 0u;
					const uint32_t n_most = __reduce_max_sync(kFullMask, n_own);
					float best = (QUERY == kClosest) ? r.t_max : __int_as_float(0x7f800000);
					uint32_t best_slot = 0u;
					bool found = false;
					for(uint32_t j = 0u; j < n_most; ++j)
					{
						const uint32_t src = (own_lo - base + j) & 31u;
						const float tj = __shfl_sync(kFullMask, cand, src);
						if(j < n_own && tj < best && !(QUERY == kShadow && found)) { best = tj; best_slot = src; found = true; }
					}
					const float bu = __shfl_sync(kFullMask, cu, best_slot), bv = __shfl_sync(kFullMask, cv, best_slot);
					const uint32_t bprim = __shfl_sync(kFullMask, cprim, best_slot);
					if(found && !hit) // (shadow: the occluder is the first candidate of the first chunk that has one)
					{
						r.best_u = bu; r.best_v = bv; r.best_prim = bprim;
						if(QUERY == kClosest) { r.t_max = best; r.t_done = best; }
						else hit = true;
					}
					__syncwarp(); // the slots are rewritten by the next chunk
				}
				if(pending)
				{
					pending = false;
					finished = hit || popNode();
				}
			}
			else if(pending)
			{
				const float4 *rec = s.tris + leaf_first;
				const bool leaf_stride4 = (leaf_count & kLeafStride4) != 0u;
				leaf_count &= kLeafCountMask;
				const float lox = r.ox, loy = r.oy, loz = r.oz, ldx = r.dx, ldy = r.dy, ldz = r.dz;
#if B200RT_LEAF_PREFETCH == 2
				cpAsyncWaitAll();
				float4 q0 = sh_leaf[0][tid], q1 = sh_leaf[1][tid], q2 = sh_leaf[2][tid];
#else
				float4 q0 = loadRecord(rec), q1 = loadRecord(rec + 1), q2 = loadRecord(rec + 2);
#endif
				if(TREELET) leaf_count = __float_as_uint(q2.w); // a leaf reference carries no count: it rides in the first record
				for(;;)
				{
					const uint32_t flags = __float_as_uint(q1.w);
					const bool quad = (flags & kFlagQuad) != 0u;
					float u, v, t;
					if(SPHERES && (flags & kFlagSphere)) { t = sphereIntersect(q0, q1.x, lox, loy, loz, ldx, ldy, ldz); u = 0.f; v = 0.f; rec += 3; }
					else if(SPHERES && (flags & (kFlagBezier | kFlagMoving))) t = motionIntersect(s, rec, q0, q1, q2, flags, quad, r.time, lox, loy, loz, ldx, ldy, ldz, u, v);
					else { t = polyIntersect(q0, q1, q2, rec + 3, quad, lox, loy, loz, ldx, ldy, ldz, u, v); rec += (quad || leaf_stride4) ? 4 : 3; }
					--leaf_count;
					// accept rules, accelerator.h:125-127 / :137-139 / :150-154
					const uint32_t need = (QUERY == kClosest) ? uint32_t(B200RT_FACE_VISIBLE) : uint32_t(B200RT_FACE_CASTS_SHADOWS);
					if(!(t <= 0.f || t < r.t_min || t >= r.t_max) && (flags & need))
					{
						const uint32_t prim = __float_as_uint(q0.w);
						r.best_u = u; r.best_v = v; r.best_prim = prim;
						if(QUERY == kClosest) { r.t_max = t; r.t_done = t; }
						else if(QUERY == kShadow) { hit = true; leaf_count = 0u; }
						else
						{
							if(!(flags & B200RT_FACE_TRANSPARENT)) { hit = true; leaf_count = 0u; } // opaque caster
							else
							{
								uint4 *list = reinterpret_cast<uint4 *>(out) + size_t(r.index) * size_t(1 + tsCapacity(max_depth)) + 1;
								bool seen = false;
								for(int k = 0; k < ts.depth; ++k) seen = seen || (*reinterpret_cast<volatile uint32_t *>(&list[k].w) == prim);
								if(!seen)
								{
									if(ts.depth >= tsDepthLimit(max_depth)) { hit = true; leaf_count = 0u; }
									else
									{
										list[ts.depth] = make_uint4(__float_as_uint(t), __float_as_uint(u), __float_as_uint(v), prim);
										++ts.depth;
									}
								}
							}
						}
					}
					if(leaf_count == 0u) break;
					q0 = loadRecord(rec); q1 = loadRecord(rec + 1); q2 = loadRecord(rec + 2);
				}
				pending = false;
				finished = hit || popNode();
			}
		}
		else
		{
			// ---------------- descend: up to kSteps nodes per lane; empty leaves are consumed here ----------------
			// Lanes that descend in this round; a lane drops out at a leaf it cannot pop past: a non-empty leaf, the end of its ray,
			// or an empty ring with lost entries.  Which of the three is decided after the loop, from state the step left untouched.
			const bool descending = alive && !pending;
			bool go = descending && !(TREELET && (r.node & kLeafRef)); // TREELET: a popped reference may be a leaf: the lane stops at once and holds it
			if constexpr(TREELET)
			{
				// One TREELET per step: the node's plane and both children's planes are known after ONE load, the ray's way through
				// the two levels -- up to four grandchildren, front to back -- is worked out without divergent branches, the first
				// one that is neither culled nor empty is entered, the others are postponed (last first), and a step that finds
				// nothing to enter pops.  A reference with kLeafRef set ends the lane's descent (non-empty leaf).
				static_assert(kBlock * sizeof(float2) == 1024, "meta bytes hold axis << 2: the byte, moved to bits 8..15, is the row offset");
#pragma unroll kTreeletUnroll
				for(int step = 0; step < kTreeletSteps; ++step)
				{
					if(!go) break;
					// room for three pushes: the oldest ring entries move to global memory first (rare: 6 % of the rays ever get here)
					if(__builtin_expect(r.sp - floor > (kShortStack - 3) * kRingStride, 0))
					{
#pragma unroll 1
						while(r.sp - floor > (kShortStack - 3) * kRingStride)
						{
							s.spill[size_t(floor / kRingStride) * s.spill_threads + gtid] = *reinterpret_cast<const uint2 *>(ring + (floor & kRingMask));
							floor += kRingStride;
						}
					}
					const uint4 A = __ldg(s.treelets + 2 * size_t(r.node)), B = __ldg(s.treelets + 2 * size_t(r.node) + 1);
					float2 oi0, oiL, oiR;
					asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(oi0.x), "=f"(oi0.y) : "r"(axis_base + __byte_perm(A.w, 0u, 0x4404u)));
					asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(oiL.x), "=f"(oiL.y) : "r"(axis_base + __byte_perm(A.w, 0u, 0x4414u)));
					asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(oiR.x), "=f"(oiR.y) : "r"(axis_base + __byte_perm(A.w, 0u, 0x4424u)));
					const float t0 = (__uint_as_float(A.x) - oi0.x) * oi0.y, tL = (__uint_as_float(A.y) - oiL.x) * oiL.y, tR = (__uint_as_float(A.z) - oiR.x) * oiR.y;
					const bool n0 = __float_as_int(oi0.y) < 0;
					// the child entered first (N) and the other one (F), and their children in the order the ray meets them
					const float tN = n0 ? tR : tL, tF = n0 ? tL : tR;
					const bool nN = __float_as_int(n0 ? oiR.y : oiL.y) < 0, nF = __float_as_int(n0 ? oiL.y : oiR.y) < 0;
					const uint32_t cN0 = n0 ? B.z : B.x, cN1 = n0 ? B.w : B.y, cF0 = n0 ? B.x : B.z, cF1 = n0 ? B.y : B.w;
					const uint32_t g0 = nN ? cN1 : cN0, g1 = nN ? cN0 : cN1, g2 = nF ? cF1 : cF0, g3 = nF ? cF0 : cF1;
					// the same strict rule as the node step, twice: t_plane > limit -> near only, t_plane < start -> far only
					const float limit = (QUERY == kClosest) ? fminf(r.seg_hi, r.t_max) : r.seg_hi;
					const bool visN = !(t0 < r.seg_lo), visF = !(t0 > limit);
					const float hiN = visF ? t0 : r.seg_hi, loF = visN ? t0 : r.seg_lo;
					const float limN = (QUERY == kClosest) ? fminf(hiN, r.t_max) : hiN;
					const bool w0 = visN && !(tN < r.seg_lo), w1 = visN && !(tN > limN), w2 = visF && !(tF < loF), w3 = visF && !(tF > limit);
					const float h0 = w1 ? tN : hiN, h2 = w3 ? tF : r.seg_hi; // ends of the intervals of g0 and g2 (g1: hiN, g3: seg_hi)
					const float l1 = w0 ? tN : r.seg_lo, l3 = w2 ? tF : loF;  // starts of g1 and g3 (g0: seg_lo, g2: loF)
					const bool v0 = w0 && g0 != kEmptyRef, v1 = w1 && g1 != kEmptyRef, v2 = w2 && g2 != kEmptyRef, v3 = w3 && g3 != kEmptyRef;
					// top of the ring, for the step that has nothing to enter
					const uint2 popped = *reinterpret_cast<const uint2 *>(ring + ((r.sp - kRingStride) & kRingMask));
					// postpone all but the first, the last one first
					if(v3 && (v0 || v1 || v2)) { *reinterpret_cast<uint2 *>(ring + (r.sp & kRingMask)) = make_uint2(g3, __float_as_uint(r.seg_hi)); r.sp += kRingStride; }
					if(v2 && (v0 || v1)) { *reinterpret_cast<uint2 *>(ring + (r.sp & kRingMask)) = make_uint2(g2, __float_as_uint(h2)); r.sp += kRingStride; }
					if(v1 && v0) { *reinterpret_cast<uint2 *>(ring + (r.sp & kRingMask)) = make_uint2(g1, __float_as_uint(hiN)); r.sp += kRingStride; }
					const bool any = v0 || v1 || v2 || v3;
					const bool closest_done = (QUERY == kClosest) && closestDone(r);
					const bool do_pop = !any && !closest_done && r.sp > floor;
					const uint32_t next = v0 ? g0 : (v1 ? g1 : (v2 ? g2 : g3));
					const float next_lo = v0 ? r.seg_lo : (v1 ? l1 : (v2 ? loF : l3));
					const float next_hi = v0 ? h0 : (v1 ? hiN : (v2 ? h2 : r.seg_hi));
					r.node = any ? next : selectu(do_pop, popped.x, r.node);
					r.seg_lo = any ? next_lo : selectf(do_pop, r.seg_hi, r.seg_lo);
					r.seg_hi = any ? next_hi : selectf(do_pop, __uint_as_float(popped.y), r.seg_hi);
					if(do_pop) r.sp -= kRingStride;
					go = (any || do_pop) && !(r.node & kLeafRef);
				}
				if(descending && !go)
				{
					if(r.node & kLeafRef)
					{
						pending = true;
						leaf_first = r.node & ~kLeafRef;
						leaf_count = 1u; // the real count is read with the first record
#if B200RT_LEAF_PREFETCH == 1
						prefetchL1(s.tris + leaf_first);
						prefetchL1(s.tris + leaf_first + 2);
#endif
					}
					else
					{
						// nothing to enter and nothing popped: the ray has ended, or its stack continues in global memory
						const bool closest_done = (QUERY == kClosest) && closestDone(r);
						if(closest_done || floor == 0) finished = true;
						else need_replay = true;
					}
				}
			}
			else
			{
#pragma unroll kUnroll
			for(int step = 0; step < kSteps; ++step)
			{
#if B200RT_BREAK
				if(!go) break;
#else
				if(go)
#endif
				{
					// One node per step, interior or leaf, without divergent branches: the interior arithmetic, the push and
					// the pop of an empty leaf are all predicated / selected.
					const uint2 nd = __ldg(&s.nodes[r.node]);
					const uint32_t axis = nd.y & 3u;
					const uint32_t payload = nd.y >> 2; // interior: right child, leaf: primitive count
					const bool is_leaf = (axis == 3u);
					const float split = __uint_as_float(nd.x);
#if B200RT_AXIS_MAD
					uint32_t row_addr;
					asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(row_addr) : "r"(axis), "r"(uint32_t(kBlock * sizeof(float2))), "r"(axis_base));
					float2 oi;
					asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(oi.x), "=f"(oi.y) : "r"(row_addr));
#else
					const float2 oi = sh_axis[axis][tid]; // row 3 (read at leaves) holds the tree interval: the value is not used there
#endif
					const float o = oi.x, inv = oi.y;
					const float t_plane = (split - o) * inv;
					// near / far child without a predicate: m = all ones for a negative direction component
					const uint32_t m = uint32_t(__float_as_int(inv) >> 31);
					const uint32_t left = r.node + 1u;
					const uint32_t swap = (left ^ payload) & m;
					const uint32_t near = left ^ swap, far = payload ^ swap;
					const float limit = (QUERY == kClosest) ? fminf(r.seg_hi, B200RT_TDONE ? r.t_done : r.t_max) : r.seg_hi;
					const bool far_only = t_plane < r.seg_lo;
					const bool both = !is_leaf && !(t_plane > limit) && !far_only;
					// top of the ring (read before this step's speculative store; different slot)
#if B200RT_RING_OR
					uint2 popped;
					asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(popped.x), "=r"(popped.y) : "r"(ring_addr | (uint32_t(r.sp - kRingStride) & uint32_t(kRingMask))));
#else
					const uint2 popped = *reinterpret_cast<const uint2 *>(ring + ((r.sp - kRingStride) & kRingMask));
#endif
					const uint32_t pop_node = popped.x;
					const float pop_far = __uint_as_float(popped.y);
					if(!is_leaf)
					{
#if B200RT_RING_OR
						asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(ring_addr | (uint32_t(r.sp) & uint32_t(kRingMask))), "r"(far), "r"(__float_as_uint(r.seg_hi)) : "memory");
#else
						*reinterpret_cast<uint2 *>(ring + (r.sp & kRingMask)) = make_uint2(far, __float_as_uint(r.seg_hi));
#endif
						floor = max(floor, r.sp + (1 - kShortStack) * kRingStride); // the store has clobbered the oldest slot of a full ring, pushed or not
					}
					// closest: once the best hit is not beyond the end of this leaf nothing nearer can follow (accelerator_kdtree_common.h:232)
					const bool closest_done = (QUERY == kClosest) && closestDone(r);
					const bool do_pop = nd.y == 3u && !closest_done && r.sp > floor; // an empty leaf (count 0, axis bits 3) with something to pop
					leaf_count = payload;
					leaf_first = nd.x;
					r.node = is_leaf ? selectu(do_pop, pop_node, r.node) : selectu(far_only, far, near);
					r.seg_lo = selectf(do_pop, r.seg_hi, r.seg_lo);
					r.seg_hi = selectf(do_pop, pop_far, selectf(both, t_plane, r.seg_hi));
					if(both) r.sp += kRingStride;
					if(do_pop) r.sp -= kRingStride;
#if B200RT_FAR_PREFETCH
					if(both) prefetchL1(&s.nodes[far]);
#endif
					go = !is_leaf || do_pop;
				}
			}
			if(descending && !go)
			{
				// stopped at a leaf (leaf_count = its primitive count); node, interval and ring are as the step found them
				if(leaf_count != 0u)
				{
					pending = true;
#if B200RT_LEAF_PREFETCH == 1
					prefetchL1(s.tris + leaf_first);
					prefetchL1(s.tris + leaf_first + 2); // a 48-byte record may straddle two sectors
#elif B200RT_LEAF_PREFETCH == 2
					cpAsync16(&sh_leaf[0][tid], s.tris + leaf_first);
					cpAsync16(&sh_leaf[1][tid], s.tris + leaf_first + 1);
					cpAsync16(&sh_leaf[2][tid], s.tris + leaf_first + 2);
					cpAsyncCommit();
#endif
				}
				else
				{
					const bool closest_done = (QUERY == kClosest) && closestDone(r);
					if(closest_done || floor == 0) finished = true;
					else need_replay = true; // ring entries were overwritten: exact kd-restart behind the leaf just left
				}
			}
			} // !TREELET
		}

		if(need_replay)
		{
			need_replay = false;
			if constexpr(TREELET) finished = globalPop();
			else finished = replayTo(r.node);
		}
		if(finished)
		{
			writeResult<QUERY>(out, r, (QUERY == kClosest) ? (r.best_prim != B200RT_MISS) : hit, ts, tsCapacity(max_depth));
			alive = false;
		}
	}
}


template <int QUERY, bool SPHERES, bool QUEUED = false, bool TREELET = false>
__global__ void __launch_bounds__(kBlock, kMinBlocks) traceKernel(const __grid_constant__ SceneView s, const b200rt_ray *__restrict__ rays, uint32_t n,
                                                                 typename OutType<QUERY>::type *__restrict__ out, uint32_t *__restrict__ cursor, int max_depth, bool tree_space,
                                                                 const float *__restrict__ times)
{
	__shared__ __align__(kShortStack * kBlock * 8) uint2 sh_stack[kShortStack][kBlock]; // x = node index, y = float bits of the far end of its interval; aligned to its size (RING_OR)
	__shared__ float2 sh_axis[4][kBlock];          // rows 0-2: (origin, inverse direction) of the lane's ray per axis; row 3: the interval inside the tree bound
#if B200RT_LEAF_PREFETCH == 2
	__shared__ float4 sh_leaf[3][kBlock];          // the first record of the leaf a lane has stopped at, staged by cp.async
#else
	float4 (*sh_leaf)[kBlock] = nullptr;
#endif
	__shared__ uint32_t sh_task[kBlock / 32][32];   // cooperative leaf phase: slot -> (owner lane, record number)
	traceWarps<QUERY, SPHERES, QUEUED, TREELET>(s, rays, n, out, cursor, max_depth, tree_space, sh_stack, sh_axis, (blockIdx.x * uint32_t(kBlock / 32) + (threadIdx.x >> 5)) * 32u, sh_leaf, times, sh_task);
}

// First pass of a two-pass batch.  One warp per 256-ray region, eight trips of 32 rays: the ray records of trip t + 1 are already
// on their way into the warp's other shared-memory buffer (one 1 KB bulk copy, completion on an mbarrier) while every lane sets
// up its ray of trip t, fully converged.  Rays that miss the tree bound are answered here; the others are appended, in batch
// order, to the region's entries of the queue (see kEntryFloats); the region's count goes to the array behind the entries.
static constexpr int kSetupBlock = 256;
#ifndef B200RT_SETUP_STAGES
#define B200RT_SETUP_STAGES 2
#endif
static constexpr int kSetupStages = B200RT_SETUP_STAGES; // 1 KB buffers per warp: bulk copies in flight ahead of the trip being set up
static_assert((kSetupStages & (kSetupStages - 1)) == 0 && kSetupStages >= 2 && kSetupStages <= 8, "stages: a power of two, 2..8");
template <int QUERY>
__global__ void __launch_bounds__(kSetupBlock) setupKernel(const __grid_constant__ SceneView s, const b200rt_ray *__restrict__ rays, uint32_t n,
                                                           typename OutType<QUERY>::type *__restrict__ out, float *__restrict__ queue, bool tree_space, const float *__restrict__ times, int max_depth)
{
	__shared__ __align__(128) float4 sh_rays[kSetupBlock / 32][kSetupStages][64]; // per warp a ring of buffers of 32 ray records
	__shared__ __align__(128) float4 sh_out[kSetupBlock / 32][32 * (kEntryFloats / 4)]; // per warp the entries of one trip, compacted, before their bulk store
	__shared__ uint64_t sh_mbar[kSetupBlock / 32][kSetupStages];
	const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
	const uint32_t n_regions = (n + uint32_t(kRegionRays) - 1u) / uint32_t(kRegionRays);
	uint32_t *const counts = reinterpret_cast<uint32_t *>(queue + size_t(n_regions) * kRegionFloats);
	const uint32_t mbar0 = uint32_t(__cvta_generic_to_shared(&sh_mbar[w][0]));
	if(lane == 0u)
		for(int k = 0; k < kSetupStages; ++k) mbarInit(mbar0 + 8u * k, 1u);
	__syncwarp();
	uint32_t parity = 0u; // bit b: phase of buffer b's mbarrier
	const uint32_t warps = gridDim.x * uint32_t(kSetupBlock / 32);
	for(uint32_t region = blockIdx.x * uint32_t(kSetupBlock / 32) + w; region < n_regions; region += warps)
	{
		const uint32_t first = region * uint32_t(kRegionRays);
		const uint32_t n_rays = min(uint32_t(kRegionRays), n - first);
		const uint32_t trips = (n_rays + 31u) / 32u;
		float4 *const entries = reinterpret_cast<float4 *>(queue + size_t(region) * kRegionFloats);
		auto fetch = [&](uint32_t trip) {
			const uint32_t buf = trip & uint32_t(kSetupStages - 1);
			bulkCopyToShared(&sh_rays[w][buf][0], rays + first + trip * 32u, min(32u, n_rays - trip * 32u) * uint32_t(sizeof(b200rt_ray)), mbar0 + 8u * buf);
		};
		__syncwarp(); // every buffer has been read (previous region)
		if(lane == 0u)
			for(uint32_t t = 0; t < uint32_t(kSetupStages - 1) && t < trips; ++t) fetch(t);
		uint32_t count = 0u; // warp-uniform: entries written so far
		for(uint32_t trip = 0; trip < trips; ++trip)
		{
			const uint32_t buf = trip & uint32_t(kSetupStages - 1);
			if(trip + uint32_t(kSetupStages - 1) < trips)
			{
				__syncwarp(); // the buffer being refilled was read in the previous trip
				if(lane == 0u) fetch(trip + uint32_t(kSetupStages - 1));
			}
			while(!mbarTryWait(mbar0 + 8u * buf, (parity >> buf) & 1u)) {}
			parity ^= 1u << buf;
			const uint32_t k = trip * 32u + lane;
			bool ready = false;
			RayState q;
			if(k < n_rays)
			{
				q.index = first + k;
				const float4 a = sh_rays[w][buf][2 * lane], b = sh_rays[w][buf][2 * lane + 1];
				ready = setupRay<QUERY>(s, a, b, q, tree_space);
				if(!ready)
				{
					TShadowState none;
					none.depth = 0;
					writeResult<QUERY>(out, q, false, none, tsCapacity(max_depth));
				}
			}
			const unsigned m_ready = __ballot_sync(kFullMask, ready);
			if(lane == 0u) bulkStoreWaitRead(); // the previous trip's bulk store has read the staging buffer
			__syncwarp();
			if(ready)
			{
				float4 *e = &sh_out[w][uint32_t(__popc(m_ready & ((1u << lane) - 1u))) * (kEntryFloats / 4)];
				e[0] = make_float4(q.ox, q.oy, q.oz, q.dx);
				e[1] = make_float4(q.dy, q.dz, q.ix, q.iy);
				e[2] = make_float4(q.iz, q.t_min, q.t_max, q.seg_lo);
				e[3] = make_float4(q.seg_hi, __uint_as_float(q.index), times ? __ldg(times + q.index) : 0.f, 0.f);
			}
			__syncwarp();
			const uint32_t n_ready = uint32_t(__popc(m_ready));
			if(lane == 0u && n_ready != 0u) bulkStoreFromShared(entries + size_t(count) * (kEntryFloats / 4), &sh_out[w][0], n_ready * uint32_t(kEntryFloats * 4));
			count += n_ready;
		}
		if(lane == 0u) counts[region] = count;
	}
	if(lane == 0u) bulkStoreWaitAll(); // the last bulk stores must have left before the block's shared memory goes away
}

// One launch for the closest, shadow and transparent-shadow rays of one flush of the renderer's ray queue
// (b200rt_trace_jobs): cursor-less, warp w of the grid takes 32 rays of whichever kind its index falls into.  Launches
// are what 16 render threads contend for in the driver, so three kinds in one launch matter more than code size.
struct MixedBatch
{
	const b200rt_ray *rays[3];
	void *out[3];
	uint32_t n[3];
	const float *times[3]; // ray times per kind, nullptr = 0
};

template <bool SPHERES>
__global__ void __launch_bounds__(kBlock, 4) traceMixedKernel(const __grid_constant__ SceneView s, MixedBatch b, int max_depth, bool tree_space)
{
	__shared__ __align__(kShortStack * kBlock * 8) uint2 sh_stack[kShortStack][kBlock];
	__shared__ float2 sh_axis[4][kBlock];
#if B200RT_LEAF_PREFETCH == 2
	__shared__ float4 sh_leaf[3][kBlock];
#else
	float4 (*sh_leaf)[kBlock] = nullptr;
#endif
	__shared__ uint32_t sh_task[kBlock / 32][32];
	const uint32_t warp = blockIdx.x * uint32_t(kBlock / 32) + (threadIdx.x >> 5);
	const uint32_t w0 = (b.n[0] + 31u) / 32u, w1 = (b.n[1] + 31u) / 32u;
	if(warp < w0) traceWarps<kClosest, SPHERES>(s, b.rays[0], b.n[0], static_cast<b200rt_hit *>(b.out[0]), nullptr, 0, tree_space, sh_stack, sh_axis, warp * 32u, sh_leaf, b.times[0], sh_task);
	else if(warp < w0 + w1) traceWarps<kShadow, SPHERES>(s, b.rays[1], b.n[1], static_cast<uint32_t *>(b.out[1]), nullptr, 0, tree_space, sh_stack, sh_axis, (warp - w0) * 32u, sh_leaf, b.times[1], sh_task);
	else traceWarps<kTShadow, SPHERES>(s, b.rays[2], b.n[2], static_cast<b200rt_tshadow *>(b.out[2]), nullptr, max_depth, tree_space, sh_stack, sh_axis, (warp - w0 - w1) * 32u, sh_leaf, b.times[2], sh_task);
}

} // namespace b200rt
#endif
