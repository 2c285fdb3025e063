/* oracle/kd_oracle.c -- TEST INFRASTRUCTURE ONLY (see kd_oracle.h for the rules and the parity status).
 *
 * Plain-C restatement of the reference's ray queries.  Compiled with -ffp-contract=off so that, like
 * the reference's Release build (x86-64, ISO C++17, no -march => no FMA), every float operation is
 * rounded separately.  Paths cited below are relative to /root/reference.
 */
#include "kd_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#define KDO_MAX_STACK 64       /* include/accelerator/accelerator_kdtree_common.h:32 */
#define KDO_MIN_RAYDIST 0.00005f /* include/accelerator/accelerator.h:84 */

enum { F_VISIBLE = 1, F_SHADOW = 2, F_TRANSPARENT = 4 };
enum { Q_NEAREST = 0, Q_SHADOW = 1, Q_TSHADOW = 2 };

/* ------------------------------------------------------------------------------------------------
 * include/geometry/bound.h:156-198  Bound<float>::cross  (Smits slab test; X assigns, Y/Z fold;
 * a zero direction component skips the axis; early reject after every axis). */
int kdo_bound_cross(const float b[6], const float from[3], const float dir[3], float t_max, float *enter, float *leave)
{
	float lmin = -FLT_MAX, lmax = FLT_MAX;
	for(int axis = 0; axis < 3; ++axis)
	{
		if(dir[axis] != 0.f)
		{
			const float p = from[axis] - b[axis];
			const float inv_dir = 1.f / dir[axis];
			float ltmin, ltmax;
			if(inv_dir > 0.f)
			{
				ltmin = -p * inv_dir;
				ltmax = ((b[3 + axis] - b[axis]) - p) * inv_dir;
			}
			else
			{
				ltmin = ((b[3 + axis] - b[axis]) - p) * inv_dir;
				ltmax = -p * inv_dir;
			}
			if(axis == 0) { lmin = ltmin; lmax = ltmax; }
			else
			{
				lmin = (ltmin < lmin) ? lmin : ltmin; /* std::max(ltmin, lmin) */
				lmax = (lmax < ltmax) ? lmax : ltmax; /* std::min(ltmax, lmax) */
			}
			if((lmax < 0.f) || (lmin > t_max)) return 0;
		}
	}
	if((lmin <= lmax) && (lmax >= 0.f) && (lmin <= t_max))
	{
		*enter = lmin;
		*leave = lmax;
		return 1;
	}
	return 0;
}

/* include/geometry/vector.h:163-164 (dot, summed left to right) and :241-245 (cross). */
static inline float dot3(const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(const float a[3], const float b[3], float o[3])
{
	o[0] = a[1] * b[2] - a[2] * b[1];
	o[1] = a[2] * b[0] - a[0] * b[2];
	o[2] = a[0] * b[1] - a[1] * b[0];
}
static inline void sub3(const float a[3], const float b[3], float o[3]) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }

/* include/geometry/shape/shape_polygon.h:126-176  ShapePolygon<float,3|4>::intersect (Moeller-Trumbore;
 * the second triangle of a quad is tried only when the first fails its u range test). */
float kdo_poly_intersect(const float *v0, const float *v1, const float *v2, const float *v3, int nv,
                         const float from[3], const float dir[3], float *out_u, float *out_v)
{
	float edge_1[3], edge_2[3], pvec_2[3];
	sub3(v1, v0, edge_1);
	sub3(v2, v0, edge_2);
	cross3(dir, edge_2, pvec_2);
	const float det_1_2 = dot3(edge_1, pvec_2);
	if(det_1_2 != 0.f)
	{
		const float inv_det_1_2 = 1.f / det_1_2;
		float tvec[3];
		sub3(from, v0, tvec);
		float u = dot3(tvec, pvec_2) * inv_det_1_2;
		if(u >= 0.f && u <= 1.f)
		{
			float qvec_1[3];
			cross3(tvec, edge_1, qvec_1);
			const float v = dot3(dir, qvec_1) * inv_det_1_2;
			if(v >= 0.f && (u + v) <= 1.f)
			{
				const float t = dot3(edge_2, qvec_1) * inv_det_1_2;
				if(t > 0.f)
				{
					if(nv == 3) { *out_u = u; *out_v = v; }
					else { *out_u = u + v; *out_v = v; }
					return t;
				}
			}
		}
		else if(nv == 4)
		{
			float edge_3[3], pvec_3[3];
			sub3(v3, v0, edge_3);
			cross3(dir, edge_3, pvec_3);
			const float det_2_3 = dot3(edge_2, pvec_3);
			if(det_2_3 != 0.f)
			{
				const float inv_det_2_3 = 1.f / det_2_3;
				u = dot3(tvec, pvec_3) * inv_det_2_3;
				if(u >= 0.f && u <= 1.f)
				{
					float qvec_2[3];
					cross3(tvec, edge_2, qvec_2);
					const float v = dot3(dir, qvec_2) * inv_det_2_3;
					if(v >= 0.f && (u + v) <= 1.f)
					{
						const float t = dot3(edge_3, qvec_2) * inv_det_2_3;
						if(t > 0.f) { *out_u = u; *out_v = u + v; return t; }
					}
				}
			}
		}
	}
	*out_u = 0.f;
	*out_v = 0.f;
	return 0.f;
}

/* SpherePrimitive::intersect, src/geometry/primitive/primitive_sphere.cc:83-102 (math::sqrt is std::sqrt on x86-64,
 * include/math/math.h:148-173,214-221).  Returns t (0 = miss); uv is {0, 0} (`return {sol, {}}`). */
float kdo_sphere_intersect(const float center[3], float radius, const float from[3], const float dir[3])
{
	float vf[3];
	sub3(from, center, vf);
	const float ea = dot3(dir, dir);
	const float eb = 2.f * dot3(vf, dir);
	const float ec = dot3(vf, vf) - radius * radius;
	float osc = eb * eb - 4.f * ea * ec;
	if(osc < 0) return 0.f;
	osc = sqrtf(osc);
	const float sol_1 = (-eb - osc) / (2.f * ea);
	const float sol_2 = (-eb + osc) / (2.f * ea);
	float sol = sol_1;
	if(sol < 0.f)
	{
		sol = sol_2;
		if(sol < 0.f) return 0.f;
	}
	return sol;
}

/* math::bezierCalculateFactors of the ray time mapped into [start, end] by math::lerpSegment(time, 0, start, 1, end)
 * (include/math/interpolation.h:50-93).  Returns 0 / 2 when the time lies at or outside the start / end of the range (time step 0 /
 * 2 is then used as is), 1 when f[] holds the three factors. */
static inline int bezierAtTime(float time, float start, float end, float f[3])
{
	if(time <= start) return 0;
	if(time >= end) return 2;
	/* lerpSegment: x == x_1 and x == x_2 were handled above; x_1 == x_2 cannot be (start < time < end) */
	const float diff_x = end - start, diff_a = time - start;
	const float x = 0.f + ((diff_a / diff_x) * (1.f - 0.f));
	const float xr = 1.f - x;
	f[0] = xr * xr;
	f[1] = 2.f * x * xr;
	f[2] = x * x;
	return 1;
}

/* SquareMatrix * Point, include/geometry/matrix.h:131-144: aux = 0; aux += m[i][j] * v[j] (j = 0..2); aux += m[i][3]. */
static inline void matPoint(const float *m, const float v[3], float o[3])
{
	for(int i = 0; i < 3; ++i)
	{
		float aux = 0.f;
		for(int j = 0; j < 3; ++j) aux += m[4 * i + j] * v[j];
		aux += m[4 * i + 3];
		o[i] = aux;
	}
}

/* Vertices of face `prim` at ray time `time` into vtx[4][3]; returns the vertex count. */
static inline int faceVerticesAtTime(const kdo_mesh *m, uint32_t prim, float time, float vtx[4][3])
{
	const uint32_t *i = m->idx + 4 * (size_t) prim;
	const int nv = (i[3] == 0xFFFFFFFFu) ? 3 : 4;
	const kdo_motion *mo = m->motion;
	const int kind = (mo && mo->kind) ? mo->kind[prim] : 0;
	if(kind == 1)
	{
		float f[3];
		const int where = bezierAtTime(time, mo->face_times[2 * (size_t) prim], mo->face_times[2 * (size_t) prim + 1], f);
		for(int k = 0; k < nv; ++k)
		{
			const float *p0 = m->xyz + 3 * (size_t) i[k], *p1 = mo->xyz1 + 3 * (size_t) i[k], *p2 = mo->xyz2 + 3 * (size_t) i[k];
			for(int a = 0; a < 3; ++a)
				/* bezierInterpolate: y[0] * f[0] + y[1] * f[1] + y[2] * f[2], left to right (interpolation.h:83-86) */
				vtx[k][a] = (where == 0) ? p0[a] : (where == 2) ? p2[a] : (f[0] * p0[a] + f[1] * p1[a]) + f[2] * p2[a];
		}
	}
	else if(kind == 2)
	{
		const float *mats = mo->matrices + 48 * (size_t) mo->face_matrix[prim];
		float f[3], mt[16];
		const int where = bezierAtTime(time, mo->face_times[2 * (size_t) prim], mo->face_times[2 * (size_t) prim + 1], f);
		const float *use = (where == 0) ? mats : (where == 2) ? mats + 32 : mt;
		if(where == 1)
			for(int e = 0; e < 16; ++e) mt[e] = (f[0] * mats[e] + f[1] * mats[16 + e]) + f[2] * mats[32 + e]; /* matrix.h:96-113 */
		for(int k = 0; k < nv; ++k) matPoint(use, m->xyz + 3 * (size_t) i[k], vtx[k]);
	}
	else
		for(int k = 0; k < nv; ++k)
			for(int a = 0; a < 3; ++a) vtx[k][a] = m->xyz[3 * (size_t) i[k] + a];
	return nv;
}

static inline float primIntersect(const kdo_mesh *m, uint32_t prim, const float from[3], const float dir[3], float time, float *u, float *v)
{
	const uint32_t *i = m->idx + 4 * (size_t) prim;
	if(i[2] == KDO_SPHERE)
	{
		*u = 0.f;
		*v = 0.f;
		return kdo_sphere_intersect(m->xyz + 3 * (size_t) i[0], m->xyz[3 * (size_t) i[1]], from, dir);
	}
	if(m->motion && m->motion->kind && m->motion->kind[prim])
	{
		float vtx[4][3];
		const int nv = faceVerticesAtTime(m, prim, time, vtx);
		return kdo_poly_intersect(vtx[0], vtx[1], vtx[2], nv == 4 ? vtx[3] : NULL, nv, from, dir, u, v);
	}
	const int nv = (i[3] == 0xFFFFFFFFu) ? 3 : 4;
	return kdo_poly_intersect(m->xyz + 3 * (size_t) i[0], m->xyz + 3 * (size_t) i[1], m->xyz + 3 * (size_t) i[2],
	                          nv == 4 ? m->xyz + 3 * (size_t) i[3] : NULL, nv, from, dir, u, v);
}

/* include/accelerator/intersect_data.h:30-39 */
typedef struct
{
	float t_hit, u, v, t_max;
	int32_t prim;
} idata;

/* state of the std::set + depth counter of the transparent-shadow query (accelerator_kdtree_common.h:114-115) */
typedef struct
{
	int depth, max_depth;
	int n_filtered, cap;
	uint32_t *filtered;
	uint32_t small[64];
} tstate;

static int tstateInsert(tstate *ts, uint32_t prim)
{
	for(int i = 0; i < ts->n_filtered; ++i) if(ts->filtered[i] == prim) return 0;
	if(ts->n_filtered == ts->cap)
	{
		const int ncap = ts->cap * 2;
		uint32_t *nf = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) ncap);
		memcpy(nf, ts->filtered, sizeof(uint32_t) * (size_t) ts->n_filtered);
		if(ts->filtered != ts->small) free(ts->filtered);
		ts->filtered = nf;
		ts->cap = ncap;
	}
	ts->filtered[ts->n_filtered++] = prim;
	return 1;
}

/* include/accelerator/accelerator.h:122-169: the three accept rules.  Return 1 = "stop traversal". */
static inline int primitiveIntersection(int query, idata *d, tstate *ts, const kdo_mesh *m, uint32_t prim,
                                        const float from[3], const float dir[3], float t_min, float t_max, float time)
{
	float u, v;
	const float t_hit = primIntersect(m, prim, from, dir, time, &u, &v);
	if(t_hit <= 0.f || t_hit < t_min || t_hit >= t_max) return 0;
	const uint8_t fl = m->flags ? m->flags[prim] : (uint8_t) (F_VISIBLE | F_SHADOW);
	if(query == Q_NEAREST) { if(!(fl & F_VISIBLE)) return 0; }
	else if(!(fl & F_SHADOW)) return 0;
	d->t_hit = t_hit;
	d->t_max = t_hit;
	d->u = u;
	d->v = v;
	d->prim = (int32_t) prim;
	if(query == Q_NEAREST) return 0;
	if(query == Q_SHADOW) return 1;
	/* transparent shadow, accelerator.h:159-168 */
	if(!(fl & F_TRANSPARENT)) return 1;
	if(tstateInsert(ts, prim))
	{
		if(ts->depth >= ts->max_depth) return 1;
		++ts->depth; /* colour *= transparency happens here in the reference; host-side material code */
	}
	return 0;
}

typedef struct
{
	float t;
	float point[3];
	int64_t node; /* -1 = "nowhere" */
	int prev;
} kdstack;

/* include/accelerator/accelerator_kdtree_common.h:107-255  kdtree::intersect<Node,Stack,TestType>.
 * Returns 1 when the query reports a hit (IntersectData::isHit(), t_hit > 0 after the query's own
 * post-processing), fills d. */
static int kdIntersect(int query, const kdo_mesh *m, const kdo_tree *tree, const float from[3], const float dir[3],
                       float ray_tmin, float t_max, float time, idata *d, tstate *ts, kdo_counters *cnt)
{
	d->t_hit = 0.f; d->u = 0.f; d->v = 0.f; d->t_max = 0.f; d->prim = -1;
	float enter, leave;
	if(!kdo_bound_cross(tree->bound, from, dir, t_max, &enter, &leave)) return 0;
	if(tree->n_nodes == 0) return 0;
	/* math::inverse, include/math/math.h:71-77 */
	float inv_dir[3];
	for(int a = 0; a < 3; ++a) inv_dir[a] = (dir[a] == 0.f) ? FLT_MAX : 1.f / dir[a];
	kdstack stack[KDO_MAX_STACK];
	int64_t curr = 0, far_child;
	int entry_id = 0;
	stack[entry_id].t = enter;
	if(enter >= 0.f) for(int a = 0; a < 3; ++a) stack[entry_id].point[a] = from[a] + dir[a] * enter;
	else for(int a = 0; a < 3; ++a) stack[entry_id].point[a] = from[a];
	int exit_id = 1;
	stack[exit_id].t = leave;
	for(int a = 0; a < 3; ++a) stack[exit_id].point[a] = from[a] + dir[a] * leave;
	stack[exit_id].node = -1;
	stack[exit_id].prev = 0; /* uninitialised in the reference; never read before the loop ends */
	d->t_max = t_max;
	/* accelerator.h:64 calculateDynamicRayBias; accelerator_kdtree_common.h:139 */
	const float bias = 0.1f * KDO_MIN_RAYDIST * fabsf(leave - enter);
	const float t_min = (query == Q_SHADOW) ? bias : (ray_tmin < bias ? bias : ray_tmin); /* std::max(ray.tmin_, bias) */

	while(curr >= 0 && stack[entry_id].t <= t_max)
	{
		while((tree->flags[curr] & 3u) != 3u)
		{
			if(cnt) ++cnt->interior;
			const int axis = (int) (tree->flags[curr] & 3u);
			const float split_val = tree->split[curr];
			if(stack[entry_id].point[axis] <= split_val)
			{
				if(stack[exit_id].point[axis] <= split_val) { ++curr; continue; }
				far_child = (int64_t) (tree->flags[curr] >> 2);
				++curr;
			}
			else
			{
				if(stack[exit_id].point[axis] > split_val) { curr = (int64_t) (tree->flags[curr] >> 2); continue; }
				far_child = curr + 1;
				curr = (int64_t) (tree->flags[curr] >> 2);
			}
			const float t = (split_val - from[axis]) * inv_dir[axis];
			const int exit_prev = exit_id;
			++exit_id;
			if(exit_id == entry_id) ++exit_id;
			const int next_axis = (axis + 1) % 3, prev_axis = (axis + 2) % 3;
			stack[exit_id].prev = exit_prev;
			stack[exit_id].t = t;
			stack[exit_id].node = far_child;
			stack[exit_id].point[axis] = split_val;
			stack[exit_id].point[next_axis] = from[next_axis] + t * dir[next_axis];
			stack[exit_id].point[prev_axis] = from[prev_axis] + t * dir[prev_axis];
		}
		const uint32_t n_prims = tree->flags[curr] >> 2;
		const uint32_t *refs = tree->refs + tree->first_ref[curr];
		if(cnt) { ++cnt->leaves; cnt->refs += n_prims; cnt->tests += n_prims; }
		for(uint32_t i = 0; i < n_prims; ++i)
		{
			const float tm = (query == Q_NEAREST) ? d->t_max : t_max;
			if(primitiveIntersection(query, d, ts, m, refs[i], from, dir, t_min, tm, time)) return 1;
		}
		if(query == Q_NEAREST && d->t_hit > 0.f && d->t_max <= stack[exit_id].t) return 1;
		entry_id = exit_id;
		curr = stack[exit_id].node;
		exit_id = stack[entry_id].prev;
	}
	if(query == Q_NEAREST) return d->t_hit > 0.f;
	d->t_hit = 0.f; /* setNoHit() / default IntersectData */
	return 0;
}

static inline float wrapperTmaxShadow(const float *r, float sfrom[3])
{
	/* accelerator.h:105-108: origin moved by dir*tmin, t_max = tmax - 2*tmin */
	for(int a = 0; a < 3; ++a) sfrom[a] = r[a] + r[4 + a] * r[3];
	return (r[7] >= 0.f) ? r[7] - 2 * r[3] : FLT_MAX;
}


/* ---- a minimal pthread parallel-for (libgomp is not in this image) ---- */
typedef struct job
{
	int query; /* Q_* or 3 = brute force */
	const kdo_mesh *mesh;
	const kdo_tree *tree;
	const float *bound6;
	const float *rays;
	const float *times; /* NULL = time 0 for every ray */
	size_t n;
	float *out_t, *out_u, *out_v;
	int32_t *out_prim;
	uint8_t *out_shadowed;
	int32_t *out_n_transparent, *out_list;
	int max_depth, max_list;
	int want_counters;
	atomic_size_t next;
	size_t chunk;
	pthread_mutex_t lock;
	kdo_counters total;
} job;

static void closestOne(job *j, size_t i, kdo_counters *local)
{
	const float *r = j->rays + 8 * i;
	const float t_max = (r[7] >= 0.f) ? r[7] : FLT_MAX; /* accelerator.h:91 */
	idata d;
	const int hit = kdIntersect(Q_NEAREST, j->mesh, j->tree, r, r + 4, r[3], t_max, j->times ? j->times[i] : 0.f, &d, NULL, j->want_counters ? local : NULL);
	if(hit && d.prim >= 0) { j->out_t[i] = d.t_max; j->out_u[i] = d.u; j->out_v[i] = d.v; j->out_prim[i] = d.prim; }
	else { j->out_t[i] = 0.f; j->out_u[i] = 0.f; j->out_v[i] = 0.f; j->out_prim[i] = -1; }
}

static void shadowOne(job *j, size_t i, kdo_counters *local)
{
	const float *r = j->rays + 8 * i;
	float sfrom[3];
	const float t_max = wrapperTmaxShadow(r, sfrom);
	idata d;
	const int hit = kdIntersect(Q_SHADOW, j->mesh, j->tree, sfrom, r + 4, r[3], t_max, j->times ? j->times[i] : 0.f, &d, NULL, j->want_counters ? local : NULL);
	j->out_shadowed[i] = (uint8_t) (hit ? 1 : 0);
	if(j->out_prim) j->out_prim[i] = hit ? d.prim : -1;
}

static void tshadowOne(job *j, size_t i)
{
	const float *r = j->rays + 8 * i;
	float sfrom[3];
	const float t_max = wrapperTmaxShadow(r, sfrom);
	idata d;
	tstate ts;
	ts.depth = 0; ts.max_depth = j->max_depth; ts.n_filtered = 0; ts.cap = 64; ts.filtered = ts.small;
	const int hit = kdIntersect(Q_TSHADOW, j->mesh, j->tree, sfrom, r + 4, r[3], t_max, j->times ? j->times[i] : 0.f, &d, &ts, NULL);
	j->out_shadowed[i] = (uint8_t) (hit ? 1 : 0);
	if(j->out_n_transparent) j->out_n_transparent[i] = ts.depth;
	if(j->out_list)
		for(int k = 0; k < j->max_list; ++k) j->out_list[i * (size_t) j->max_list + (size_t) k] = (k < ts.n_filtered) ? (int32_t) ts.filtered[k] : -1;
	if(ts.filtered != ts.small) free(ts.filtered);
}

static void bruteOne(job *j, size_t i)
{
	const float *r = j->rays + 8 * i;
	const float t_max = (r[7] >= 0.f) ? r[7] : FLT_MAX;
	j->out_t[i] = 0.f; j->out_u[i] = 0.f; j->out_v[i] = 0.f; j->out_prim[i] = -1;
	float enter, leave;
	if(!kdo_bound_cross(j->bound6, r, r + 4, t_max, &enter, &leave)) return;
	const float bias = 0.1f * KDO_MIN_RAYDIST * fabsf(leave - enter);
	const float t_min = (r[3] < bias) ? bias : r[3];
	idata d;
	d.t_hit = 0.f; d.t_max = t_max; d.prim = -1; d.u = d.v = 0.f;
	for(size_t p = 0; p < j->mesh->n_faces; ++p) primitiveIntersection(Q_NEAREST, &d, NULL, j->mesh, (uint32_t) p, r, r + 4, t_min, d.t_max, j->times ? j->times[i] : 0.f);
	if(d.t_hit > 0.f) { j->out_t[i] = d.t_max; j->out_u[i] = d.u; j->out_v[i] = d.v; j->out_prim[i] = d.prim; }
}

static void *worker(void *arg)
{
	job *j = (job *) arg;
	kdo_counters local = {0, 0, 0, 0, 0};
	for(;;)
	{
		const size_t b = atomic_fetch_add(&j->next, j->chunk);
		if(b >= j->n) break;
		const size_t e = (b + j->chunk < j->n) ? b + j->chunk : j->n;
		for(size_t i = b; i < e; ++i)
		{
			switch(j->query)
			{
				case Q_NEAREST: closestOne(j, i, &local); break;
				case Q_SHADOW: shadowOne(j, i, &local); break;
				case Q_TSHADOW: tshadowOne(j, i); break;
				default: bruteOne(j, i); break;
			}
		}
		local.rays += e - b;
	}
	pthread_mutex_lock(&j->lock);
	j->total.rays += local.rays; j->total.interior += local.interior; j->total.leaves += local.leaves;
	j->total.refs += local.refs; j->total.tests += local.tests;
	pthread_mutex_unlock(&j->lock);
	return NULL;
}

static void runJob(job *j, int n_threads, kdo_counters *counters)
{
	if(n_threads < 1) n_threads = 1;
	if(n_threads > 256) n_threads = 256;
	atomic_init(&j->next, 0);
	j->chunk = (j->query == 3) ? 64 : 4096;
	j->want_counters = counters != NULL;
	memset(&j->total, 0, sizeof(j->total));
	pthread_mutex_init(&j->lock, NULL);
	if(n_threads == 1 || j->n <= j->chunk) worker(j);
	else
	{
		pthread_t th[256];
		for(int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, worker, j);
		for(int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
	}
	pthread_mutex_destroy(&j->lock);
	if(counters) *counters = j->total;
}

void kdo_trace_closest(const kdo_mesh *mesh, const kdo_tree *tree, const float *rays, size_t n,
                       float *out_t, float *out_u, float *out_v, int32_t *out_prim, int n_threads, kdo_counters *counters)
{
	job j;
	memset(&j, 0, sizeof(j));
	j.query = Q_NEAREST; j.mesh = mesh; j.tree = tree; j.rays = rays; j.n = n;
	j.out_t = out_t; j.out_u = out_u; j.out_v = out_v; j.out_prim = out_prim;
	runJob(&j, n_threads, counters);
}

void kdo_trace_shadow(const kdo_mesh *mesh, const kdo_tree *tree, const float *rays, size_t n,
                      uint8_t *out_shadowed, int32_t *out_prim, int n_threads, kdo_counters *counters)
{
	job j;
	memset(&j, 0, sizeof(j));
	j.query = Q_SHADOW; j.mesh = mesh; j.tree = tree; j.rays = rays; j.n = n;
	j.out_shadowed = out_shadowed; j.out_prim = out_prim;
	runJob(&j, n_threads, counters);
}

void kdo_trace_tshadow(const kdo_mesh *mesh, const kdo_tree *tree, const float *rays, size_t n, int max_depth,
                       uint8_t *out_shadowed, int32_t *out_n_transparent, int32_t *out_list, int max_list, int n_threads)
{
	job j;
	memset(&j, 0, sizeof(j));
	j.query = Q_TSHADOW; j.mesh = mesh; j.tree = tree; j.rays = rays; j.n = n; j.max_depth = max_depth;
	j.out_shadowed = out_shadowed; j.out_n_transparent = out_n_transparent; j.out_list = out_list; j.max_list = max_list;
	runJob(&j, n_threads, NULL);
}

void kdo_brute_closest(const kdo_mesh *mesh, const float bound6[6], const float *rays, size_t n,
                       float *out_t, float *out_u, float *out_v, int32_t *out_prim, int n_threads)
{
	job j;
	memset(&j, 0, sizeof(j));
	j.query = 3; j.mesh = mesh; j.bound6 = bound6; j.rays = rays; j.n = n;
	j.out_t = out_t; j.out_u = out_u; j.out_v = out_v; j.out_prim = out_prim;
	runJob(&j, n_threads, NULL);
}

void kdo_trace_closest_timed(const kdo_mesh *mesh, const kdo_tree *tree, const float *rays, const float *times, size_t n,
                             float *out_t, float *out_u, float *out_v, int32_t *out_prim, int n_threads)
{
	job j;
	memset(&j, 0, sizeof(j));
	j.query = Q_NEAREST; j.mesh = mesh; j.tree = tree; j.rays = rays; j.times = times; j.n = n;
	j.out_t = out_t; j.out_u = out_u; j.out_v = out_v; j.out_prim = out_prim;
	runJob(&j, n_threads, NULL);
}

void kdo_trace_shadow_timed(const kdo_mesh *mesh, const kdo_tree *tree, const float *rays, const float *times, size_t n,
                            uint8_t *out_shadowed, int32_t *out_prim, int n_threads)
{
	job j;
	memset(&j, 0, sizeof(j));
	j.query = Q_SHADOW; j.mesh = mesh; j.tree = tree; j.rays = rays; j.times = times; j.n = n;
	j.out_shadowed = out_shadowed; j.out_prim = out_prim;
	runJob(&j, n_threads, NULL);
}

void kdo_trace_tshadow_timed(const kdo_mesh *mesh, const kdo_tree *tree, const float *rays, const float *times, size_t n, int max_depth,
                             uint8_t *out_shadowed, int32_t *out_n_transparent, int32_t *out_list, int max_list, int n_threads)
{
	job j;
	memset(&j, 0, sizeof(j));
	j.query = Q_TSHADOW; j.mesh = mesh; j.tree = tree; j.rays = rays; j.times = times; j.n = n; j.max_depth = max_depth;
	j.out_shadowed = out_shadowed; j.out_n_transparent = out_n_transparent; j.out_list = out_list; j.max_list = max_list;
	runJob(&j, n_threads, NULL);
}

void kdo_brute_closest_timed(const kdo_mesh *mesh, const float bound6[6], const float *rays, const float *times, size_t n,
                             float *out_t, float *out_u, float *out_v, int32_t *out_prim, int n_threads)
{
	job j;
	memset(&j, 0, sizeof(j));
	j.query = 3; j.mesh = mesh; j.bound6 = bound6; j.rays = rays; j.times = times; j.n = n;
	j.out_t = out_t; j.out_u = out_u; j.out_v = out_v; j.out_prim = out_prim;
	runJob(&j, n_threads, NULL);
}

/* ------------------------------------------------------------------------------------------------
 * Tree bound, src/accelerator/accelerator_kdtree_original.cc:88-103 (+ FacePrimitive::getBound =
 * min/max over the face's vertices). */
static void primBound(const kdo_mesh *m, size_t f, float lo[3], float hi[3])
{
	const uint32_t *i = m->idx + 4 * f;
	if(i[2] == KDO_SPHERE)
	{
		/* SpherePrimitive::getBound, src/geometry/primitive/primitive_sphere.cc:71-75: r = radius * 1.0001f, centre -+ r */
		const float r = m->xyz[3 * (size_t) i[1]] * 1.0001f;
		for(int a = 0; a < 3; ++a)
		{
			lo[a] = m->xyz[3 * (size_t) i[0] + a] - r;
			hi[a] = m->xyz[3 * (size_t) i[0] + a] + r;
		}
		return;
	}
	const int nv = (i[3] == 0xFFFFFFFFu) ? 3 : 4;
	const kdo_motion *mo = m->motion;
	const int kind = (mo && mo->kind) ? mo->kind[f] : 0;
	if(kind)
	{
		/* FacePrimitive::getBoundTimeSteps (primitive_face.h:155-170): min/max over the vertices of all three time steps;
		 * PrimitiveInstance::getBound (primitive_instance.h:119-128): the union of the base face's bound under each matrix */
		int first = 1;
		for(int step = 0; step < 3; ++step)
			for(int k = 0; k < nv; ++k)
			{
				float p[3];
				if(kind == 1)
				{
					const float *src = (step == 0 ? m->xyz : step == 1 ? mo->xyz1 : mo->xyz2) + 3 * (size_t) i[k];
					p[0] = src[0]; p[1] = src[1]; p[2] = src[2];
				}
				else matPoint(mo->matrices + 48 * (size_t) mo->face_matrix[f] + 16 * step, m->xyz + 3 * (size_t) i[k], p);
				for(int a = 0; a < 3; ++a)
				{
					if(first || p[a] < lo[a]) lo[a] = p[a];
					if(first || p[a] > hi[a]) hi[a] = p[a];
				}
				first = 0;
			}
		return;
	}
	for(int a = 0; a < 3; ++a) lo[a] = hi[a] = m->xyz[3 * (size_t) i[0] + a];
	for(int k = 1; k < nv; ++k)
		for(int a = 0; a < 3; ++a)
		{
			const float x = m->xyz[3 * (size_t) i[k] + a];
			if(x < lo[a]) lo[a] = x;
			if(x > hi[a]) hi[a] = x;
		}
}

void kdo_tree_bound(const kdo_mesh *mesh, float out6[6])
{
	for(int a = 0; a < 6; ++a) out6[a] = 0.f;
	for(size_t f = 0; f < mesh->n_faces; ++f)
	{
		float lo[3], hi[3];
		primBound(mesh, f, lo, hi);
		for(int a = 0; a < 3; ++a)
		{
			if(f == 0 || lo[a] < out6[a]) out6[a] = lo[a];
			if(f == 0 || hi[a] > out6[3 + a]) out6[3 + a] = hi[a];
		}
	}
	for(int a = 0; a < 3; ++a)
	{
		const double offset = (out6[3 + a] - out6[a]) * 0.001; /* float subtraction, double multiply */
		out6[a] -= (float) offset;
		out6[3 + a] += (float) offset;
	}
}

/* ------------------------------------------------------------------------------------------------
 * Stand-alone builder (own design, see header): binned SAH on primitive bounding boxes, DFS node order
 * with the reference's node encoding. */
typedef struct
{
	float *split;
	uint32_t *flags, *first_ref, *refs;
	size_t n_nodes, cap_nodes, n_refs, cap_refs;
	float bound[6];
	const float *plo, *phi; /* per primitive bounds */
	int max_leaf, max_depth;
} built;

static size_t newNode(built *b)
{
	if(b->n_nodes == b->cap_nodes)
	{
		b->cap_nodes = b->cap_nodes ? b->cap_nodes * 2 : 1024;
		b->split = (float *) realloc(b->split, b->cap_nodes * sizeof(float));
		b->flags = (uint32_t *) realloc(b->flags, b->cap_nodes * sizeof(uint32_t));
		b->first_ref = (uint32_t *) realloc(b->first_ref, b->cap_nodes * sizeof(uint32_t));
	}
	return b->n_nodes++;
}

static void makeLeaf(built *b, size_t node, const uint32_t *prims, size_t n)
{
	if(b->n_refs + n > b->cap_refs)
	{
		while(b->n_refs + n > b->cap_refs) b->cap_refs = b->cap_refs ? b->cap_refs * 2 : 4096;
		b->refs = (uint32_t *) realloc(b->refs, b->cap_refs * sizeof(uint32_t));
	}
	b->split[node] = 0.f;
	b->flags[node] = ((uint32_t) n << 2) | 3u;
	b->first_ref[node] = (uint32_t) b->n_refs;
	memcpy(b->refs + b->n_refs, prims, n * sizeof(uint32_t));
	b->n_refs += n;
}

#define KDO_BINS 32
static void buildRec(built *b, uint32_t *prims, size_t n, const float lo[3], const float hi[3], int depth)
{
	const size_t node = newNode(b);
	if(n <= (size_t) b->max_leaf || depth >= b->max_depth) { makeLeaf(b, node, prims, n); return; }
	const float ext[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
	const double area = 2.0 * ((double) ext[0] * ext[1] + (double) ext[1] * ext[2] + (double) ext[2] * ext[0]);
	double best_cost = 1.0 * (double) n; /* leaf cost, intersection cost 1, traversal cost 0.5 below */
	int best_axis = -1;
	float best_pos = 0.f;
	for(int axis = 0; axis < 3; ++axis)
	{
		if(!(ext[axis] > 0.f)) continue;
		size_t start[KDO_BINS + 1], end[KDO_BINS + 1];
		memset(start, 0, sizeof(start));
		memset(end, 0, sizeof(end));
		const double scale = KDO_BINS / (double) ext[axis];
		for(size_t i = 0; i < n; ++i)
		{
			const float pl = b->plo[3 * (size_t) prims[i] + axis], ph = b->phi[3 * (size_t) prims[i] + axis];
			int bl = (int) (((double) pl - lo[axis]) * scale), bh = (int) (((double) ph - lo[axis]) * scale);
			if(bl < 0) bl = 0;
			if(bl > KDO_BINS - 1) bl = KDO_BINS - 1;
			if(bh < 0) bh = 0;
			if(bh > KDO_BINS - 1) bh = KDO_BINS - 1;
			++start[bl];
			++end[bh];
		}
		const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
		size_t n_left = 0, n_right = n;
		for(int k = 1; k < KDO_BINS; ++k)
		{
			n_left += start[k - 1];
			n_right -= end[k - 1];
			const double w = (double) ext[axis] * k / KDO_BINS;
			const double al = 2.0 * ((double) ext[a1] * ext[a2] + w * ((double) ext[a1] + ext[a2]));
			const double ar = 2.0 * ((double) ext[a1] * ext[a2] + ((double) ext[axis] - w) * ((double) ext[a1] + ext[a2]));
			double cost = 0.5 + (al * (double) n_left + ar * (double) n_right) / area;
			if(n_left == 0 || n_right == 0) cost *= 0.8;
			if(cost < best_cost)
			{
				best_cost = cost;
				best_axis = axis;
				best_pos = lo[axis] + (float) w;
			}
		}
	}
	if(best_axis < 0 || !(best_pos > lo[best_axis]) || !(best_pos < hi[best_axis])) { makeLeaf(b, node, prims, n); return; }
	uint32_t *left = (uint32_t *) malloc(n * sizeof(uint32_t)), *right = (uint32_t *) malloc(n * sizeof(uint32_t));
	size_t nl = 0, nr = 0;
	for(size_t i = 0; i < n; ++i)
	{
		const float pl = b->plo[3 * (size_t) prims[i] + best_axis], ph = b->phi[3 * (size_t) prims[i] + best_axis];
		if(pl <= best_pos) left[nl++] = prims[i];  /* a primitive touching the plane goes to both sides */
		if(ph >= best_pos) right[nr++] = prims[i];
	}
	if(nl == n && nr == n) { free(left); free(right); makeLeaf(b, node, prims, n); return; }
	float lhi[3] = {hi[0], hi[1], hi[2]}, rlo[3] = {lo[0], lo[1], lo[2]};
	lhi[best_axis] = best_pos;
	rlo[best_axis] = best_pos;
	b->split[node] = best_pos;
	b->first_ref[node] = 0;
	buildRec(b, left, nl, lo, lhi, depth + 1);
	free(left);
	b->flags[node] = ((uint32_t) b->n_nodes << 2) | (uint32_t) best_axis; /* right child index, known after the left subtree */
	buildRec(b, right, nr, rlo, hi, depth + 1);
	free(right);
}

void *kdo_build(const kdo_mesh *mesh, int max_leaf, int max_depth)
{
	built *b = (built *) calloc(1, sizeof(built));
	const size_t n = mesh->n_faces;
	float *plo = (float *) malloc(3 * sizeof(float) * (n ? n : 1)), *phi = (float *) malloc(3 * sizeof(float) * (n ? n : 1));
	uint32_t *prims = (uint32_t *) malloc(sizeof(uint32_t) * (n ? n : 1));
	for(size_t f = 0; f < n; ++f) { primBound(mesh, f, plo + 3 * f, phi + 3 * f); prims[f] = (uint32_t) f; }
	b->plo = plo;
	b->phi = phi;
	b->max_leaf = max_leaf > 0 ? max_leaf : 2;
	if(max_depth <= 0) max_depth = (int) (8.0 + 1.3 * log2((double) (n ? n : 1)));
	b->max_depth = max_depth > KDO_MAX_STACK - 4 ? KDO_MAX_STACK - 4 : max_depth;
	kdo_tree_bound(mesh, b->bound);
	if(n > 0) buildRec(b, prims, n, b->bound, b->bound + 3, 0);
	free(plo);
	free(phi);
	free(prims);
	b->plo = b->phi = NULL;
	return b;
}

void kdo_built_view(void *h, kdo_tree *out)
{
	const built *b = (const built *) h;
	out->split = b->split;
	out->flags = b->flags;
	out->first_ref = b->first_ref;
	out->refs = b->refs;
	out->n_nodes = b->n_nodes;
	memcpy(out->bound, b->bound, sizeof(b->bound));
}

size_t kdo_built_num_refs(void *h) { return ((const built *) h)->n_refs; }

void kdo_built_free(void *h)
{
	built *b = (built *) h;
	if(!b) return;
	free(b->split);
	free(b->flags);
	free(b->first_ref);
	free(b->refs);
	free(b);
}
