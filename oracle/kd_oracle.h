/* oracle/kd_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of libYafaRay's kd-tree ray queries (closest hit, shadow, transparent shadow) used as
 * the checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (libyafaray_b200/, include/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement bit-for-bit (ties included)
 * against the unmodified reference (oracle/_ref/libyafref.so, built by oracle/Makefile) when that
 * library is present, and against the committed golden vectors under tests/golden/ (generated from the
 * unmodified reference by tests/golden/make_golden.py) everywhere else.  The reference's own tests/
 * hold no golden vectors for this path.
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 */
#ifndef KD_ORACLE_H
#define KD_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Flat mesh, the same arrays libb200rt takes (include/b200rt.h):
 *   xyz  float[3*n_verts]; idx uint32[4*n_faces] with idx[4f+3]==0xFFFFFFFF for a triangle;
 *   flags uint8[n_faces]: bit0 Visible, bit1 CastsShadows (object AND material, accelerator.h:126-127),
 *                         bit2 material isTransparent() (material.h:85). */
/* A face with idx[4f+2] == KDO_SPHERE is a sphere (SpherePrimitive, src/geometry/primitive/primitive_sphere.cc):
 * vertex idx[4f+0] is its centre, the x component of vertex idx[4f+1] its radius, idx[4f+3] = 0xFFFFFFFF. */
#define KDO_SPHERE 0xFFFFFFFEu
/* Motion blur, per face (NULL = static scene).  kind 1 = face of a Bezier motion-blur mesh: its vertices exist at three time
 * steps (step 0 in kdo_mesh.xyz, steps 1 and 2 in xyz1 / xyz2, same indices) and the shape at ray time t is built from
 * vertices interpolated with the quadratic Bezier factors of t mapped into the mesh's time range
 * (include/geometry/primitive/primitive_polygon.h:238-257, primitive_face.h:86-98, include/math/interpolation.h:50-93).
 * kind 2 = face of a MOVING instance (three obj_to_world matrices): vertices = M(t) * vertex with M(t) interpolated the same way,
 * element by element (include/geometry/instance.h:72-90, primitive_instance.h:83-86, include/geometry/matrix.h:96-144).
 * face_times: 2 floats per face, the time range (the mesh's / the instance's); face_matrix: per face the index of its instance in
 * `matrices` (3 x 16 floats per instance, row major).  A primitive's bound covers all three time steps / matrices
 * (primitive_face.h:155-170, primitive_instance.h:119-128). */
typedef struct kdo_motion
{
	const uint8_t *kind;
	const float *xyz1, *xyz2;
	const float *face_times;
	const uint32_t *face_matrix;
	const float *matrices;
} kdo_motion;

typedef struct kdo_mesh
{
	const float *xyz;
	size_t n_verts;
	const uint32_t *idx;
	size_t n_faces;
	const uint8_t *flags;
	const kdo_motion *motion; /* NULL = every face static */
} kdo_mesh;

/* kd-tree in the reference's own node encoding (accelerator_kdtree_original.h:106-125):
 *   flags[i] bits0-1 = split axis, 3 = leaf; bits 2..31 = right-child index (interior) or primitive
 *   count (leaf); left child = i+1.  split[i] = splitting plane.  Leaf primitives are
 *   refs[first_ref[i] .. first_ref[i]+count).  bound = lo xyz, hi xyz of the (inflated) tree bound. */
typedef struct kdo_tree
{
	const float *split;
	const uint32_t *flags;
	const uint32_t *first_ref;
	const uint32_t *refs;
	size_t n_nodes;
	float bound[6];
} kdo_tree;

typedef struct kdo_counters
{
	uint64_t rays, interior, leaves, refs, tests;
} kdo_counters;

/* accelerator_kdtree_original.cc:88-103: union of primitive bounds, +0.1 % per axis (double multiply). */
void kdo_tree_bound(const kdo_mesh *mesh, float out6[6]);

/* A small self-contained SAH builder (NOT a restatement of the reference builder: closest hits do not
 * depend on the tree except for exact-t ties).  Produces the encoding above.  Returns an
 * opaque handle owning the arrays; kdo_built_view fills a kdo_tree pointing into it. */
void *kdo_build(const kdo_mesh *mesh, int max_leaf, int max_depth);
void kdo_built_view(void *built, kdo_tree *out);
size_t kdo_built_num_refs(void *built);
void kdo_built_free(void *built);

/* Ray record: 8 floats ox oy oz tmin dx dy dz tmax (tmax < 0 => unbounded, accelerator.h:91).
 * Closest hit: wrapper rule of Accelerator::intersect(ray,camera) (accelerator.h:89-101) around
 * kdtree::intersect<Nearest> (accelerator_kdtree_common.h:107-255).  out_prim = -1 on a miss. */
void kdo_trace_closest(const kdo_mesh *mesh, const kdo_tree *tree, const float *rays, size_t n,
                       float *out_t, float *out_u, float *out_v, int32_t *out_prim,
                       int n_threads, kdo_counters *counters);

/* Accelerator::isShadowed (accelerator.h:103-111) around kdtree::intersect<Shadow>. */
void kdo_trace_shadow(const kdo_mesh *mesh, const kdo_tree *tree, const float *rays, size_t n,
                      uint8_t *out_shadowed, int32_t *out_prim, int n_threads, kdo_counters *counters);

/* Accelerator::isShadowedTransparentShadow (accelerator.h:113-120,147-169).  Material evaluation is not
 * part of this path: instead of a colour the oracle returns the number of DISTINCT transparent
 * occluders it multiplied in (out_n_transparent) and up to `max_list` of their primitive ids in visit
 * order (out_list, may be NULL). */
void kdo_trace_tshadow(const kdo_mesh *mesh, const kdo_tree *tree, const float *rays, size_t n, int max_depth,
                       uint8_t *out_shadowed, int32_t *out_n_transparent, int32_t *out_list, int max_list,
                       int n_threads);

/* The same three queries with a ray time per ray (Ray::time_, include/geometry/ray.h:49); times == NULL means time 0. */
void kdo_trace_closest_timed(const kdo_mesh *mesh, const kdo_tree *tree, const float *rays, const float *times, size_t n,
                             float *out_t, float *out_u, float *out_v, int32_t *out_prim, int n_threads);
void kdo_trace_shadow_timed(const kdo_mesh *mesh, const kdo_tree *tree, const float *rays, const float *times, size_t n,
                            uint8_t *out_shadowed, int32_t *out_prim, int n_threads);
void kdo_trace_tshadow_timed(const kdo_mesh *mesh, const kdo_tree *tree, const float *rays, const float *times, size_t n, int max_depth,
                             uint8_t *out_shadowed, int32_t *out_n_transparent, int32_t *out_list, int max_list, int n_threads);
void kdo_brute_closest_timed(const kdo_mesh *mesh, const float bound6[6], const float *rays, const float *times, size_t n,
                             float *out_t, float *out_u, float *out_v, int32_t *out_prim, int n_threads);

/* Tree-free ground truth for small scenes: test every primitive in index order with the same accept
 * rules; the first primitive reaching the minimum t wins. */
void kdo_brute_closest(const kdo_mesh *mesh, const float bound6[6], const float *rays, size_t n,
                       float *out_t, float *out_u, float *out_v, int32_t *out_prim, int n_threads);

/* One polygon test, exported for unit tests (shape_polygon.h:126-176).  nv = 3 or 4.  Returns t (0 = miss). */
float kdo_poly_intersect(const float *v0, const float *v1, const float *v2, const float *v3, int nv,
                         const float from[3], const float dir[3], float *u, float *v);

/* One sphere test, exported for unit tests (primitive_sphere.cc:83-102).  Returns t (0 = miss). */
float kdo_sphere_intersect(const float center[3], float radius, const float from[3], const float dir[3]);

/* Bound<float>::cross (bound.h:156-198). Returns crossed; enter/leave written when crossed. */
int kdo_bound_cross(const float bound6[6], const float from[3], const float dir[3], float t_max, float *enter, float *leave);

#ifdef __cplusplus
}
#endif
#endif
