/* include/b200rt.h -- C ABI of libb200rt: the B200 (sm_100a) ray-scene intersection path for libYafaRay.
 *
 * This is the drop-in boundary.  The entry points below are what a libYafaRay `Accelerator` subclass
 * binds to replace the reference's CPU kd-tree (INTEGRATION.md shows that class and the two-file patch
 * that registers it).  Paths cited are relative to the reference tree.
 *
 *   reference interface                                                    replaced by
 *   -------------------------------------------------------------------    ---------------------------------
 *   AcceleratorKdTree ctor / Accelerator::factory                          b200rt_create + b200rt_add_mesh
 *     (src/accelerator/accelerator_kdtree_original.cc:57-141,               [+ b200rt_add_spheres] + b200rt_build
 *      src/accelerator/accelerator.cc:44-55)
 *   Accelerator::getBound()  (include/accelerator/accelerator.h:53)        b200rt_get_bound
 *   Accelerator::intersect(ray,t_max) / intersect(ray,camera)              b200rt_trace_closest[_device]
 *     (include/accelerator/accelerator.h:50,89-101;
 *      include/accelerator/accelerator_kdtree_common.h:107-255 <Nearest>)
 *   Accelerator::intersectShadow / isShadowed                              b200rt_trace_shadow[_device]
 *     (accelerator.h:51,103-111; accelerator_kdtree_common.h <Shadow>)
 *   Accelerator::intersectTransparentShadow / isShadowedTransparentShadow  b200rt_trace_tshadow[_device]
 *     (accelerator.h:52,113-120,147-169; <TransparentShadow>)
 *   Params depth / max_leaf_size_ / cost_ratio / empty_bonus               b200rt_build_params
 *     (include/accelerator/accelerator_kdtree_original.h:56-59)
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a negative
 * B200RT_E_* code on failure (b200rt_last_error() gives the text for the calling thread); no C++
 * exception crosses this boundary; there is NO CPU fallback -- without a usable CUDA device
 * b200rt_create fails.  Queries are thread-safe: any number of host threads may trace against one
 * built scene concurrently (the reference's render workers do, integrator_tiled.cc:232-248).
 */
#ifndef B200RT_H
#define B200RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200RT_VERSION 3
#define B200RT_MISS 0xFFFFFFFFu

enum
{
	B200RT_OK = 0,
	B200RT_E_INVALID = -1,  /* bad argument / wrong call order */
	B200RT_E_CUDA = -2,     /* CUDA runtime error (text in b200rt_last_error) */
	B200RT_E_NO_DEVICE = -3,
	B200RT_E_MEMORY = -4
};

/* per-face flag bits (b200rt_add_mesh).  Visible / CastsShadows must already be the AND of the object's
 * and the material's visibility (include/accelerator/accelerator.h:126-127,138-139); TRANSPARENT is
 * Material::isTransparent() (include/material/material.h:85). */
enum
{
	B200RT_FACE_VISIBLE = 1,
	B200RT_FACE_CASTS_SHADOWS = 2,
	B200RT_FACE_TRANSPARENT = 4
};

/* Ray: the fields of yafaray::Ray the path reads (include/geometry/ray.h:46-50).  tmax < 0 means
 * unbounded (accelerator.h:91).  dir need not be normalised.  32 bytes. */
typedef struct b200rt_ray
{
	float ox, oy, oz, tmin;
	float dx, dy, dz, tmax;
} b200rt_ray;

/* Closest hit: IntersectData's t_hit_/uv_/primitive_ (include/accelerator/intersect_data.h:30-39).
 * prim = index of the face in upload order (B200RT_MISS and t = 0 on a miss).  16 bytes. */
typedef struct b200rt_hit
{
	float t, u, v;
	uint32_t prim;
} b200rt_hit;

/* Transparent shadow result.  The reference multiplies the ray colour by
 * Material::getTransparency of every DISTINCT transparent shadow caster it meets and reports
 * "shadowed" for an opaque caster or for more than max_depth distinct transparent ones
 * (accelerator.h:147-169).  Material code stays on the host, so the kernel returns the casters:
 * shadowed as in the reference; otherwise n_transparent (<= max_depth <= B200RT_TSHADOW_MAX) entries
 * of prim/t/u/v to evaluate getSurface()/getTransparency() on.  144 bytes. */
#define B200RT_TSHADOW_MAX 8
typedef struct b200rt_tshadow
{
	uint32_t shadowed;
	uint32_t n_transparent;
	uint32_t occluder; /* last accepted caster (IntersectData::primitive_), B200RT_MISS if none */
	uint32_t pad_;
	b200rt_hit transparent[B200RT_TSHADOW_MAX];
} b200rt_tshadow;

/* Build parameters, named after the reference's (accelerator_kdtree_original.h:56-59).  Zero / negative
 * values select the library default for that field. */
typedef struct b200rt_build_params
{
	int max_depth;      /* "depth": 0 = automatic */
	int max_leaf_size;  /* "max_leaf_size_" */
	float cost_ratio;   /* "cost_ratio": node traversal cost / primitive test cost */
	float empty_bonus;  /* "empty_bonus" */
	int build_threads;  /* host threads for the build, 0 = all */
	int reserved_[3];
} b200rt_build_params;

typedef struct b200rt_stats
{
	uint64_t n_faces, n_triangles, n_quads;
	uint64_t n_nodes, n_interior, n_leaves, n_empty_leaves, n_leaf_refs;
	uint32_t max_depth, max_leaf_prims;
	double build_seconds, upload_seconds;
	uint64_t device_bytes;
	uint64_t n_spheres;
	uint64_t n_bezier_faces, n_moving_faces;
} b200rt_stats;

typedef struct b200rt_scene b200rt_scene;

/* Number of CUDA devices (0 with an error code when there is no driver / device). */
int b200rt_device_count(int *count);

/* Create an empty scene bound to CUDA device `device`.  params may be NULL. */
int b200rt_create(int device, const b200rt_build_params *params, b200rt_scene **out);
void b200rt_destroy(b200rt_scene *scene);

/* Append one mesh.  xyz: 3 floats per vertex.  idx: 4 uint32 per face, idx[4f+3] == 0xFFFFFFFF marks a
 * triangle, anything else a quad (v0 v1 v2 v3; the reference tests (v0,v1,v2) then (v0,v2,v3),
 * include/geometry/shape/shape_polygon.h:126-176).  flags: one byte per face (NULL = visible shadow
 * caster).  Face ids continue across calls in upload order, like the primitive vector Scene::preprocess
 * hands to the factory (src/scene/scene.cc:320-341). */
int b200rt_add_mesh(b200rt_scene *scene, const float *xyz, size_t n_verts, const uint32_t *idx, size_t n_faces, const uint8_t *flags);

/* Append spheres (SpherePrimitive, src/geometry/primitive/primitive_sphere.cc:71-102; objects of type "sphere",
 * src/geometry/object/object.cc:80-90).  center_radius: 4 floats per sphere (cx cy cz radius).  flags as for faces.
 * Every sphere takes the next face id, in call order with b200rt_add_mesh.  A hit on a sphere reports u = v = 0 like the
 * reference.  An instance of a sphere is the same sphere: the reference ignores the instance matrix for spheres
 * (primitive_sphere.cc:104-122), so upload it unchanged. */
int b200rt_add_spheres(b200rt_scene *scene, const float *center_radius, size_t n_spheres, const uint8_t *flags);

/* ---- motion blur (Ray::time_, include/geometry/ray.h:49).  Rays keep their 32-byte record: the timed queries below take the ray
 * times in a parallel float array.  Faces added here take the next face ids like any other face.
 *
 * Bezier motion-blur mesh (MeshObject with "motion_blur_bezier", three time steps): at ray time t a face is the polygon of its
 * vertices interpolated with the quadratic Bezier factors of t mapped into [time_start, time_end]; at or outside the ends of the
 * range time step 0 / 2 is used as is (include/geometry/primitive/primitive_polygon.h:238-257, primitive_face.h:86-98,
 * include/math/interpolation.h:50-93).  xyz0 / xyz1 / xyz2 are the three time steps AS THE MESH STORES THEM, i.e. xyz1 holds the
 * Bezier control points MeshObject::convertToBezierControlPoints has put there (src/geometry/object/object_mesh.cc:268-277) --
 * what FacePrimitive::getVertex(v, 1) returns.  A face's bound covers all three steps (primitive_face.h:155-170). */
int b200rt_add_mesh_bezier(b200rt_scene *scene, const float *xyz0, const float *xyz1, const float *xyz2, size_t n_verts, const uint32_t *idx, size_t n_faces,
                           const uint8_t *flags, float time_start, float time_end);
/* Faces of a MOVING instance (Instance with three obj_to_world matrices, include/geometry/instance.h:48,72-90): at ray time t
 * the face's vertices are matrix(t) * vertex, matrix(t) interpolated element by element with the same Bezier factors
 * (primitive_instance.h:83-86, include/geometry/matrix.h:96-144).  xyz = the base object's vertices; matrices = 3 x 16 floats,
 * row major, at time_start, mid-time and time_end.  A face's bound is the union of its bounds under the three matrices
 * (primitive_instance.h:119-128). */
int b200rt_add_mesh_moving(b200rt_scene *scene, const float *xyz, size_t n_verts, const uint32_t *idx, size_t n_faces, const uint8_t *flags,
                           const float matrices[48], float time_start, float time_end);

/* Build the kd-tree on the host, flatten it and upload it.  Must precede any trace call; calling it
 * again after more b200rt_add_mesh calls rebuilds. */
int b200rt_build(b200rt_scene *scene);

/* Tree bound after the reference's 0.1 % inflation (accelerator_kdtree_original.cc:88-103): lo xyz, hi xyz. */
int b200rt_get_bound(const b200rt_scene *scene, float out6[6]);
int b200rt_get_stats(const b200rt_scene *scene, b200rt_stats *out);

/* Re-derive the per-face flag bytes without rebuilding (materials changed between renders). */
int b200rt_update_face_flags(b200rt_scene *scene, const uint8_t *flags, size_t n_faces);

/* ---- host-buffer queries: rays and results live in host memory; the call stages them through pinned
 * buffers in chunks, overlapping H2D, kernel and D2H, and returns when `out` is complete.
 * b200rt_trace_shadow writes the occluding face id or B200RT_MISS (not shadowed) per ray. */
int b200rt_trace_closest(b200rt_scene *scene, const b200rt_ray *rays, size_t n, b200rt_hit *out);
int b200rt_trace_shadow(b200rt_scene *scene, const b200rt_ray *rays, size_t n, uint32_t *out);
int b200rt_trace_tshadow(b200rt_scene *scene, const b200rt_ray *rays, size_t n, int max_depth, b200rt_tshadow *out);

/* ---- device-buffer queries: rays/out are device pointers on the scene's device; the kernel is
 * enqueued on `stream` (a cudaStream_t passed as void*, NULL = legacy default stream) and the call
 * returns without synchronising. */
int b200rt_trace_closest_device(b200rt_scene *scene, const b200rt_ray *d_rays, size_t n, b200rt_hit *d_out, void *stream);
int b200rt_trace_shadow_device(b200rt_scene *scene, const b200rt_ray *d_rays, size_t n, uint32_t *d_out, void *stream);
int b200rt_trace_tshadow_device(b200rt_scene *scene, const b200rt_ray *d_rays, size_t n, int max_depth, b200rt_tshadow *d_out, void *stream);

/* ---- generic entry points: query kind + flags.  The functions above are b200rt_trace[_device] with flags = 0.
 *
 * B200RT_RAYS_TREE_SPACE: the rays are what the reference's VIRTUAL queries receive -- Accelerator::intersect(ray,
 * t_max), intersectShadow(ray, t_max), intersectTransparentShadow(ray, max_depth, t_max, camera)
 * (include/accelerator/accelerator.h:50-52) -- i.e. the public wrappers (accelerator.h:89-120) have already run:
 * the origin is used as is and the `tmax` field IS the t_max argument (negative = unbounded); `tmin` is
 * Ray::tmin_ (closest and transparent shadow use max(tmin, bias); shadow ignores it,
 * accelerator_kdtree_common.h:139).  Without the flag the library applies the wrappers itself. */
enum
{
	B200RT_QUERY_CLOSEST = 0, /* out: b200rt_hit[n] */
	B200RT_QUERY_SHADOW = 1,  /* out: uint32_t[n] */
	B200RT_QUERY_TSHADOW = 2  /* out: b200rt_tshadow[n] */
};
enum
{
	B200RT_RAYS_TREE_SPACE = 1,
	B200RT_BUFFERS_PINNED = 2 /* host-buffer calls: rays and out come from b200rt_host_alloc (skips the pointer query) */
};
int b200rt_trace(b200rt_scene *scene, int query, unsigned flags, const b200rt_ray *rays, size_t n, void *out, int max_depth);
int b200rt_trace_device(b200rt_scene *scene, int query, unsigned flags, const b200rt_ray *d_rays, size_t n, void *d_out, int max_depth, void *stream);
/* The same with one ray time per ray (Ray::time_); times == NULL means time 0 for every ray, which is what the untimed entry
 * points trace with.  Static faces and spheres ignore the time. */
int b200rt_trace_timed(b200rt_scene *scene, int query, unsigned flags, const b200rt_ray *rays, const float *times, size_t n, void *out, int max_depth);
int b200rt_trace_timed_device(b200rt_scene *scene, int query, unsigned flags, const b200rt_ray *d_rays, const float *d_times, size_t n, void *d_out, int max_depth, void *stream);

/* Transparent shadows through more than B200RT_TSHADOW_MAX distinct transparent casters (an integrator "shadow_depth" above 8;
 * the reference has no limit, include/accelerator/accelerator.h:147-169).  The result record of ray i starts at byte
 * i * (16 + 16 * capacity) of `out`: the first 16 bytes are the first four fields of b200rt_tshadow, then `capacity` b200rt_hit
 * entries of which the first n_transparent are written.  max_depth <= capacity <= 4096; n <= 2^26 rays per call.  `times` may
 * be NULL.  The host variant copies rays and results through device memory it allocates for the call. */
int b200rt_trace_tshadow_deep(b200rt_scene *scene, unsigned flags, const b200rt_ray *rays, const float *times, size_t n, int max_depth, int capacity, void *out);
int b200rt_trace_tshadow_deep_device(b200rt_scene *scene, unsigned flags, const b200rt_ray *d_rays, const float *d_times, size_t n, int max_depth, int capacity, void *d_out, void *stream);

/* Several host-buffer batches in one call -- what one flush of the renderer's wavefront ray queue holds (closest,
 * shadow and transparent-shadow rays of the pixels in flight on one render thread; Accelerator::intersect / isShadowed /
 * isShadowedTransparentShadow call sites of integrator_montecarlo.cc:148,240,362, integrator_path_tracer.cc:145,210,251).
 * Batches in pinned memory of up to 65 536 rays are traced in place (the kernels read the rays and write the results
 * across PCIe, no staging copies), all jobs concurrently on their own streams, one wait per job at the end; larger or
 * unpinned batches take the staged path of b200rt_trace one after the other.  Returns the first error. */
typedef struct b200rt_job
{
	b200rt_scene *scene;
	int query;          /* B200RT_QUERY_* */
	unsigned flags;     /* B200RT_RAYS_TREE_SPACE | B200RT_BUFFERS_PINNED */
	const b200rt_ray *rays;
	size_t n;
	void *out;
	int max_depth;      /* transparent shadows only */
	const float *times; /* one ray time per ray (same memory kind as rays), or NULL = time 0 */
} b200rt_job;
int b200rt_trace_jobs(const b200rt_job *jobs, size_t n_jobs);
/* The same in two halves, so that a render thread can shade one group of pixels while the rays of another group are
 * on the GPU: _begin enqueues the in-place jobs and returns (jobs that need staging are traced before it returns);
 * _end waits for them, releases the flight and returns the first error of either half.  The job buffers must stay
 * untouched between the two calls; the job array itself may go away after _begin.
 * Environment (read by b200rt_build): B200RT_COMBINE=1 lets the in-place jobs of ALL calling threads of a scene leave in shared
 * launches (up to B200RT_COMBINE_RAYS rays each, default 2048; _end launches what is pending if its own job still waits).
 * Off by default: fewer, larger launches, but a longer wait for each caller. */
typedef struct b200rt_flight b200rt_flight;
int b200rt_trace_jobs_begin(const b200rt_job *jobs, size_t n_jobs, b200rt_flight **flight);
int b200rt_trace_jobs_end(b200rt_flight *flight);

/* Pinned host memory for ray / result buffers (makes the host-buffer queries copy at full PCIe rate). */
int b200rt_host_alloc(void **ptr, size_t bytes);
int b200rt_host_free(void *ptr);

/* ---- diagnostics: the host-side builder alone (no CUDA device needed).  Used by the CPU test-suite to
 * validate the tree (every face reachable, same hits as the reference traversal run over it) and by tools.
 * Export format: node i = (a[i], b[i]); interior: a = float bits of the split, b = (right child << 2) | axis,
 * left child = i + 1; leaf: a = first index into refs, b = (count << 2) | 3; refs = face ids.
 * Here (and only here) a sphere is passed inside the mesh arrays: a face with idx[4f+2] == 0xFFFFFFFE and
 * idx[4f+3] == 0xFFFFFFFF whose vertex idx[4f+0] is the centre and whose vertex idx[4f+1] carries the radius in x --
 * the layout b200rt_add_spheres stores internally. */
typedef struct b200rt_host_tree b200rt_host_tree;
int b200rt_host_tree_build(const float *xyz, size_t n_verts, const uint32_t *idx, size_t n_faces, const b200rt_build_params *params, b200rt_host_tree **out);
int b200rt_host_tree_sizes(const b200rt_host_tree *tree, size_t *n_nodes, size_t *n_refs);
int b200rt_host_tree_export(const b200rt_host_tree *tree, uint32_t *node_a, uint32_t *node_b, uint32_t *refs, float bound6[6]);
void b200rt_host_tree_destroy(b200rt_host_tree *tree);

/* Kernels launched by this library in this process so far (all scenes); bench.py's gpu_launches. */
uint64_t b200rt_launch_count(void);

/* Text of the last error raised on the calling thread ("" if none). */
const char *b200rt_last_error(void);
int b200rt_version(void);

#ifdef __cplusplus
}
#endif
#endif
