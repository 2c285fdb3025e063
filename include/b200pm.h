/* include/b200pm.h -- C ABI of libb200rt's photon-map queries: the reference's PhotonMap lookups on B200.
 *
 * Part of libb200rt.so (include/b200rt.h: same conventions -- plain pointers and sizes, 0 on success / negative B200RT_E_* code,
 * b200rt_last_error() for the text, no CPU fallback).  Paths cited are relative to the reference tree.
 *
 *   reference interface                                                        replaced by
 *   -----------------------------------------------------------------------    ------------------------------
 *   PhotonMap::updateTree -> kdtree::PointKdTree<Photon> constructor            b200pm_create
 *     (src/photon/photon.cc:46-56, include/photon/pkdtree.h:105-218)
 *   PhotonMap::gather(p, found, k, sq_radius)                                   b200pm_gather[_device]
 *     (src/photon/photon.cc:58-64; PhotonGather :26-44;
 *      PointKdTree::lookup include/photon/pkdtree.h:221-279)
 *   PhotonMap::findNearest(p, n, dist)                                          b200pm_find_nearest[_device]
 *     (src/photon/photon.cc:66-72; NearestPhoton include/photon/photon.h:101-109)
 *
 * The queries are BATCHED: one call answers the lookups of many points (the reference's natural batches: the radiance-map
 * precompute PhotonIntegrator::preGatherWorker, src/integrator/surface/integrator_photon_mapping.cc:98-160, one gather per
 * radiance point; the hit points of an SPPM pass, src/integrator/surface/integrator_sppm.cc:640-740).
 *
 * Results are the reference's, bit for bit: the tree is the reference's tree (same node array), every lookup walks it in the
 * reference's order, and the `found` array of a gather comes back in the reference's own order -- insertion order while fewer
 * than k photons were found, the libstdc++ max-heap the reference keeps (std::make_heap / pop_heap / push_heap with
 * FoundPhoton::operator<) once k were found -- because the integrators sum photon contributions in that order.
 */
#ifndef B200PM_H
#define B200PM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200PM_NONE 0xFFFFFFFFu

/* FoundPhoton (include/photon/photon.h:50-55): the photon as its index in the array given to b200pm_create. 8 bytes. */
typedef struct b200pm_found
{
	uint32_t photon;
	float dist_square;
} b200pm_found;

typedef struct b200pm_stats
{
	uint64_t n_photons, n_nodes;
	uint32_t depth;        /* levels of the tree (1 = a single leaf) */
	uint32_t reserved_;
	double build_seconds;  /* host median-split build */
	double upload_seconds;
	uint64_t device_bytes;
} b200pm_stats;

typedef struct b200pm_map b200pm_map;

/* Build the reference's point kd-tree over n photons (1 <= n < 2^29, finite positions) and upload it to CUDA device `device`.
 * pos: 3 floats per photon (Photon::pos_).  dir: 3 floats per photon (Photon::dir_), needed by b200pm_find_nearest only; may be
 * NULL.  build_threads: host threads for the build (0 = all); the tree does not depend on it, as in the reference. */
int b200pm_create(int device, const float *pos, const float *dir, size_t n, int build_threads, b200pm_map **out);
void b200pm_destroy(b200pm_map *map);
int b200pm_get_stats(const b200pm_map *map, b200pm_stats *out);

/* PhotonMap::gather for n_points points (3 floats each).  k = the reference's `k` (capacity of `found` per point, >= 1).
 * sq_radii: one squared search radius per point, or NULL = sq_radius for every point.
 *   found[i * k + j], j < n_found[i]   the reference's found[j] of point i (entries past n_found[i] are unspecified)
 *   n_found[i]                         the reference's return value
 *   sq_radius_out[i]                   the reference's sq_radius argument after the call (the distance of the farthest photon
 *                                      kept once k were found, otherwise unchanged); may be NULL
 * Host-buffer variant: points / results in host memory, copied inside the call, returns when the results are complete. */
int b200pm_gather(b200pm_map *map, const float *points, size_t n_points, uint32_t k, float sq_radius, const float *sq_radii,
                  b200pm_found *found, uint32_t *n_found, float *sq_radius_out);
/* Device-buffer variant: all pointers are device pointers on the map's device; enqueued on `stream` (cudaStream_t as void*,
 * NULL = legacy default stream), returns without synchronising. */
int b200pm_gather_device(b200pm_map *map, const float *d_points, size_t n_points, uint32_t k, float sq_radius, const float *d_sq_radii,
                         b200pm_found *d_found, uint32_t *d_n_found, float *d_sq_radius_out, void *stream);

/* PhotonMap::findNearest for n_points points with their normals (3 floats each): the photon the reference returns -- the LAST
 * photon the lookup accepted, i.e. the nearest one with dir * normal > 0 within `dist`, which the reference uses as the SQUARED
 * search radius (photon.cc:66-72) -- as its index, or B200PM_NONE.  Needs the `dir` array at b200pm_create. */
int b200pm_find_nearest(b200pm_map *map, const float *points, const float *normals, size_t n_points, float dist, uint32_t *out_photon);
int b200pm_find_nearest_device(b200pm_map *map, const float *d_points, const float *d_normals, size_t n_points, float dist, uint32_t *d_out_photon, void *stream);

/* ---- diagnostics: the host-side builder alone (no CUDA device needed).  Node i = (a[i], b[i]) in the reference's KdNode
 * terms (pkdtree.h:41-67): b = flags_ ((right child << 2) | axis for an interior node, 3 for a leaf); a = float bits of the
 * split for an interior node, the photon index for a leaf; left child = i + 1.  a / b hold 2 n - 1 entries. */
int b200pm_host_tree_build(const float *pos, size_t n, int build_threads, uint32_t *a, uint32_t *b);

/* Tuning aid (tools/pm_sweep.py), process-wide, not for production use; a negative argument keeps the current value.
 * kernel: gather kernel -- 0 = the plain per-thread loop, 1 = the phased state machine, 2 = phased with one stack pop per step,
 * 3 = the default again (2 above smem_k, 0 with shared-memory heaps up to smem_k);
 * round_steps: node visits per round of the phased kernels; smem_k: largest k whose heaps live in shared memory (0 = always in
 * `found`, at most 256); patience: lanes of a warp that must have a make_heap pending before it is done (1 = no waiting).  The
 * environment variables B200PM_KERNEL=plain|phased1|phased, B200PM_ROUND, B200PM_SMEM_K, B200PM_PATIENCE set the same at load
 * time.  Results do not depend on any of them. */
int b200pm_debug_set_tuning(int kernel, int round_steps, int smem_k, int patience);

#ifdef __cplusplus
}
#endif
#endif
