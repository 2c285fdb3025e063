/* oracle/ref_pm_driver.cc -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product path).
 *
 * A small C ABI ("yref_pm_*") around the UNMODIFIED reference's photon map (libYafaRay, compiled in place from
 * /root/reference by oracle/Makefile):
 *   - PhotonMap::updateTree -> kdtree::PointKdTree<Photon> ctor (src/photon/photon.cc:46-56, include/photon/pkdtree.h:105-218)
 *   - PhotonMap::gather(p, found, k, sq_radius)                 (src/photon/photon.cc:58-64, PhotonGather :26-44)
 *   - PhotonMap::findNearest(p, n, dist)                        (src/photon/photon.cc:66-72, NearestPhoton include/photon/photon.h:101-109)
 * so that tests/ can pin the C restatement (pm_oracle.c) and the CUDA path (b200pm_*) against the real thing, and time the
 * real thing on the host cores.  The node array of the reference's own tree is exported for a bit-for-bit comparison of the
 * tree (private members: this file is compiled with -fno-access-control instead of editing any reference header).
 */
#include "photon/photon.h"
#include "common/logger.h"
#include "render/render_control.h"
#include "render/render_monitor.h"
#include "render/progress_bar.h"

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

namespace {

struct RefMap
{
	yafaray::Logger logger{"yref_pm", nullptr, nullptr, YAFARAY_DISPLAY_CONSOLE_HIDDEN};
	yafaray::RenderMonitor monitor;
	yafaray::RenderControl control;
	std::unique_ptr<yafaray::PhotonMap> map;
	const yafaray::Photon *first = nullptr;
	double build_seconds = 0.0;
};

template <typename F>
void parallelBlocks(size_t n, int n_threads, F &&f)
{
	if(n_threads <= 1) { f(size_t(0), n); return; }
	std::atomic<size_t> next{0};
	constexpr size_t block = 1024;
	std::vector<std::thread> pool;
	for(int t = 0; t < n_threads; ++t)
		pool.emplace_back([&] {
			for(;;)
			{
				const size_t b = next.fetch_add(block);
				if(b >= n) return;
				f(b, std::min(n, b + block));
			}
		});
	for(auto &t : pool) t.join();
}

} // namespace

extern "C" {

/* pos / dir: 3 floats per photon (dir may be NULL = zero).  threads = PhotonMap::threads_pkd_tree_. */
void *yref_pm_create(const float *pos, const float *dir, size_t n, int threads)
{
	auto *m = new RefMap;
	/* nothing to format: the statically linked libstdc++ of this .so must not touch its locale facets inside a Python process */
	m->logger.setConsoleMasterVerbosity(YAFARAY_LOG_LEVEL_MUTE);
	m->logger.setLogMasterVerbosity(YAFARAY_LOG_LEVEL_MUTE);
	m->map = std::make_unique<yafaray::PhotonMap>(m->logger, "yref", threads);
	std::vector<yafaray::Photon> photons(n);
	for(size_t i = 0; i < n; ++i)
	{
		photons[i].pos_ = yafaray::Point3f{{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]}};
		if(dir) photons[i].dir_ = yafaray::Vec3f{{dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]}};
		else photons[i].dir_ = yafaray::Vec3f{{0.f, 0.f, 0.f}};
		photons[i].col_ = yafaray::Rgb{1.f};
	}
	m->map->swapVector(photons);
	const auto t0 = std::chrono::steady_clock::now();
	m->map->updateTree(m->monitor, m->control);
	m->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	m->first = m->map->photons_.data();
	return m;
}

void yref_pm_destroy(void *h) { delete static_cast<RefMap *>(h); }
double yref_pm_build_seconds(void *h) { return static_cast<RefMap *>(h)->build_seconds; }

/* Node array of the reference's tree: a[i] = float bits of the split (interior) or the photon index (leaf),
 * b[i] = KdNode::flags_.  Returns the node count; call with a == NULL to size the arrays. */
int64_t yref_pm_tree_export(void *h, uint32_t *a, uint32_t *b)
{
	auto *m = static_cast<RefMap *>(h);
	if(!m->map->tree_) return 0;
	const auto &tree = *m->map->tree_;
	const int64_t n_nodes = tree.next_free_node_;
	if(!a) return n_nodes;
	for(int64_t i = 0; i < n_nodes; ++i)
	{
		const auto &node = tree.nodes_[i];
		b[i] = node.flags_;
		if(node.isLeaf()) a[i] = static_cast<uint32_t>(node.data_ - m->first);
		else std::memcpy(&a[i], &node.division_, 4);
	}
	return n_nodes;
}

/* PhotonMap::gather for n points (3 floats each).  sq_radii: one squared radius per point, or NULL = sq_radius for all.
 * found_idx / found_d2: k entries per point, the first n_found[i] are the reference's `found` array in its own order
 * (a max-heap once k photons were found, insertion order before); sq_radius_out[i] = the radius argument after the call.
 * Returns the seconds spent in the gather calls (wall clock around the parallel loop). */
double yref_pm_gather(void *h, const float *points, size_t n, uint32_t k, float sq_radius, const float *sq_radii,
                      uint32_t *found_idx, float *found_d2, uint32_t *n_found, float *sq_radius_out, int n_threads)
{
	auto *m = static_cast<RefMap *>(h);
	const auto t0 = std::chrono::steady_clock::now();
	parallelBlocks(n, n_threads, [&](size_t b, size_t e) {
		std::vector<yafaray::FoundPhoton> found(k);
		for(size_t i = b; i < e; ++i)
		{
			const yafaray::Point3f p{{points[3 * i], points[3 * i + 1], points[3 * i + 2]}};
			float radius = sq_radii ? sq_radii[i] : sq_radius;
			const int got = m->map->gather(p, found.data(), k, radius);
			if(n_found) n_found[i] = static_cast<uint32_t>(got);
			if(sq_radius_out) sq_radius_out[i] = radius;
			for(int j = 0; j < got; ++j)
			{
				if(found_idx) found_idx[i * k + j] = static_cast<uint32_t>(found[j].photon_ - m->first);
				if(found_d2) found_d2[i * k + j] = found[j].dist_square_;
			}
		}
	});
	return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

/* PhotonMap::findNearest for n points with their normals; `dist` is what the reference passes as the third argument
 * (used as the squared search radius).  out_idx[i] = photon index or 0xFFFFFFFF. */
double yref_pm_nearest(void *h, const float *points, const float *normals, size_t n, float dist, uint32_t *out_idx, int n_threads)
{
	auto *m = static_cast<RefMap *>(h);
	const auto t0 = std::chrono::steady_clock::now();
	parallelBlocks(n, n_threads, [&](size_t b, size_t e) {
		for(size_t i = b; i < e; ++i)
		{
			const yafaray::Point3f p{{points[3 * i], points[3 * i + 1], points[3 * i + 2]}};
			const yafaray::Vec3f nrm{{normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]}};
			const yafaray::Photon *photon = m->map->findNearest(p, nrm, dist);
			out_idx[i] = photon ? static_cast<uint32_t>(photon - m->first) : 0xFFFFFFFFu;
		}
	});
	return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

} // extern "C"
