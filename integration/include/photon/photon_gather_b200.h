/* integration/include/photon/photon_gather_b200.h -- the photon-map gather of the radiance-map precompute as ONE batched call
 * into libb200rt (C ABI: include/b200pm.h).
 *
 * The reference precomputes the radiance map of final gathering with `threads` OS threads that each take 32 radiance points at a
 * time and call PhotonMap::gather once per point (PhotonIntegrator::preGatherWorker,
 * src/integrator/surface/integrator_photon_mapping.cc:98-147, launched at :511-514).  All points are known before the first
 * gather: b200::preGather uploads the diffuse map once (b200pm_create builds the reference's own point kd-tree, so every lookup
 * walks the same nodes in the same order), gathers all points in one b200pm_gather call and then runs the reference's per-point
 * sum over the `found` arrays, which come back in the reference's own order -- the radiance photons are the same floats.
 *
 * Used only when the scene's accelerator is the b200-kdtree one (the device is taken from it); returns false -- and the caller
 * runs the reference's threads -- with any other accelerator, with B200_PHOTON_GATHER=0, or if a libb200rt call fails (logged).
 */
#ifndef LIBYAFARAY_PHOTON_GATHER_B200_H
#define LIBYAFARAY_PHOTON_GATHER_B200_H

#include "accelerator/accelerator_b200.h"
#include "photon/photon.h"
#include "common/logger.h"
#include "b200pm.h"
#include "b200rt.h"
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

namespace yafaray::b200 {

/*! The photons of a reference PhotonMap on the device (b200pm_map), for batched lookups. */
class DevicePhotonMap final
{
	public:
		DevicePhotonMap(int device, const PhotonMap &map, int build_threads)
		{
			const std::vector<Photon> &photons{map.photons()};
			std::vector<float> pos(3 * photons.size()), dir(3 * photons.size());
			for(size_t i = 0; i < photons.size(); ++i)
				for(size_t c = 0; c < 3; ++c)
				{
					pos[3 * i + c] = photons[i].pos_[c];
					dir[3 * i + c] = photons[i].dir_[c];
				}
			if(b200pm_create(device, pos.data(), dir.data(), photons.size(), build_threads, &map_) != B200RT_OK) error_ = b200rt_last_error();
		}
		~DevicePhotonMap() { if(map_) b200pm_destroy(map_); }
		DevicePhotonMap(const DevicePhotonMap &) = delete;
		DevicePhotonMap &operator=(const DevicePhotonMap &) = delete;
		[[nodiscard]] bool ok() const { return map_ != nullptr; }
		[[nodiscard]] const std::string &error() const { return error_; }
		/*! PhotonMap::gather (src/photon/photon.cc:58-64) for n points: see b200pm_gather. */
		bool gather(const float *points, size_t n, unsigned int k, float sq_radius, b200pm_found *found, uint32_t *n_found, float *sq_radius_out)
		{
			if(b200pm_gather(map_, points, n, k, sq_radius, nullptr, found, n_found, sq_radius_out) == B200RT_OK) return true;
			error_ = b200rt_last_error();
			return false;
		}

	private:
		b200pm_map *map_ = nullptr;
		std::string error_;
};

/*! PhotonIntegrator's radiance-map precompute (integrator_photon_mapping.cc:505-514) with the gathers on the GPU.
 *  gdata: the integrator's PreGatherData, rad_points_ final and radiance_vec_ sized.  true = radiance_vec_ is filled. */
template <typename PreGatherDataT>
bool preGather(Logger &logger, const Accelerator *accelerator, PreGatherDataT &gdata, float ds_rad, int n_search, int threads)
{
	const auto *b200{dynamic_cast<const AcceleratorB200 *>(accelerator)};
	if(!b200 || !b200->ok()) return false;
	if(const char *e{std::getenv("B200_PHOTON_GATHER")}; e && std::atoi(e) == 0) return false;
	const PhotonMap &map{*gdata.getDiffuseMap()};
	const size_t n_points{gdata.rad_points_.size()};
	if(map.nPhotons() <= 0 || n_points == 0 || n_search < 1) return false;
	const auto t_0{std::chrono::steady_clock::now()};
	threads = std::max(1, threads);
	DevicePhotonMap device_map{b200->device(), map, threads};
	if(!device_map.ok())
	{
		logger.logWarning("b200pm: the diffuse photon map could not be put on the device (", device_map.error(), "); gathering on the host");
		return false;
	}
	const std::vector<Photon> &photons{map.photons()};
	const float ds_radius_2{ds_rad * ds_rad};
	const float i_scale{1.f / (static_cast<float>(map.nPaths()) * math::num_pi<>)};
	// slabs bound the host memory of the `found` arrays (n_search entries of 8 bytes per point)
	const size_t slab{std::max<size_t>(4096, (size_t{256} << 20) / (sizeof(b200pm_found) * static_cast<size_t>(n_search)))};
	std::vector<float> points(3 * std::min(slab, n_points)), radius(std::min(slab, n_points));
	std::vector<b200pm_found> found(std::min(slab, n_points) * static_cast<size_t>(n_search));
	std::vector<uint32_t> n_found(std::min(slab, n_points));
	for(size_t begin = 0; begin < n_points; begin += slab)
	{
		const size_t count{std::min(slab, n_points - begin)};
		for(size_t i = 0; i < count; ++i)
			for(size_t c = 0; c < 3; ++c) points[3 * i + c] = gdata.rad_points_[begin + i].pos_[c];
		if(!device_map.gather(points.data(), count, static_cast<unsigned int>(n_search), ds_radius_2, found.data(), n_found.data(), radius.data()))
		{
			logger.logWarning("b200pm: gather failed (", device_map.error(), "); gathering on the host");
			return false;
		}
		// the reference's sum over the gathered photons (integrator_photon_mapping.cc:121-138), same order, same operations
		const auto sum_range{[&](size_t first, size_t last) {
			for(size_t i = first; i < last; ++i)
			{
				const auto &rad_point{gdata.rad_points_[begin + i]};
				const Vec3f rnorm{rad_point.normal_};
				Rgb sum(0.0);
				const int n_gathered{static_cast<int>(n_found[i])};
				if(n_gathered > 0)
				{
					const float scale{i_scale / radius[i]};
					const b200pm_found *gathered{&found[i * static_cast<size_t>(n_search)]};
					for(int j = 0; j < n_gathered; ++j)
					{
						const Photon &photon{photons[gathered[j].photon]};
						const Vec3f pdir{photon.dir_};
						if(rnorm * pdir > 0.f) sum += rad_point.refl_ * scale * photon.col_;
						else sum += rad_point.transm_ * scale * photon.col_;
					}
				}
				gdata.radiance_vec_[begin + i] = Photon{rnorm, rad_point.pos_, sum, rad_point.time_};
			}
		}};
		std::vector<std::thread> workers;
		const size_t per_thread{(count + static_cast<size_t>(threads) - 1) / static_cast<size_t>(threads)};
		for(int t = 0; t < threads; ++t)
		{
			const size_t first{std::min(count, static_cast<size_t>(t) * per_thread)}, last{std::min(count, first + per_thread)};
			if(first < last) workers.emplace_back(sum_range, first, last);
		}
		for(auto &worker : workers) worker.join();
	}
	const double seconds{std::chrono::duration<double>(std::chrono::steady_clock::now() - t_0).count()};
	logger.logInfo("b200pm: radiance pre-gather on the device: ", n_points, " points x ", n_search, " photons from a map of ", map.nPhotons(), " in ", seconds, " s");
	return true;
}

} //namespace yafaray::b200

#endif //LIBYAFARAY_PHOTON_GATHER_B200_H
