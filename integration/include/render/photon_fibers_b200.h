/* integration/include/render/photon_fibers_b200.h -- photon shooting through the wavefront ray queue (SURVEY.md 8f row N4,
 * BASELINE.json configs[4]: "photon + gather rays through the GPU accelerator").
 *
 * The reference shoots photons from `threads_photons` OS threads, one worker function call per thread:
 *   PhotonIntegrator::diffuseWorker / causticWorker   src/integrator/surface/integrator_photon_mapping.cc:157-280,373,452
 *   CausticPhotonIntegrator::causticWorker            src/integrator/surface/integrator_photon_caustic.cc:100-225,264
 *   SppmIntegrator::photonWorker                      src/integrator/surface/integrator_sppm.cc:353-505,564
 * Worker i shoots photons [i * n/T, (i + 1) * n/T) of the Halton sequence (T = num_threads_photons_) and follows each one
 * bounce by bounce with Accelerator::intersect -- one ray at a time, which on a GPU accelerator would be one kernel launch per
 * bounce.  Here the SAME worker functions run as many more, smaller LOGICAL workers: num_threads_photons_ is raised to
 * os_threads x fibers while the photons are shot, every OS thread runs its share of the logical workers on the fibers of a
 * b200::RayQueue (render/wavefront_b200.h), and each Accelerator::intersect parks its fiber until the thread's batch has been
 * traced.  The photon sequence is partitioned exactly as the reference partitions it for that thread count; which photons land in the
 * map does not depend on the partition (index-keyed Halton samples), only their order in the map and the draws from the
 * shared `fast_random_` do -- as they already do between two runs of the reference.
 *
 * With any other accelerator (or wavefront_fibers = 0) run() starts one std::thread per worker, the reference's own code path.
 */
#ifndef LIBYAFARAY_PHOTON_FIBERS_B200_H
#define LIBYAFARAY_PHOTON_FIBERS_B200_H

#include "accelerator/accelerator_b200.h"
#include "render/wavefront_b200.h"
#include "render/photon_mutex_b200.h"
#include <algorithm>
#include <cstdlib>
#include <functional>
#include <thread>
#include <vector>

namespace yafaray::b200 {

class PhotonWorkers final
{
	public:
		/*! num_threads: the integrator's num_threads_photons_ (raised here, restored by run() / the destructor);
		 *  n_photons: how many photons this pass shoots (a logical worker gets at least kMinPhotonsPerWorker of them). */
		PhotonWorkers(const Accelerator *accelerator, int &num_threads, int n_photons, PhotonMutex *per_photon_lock = nullptr) : num_threads_{num_threads}, os_threads_{std::max(1, num_threads)}, per_photon_lock_{per_photon_lock}
		{
			b200_ = dynamic_cast<const AcceleratorB200 *>(accelerator);
			if(b200_) b200_->refreshFaceFlags(); //materials may have been replaced since the accelerator was built
			if(b200_ && b200_->wavefrontFibers() > 0)
			{
				const int per_thread{std::clamp(n_photons / (os_threads_ * minPhotonsPerWorker()), 1, b200_->wavefrontFibers())};
				fibers_per_thread_ = per_thread;
				num_threads_ = os_threads_ * per_thread;
			}
			else b200_ = nullptr;
		}
		~PhotonWorkers() { restore(); }
		PhotonWorkers(const PhotonWorkers &) = delete;
		PhotonWorkers &operator=(const PhotonWorkers &) = delete;
		[[nodiscard]] int osThreads() const { return os_threads_; }
		[[nodiscard]] int logicalWorkers() const { return b200_ ? os_threads_ * fibers_per_thread_ : os_threads_; }
		void restore() { num_threads_ = os_threads_; }

		/*! Calls worker(i) for every i in [0, logicalWorkers()) and returns when all have returned. */
		void run(const std::function<void(int)> &worker)
		{
			std::vector<std::thread> threads;
			threads.reserve(os_threads_);
			if(!b200_)
			{
				for(int i = 0; i < os_threads_; ++i) threads.emplace_back(worker, i);
			}
			else
			{
				if(per_photon_lock_) per_photon_lock_->spin(true); //the workers arrive at it back to back (photon_mutex_b200.h)
				for(int t = 0; t < os_threads_; ++t) threads.emplace_back([this, t, &worker]() {
					std::unique_ptr<RayQueue> queue{b200_->acquireRayQueue()};
					int next{t * fibers_per_thread_};
					const int end{next + fibers_per_thread_};
					//every fiber takes logical workers until none is left; fibers of one OS thread never run concurrently
					const bool ran{queue && queue->run([&]() { while(next < end) worker(next++); })};
					if(!ran)
					{
						//the queue was unusable (pinned / stack allocation failed) or a libb200rt call failed in flight: say so, and give
						//the logical workers that never started to the per-ray path -- it still traces on the GPU, one launch per ray
						b200_->logQueueError(queue ? queue->error() : std::string{"no ray queue"});
						while(next < end) worker(next++);
					}
					if(queue) b200_->releaseRayQueue(std::move(queue));
				});
			}
			for(auto &thread : threads) thread.join();
			if(b200_ && per_photon_lock_) per_photon_lock_->spin(false);
			if(b200_) b200_->logWavefrontStats();
			restore();
		}

	private:
		static constexpr int kMinPhotonsPerWorker = 16;
		//! photons a logical worker shoots at least; B200_MIN_PHOTONS_PER_WORKER overrides it (tuning aid, profiles/r3l_*)
		static int minPhotonsPerWorker()
		{
			static const int value{[] { const char *e{std::getenv("B200_MIN_PHOTONS_PER_WORKER")}; const int n{e ? std::atoi(e) : 0}; return n > 0 ? n : kMinPhotonsPerWorker; }()};
			return value;
		}
		int &num_threads_;
		const int os_threads_;
		int fibers_per_thread_ = 1;
		const AcceleratorB200 *b200_ = nullptr;
		PhotonMutex *const per_photon_lock_;
};

} //namespace yafaray::b200

#endif //LIBYAFARAY_PHOTON_FIBERS_B200_H
