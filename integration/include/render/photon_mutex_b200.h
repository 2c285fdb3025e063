/* integration/include/render/photon_mutex_b200.h -- the lock around SppmIntegrator's shared Halton sequences.
 *
 * SppmIntegrator::photonWorker takes `mutex_` once per photon for four Halton::getNext() calls
 * (src/integrator/surface/integrator_sppm.cc:395-400): ~50 ns of work behind a std::mutex.  With the stock CPU
 * accelerators the sixteen photon threads spend a microsecond tracing between two acquisitions and mostly find the lock free;
 * with the b200-kdtree accelerator the tracing is off the CPU, the threads arrive back to back, most acquisitions go through
 * futex sleep / wake, and the photon pass becomes hand-off bound (measured with render_bench's B200_PROF sampling profile on
 * B200 + 16 vCPUs: 55 % of the process's CPU samples in __lll_lock_wait_private, 380 ns per photon -- profiles/r4e_*).
 * PhotonMutex is a std::mutex until b200::PhotonWorkers switches it to a test-and-test-and-set spin lock for the duration of a
 * pass that runs on fibers (one OS thread per core, nothing to yield to, hand-off = one cache-line transfer).
 */
#ifndef LIBYAFARAY_PHOTON_MUTEX_B200_H
#define LIBYAFARAY_PHOTON_MUTEX_B200_H

#include <atomic>
#include <mutex>
#include <thread>

namespace yafaray::b200 {

class PhotonMutex final
{
	public:
		void lock()
		{
			if(!spin_.load(std::memory_order_relaxed)) { mutex_.lock(); return; }
			for(unsigned k = 0;;)
			{
				if(!held_.exchange(true, std::memory_order_acquire)) return;
				while(held_.load(std::memory_order_relaxed))
				{
#if defined(__x86_64__)
					__builtin_ia32_pause();
#endif
					if((++k & 4095u) == 0u) std::this_thread::yield(); //an oversubscribed host: let the holder run
				}
			}
		}
		void unlock()
		{
			if(!spin_.load(std::memory_order_relaxed)) mutex_.unlock();
			else held_.store(false, std::memory_order_release);
		}
		/*! The per-photon critical section of SppmIntegrator::photonWorker: one tuple (a.getNext(), b.getNext(), c.getNext(),
		 *  d.getNext()) drawn atomically.  As a std::mutex it is exactly the reference's lock / four draws / unlock.  In spin mode the
		 *  calling OS thread draws kBatch tuples per acquisition and serves its fibers from them: which photon gets which tuple of
		 *  the shared sequences depends on thread timing in the reference too, every tuple is still drawn under the lock and used
		 *  once -- but the lock and the sequences' cache lines change cores 64 times less often (the spin lock alone left the photon
		 *  pass hand-off bound: 0.24 s for 500 k photons against 0.16 s with the stock accelerator, profiles/r4f_*). */
		template <typename Sequence>
		void draw(Sequence &a, Sequence &b, Sequence &c, Sequence &d, float &s_1, float &s_2, float &s_3, float &s_4)
		{
			if(!spin_.load(std::memory_order_relaxed))
			{
				mutex_.lock();
				s_1 = a.getNext(); s_2 = b.getNext(); s_3 = c.getNext(); s_4 = d.getNext();
				mutex_.unlock();
				return;
			}
			struct Cache { const PhotonMutex *owner = nullptr; int next = 0, count = 0; float s[kBatch][4]; };
			thread_local Cache cache; //fibers of one OS thread share it; they never run concurrently
			if(cache.owner != this || cache.next == cache.count)
			{
				lock();
				for(int k = 0; k < kBatch; ++k) { cache.s[k][0] = a.getNext(); cache.s[k][1] = b.getNext(); cache.s[k][2] = c.getNext(); cache.s[k][3] = d.getNext(); }
				unlock();
				cache.owner = this;
				cache.next = 0;
				cache.count = kBatch;
			}
			const float *t = cache.s[cache.next++];
			s_1 = t[0]; s_2 = t[1]; s_3 = t[2]; s_4 = t[3];
		}
		/*! To be called while no thread holds or waits for the lock (before the workers start, after they have joined). */
		void spin(bool on) { spin_.store(on, std::memory_order_seq_cst); }

	private:
		static constexpr int kBatch = 64;
		std::mutex mutex_;
		std::atomic<bool> held_{false};
		std::atomic<bool> spin_{false};
};

} //namespace yafaray::b200

#endif //LIBYAFARAY_PHOTON_MUTEX_B200_H
