/* integration/include/render/wavefront_b200.h -- the wavefront ray queue that feeds libb200rt from libYafaRay's
 * UNMODIFIED integrators (north-star (c)).
 *
 * The reference's integrators are recursive, one-ray-at-a-time code: integrate() calls Accelerator::intersect /
 * isShadowed / isShadowedTransparentShadow and needs the answer before its next line
 * (src/integrator/surface/integrator_direct_light.cc:101, integrator_path_tracer.cc:145,210,251,
 * integrator_montecarlo.cc:148,240,362, integrator_tiled.cc:630-678).  A GPU needs thousands of rays per launch.
 * Instead of rewriting every integrator into stages, each render thread runs its pixels on FIBERS (cooperative
 * user-level contexts): a fiber executes the reference's own renderTile() on a small block of pixels; when it reaches
 * a ray query, AcceleratorB200 parks the ray in this queue (pinned host memory) and switches to the next fiber.  When
 * every fiber of a group is parked, the queue hands the group's rays to libb200rt (one job per query kind, traced in
 * place from the pinned buffers) and meanwhile runs the next group; a group's fibers resume with their answers when its
 * flight has landed.  The per-ray API is kept, the integrators are untouched, the rays of one thread reach the GPU
 * n_fibers / n_groups at a time, and the GPU latency hides behind the shading of the other group.
 *
 * One RayQueue per render thread; no locks on the hot path (fibers of a queue never run concurrently).
 */
#ifndef LIBYAFARAY_WAVEFRONT_B200_H
#define LIBYAFARAY_WAVEFRONT_B200_H

#include "b200rt.h"
#include <cstddef>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace yafaray::b200 {

class RayQueue final
{
	public:
		struct Stats
		{
			uint64_t rays[3] = {0, 0, 0};   //!< per query kind (B200RT_QUERY_*)
			uint64_t batches = 0;           //!< flushes
			uint64_t calls = 0;             //!< libb200rt jobs (one per query kind, scene and shadow depth present in a flush)
			uint64_t switches = 0;          //!< fiber switches
			double trace_seconds = 0.0;     //!< wall time spent inside libb200rt
			double run_seconds = 0.0;       //!< wall time of run() (shading on fibers + tracing)
			int peak_fibers = 0;
		};
		/*! n_fibers = rays in flight per render thread, split into n_groups groups that take turns: while the rays of one
		 *  group are on the GPU the fibers of the next group shade (2 hides the launch + PCIe + kernel latency behind CPU
		 *  work; 1 = trace and shade alternate).  stack_bytes per fiber (mapped lazily, one guard page each). */
		RayQueue(int n_fibers, int n_groups, size_t stack_bytes);
		~RayQueue();
		RayQueue(const RayQueue &) = delete;
		RayQueue &operator=(const RayQueue &) = delete;
		[[nodiscard]] bool ok() const { return error_.empty(); }
		[[nodiscard]] const std::string &error() const { return error_; }
		[[nodiscard]] const Stats &stats() const { return stats_; }
		void resetStats() { stats_ = Stats{}; }
		[[nodiscard]] int numFibers() const { return n_fibers_; }

		/*! The queue whose fiber is executing on the calling OS thread (nullptr outside a fiber). */
		static RayQueue *current() { return current_; }

		/*! To be called on a fiber: park one ray and return its answer once the batch it joined has been traced.
		 *  Rays are as the Accelerator virtuals receive them (B200RT_RAYS_TREE_SPACE). */
		b200rt_hit closest(b200rt_scene *scene, const b200rt_ray &ray, float time);
		uint32_t shadow(b200rt_scene *scene, const b200rt_ray &ray, float time);
		const b200rt_tshadow &transparentShadow(b200rt_scene *scene, const b200rt_ray &ray, float time, int max_depth);

		/*! Runs `body` once on every fiber (started one after the other; a fiber whose body returns without ever
		 *  having parked a ray tells the queue that no work is left, and no further fiber is started), then keeps
		 *  flushing batches and resuming fibers until all bodies have returned.  False on a libb200rt failure
		 *  (every affected ray then reads as a miss; nothing is traced on the CPU). */
		bool run(const std::function<void()> &body);

	private:
		struct Group;
		struct Fiber
		{
			void *sp = nullptr;        //!< saved stack pointer while switched out
			void *stack = nullptr;     //!< lowest address of its stack (canary word first)
			size_t stack_bytes = 0;
			Group *group = nullptr;
			uint32_t slot = 0;         //!< index of its parked ray within its group and query kind
			bool started = false, done = false;
		};
		struct Request
		{
			b200rt_scene *scene;
			int max_depth;
		};
		static constexpr size_t kMaxGroupsPerKind = 8; //!< distinct (scene, shadow depth) pairs of one query kind per flight
		//! The fibers that park and resume together, with the pinned buffers their rays and answers live in.
		struct Group
		{
			uint32_t capacity = 0;
			std::vector<Fiber *> fibers, parked;
			b200rt_ray *rays[3] = {nullptr, nullptr, nullptr};
			float *times[3] = {nullptr, nullptr, nullptr}; //!< Ray::time_ of the parked rays (motion blur)
			void *outs[3] = {nullptr, nullptr, nullptr};   //!< b200rt_hit[], uint32_t[], b200rt_tshadow[]
			std::vector<Request> requests[3];
			uint32_t count[3] = {0, 0, 0};
			// a flight whose rays of one kind belong to different scenes or shadow depths is traced from sorted copies (rare)
			b200rt_ray *sorted_rays[3] = {nullptr, nullptr, nullptr};
			float *sorted_times[3] = {nullptr, nullptr, nullptr};
			void *sorted_out[3] = {nullptr, nullptr, nullptr};
			std::vector<uint32_t> order[3];
			bool mixed[3] = {false, false, false};
			uint32_t in_flight[3] = {0, 0, 0};
			b200rt_flight *flight = nullptr;
			bool flying = false;
		};
		void park(int kind, b200rt_scene *scene, const b200rt_ray &ray, float time, int max_depth);
		void submit(Group &group);
		void land(Group &group);
		void resume(Fiber &fiber);
		static void entry();
		static void checkStack(const Fiber &fiber);
		static constexpr uint64_t kStackCanary = 0xB200CA9A57ACC0DEull; //!< lowest word of every fiber stack

		static thread_local RayQueue *current_;
		const int n_fibers_;
		std::vector<Fiber> fibers_;
		std::vector<Group> groups_;
		char *slab_ = nullptr;     //!< the pinned memory every group's ray / answer arrays are carved from
		char *stacks_ = nullptr;   //!< one mapping: guard page, then the fibers' stacks
		size_t stacks_bytes_ = 0;
		Fiber *running_ = nullptr;
		void *scheduler_sp_ = nullptr;
		const std::function<void()> *body_ = nullptr;
		std::vector<Fiber *> resuming_;
		std::string error_;
		Stats stats_;
};

} //namespace yafaray::b200

#endif //LIBYAFARAY_WAVEFRONT_B200_H
