/* integration/accelerator_b200.h -- the libYafaRay `Accelerator` that puts libb200rt behind the reference's own
 * accelerator factory (type "b200-kdtree").  This file is compiled INSIDE a libYafaRay source tree (it includes the
 * reference's headers); INTEGRATION.md and integration/b200-kdtree.patch show the registration.
 *
 * It mirrors the interface of the reference's kd-tree accelerators
 * (include/accelerator/accelerator_kdtree_original.h:37-104): same factory signature, same ParamMap keys
 * ("depth", "max_leaf_size_", "cost_ratio", "empty_bonus"), same virtual queries, plus batched entry points.
 * All ray work is done by the CUDA kernels of libb200rt; there is no CPU traversal in this class.
 */
#ifndef LIBYAFARAY_ACCELERATOR_B200_H
#define LIBYAFARAY_ACCELERATOR_B200_H

#include "accelerator/accelerator.h"
#include "render/wavefront_b200.h"
#include <atomic>
#include <memory>
#include <mutex>
#include <vector>

struct b200rt_scene;

namespace yafaray {

class AcceleratorB200 final : public Accelerator
{
		using ThisClassType_t = AcceleratorB200; using ParentClassType_t = Accelerator;

	public:
		inline static std::string getClassName() { return "AcceleratorB200"; }
		static std::pair<std::unique_ptr<Accelerator>, ParamResult> factory(Logger &logger, const RenderControl *render_control, const std::vector<const Primitive *> &primitives, const ParamMap &params);
		[[nodiscard]] std::map<std::string, const ParamMeta *> getParamMetaMap() const override { return params_.getParamMetaMap(); }
		static std::string printMeta(const std::vector<std::string> &excluded_params) { return class_meta::print<Params>(excluded_params); }
		AcceleratorB200(Logger &logger, ParamResult &param_result, const RenderControl *render_control, const std::vector<const Primitive *> &primitives, const ParamMap &param_map);
		~AcceleratorB200() override;
		bool ok() const { return scene_ != nullptr; }
		int device() const { return params_.device_; } //!< CUDA device of the scene (photon/photon_gather_b200.h puts its photon map there)

		/* The batched entry of north-star (c) is the wavefront ray queue below: every Accelerator::intersect / isShadowed /
		 * isShadowedTransparentShadow call an integrator makes on a fiber joins the batch its render thread flushes through
		 * b200rt_trace_jobs (one mixed-kind launch).  (Round 1 also had intersectBatch / isShadowedBatch methods nothing called;
		 * they are gone.) */
		/* Wavefront rendering (render/wavefront_b200.h): TiledIntegrator::renderWorkerWavefront runs the reference's
		 * renderTile() on wavefrontFibers() fibers per render thread; the three per-ray virtuals below then park their ray
		 * in the calling fiber's queue instead of launching a one-ray kernel.  0 fibers = the per-ray path only. */
		int wavefrontFibers() const { return scene_ ? params_.wavefront_fibers_ : 0; }
		int wavefrontGroups() const { return params_.wavefront_groups_; }
		int wavefrontBlock() const { return params_.wavefront_block_; }
		int tileShardIndex() const { return params_.tile_shard_index_; }
		int tileShardCount() const { return params_.tile_shard_count_; }
		size_t wavefrontStackBytes() const { return static_cast<size_t>(std::max(64, params_.wavefront_stack_kb_)) << 10; }
		/* Ray queues (fiber stacks + pinned buffers) are kept between render passes: a render worker borrows one and returns it. */
		std::unique_ptr<b200::RayQueue> acquireRayQueue() const;
		void releaseRayQueue(std::unique_ptr<b200::RayQueue> queue) const;
		void addWavefrontStats(const b200::RayQueue::Stats &stats) const;
		void logWavefrontStats() const; //!< one Info line with the totals since the last call, then resets them
		void logQueueError(const std::string &what) const; //!< a ray queue could not run (photon_fibers_b200.h, integrator_tiled_b200.cc)
		/*! The reference reads object / material visibility and transparency LIVE at every hit (accelerator.h:126-127,138-139,152-154),
		 *  while the GPU scene bakes them per face; Scene::preprocess rebuilds the accelerator for OBJECTS / accelerator-parameter
		 *  changes only (src/scene/scene.cc:318), not when a material is replaced.  Called at the start of every render pass and
		 *  photon pass: re-derives the flag bytes (three virtual calls per primitive) and, if any changed, patches them on the
		 *  device with b200rt_update_face_flags -- no rebuild.  Not to be called while queries are in flight. */
		void refreshFaceFlags() const;

	private:
		[[nodiscard]] Type type() const override { return Type::B200KdTree; }
		const struct Params
		{
			Params(ParamResult &param_result, const ParamMap &param_map);
			static std::map<std::string, const ParamMeta *> getParamMetaMap();
			PARAM_DECL(int, max_depth_, 0, "depth", "0 = automatic");
			PARAM_DECL(int, max_leaf_size_, 0, "max_leaf_size_", "0 = library default");
			PARAM_DECL(float, cost_ratio_, 0.f, "cost_ratio", "node traversal cost divided by primitive intersection cost; 0 = library default");
			PARAM_DECL(float, empty_bonus_, 0.f, "empty_bonus", "0 = library default");
			PARAM_DECL(int, device_, 0, "device", "CUDA device index");
			PARAM_DECL(int, num_threads_, 0, "accelerator_threads", "host threads for the tree build; 0 = all");
			PARAM_DECL(int, wavefront_fibers_, 512, "wavefront_fibers", "rays in flight per render thread (fibers running renderTile on pixel blocks); 0 = per-ray calls only");
			PARAM_DECL(int, wavefront_groups_, 2, "wavefront_groups", "groups the fibers of a thread are split into; one group shades while the rays of another are on the GPU");
			PARAM_DECL(int, wavefront_block_, 2, "wavefront_block", "side of the pixel block one fiber renders");
			PARAM_DECL(int, wavefront_stack_kb_, 256, "wavefront_stack_kb", "stack per fiber, KiB (mapped lazily)");
			PARAM_DECL(int, tile_shard_index_, 0, "tile_shard_index", "multi-GPU rendering: which share of the frame's tiles this process renders (render/tile_shard_b200.h)");
			PARAM_DECL(int, tile_shard_count_, 1, "tile_shard_count", "multi-GPU rendering: number of processes (one per GPU) sharing the frame; 1 = render every tile");
		} params_;
		[[nodiscard]] ParamMap getAsParamMap(bool only_non_default) const override;

		IntersectData intersect(const Ray &ray, float t_max) const override;
		IntersectData intersectShadow(const Ray &ray, float t_max) const override;
		IntersectData intersectTransparentShadow(const Ray &ray, int max_depth, float dist, const Camera *camera) const override;
		Bound<float> getBound() const override { return bound_; }

		std::vector<const Primitive *> primitives_; //!< face id (upload order) -> primitive; copies the factory's temporary vector
		b200rt_scene *scene_ = nullptr;             //!< owned; device memory lives behind this handle
		Bound<float> bound_{{{0.f, 0.f, 0.f}}, {{0.f, 0.f, 0.f}}};
		mutable std::vector<uint8_t> face_flags_;   //!< the flag byte every primitive was uploaded with (refreshFaceFlags)
		mutable std::atomic<bool> depth_clamp_logged_{false};
		mutable std::mutex queues_mutex_;
		mutable std::vector<std::unique_ptr<b200::RayQueue>> idle_queues_;
		mutable std::atomic<uint64_t> wf_rays_[3]{}, wf_batches_{0}, wf_calls_{0}, wf_switches_{0}, wf_trace_us_{0}, wf_run_us_{0}, wf_per_ray_calls_{0}, wf_launches_seen_{0};
};

} //namespace yafaray

#endif //LIBYAFARAY_ACCELERATOR_B200_H
