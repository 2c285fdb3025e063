/* integration/src/integrator/surface/integrator_tiled_b200.cc -- TiledIntegrator::renderWorkerWavefront, the render
 * worker used when the scene's accelerator is the "b200-kdtree" one (compiled inside a libYafaRay
 * tree, declared by integration/b200-kdtree.patch in include/integrator/surface/integrator_tiled.h).
 *
 * It does what TiledIntegrator::renderWorker does (src/integrator/surface/integrator_tiled.cc:45-63): take areas from
 * the film until there are none, render them, report each finished area to the thread that flushes the film.  The
 * difference is how an area is rendered: it is cut into blocks of wavefront_block x wavefront_block pixels and the
 * reference's own (virtual) renderTile() runs on every block ON A FIBER of a b200::RayQueue, so that each ray query
 * inside integrate() parks its fiber instead of waiting for a one-ray kernel launch, and the rays of n_fibers blocks
 * travel to the GPU together.  A block keeps its area's id and lock-free "safe" rectangle: ImageFilm::addSample only
 * compares the splat footprint with that rectangle (src/render/imagefilm.cc:804), and fibers of one thread never run
 * concurrently, so the film sees the same accesses as from renderTile() on the whole area.
 */
#include "integrator/surface/integrator_tiled.h"
#include "accelerator/accelerator_b200.h"
#include "common/logger.h"
#include "render/imagefilm.h"
#include "render/imagesplitter.h"
#include "render/render_control.h"
#include "render/wavefront_b200.h"
#include "render/tile_shard_b200.h"
#include <deque>

namespace yafaray {

void TiledIntegrator::renderWorkerWavefront(const AcceleratorB200 &accelerator, ThreadControl *control, std::vector<int> &correlative_sample_number, int thread_id, int samples, int offset, bool adaptive, int aa_pass, float aa_light_sample_multiplier, float aa_indirect_sample_multiplier, const RenderMonitor &render_monitor, const RenderControl &render_control)
{
	std::unique_ptr<b200::RayQueue> queue_holder{accelerator.acquireRayQueue()};
	b200::RayQueue &queue{*queue_holder};
	struct Area
	{
		RenderArea area_;
		int blocks_left_;
	};
	struct Block
	{
		RenderArea block_;
		Area *area_;
	};
	std::deque<Area> areas;      //stable addresses; an area stays here until the pass of this thread ends (a few dozen bytes each)
	std::vector<Block> blocks;   //blocks of the areas taken so far that no fiber has claimed yet
	size_t next_block = 0;
	const int block_size{std::max(1, accelerator.wavefrontBlock())};
	const b200::TileShard shard{b200::TileShard::make(*image_film_, accelerator.tileShardIndex(), accelerator.tileShardCount())}; //multi-GPU rendering: areas of other processes are skipped

	const auto area_finished{[&](const RenderArea &a) {
		std::unique_lock<std::mutex> lk(control->m_);
		control->areas_.emplace_back(a);
		control->c_.notify_one();
	}};
	//called on a fiber; only the film's own nextArea() is shared with other threads
	const auto claim_block{[&](Block &out) -> bool {
		while(next_block == blocks.size())
		{
			blocks.clear();
			next_block = 0;
			RenderArea a;
			if(render_control.canceled() || !image_film_->nextArea(a)) return false;
			if(!shard.owns(a)) continue;
			areas.push_back({a, 0});
			Area &area{areas.back()};
			for(int y = a.y_; y < a.y_ + a.h_; y += block_size)
				for(int x = a.x_; x < a.x_ + a.w_; x += block_size)
				{
					RenderArea b{a}; //same id and safe rectangle
					b.x_ = x;
					b.y_ = y;
					b.w_ = std::min(block_size, a.x_ + a.w_ - x);
					b.h_ = std::min(block_size, a.y_ + a.h_ - y);
					blocks.push_back({b, &area});
					++area.blocks_left_;
				}
			if(area.blocks_left_ == 0) area_finished(a);
		}
		out = blocks[next_block++];
		return true;
	}};
	const bool ok{queue.run([&]() {
		Block job;
		while(claim_block(job))
		{
			renderTile(correlative_sample_number, job.block_, samples, offset, adaptive, thread_id, aa_pass, aa_light_sample_multiplier, aa_indirect_sample_multiplier, render_monitor, render_control);
			if(--job.area_->blocks_left_ == 0) area_finished(job.area_->area_);
		}
	})};
	if(!ok) logger_.logError(getName(), ": b200 wavefront ray queue: ", queue.error());
	if(!ok) accelerator.addWavefrontStats(queue.stats());
	accelerator.releaseRayQueue(std::move(queue_holder));
	std::unique_lock<std::mutex> lk(control->m_);
	++(control->finished_threads_);
	control->c_.notify_one();
}

} //namespace yafaray
