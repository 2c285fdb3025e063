/* integration/include/render/tile_shard_b200.h -- which render areas THIS process renders when one frame is shared
 * by several processes, one per GPU (BASELINE.json configs[2]: "tile-sharded at 1/2/4/8 GPUs").
 *
 * Every process builds the same scene and the same film and walks the same ImageSplitter areas
 * (src/render/imagefilm.cc:549-570), but renders only the areas it owns; its film then holds the weighted sums of its
 * own tiles (filter splats reach ifilterw pixels into neighbouring tiles, src/render/imagefilm.cc:771-822).  The films
 * are SUMMED afterwards -- weights and every layer, exactly what the reference does when it merges ".film" files
 * (src/render/imagefilm.cc:1072-1090) -- by libyafaray_b200/film.py over NCCL (tools/render_sharded.py).
 *
 * Ownership is a function of the RAW tile an area lies in, not of the area id: the splitter shuffles tiles with
 * std::random_device and subdivides the last 2 x threads tiles (src/render/imagesplitter.cc:51-112), so ids and even the
 * set of areas differ between processes; the raw tile grid (tile_size, film origin) is the same everywhere.
 *
 * Selected by the accelerator parameters tile_shard_index / tile_shard_count of "b200-kdtree", or -- for any accelerator,
 * which is what the CPU tests use -- by the environment variable B200_TILE_SHARD="index/count".
 * Only AA_passes = 1 renders shard exactly: an adaptive pass decides what to resample from the whole film
 * (src/render/imagefilm.cc:396-520), which a shard does not have until the films are summed.
 */
#ifndef LIBYAFARAY_TILE_SHARD_B200_H
#define LIBYAFARAY_TILE_SHARD_B200_H

#include "render/imagefilm.h"
#include "render/imagesplitter.h"
#include "param/param.h"
#include <cstdio>
#include <cstdlib>

namespace yafaray::b200 {

struct TileShard final
{
	int index_ = 0, count_ = 1;
	int tile_size_ = 32, x_0_ = 0, y_0_ = 0;

	[[nodiscard]] bool sharded() const { return count_ > 1; }
	/*! Owner of the raw tile (tx, ty): (tx + 3 ty) mod count -- a checkerboard for 2 processes, diagonals for 4 and 8,
	 *  so that every process gets tiles from all over the frame (the cost of a tile depends on what it shows). */
	[[nodiscard]] static int owner(int tx, int ty, int count) { return (tx + 3 * ty) % count; }
	[[nodiscard]] bool owns(const RenderArea &a) const
	{
		if(count_ <= 1) return true;
		return owner((a.x_ - x_0_) / tile_size_, (a.y_ - y_0_) / tile_size_, count_) == index_;
	}
	/*! index/count from the arguments when count > 1, else from B200_TILE_SHARD; the tile grid from the film. */
	static TileShard make(const ImageFilm &film, int index, int count)
	{
		TileShard shard;
		if(count > 1) { shard.index_ = index; shard.count_ = count; }
		else if(const char *env = std::getenv("B200_TILE_SHARD"))
		{
			int i = 0, n = 1;
			if(std::sscanf(env, "%d/%d", &i, &n) == 2 && n >= 1 && i >= 0 && i < n) { shard.index_ = i; shard.count_ = n; }
		}
		if(shard.index_ < 0 || shard.index_ >= shard.count_) { shard.index_ = 0; shard.count_ = 1; }
		int tile_size = 32;
		film.getAsParamMap(false).getParam("tile_size", tile_size);
		shard.tile_size_ = tile_size > 0 ? tile_size : 32;
		shard.x_0_ = film.getCx0();
		shard.y_0_ = film.getCy0();
		return shard;
	}
};

} //namespace yafaray::b200

#endif //LIBYAFARAY_TILE_SHARD_B200_H
