// libyafaray_b200/csrc/kd_build.h -- host-side SAH kd-tree builder for the GPU traversal kernels.
//
// Own design (not a restatement of the reference builder, src/accelerator/accelerator_kdtree_original.cc:
// 154-645): closest-hit and shadow results do not depend on the tree except for exact-t ties,
// only the tree BOUND does (it fixes the ray bias t_min), and that is reproduced exactly in treeBound().
#ifndef B200RT_KD_BUILD_H
#define B200RT_KD_BUILD_H

#include <cstddef>
#include <cstdint>
#include <vector>

namespace b200rt {

struct BuildConfig
{
	int max_depth = 0;        // 0 = 8 + 1.3 log2(N), clamped to kMaxTreeDepth
	int max_leaf_size = 0;    // 0 = default (2): stop splitting at or below this many primitives
	float cost_ratio = 0.f;   // traversal cost / primitive test cost; <= 0 = default
	float empty_bonus = -1.f; // SAH discount for a split that cuts off empty space; < 0 = default
	int threads = 0;          // 0 = hardware concurrency
};

static constexpr int kMaxTreeDepth = 60; // traversal stack in the kernels holds 64 entries

// Node, 8 bytes, depth-first order, left child = this + 1.
//   interior: a = float bits of the split position, b = (right child index << 2) | axis
//   leaf:     a = index of its first entry in leaf_refs,  b = (count << 2) | 3
struct HostNode
{
	uint32_t a, b;
};

struct HostTree
{
	std::vector<HostNode> nodes;
	std::vector<uint32_t> leaf_refs; // face ids, concatenated in leaf (depth-first) order
	float bound[6] = {0, 0, 0, 0, 0, 0};
	uint64_t n_interior = 0, n_leaves = 0, n_empty_leaves = 0;
	uint32_t depth = 0, max_leaf_prims = 0;
};

// Flat mesh view shared by builder and flattener.  idx: 4 per face, idx[3] == 0xFFFFFFFF => triangle.
// A face with idx[2] == 0xFFFFFFFE is a sphere: vertex idx[0] is its centre, the x of vertex idx[1] its radius.
// A face with idx[2] == 0xFFFFFFFD is a box: vertices idx[0] / idx[1] are the corners of the primitive's bound (motion-blur
// primitives: the builder sees the union of their bounds over the time steps and never clips them).
struct MeshView
{
	const float *xyz;
	size_t n_verts;
	const uint32_t *idx;
	size_t n_faces;
};

// Tree bound: union of face bounds, then per axis offset = double(hi - lo) * 0.001, lo -= float(offset),
// hi += float(offset) -- the arithmetic of src/accelerator/accelerator_kdtree_original.cc:88-103, which
// the ray bias (include/accelerator/accelerator.h:64) depends on.  Empty mesh => all zero.
void treeBound(const MeshView &mesh, float out6[6]);

void buildKdTree(const MeshView &mesh, const BuildConfig &config, HostTree &out);

} // namespace b200rt
#endif
