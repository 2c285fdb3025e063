// libyafaray_b200/csrc/kd_build.cc -- host-side SAH kd-tree builder (see kd_build.h).
//
// Top-down surface-area-heuristic build over per-reference bounding boxes:
//   * > kExactThreshold references: min/max binning with kBins bins per axis,
//   * otherwise an exact sweep over the sorted box edges,
//   * <= kClipThreshold references: straddling polygons are re-clipped to the child box in double
//     precision ("perfect splits"), which tightens the boxes the next SAH decision sees,
//   * large subtrees are built by separate threads and spliced together in depth-first order.
#include "kd_build.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <future>
#include <limits>
#include <thread>

namespace b200rt {

namespace {

constexpr int kBins = 256;
constexpr size_t kExactThreshold = 384;
constexpr size_t kClipThreshold = 64;
constexpr size_t kSpawnThreshold = 40000; // references; below this a subtree is built serially
constexpr uint32_t kTriangle = 0xFFFFFFFFu;
constexpr uint32_t kSphere = 0xFFFFFFFEu; // idx[2] of a sphere "face": vertex idx[0] = centre, x of vertex idx[1] = radius (kd_build.h)
constexpr uint32_t kBox = 0xFFFFFFFDu;    // idx[2] of a box "face" (a primitive the builder only knows the bound of): vertices idx[0], idx[1] = lo, hi corner

void faceBound(const MeshView &mesh, size_t f, float lo[3], float hi[3]);

struct Ref
{
	uint32_t prim;
	float lo[3], hi[3];
};

struct Box
{
	float lo[3], hi[3];
};

struct Subtree
{
	// right child stored RELATIVE to the node (b = rel << 2 | axis) and leaf first-ref relative to `refs`
	std::vector<HostNode> nodes;
	std::vector<uint32_t> refs;
	uint64_t n_interior = 0, n_leaves = 0, n_empty = 0;
	uint32_t depth = 0, max_leaf = 0;
};

struct Context
{
	MeshView mesh;
	int max_depth;
	size_t max_leaf_size;
	float cost_ratio, empty_bonus;
	double clip_pad; // absolute padding of the clip box (scene scale)
	int max_spawn_depth;
};

inline uint32_t floatBits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

inline double halfArea(const double e[3]) { return e[0] * e[1] + e[1] * e[2] + e[2] * e[0]; }

struct Split
{
	int axis = -1;
	float pos = 0.f;
	double cost = std::numeric_limits<double>::infinity();
};

// SAH cost of splitting `box` at `pos` on `axis` with n_left / n_right references.
inline double splitCost(const Context &c, const double ext[3], double inv_area, int axis, double w_left, size_t n_left, size_t n_right)
{
	const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
	const double cap = ext[a1] * ext[a2], rim = ext[a1] + ext[a2];
	const double area_l = cap + w_left * rim, area_r = cap + (ext[axis] - w_left) * rim;
	double cost = (area_l * double(n_left) + area_r * double(n_right)) * inv_area;
	if(n_left == 0 || n_right == 0) cost *= (1.0 - c.empty_bonus);
	return c.cost_ratio + cost;
}

Split findSplitBinned(const Context &c, const std::vector<Ref> &refs, const Box &box)
{
	Split best;
	const double ext[3] = {double(box.hi[0]) - box.lo[0], double(box.hi[1]) - box.lo[1], double(box.hi[2]) - box.lo[2]};
	const double area = halfArea(ext);
	if(!(area > 0.0)) return best;
	const double inv_area = 1.0 / area;
	const size_t n = refs.size();
	for(int axis = 0; axis < 3; ++axis)
	{
		if(!(ext[axis] > 0.0)) continue;
		uint32_t starts[kBins], ends[kBins];
		std::memset(starts, 0, sizeof(starts));
		std::memset(ends, 0, sizeof(ends));
		const float origin = box.lo[axis];
		const float scale = float(kBins / ext[axis]);
		for(const Ref &r : refs)
		{
			int bl = int((r.lo[axis] - origin) * scale), bh = int((r.hi[axis] - origin) * scale);
			bl = std::min(std::max(bl, 0), kBins - 1);
			bh = std::min(std::max(bh, 0), kBins - 1);
			++starts[bl];
			++ends[bh];
		}
		size_t n_left = 0, n_gone = 0;
		for(int k = 1; k < kBins; ++k)
		{
			n_left += starts[k - 1];
			n_gone += ends[k - 1];
			const double w = ext[axis] * k / kBins;
			const double cost = splitCost(c, ext, inv_area, axis, w, n_left, n - n_gone);
			if(cost < best.cost)
			{
				const float pos = float(double(origin) + w);
				if(pos > box.lo[axis] && pos < box.hi[axis]) { best.cost = cost; best.axis = axis; best.pos = pos; }
			}
		}
	}
	return best;
}

// An edge of the exact sweep as ONE integer: the float position mapped to an unsigned integer of the same order (sign bit flipped for
// positive values, all bits for negative ones) in the high bits, the kind (0 = end, 1 = planar, 2 = start) in the low two.  Sorting
// the integers is the (position, kind) order the sweep needs, with one compare per pair instead of two float compares and a branch;
// equal keys are indistinguishable, so any sorting algorithm gives the same sweep.
inline uint64_t edgeKey(float pos, uint32_t kind)
{
	uint32_t u = floatBits(pos);
	u ^= (u & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u;
	return (uint64_t(u) << 2) | kind;
}
inline float edgePos(uint64_t key)
{
	uint32_t u = uint32_t(key >> 2);
	u ^= (u & 0x80000000u) ? 0x80000000u : 0xFFFFFFFFu;
	float f;
	std::memcpy(&f, &u, 4);
	return f;
}

Split findSplitExact(const Context &c, const std::vector<Ref> &refs, const Box &box, std::vector<uint64_t> &edges)
{
	Split best;
	const double ext[3] = {double(box.hi[0]) - box.lo[0], double(box.hi[1]) - box.lo[1], double(box.hi[2]) - box.lo[2]};
	const double area = halfArea(ext);
	if(!(area > 0.0)) return best;
	const double inv_area = 1.0 / area;
	const size_t n = refs.size();
	edges.resize(2 * n);
	for(int axis = 0; axis < 3; ++axis)
	{
		if(!(ext[axis] > 0.0)) continue;
		size_t n_edges = 0;
		for(const Ref &r : refs)
		{
			if(r.lo[axis] == r.hi[axis]) edges[n_edges++] = edgeKey(r.lo[axis], 1u);
			else { edges[n_edges++] = edgeKey(r.lo[axis], 2u); edges[n_edges++] = edgeKey(r.hi[axis], 0u); }
		}
		std::sort(edges.begin(), edges.begin() + n_edges);
		size_t started = 0, ended = 0; // references with lo < p / hi < p
		for(size_t i = 0; i < n_edges;)
		{
			const float p = edgePos(edges[i]);
			size_t n_end = 0, n_planar = 0, n_start = 0;
			// one plane = one float value: -0 and +0 are different keys but the same plane
			for(; i < n_edges && edgePos(edges[i]) == p; ++i)
			{
				const uint32_t kind = uint32_t(edges[i] & 3u);
				if(kind == 0u) ++n_end;
				else if(kind == 1u) ++n_planar;
				else ++n_start;
			}
			if(p > box.lo[axis] && p < box.hi[axis])
			{
				// planar references lying in the plane go to the left child (see partition())
				const size_t n_left = started + n_planar, n_right = n - ended - n_end - n_planar;
				const double cost = splitCost(c, ext, inv_area, axis, double(p) - box.lo[axis], n_left, n_right);
				if(cost < best.cost) { best.cost = cost; best.axis = axis; best.pos = p; }
			}
			started += n_start + n_planar;
			ended += n_end + n_planar;
		}
	}
	return best;
}

// ---- polygon clipping against an axis-aligned box, double precision (Sutherland-Hodgman) ----
struct Poly
{
	double v[12][3];
	int n = 0;
};

void clipPlane(const Poly &in, Poly &out, int axis, double pos, bool keep_below)
{
	out.n = 0;
	for(int i = 0; i < in.n; ++i)
	{
		const double *a = in.v[i], *b = in.v[(i + 1) % in.n];
		const bool a_in = keep_below ? (a[axis] <= pos) : (a[axis] >= pos);
		const bool b_in = keep_below ? (b[axis] <= pos) : (b[axis] >= pos);
		if(a_in) { std::memcpy(out.v[out.n++], a, sizeof(double) * 3); }
		if(a_in != b_in)
		{
			const double t = (pos - a[axis]) / (b[axis] - a[axis]);
			double *o = out.v[out.n++];
			for(int k = 0; k < 3; ++k) o[k] = a[k] + t * (b[k] - a[k]);
			o[axis] = pos;
		}
	}
}

// Bounds of (triangle v0 v1 v2) intersected with [lo,hi]; returns false when empty.
bool clipTriangle(const float *v0, const float *v1, const float *v2, const double lo[3], const double hi[3], double out_lo[3], double out_hi[3])
{
	Poly a, b;
	a.n = 3;
	for(int k = 0; k < 3; ++k) { a.v[0][k] = v0[k]; a.v[1][k] = v1[k]; a.v[2][k] = v2[k]; }
	for(int axis = 0; axis < 3; ++axis)
	{
		clipPlane(a, b, axis, lo[axis], false);
		if(b.n == 0) return false;
		clipPlane(b, a, axis, hi[axis], true);
		if(a.n == 0) return false;
	}
	for(int k = 0; k < 3; ++k) { out_lo[k] = a.v[0][k]; out_hi[k] = a.v[0][k]; }
	for(int i = 1; i < a.n; ++i)
		for(int k = 0; k < 3; ++k)
		{
			out_lo[k] = std::min(out_lo[k], a.v[i][k]);
			out_hi[k] = std::max(out_hi[k], a.v[i][k]);
		}
	return true;
}

inline float roundDown(double x) { float f = float(x); return (double(f) > x) ? std::nextafter(f, -std::numeric_limits<float>::infinity()) : f; }
inline float roundUp(double x) { float f = float(x); return (double(f) < x) ? std::nextafter(f, std::numeric_limits<float>::infinity()) : f; }

// Tightened bounds of face `prim` inside `box` (padded so that a hit the float Moeller-Trumbore test accepts
// a hair outside the exact polygon still finds the face in the leaf that contains the hit point).
// Returns false when the face does not reach the padded box at all.
bool clipRef(const Context &c, uint32_t prim, const Box &box, Ref &out)
{
	const uint32_t *id = c.mesh.idx + 4 * size_t(prim);
	if(id[2] == kSphere || id[2] == kBox)
	{
		// no clipping for spheres and motion-blur primitives (no clippingSupport() in the reference either): the bound cut to the box
		out.prim = prim;
		faceBound(c.mesh, prim, out.lo, out.hi);
		for(int k = 0; k < 3; ++k)
		{
			out.lo[k] = std::max(out.lo[k], box.lo[k]);
			out.hi[k] = std::min(out.hi[k], box.hi[k]);
			if(out.lo[k] > out.hi[k]) return false;
		}
		return true;
	}
	const float *v0 = c.mesh.xyz + 3 * size_t(id[0]), *v1 = c.mesh.xyz + 3 * size_t(id[1]), *v2 = c.mesh.xyz + 3 * size_t(id[2]);
	double lo[3], hi[3];
	for(int k = 0; k < 3; ++k)
	{
		const double pad = c.clip_pad + 1e-5 * (double(box.hi[k]) - box.lo[k]);
		lo[k] = box.lo[k] - pad;
		hi[k] = box.hi[k] + pad;
	}
	double alo[3], ahi[3];
	bool any = clipTriangle(v0, v1, v2, lo, hi, alo, ahi);
	if(id[3] != kTriangle)
	{
		const float *v3 = c.mesh.xyz + 3 * size_t(id[3]);
		double blo[3], bhi[3];
		if(clipTriangle(v0, v2, v3, lo, hi, blo, bhi))
		{
			if(any) for(int k = 0; k < 3; ++k) { alo[k] = std::min(alo[k], blo[k]); ahi[k] = std::max(ahi[k], bhi[k]); }
			else for(int k = 0; k < 3; ++k) { alo[k] = blo[k]; ahi[k] = bhi[k]; }
			any = true;
		}
	}
	if(!any) return false;
	out.prim = prim;
	for(int k = 0; k < 3; ++k)
	{
		out.lo[k] = std::max(roundDown(alo[k]), box.lo[k]);
		out.hi[k] = std::min(roundUp(ahi[k]), box.hi[k]);
		if(out.lo[k] > out.hi[k]) return false;
	}
	return true;
}

void makeLeaf(Subtree &st, const std::vector<Ref> &refs, int depth)
{
	const uint32_t n = uint32_t(refs.size());
	st.nodes.push_back({uint32_t(st.refs.size()), (n << 2) | 3u});
	for(const Ref &r : refs) st.refs.push_back(r.prim);
	++st.n_leaves;
	if(n == 0) ++st.n_empty;
	st.max_leaf = std::max(st.max_leaf, n);
	st.depth = std::max(st.depth, uint32_t(depth));
}

void appendSubtree(Subtree &dst, const Subtree &src)
{
	const uint32_t ref_base = uint32_t(dst.refs.size());
	const size_t node_base = dst.nodes.size();
	dst.nodes.insert(dst.nodes.end(), src.nodes.begin(), src.nodes.end());
	if(ref_base)
		for(size_t i = node_base; i < dst.nodes.size(); ++i)
			if((dst.nodes[i].b & 3u) == 3u) dst.nodes[i].a += ref_base;
	dst.refs.insert(dst.refs.end(), src.refs.begin(), src.refs.end());
	dst.n_interior += src.n_interior;
	dst.n_leaves += src.n_leaves;
	dst.n_empty += src.n_empty;
	dst.depth = std::max(dst.depth, src.depth);
	dst.max_leaf = std::max(dst.max_leaf, src.max_leaf);
}

void buildNode(const Context &c, Subtree &st, std::vector<Ref> &refs, const Box &box, int depth, int bad_refines, int spawn_depth)
{
	const size_t n = refs.size();
	if(n <= c.max_leaf_size || depth >= c.max_depth) { makeLeaf(st, refs, depth); return; }

	Split split;
	if(n > kExactThreshold) split = findSplitBinned(c, refs, box);
	else
	{
		static thread_local std::vector<uint64_t> edges; // scratch of the exact sweep, one per build thread
		split = findSplitExact(c, refs, box, edges);
	}
	const double leaf_cost = double(n);
	if(split.axis < 0) { makeLeaf(st, refs, depth); return; }
	if(split.cost >= leaf_cost)
	{
		// a non-improving split: tolerate a few in a row for larger nodes, as SAH plateaus are common
		++bad_refines;
		if((split.cost > 1.5 * leaf_cost && n < 16) || bad_refines >= 3) { makeLeaf(st, refs, depth); return; }
	}

	// partition: strictly-left / strictly-right by the reference's box; a reference lying IN the plane goes left
	const int axis = split.axis;
	const float pos = split.pos;
	Box lbox = box, rbox = box;
	lbox.hi[axis] = pos;
	rbox.lo[axis] = pos;
	std::vector<Ref> left, right;
	left.reserve(n / 2 + 8);
	right.reserve(n / 2 + 8);
	const bool clip = n <= kClipThreshold;
	for(const Ref &r : refs)
	{
		const bool planar_on_plane = (r.lo[axis] == pos && r.hi[axis] == pos);
		const bool go_left = r.lo[axis] < pos || planar_on_plane;
		const bool go_right = r.hi[axis] > pos;
		if(go_left && go_right)
		{
			Ref a = r, b = r;
			a.hi[axis] = pos;
			b.lo[axis] = pos;
			if(clip)
			{
				Ref t;
				if(clipRef(c, r.prim, lbox, t)) left.push_back(t);
				if(clipRef(c, r.prim, rbox, t)) right.push_back(t);
			}
			else { left.push_back(a); right.push_back(b); }
		}
		else if(go_left) left.push_back(r);
		else if(go_right) right.push_back(r);
		else left.push_back(r); // lo == hi == something else cannot happen; keep the reference anyway
	}
	if(left.size() == n && right.size() == n) { makeLeaf(st, refs, depth); return; }
	std::vector<Ref>().swap(refs); // release the parent's list before recursing

	const size_t node = st.nodes.size();
	st.nodes.push_back({floatBits(pos), uint32_t(axis)});
	++st.n_interior;
	if(spawn_depth < c.max_spawn_depth && std::min(left.size(), right.size()) >= kSpawnThreshold)
	{
		Subtree lsub, rsub;
		auto fut = std::async(std::launch::async, [&]() { buildNode(c, lsub, left, lbox, depth + 1, bad_refines, spawn_depth + 1); });
		buildNode(c, rsub, right, rbox, depth + 1, bad_refines, spawn_depth + 1);
		fut.get();
		appendSubtree(st, lsub);
		st.nodes[node].b |= uint32_t(st.nodes.size() - node) << 2;
		appendSubtree(st, rsub);
	}
	else
	{
		buildNode(c, st, left, lbox, depth + 1, bad_refines, spawn_depth);
		st.nodes[node].b |= uint32_t(st.nodes.size() - node) << 2;
		buildNode(c, st, right, rbox, depth + 1, bad_refines, spawn_depth);
	}
}

void faceBound(const MeshView &mesh, size_t f, float lo[3], float hi[3])
{
	const uint32_t *id = mesh.idx + 4 * f;
	if(id[2] == kSphere)
	{
		// SpherePrimitive::getBound (src/geometry/primitive/primitive_sphere.cc:71-75): r = radius * 1.0001f, centre -+ r.
		// Part of the tree bound, hence of the ray bias: the same float operations.
		volatile float r = mesh.xyz[3 * size_t(id[1])] * 1.0001f;
		for(int k = 0; k < 3; ++k)
		{
			volatile float a = mesh.xyz[3 * size_t(id[0]) + k] - r, g = mesh.xyz[3 * size_t(id[0]) + k] + r;
			lo[k] = a;
			hi[k] = g;
		}
		return;
	}
	if(id[2] == kBox)
	{
		for(int k = 0; k < 3; ++k) { lo[k] = mesh.xyz[3 * size_t(id[0]) + k]; hi[k] = mesh.xyz[3 * size_t(id[1]) + k]; }
		return;
	}
	const int nv = (id[3] == kTriangle) ? 3 : 4;
	for(int k = 0; k < 3; ++k) lo[k] = hi[k] = mesh.xyz[3 * size_t(id[0]) + k];
	for(int v = 1; v < nv; ++v)
		for(int k = 0; k < 3; ++k)
		{
			const float x = mesh.xyz[3 * size_t(id[v]) + k];
			lo[k] = std::min(lo[k], x);
			hi[k] = std::max(hi[k], x);
		}
}

} // namespace

void treeBound(const MeshView &mesh, float out6[6])
{
	for(int k = 0; k < 6; ++k) out6[k] = 0.f;
	for(size_t f = 0; f < mesh.n_faces; ++f)
	{
		float lo[3], hi[3];
		faceBound(mesh, f, lo, hi);
		for(int k = 0; k < 3; ++k)
		{
			if(f == 0 || lo[k] < out6[k]) out6[k] = lo[k];
			if(f == 0 || hi[k] > out6[3 + k]) out6[3 + k] = hi[k];
		}
	}
	for(int k = 0; k < 3; ++k)
	{
		const double offset = double(out6[3 + k] - out6[k]) * 0.001; // float difference, double product
		out6[k] -= float(offset);
		out6[3 + k] += float(offset);
	}
}

void buildKdTree(const MeshView &mesh, const BuildConfig &config, HostTree &out)
{
	out = HostTree{};
	treeBound(mesh, out.bound);
	const size_t n = mesh.n_faces;
	if(n == 0)
	{
		out.nodes.push_back({0u, 3u}); // a single empty leaf keeps the kernels branch-free
		out.n_leaves = out.n_empty_leaves = 1;
		return;
	}
	Context c;
	c.mesh = mesh;
	int depth = config.max_depth > 0 ? config.max_depth : int(8.0 + 1.3 * std::log2(double(n)));
	c.max_depth = std::min(depth, kMaxTreeDepth);
	c.max_leaf_size = size_t(config.max_leaf_size > 0 ? config.max_leaf_size : 2);
	c.cost_ratio = config.cost_ratio > 0.f ? config.cost_ratio : 1.0f;
	c.empty_bonus = (config.empty_bonus >= 0.f && config.empty_bonus < 1.f) ? config.empty_bonus : 0.3f;
	double diag = 0.0;
	for(int k = 0; k < 3; ++k) diag = std::max(diag, double(out.bound[3 + k]) - out.bound[k]);
	c.clip_pad = 1e-6 * diag;
	const int threads = config.threads > 0 ? config.threads : int(std::max(1u, std::thread::hardware_concurrency()));
	c.max_spawn_depth = 0;
	while((1 << c.max_spawn_depth) < 2 * threads && c.max_spawn_depth < 8) ++c.max_spawn_depth;
	if(threads <= 1) c.max_spawn_depth = 0;

	std::vector<Ref> refs(n);
	for(size_t f = 0; f < n; ++f)
	{
		refs[f].prim = uint32_t(f);
		faceBound(mesh, f, refs[f].lo, refs[f].hi);
	}
	Box box;
	for(int k = 0; k < 3; ++k) { box.lo[k] = out.bound[k]; box.hi[k] = out.bound[3 + k]; }
	Subtree st;
	st.nodes.reserve(2 * n);
	st.refs.reserve(n + n / 2);
	buildNode(c, st, refs, box, 0, 0, 0);

	// relative -> absolute right-child indices
	for(size_t i = 0; i < st.nodes.size(); ++i)
		if((st.nodes[i].b & 3u) != 3u)
		{
			const uint32_t rel = st.nodes[i].b >> 2;
			st.nodes[i].b = (uint32_t(i + rel) << 2) | (st.nodes[i].b & 3u);
		}
	out.nodes = std::move(st.nodes);
	out.leaf_refs = std::move(st.refs);
	out.n_interior = st.n_interior;
	out.n_leaves = st.n_leaves;
	out.n_empty_leaves = st.n_empty;
	out.depth = st.depth;
	out.max_leaf_prims = st.max_leaf;
}

} // namespace b200rt
