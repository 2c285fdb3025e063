/* oracle/pm_oracle.c -- TEST INFRASTRUCTURE ONLY: never linked, imported or executed by the product path
 * (libyafaray_b200/, include/).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it.
 *
 * Plain-C restatement of the reference's photon-map queries:
 *   - kdtree::PointKdTree<Photon> construction           include/photon/pkdtree.h:105-218
 *       (balanced median split on the largest axis of the CLIPPED node bound, median goes right,
 *        ties broken by the element's address = its index; one photon per leaf; preorder node array)
 *   - PointKdTree::lookup (non-recursive)                include/photon/pkdtree.h:221-291
 *   - PhotonGather (k nearest, libstdc++ heap calls)     src/photon/photon.cc:26-44
 *   - PhotonMap::gather / findNearest                    src/photon/photon.cc:58-72
 *   - NearestPhoton                                      include/photon/photon.h:101-109
 *
 * PARITY STATUS: PINNED -- tests/test_pm_oracle.py compares the node array, every gather result (photon ids in the
 * reference's own `found` order, squared distances bit for bit, counts, final radii) and every findNearest result with the
 * UNMODIFIED reference (oracle/_ref/libyafref.so through ref_pm_driver.cc) and with tests/golden/pm/pm_*.npz, which
 * tests/golden/pm/make_pm_golden.py generated from that same unmodified reference.
 *
 * std::make_heap / pop_heap / push_heap are not in the reference tree: they come from libstdc++ (GCC 13 here,
 * bits/stl_heap.h __push_heap / __adjust_heap / __make_heap / __pop_heap); the three functions heap_* below restate that
 * published algorithm for the comparator FoundPhoton::operator< (include/photon/photon.h:52).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct pmo_map
{
	size_t n;
	float *pos; /* 3 per photon */
	float *dir; /* 3 per photon */
	uint32_t *node_a, *node_b;
	size_t n_nodes;
} pmo_map;

/* ---- construction -------------------------------------------------------------------------------------------- */

/* CompareNode (pkdtree.h:71-79): by coordinate, then by address (= index in the photon vector) */
static int less_than(const float *pos, int axis, uint32_t i, uint32_t j)
{
	const float a = pos[3 * (size_t) i + axis], b = pos[3 * (size_t) j + axis];
	return a == b ? (i < j) : (a < b);
}

/* std::nth_element's postcondition under a strict total order: position `nth` holds the element of that rank, everything
 * before it is smaller, everything after it larger.  The SETS on either side are therefore fixed whatever the algorithm;
 * this is a plain quickselect. */
static void select_nth(const float *pos, int axis, uint32_t *v, size_t lo, size_t hi /* exclusive */, size_t nth)
{
	while(hi - lo > 1)
	{
		const size_t mid = lo + (hi - lo) / 2;
		/* median of three to v[lo] */
		uint32_t a = v[lo], b = v[mid], c = v[hi - 1];
		uint32_t pivot;
		if(less_than(pos, axis, a, b)) pivot = less_than(pos, axis, b, c) ? b : (less_than(pos, axis, a, c) ? c : a);
		else pivot = less_than(pos, axis, a, c) ? a : (less_than(pos, axis, b, c) ? c : b);
		size_t i = lo, j = hi - 1;
		for(;;)
		{
			while(less_than(pos, axis, v[i], pivot)) ++i;
			while(less_than(pos, axis, pivot, v[j])) --j;
			if(i >= j) break;
			const uint32_t t = v[i]; v[i] = v[j]; v[j] = t;
			++i; --j;
		}
		if(i == j)
		{
			/* both scans stopped on the pivot itself: it is in its final place */
			if(nth == j) return;
			if(nth < j) hi = j;
			else lo = j + 1;
		}
		else
		{
			/* v[lo..j] < v[j+1..hi) (all keys distinct), both parts non-empty */
			if(nth <= j) hi = j + 1;
			else lo = j + 1;
		}
	}
}

static int largest_axis(const float lo[3], const float hi[3])
{
	/* Bound::largestAxis (include/geometry/bound.h:100-104) */
	const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
	return (dx > dy) ? ((dx > dz) ? 0 : 2) : ((dy > dz) ? 1 : 2);
}

/* buildTreeWorker (pkdtree.h:144-218); the threaded levels only build subtrees apart and splice them back in preorder, the
 * node array is the same as the sequential one */
static void build(pmo_map *m, uint32_t *prims, size_t start, size_t end, const float lo[3], const float hi[3])
{
	if(end - start == 1)
	{
		m->node_a[m->n_nodes] = prims[start];
		m->node_b[m->n_nodes] = 3u;
		++m->n_nodes;
		return;
	}
	const int axis = largest_axis(lo, hi);
	const size_t split_el = (start + end) / 2;
	select_nth(m->pos, axis, prims, start, end, split_el);
	const size_t cur = m->n_nodes++;
	const float split = m->pos[3 * (size_t) prims[split_el] + axis];
	memcpy(&m->node_a[cur], &split, 4);
	float lo_r[3] = {lo[0], lo[1], lo[2]}, hi_l[3] = {hi[0], hi[1], hi[2]};
	hi_l[axis] = split;
	lo_r[axis] = split;
	build(m, prims, start, split_el, lo, hi_l);
	m->node_b[cur] = ((uint32_t) m->n_nodes << 2) | (uint32_t) axis;
	build(m, prims, split_el, end, lo_r, hi);
}

pmo_map *pmo_create(const float *pos, const float *dir, size_t n)
{
	if(!n) return NULL;
	pmo_map *m = (pmo_map *) calloc(1, sizeof(pmo_map));
	m->n = n;
	m->pos = (float *) malloc(12 * n);
	m->dir = (float *) calloc(3 * n, 4);
	memcpy(m->pos, pos, 12 * n);
	if(dir) memcpy(m->dir, dir, 12 * n);
	m->node_a = (uint32_t *) malloc(4 * 2 * n);
	m->node_b = (uint32_t *) malloc(4 * 2 * n);
	uint32_t *prims = (uint32_t *) malloc(4 * n);
	float lo[3], hi[3];
	for(int c = 0; c < 3; ++c) lo[c] = hi[c] = pos[c];
	for(size_t i = 0; i < n; ++i)
	{
		prims[i] = (uint32_t) i;
		for(int c = 0; c < 3; ++c)
		{
			/* Bound::include (bound.h): min / max per component */
			const float v = pos[3 * i + c];
			if(v < lo[c]) lo[c] = v;
			if(v > hi[c]) hi[c] = v;
		}
	}
	build(m, prims, 0, n, lo, hi);
	free(prims);
	return m;
}

void pmo_destroy(pmo_map *m)
{
	if(!m) return;
	free(m->pos); free(m->dir); free(m->node_a); free(m->node_b); free(m);
}

int64_t pmo_tree_export(const pmo_map *m, uint32_t *a, uint32_t *b)
{
	if(a)
	{
		memcpy(a, m->node_a, 4 * m->n_nodes);
		memcpy(b, m->node_b, 4 * m->n_nodes);
	}
	return (int64_t) m->n_nodes;
}

/* ---- libstdc++ heap algorithms on (photon, dist2) pairs, comparator dist2 < dist2 ------------------------------- */

typedef struct { uint32_t photon; float d2; } found_t;

static void heap_push(found_t *first, long hole, long top, found_t value)
{
	long parent = (hole - 1) / 2;
	while(hole > top && first[parent].d2 < value.d2)
	{
		first[hole] = first[parent];
		hole = parent;
		parent = (hole - 1) / 2;
	}
	first[hole] = value;
}

static void heap_adjust(found_t *first, long hole, long len, found_t value)
{
	const long top = hole;
	long second = hole;
	while(second < (len - 1) / 2)
	{
		second = 2 * (second + 1);
		if(first[second].d2 < first[second - 1].d2) --second;
		first[hole] = first[second];
		hole = second;
	}
	if((len & 1) == 0 && second == (len - 2) / 2)
	{
		second = 2 * (second + 1);
		first[hole] = first[second - 1];
		hole = second - 1;
	}
	heap_push(first, hole, top, value);
}

static void heap_make(found_t *first, long len)
{
	if(len < 2) return;
	long parent = (len - 2) / 2;
	for(;;)
	{
		const found_t value = first[parent];
		heap_adjust(first, parent, len, value);
		if(parent == 0) return;
		--parent;
	}
}

/* ---- lookup ---------------------------------------------------------------------------------------------------- */

typedef struct
{
	int kind; /* 0 = PhotonGather, 1 = NearestPhoton */
	/* gather */
	found_t *found;
	uint32_t n_lookup, n_found;
	/* nearest */
	const float *normal;
	uint32_t nearest;
} proc_t;

static void proc_call(const pmo_map *m, proc_t *proc, uint32_t photon, float dist_2, float *max_dist_squared)
{
	if(proc->kind == 0)
	{
		/* PhotonGather::operator() (photon.cc:26-44) */
		const long n = (long) proc->n_lookup;
		if(proc->n_found < proc->n_lookup)
		{
			proc->found[proc->n_found].photon = photon;
			proc->found[proc->n_found].d2 = dist_2;
			++proc->n_found;
			if(proc->n_found == proc->n_lookup)
			{
				heap_make(proc->found, n);
				*max_dist_squared = proc->found[0].d2;
			}
		}
		else
		{
			/* std::pop_heap(first, first + n): the top goes to first[n - 1], the old last element is sifted in from the root */
			if(n > 1)
			{
				const found_t value = proc->found[n - 1];
				proc->found[n - 1] = proc->found[0];
				heap_adjust(proc->found, 0, n - 1, value);
			}
			proc->found[n - 1].photon = photon;
			proc->found[n - 1].d2 = dist_2;
			/* std::push_heap(first, first + n) */
			heap_push(proc->found, n - 1, 0, proc->found[n - 1]);
			*max_dist_squared = proc->found[0].d2;
		}
	}
	else
	{
		/* NearestPhoton::operator() (photon.h:103-106) */
		const float *d = &m->dir[3 * (size_t) photon];
		if(d[0] * proc->normal[0] + d[1] * proc->normal[1] + d[2] * proc->normal[2] > 0.f)
		{
			proc->nearest = photon;
			*max_dist_squared = dist_2;
		}
	}
}

/* PointKdTree::lookup, NON_REC_LOOKUP branch (pkdtree.h:224-279) */
static void lookup(const pmo_map *m, const float p[3], proc_t *proc, float *max_dist_squared)
{
	struct { int64_t node; float s; int axis; } stack[64];
	int64_t curr = 0;
	int stack_ptr = 1;
	stack[stack_ptr].node = -1;
	for(;;)
	{
		while((m->node_b[curr] & 3u) != 3u)
		{
			const int axis = (int) (m->node_b[curr] & 3u);
			float split_val;
			memcpy(&split_val, &m->node_a[curr], 4);
			int64_t far_child;
			if(p[axis] <= split_val)
			{
				far_child = (int64_t) (m->node_b[curr] >> 2);
				curr = curr + 1;
			}
			else
			{
				far_child = curr + 1;
				curr = (int64_t) (m->node_b[curr] >> 2);
			}
			++stack_ptr;
			stack[stack_ptr].node = far_child;
			stack[stack_ptr].axis = axis;
			stack[stack_ptr].s = split_val;
		}
		const uint32_t photon = m->node_a[curr];
		const float *q = &m->pos[3 * (size_t) photon];
		const float vx = q[0] - p[0], vy = q[1] - p[1], vz = q[2] - p[2];
		float dist_2 = vx * vx + vy * vy + vz * vz;
		if(dist_2 < *max_dist_squared) proc_call(m, proc, photon, dist_2, max_dist_squared);
		if(stack[stack_ptr].node < 0) return;
		int axis = stack[stack_ptr].axis;
		dist_2 = p[axis] - stack[stack_ptr].s;
		dist_2 *= dist_2;
		while(dist_2 > *max_dist_squared)
		{
			--stack_ptr;
			if(stack[stack_ptr].node < 0) return;
			axis = stack[stack_ptr].axis;
			dist_2 = p[axis] - stack[stack_ptr].s;
			dist_2 *= dist_2;
		}
		curr = stack[stack_ptr].node;
		--stack_ptr;
	}
}

/* The lookup WITHOUT the split-plane pruning: every leaf, near child first -- the order PointKdTree::lookup visits the leaves it
 * does visit.  Not a restatement of anything in the reference: a cross-check of the argument DESIGN.md 11 rests on (pruning only
 * skips photons the distance test would reject anyway, so any traversal that keeps the leaf order gives the same results). */
static void lookup_unpruned(const pmo_map *m, const float p[3], proc_t *proc, float *max_dist_squared, int64_t node)
{
	if((m->node_b[node] & 3u) == 3u)
	{
		const uint32_t photon = m->node_a[node];
		const float *q = &m->pos[3 * (size_t) photon];
		const float vx = q[0] - p[0], vy = q[1] - p[1], vz = q[2] - p[2];
		const float dist_2 = vx * vx + vy * vy + vz * vz;
		if(dist_2 < *max_dist_squared) proc_call(m, proc, photon, dist_2, max_dist_squared);
		return;
	}
	const int axis = (int) (m->node_b[node] & 3u);
	float split_val;
	memcpy(&split_val, &m->node_a[node], 4);
	const int64_t left = node + 1, right = (int64_t) (m->node_b[node] >> 2);
	if(p[axis] <= split_val) { lookup_unpruned(m, p, proc, max_dist_squared, left); lookup_unpruned(m, p, proc, max_dist_squared, right); }
	else { lookup_unpruned(m, p, proc, max_dist_squared, right); lookup_unpruned(m, p, proc, max_dist_squared, left); }
}

static int g_unpruned = 0;
void pmo_set_unpruned(int on) { g_unpruned = on; }

/* PhotonMap::gather for n points; same argument meaning as yref_pm_gather (ref_pm_driver.cc) */
void pmo_gather(const pmo_map *m, const float *points, size_t n, uint32_t k, float sq_radius, const float *sq_radii,
                uint32_t *found_idx, float *found_d2, uint32_t *n_found, float *sq_radius_out)
{
	found_t *found = (found_t *) malloc(sizeof(found_t) * (k ? k : 1));
	for(size_t i = 0; i < n; ++i)
	{
		proc_t proc;
		memset(&proc, 0, sizeof proc);
		proc.kind = 0;
		proc.found = found;
		proc.n_lookup = k;
		float radius = sq_radii ? sq_radii[i] : sq_radius;
		if(g_unpruned) lookup_unpruned(m, &points[3 * i], &proc, &radius, 0);
		else lookup(m, &points[3 * i], &proc, &radius);
		if(n_found) n_found[i] = proc.n_found;
		if(sq_radius_out) sq_radius_out[i] = radius;
		for(uint32_t j = 0; j < proc.n_found; ++j)
		{
			if(found_idx) found_idx[i * k + j] = found[j].photon;
			if(found_d2) found_d2[i * k + j] = found[j].d2;
		}
	}
	free(found);
}

/* PhotonMap::findNearest for n points */
void pmo_nearest(const pmo_map *m, const float *points, const float *normals, size_t n, float dist, uint32_t *out_idx)
{
	for(size_t i = 0; i < n; ++i)
	{
		proc_t proc;
		memset(&proc, 0, sizeof proc);
		proc.kind = 1;
		proc.normal = &normals[3 * i];
		proc.nearest = 0xFFFFFFFFu;
		float radius = dist;
		if(g_unpruned) lookup_unpruned(m, &points[3 * i], &proc, &radius, 0);
		else lookup(m, &points[3 * i], &proc, &radius);
		out_idx[i] = proc.nearest;
	}
}
