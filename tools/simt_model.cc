// tools/simt_model.cc -- design aid, not product code: a warp-level cost model of the traversal kernel's control structure.
//
// It replays the kernel's closest-hit traversal (t-interval rule of kd_kernels.cuh, plain float arithmetic -- a model, not a
// parity checker) for a sample of rays over the tree libb200rt's builder exported, records per ray the sequence
// "k node steps, then a leaf with c primitives" ..., and then runs scheduling policies over those traces, counting warp
// instructions with per-phase costs taken from the SASS (tools/sass_loop.py, profiles/*regions*):
//   A  the shipped kernel: 32 rays per warp, refill at >= 8 idle lanes, leaf phase at >= 8 pending lanes, 16-step rounds
//   B  two rays per lane, the parked one swapped in at round boundaries
//   C  a pool of `slots` rays per warp, any lane may take any ray each round (upper bound for slot-based designs)
// Calibrated against ncu (policy A must land near 130 warp instructions per ray at 16 lanes).
//
//   g++ -O2 -o /tmp/simt_model tools/simt_model.cc && python tools/simt_model.py
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct Seg { uint16_t steps; uint16_t count; };   // `steps` node visits ending at a leaf with `count` primitives (0: the ray ended there)
struct Trace { std::vector<Seg> segs; bool missed_bound = false; };

static std::vector<uint32_t> na, nb, refs;
static std::vector<float> xyz, rays;
static std::vector<uint32_t> idx;
static float bound[6];

template <typename T> static std::vector<T> load(const std::string &path)
{
	FILE *f = fopen(path.c_str(), "rb");
	if(!f) { perror(path.c_str()); exit(1); }
	fseek(f, 0, SEEK_END);
	const long n = ftell(f);
	fseek(f, 0, SEEK_SET);
	std::vector<T> v(n / sizeof(T));
	if(fread(v.data(), sizeof(T), v.size(), f) != v.size()) exit(1);
	fclose(f);
	return v;
}

static float triTest(uint32_t face, const float *o, const float *d)
{
	const float *v0 = &xyz[3 * idx[4 * face]], *v1 = &xyz[3 * idx[4 * face + 1]], *v2 = &xyz[3 * idx[4 * face + 2]];
	const float e1[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]}, e2[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
	const float p[3] = {d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0]};
	const float det = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
	if(det == 0.f) return 0.f;
	const float inv = 1.f / det, t[3] = {o[0] - v0[0], o[1] - v0[1], o[2] - v0[2]};
	const float u = (t[0] * p[0] + t[1] * p[1] + t[2] * p[2]) * inv;
	if(u < 0.f || u > 1.f) return 0.f;
	const float q[3] = {t[1] * e1[2] - t[2] * e1[1], t[2] * e1[0] - t[0] * e1[2], t[0] * e1[1] - t[1] * e1[0]};
	const float v = (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]) * inv;
	if(v < 0.f || u + v > 1.f) return 0.f;
	const float tt = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) * inv;
	return tt > 0.f ? tt : 0.f;
}

static Trace traceRay(const float *r, bool shadow)
{
	Trace tr;
	const float o[3] = {r[0], r[1], r[2]}, d[3] = {r[4], r[5], r[6]};
	float t_max = r[7] >= 0.f ? r[7] : FLT_MAX, inv[3], lo = -FLT_MAX, hi = FLT_MAX;
	for(int a = 0; a < 3; ++a)
	{
		inv[a] = d[a] == 0.f ? FLT_MAX : 1.f / d[a];
		if(d[a] != 0.f)
		{
			const float t0 = (bound[a] - o[a]) * inv[a], t1 = (bound[3 + a] - o[a]) * inv[a];
			lo = std::max(lo, std::min(t0, t1));
			hi = std::min(hi, std::max(t0, t1));
		}
	}
	if(!(lo <= hi) || hi < 0.f || lo > t_max) { tr.missed_bound = true; return tr; }
	const float t_min = std::max(r[3], 5e-6f * std::fabs(hi - lo));
	float seg_lo = std::max(lo, 0.f), seg_hi = std::min(hi, t_max);
	std::vector<std::pair<uint32_t, float>> stack;
	uint32_t node = 0, steps = 0;
	bool hit = false;
	for(;;)
	{
		++steps;
		const uint32_t a = na[node], b = nb[node];
		if((b & 3u) != 3u)
		{
			const uint32_t axis = b & 3u, right = b >> 2, left = node + 1;
			float split;
			memcpy(&split, &a, 4);
			const float tp = (split - o[axis]) * inv[axis];
			const bool neg = inv[axis] < 0.f;
			const uint32_t near = neg ? right : left, far = neg ? left : right;
			const float limit = shadow ? seg_hi : std::min(seg_hi, t_max);
			if(tp > limit) node = near;
			else if(tp < seg_lo) node = far;
			else { stack.push_back({far, seg_hi}); node = near; seg_hi = tp; }
			continue;
		}
		const uint32_t count = b >> 2;
		bool done = false;
		if(count)
		{
			tr.segs.push_back({uint16_t(steps), uint16_t(count)});
			steps = 0;
			for(uint32_t k = 0; k < count; ++k)
			{
				const float t = triTest(refs[a + k], o, d);
				if(t > 0.f && t >= t_min && t < t_max)
				{
					hit = true;
					if(shadow) { done = true; break; }
					t_max = t;
				}
			}
		}
		if(!done && !shadow && hit && t_max <= seg_hi) done = true;
		if(!done && stack.empty()) done = true;
		if(done)
		{
			if(!count) tr.segs.push_back({uint16_t(steps), 0});
			else tr.segs.push_back({0, 0}); // ended right after a leaf test
			return tr;
		}
		node = stack.back().first;
		seg_lo = seg_hi;
		seg_hi = stack.back().second;
		stack.pop_back();
	}
}

// ---- cost constants (warp instructions), from the SASS of the shipped kernel --------------------------------------------
struct Costs
{
	double round_fixed = 14;  // votes and branches of one trip of the outer loop
	double pool = 20;         // cursor atomic + shuffle when a pool runs dry (once per 256 rays)
	double setup = 175;       // ray loads, 3 IEEE divisions, slab test, shared-memory rows
	double step = 46;         // one trip of the descent loop
	double descent_fixed = 18; // loop prologue / stop-reason code after the loop
	double test = 85;         // one trip of the leaf loop (union of the early-out paths over the lanes)
	double leaf_fixed = 40;   // record address, pop after the leaf
	double write = 12;        // result store
	double swap = 50;         // policy B: park / unpark the lane's other ray (TMEM round trip + selects)
	double take = 34;         // policy C: load / store a slot's hot state around a phase
};

struct Tally
{
	double inst[5] = {0, 0, 0, 0, 0}, lanes[5] = {0, 0, 0, 0, 0}; // 0 refill+votes 1 descent 2 leaf 3 write/other 4 swap/take
	uint64_t rays = 0;
	void add(int region, double n_inst, double active) { inst[region] += n_inst; lanes[region] += n_inst * active; }
	void print(const char *name) const
	{
		double ti = 0, tl = 0;
		for(int k = 0; k < 5; ++k) { ti += inst[k]; tl += lanes[k]; }
		printf("%-34s %7.1f warp-inst/ray at %5.2f lanes |", name, ti / rays, tl / ti);
		const char *rn[5] = {"refill+votes", "descent", "leaf", "write", "swap/take"};
		for(int k = 0; k < 5; ++k) if(inst[k] > 0) printf(" %s %5.1f@%4.1f", rn[k], inst[k] / rays, lanes[k] / inst[k]);
		printf("\n");
	}
};

struct RayCursor
{
	const Trace *tr = nullptr;
	size_t seg = 0;
	int steps_left = 0;
	int state = 0; // 0 idle, 1 descending, 2 pending (leaf count in `count`)
	int count = 0;
	void start(const Trace *t) { tr = t; seg = 0; steps_left = t->segs[0].steps; state = 1; count = 0; }
	// one node step; returns true when the lane stops (leaf reached or ray ended)
	bool step(bool &finished)
	{
		finished = false;
		if(--steps_left > 0) return false;
		const Seg &s = tr->segs[seg];
		if(s.count == 0) { finished = true; state = 0; return true; }
		state = 2;
		count = s.count;
		return true;
	}
	// after the leaf tests of the current segment; returns true when the ray ended
	bool afterLeaf()
	{
		++seg;
		const Seg &s = tr->segs[seg];
		if(s.steps == 0 && s.count == 0) { state = 0; return true; }
		steps_left = s.steps;
		state = 1;
		return false;
	}
};

// Policy A: the shipped kernel.
static Tally policyA(const std::vector<Trace> &traces, const Costs &c, int refill = 8, int leaf_batch = 8, int k_steps = 16)
{
	Tally t;
	const size_t per_warp = 4096;
	for(size_t base = 0; base + per_warp <= traces.size(); base += per_warp)
	{
		size_t next = base;
		const size_t end = base + per_warp;
		RayCursor lane[32];
		for(;;)
		{
			int idle = 0;
			for(auto &l : lane) idle += l.state == 0;
			t.add(0, c.round_fixed, 32);
			if(next < end && idle >= refill)
			{
				int takers = 0, missed = 0;
				for(auto &l : lane)
					if(l.state == 0 && next < end)
					{
						const Trace &tr = traces[next++];
						++takers;
						++t.rays;
						if(tr.missed_bound) ++missed; else l.start(&tr);
					}
				t.add(0, c.setup + c.pool * takers / 256.0, takers);
				if(missed) t.add(3, c.write, missed);
			}
			int alive = 0, pending = 0;
			for(auto &l : lane) { alive += l.state != 0; pending += l.state == 2; }
			if(!alive) { if(next >= end) break; continue; }
			int finished = 0;
			if(pending >= leaf_batch || pending == alive)
			{
				int maxc = 0;
				for(auto &l : lane) if(l.state == 2) maxc = std::max(maxc, l.count);
				for(int k = 0; k < maxc; ++k)
				{
					int act = 0;
					for(auto &l : lane) act += (l.state == 2 && l.count > k);
					t.add(2, c.test, act);
				}
				t.add(2, c.leaf_fixed, pending);
				for(auto &l : lane) if(l.state == 2) finished += l.afterLeaf();
			}
			else
			{
				bool go[32];
				int going = 0, stopped = 0;
				for(int i = 0; i < 32; ++i) { go[i] = lane[i].state == 1; going += go[i]; }
				const int started = going;
				for(int s = 0; s < k_steps && going; ++s)
				{
					t.add(1, c.step, going);
					for(int i = 0; i < 32; ++i)
						if(go[i])
						{
							bool fin;
							if(lane[i].step(fin)) { go[i] = false; --going; ++stopped; finished += fin; }
						}
				}
				t.add(1, c.descent_fixed, started);
			}
			if(finished) t.add(3, c.write, finished);
		}
	}
	return t;
}

// Policy A2: the shipped kernel with a CONVERGED setup: when a lane is idle and the warp's ready queue is empty, all 32 lanes set
// up the next 32 rays at once into a shared-memory queue (rays that miss the bound are answered there); idle lanes then take a
// ready ray every round (`take_q` instructions).
static Tally policyA2(const std::vector<Trace> &traces, const Costs &c, int leaf_batch, int k_steps, double take_q, int take_at = 1)
{
	Tally t;
	const size_t per_warp = 4096;
	for(size_t base = 0; base + per_warp <= traces.size(); base += per_warp)
	{
		size_t next = base;
		const size_t end = base + per_warp;
		RayCursor lane[32];
		std::vector<const Trace *> queue;
		for(;;)
		{
			int idle = 0;
			for(auto &l : lane) idle += l.state == 0;
			t.add(0, c.round_fixed, 32);
			if(idle >= take_at && queue.empty() && next < end)
			{
				int missed = 0, taken = 0;
				for(int k = 0; k < 32 && next < end; ++k)
				{
					const Trace &tr = traces[next++];
					++t.rays;
					++taken;
					if(tr.missed_bound) ++missed; else queue.push_back(&tr);
				}
				t.add(0, c.setup + 20 + c.pool * taken / 256.0, taken); // + queue stores
				if(missed) t.add(3, c.write, missed);
			}
			if(idle >= take_at && !queue.empty())
			{
				int takers = 0;
				for(auto &l : lane)
					if(l.state == 0 && !queue.empty()) { l.start(queue.back()); queue.pop_back(); ++takers; }
				t.add(0, take_q, takers);
			}
			int alive = 0, pending = 0;
			for(auto &l : lane) { alive += l.state != 0; pending += l.state == 2; }
			if(!alive) { if(next >= end && queue.empty()) break; continue; }
			int finished = 0;
			if(pending >= leaf_batch || pending == alive)
			{
				int maxc = 0;
				for(auto &l : lane) if(l.state == 2) maxc = std::max(maxc, l.count);
				for(int k = 0; k < maxc; ++k)
				{
					int act = 0;
					for(auto &l : lane) act += (l.state == 2 && l.count > k);
					t.add(2, c.test, act);
				}
				t.add(2, c.leaf_fixed, pending);
				for(auto &l : lane) if(l.state == 2) finished += l.afterLeaf();
			}
			else
			{
				bool go[32];
				int going = 0;
				for(int i = 0; i < 32; ++i) { go[i] = lane[i].state == 1; going += go[i]; }
				const int started = going;
				for(int s = 0; s < k_steps && going; ++s)
				{
					t.add(1, c.step, going);
					for(int i = 0; i < 32; ++i)
						if(go[i])
						{
							bool fin;
							if(lane[i].step(fin)) { go[i] = false; --going; finished += fin; }
						}
				}
				t.add(1, c.descent_fixed, started);
			}
			if(finished) t.add(3, c.write, finished);
		}
	}
	return t;
}

// Policy B: two rays per lane (slot 0 active in registers, slot 1 parked); at every round boundary a lane may swap.
static Tally policyB(const std::vector<Trace> &traces, const Costs &c, int leaf_lanes, int setup_lanes, int k_steps)
{
	Tally t;
	const size_t per_warp = 4096;
	for(size_t base = 0; base + per_warp <= traces.size(); base += per_warp)
	{
		size_t next = base;
		const size_t end = base + per_warp;
		RayCursor slot[32][2];
		for(;;)
		{
			t.add(0, c.round_fixed + 6, 32);
			int nD = 0, nP = 0, nI = 0, total_alive = 0;
			for(auto &l : slot)
			{
				nD += (l[0].state == 1 || l[1].state == 1);
				nP += (l[0].state == 2 || l[1].state == 2);
				nI += (l[0].state == 0 || l[1].state == 0);
				total_alive += (l[0].state != 0) + (l[1].state != 0);
			}
			const bool can_fill = next < end;
			if(!total_alive && !can_fill) break;
			int phase; // 0 setup 1 descent 2 leaf
			if(nP >= leaf_lanes) phase = 2;
			else if(can_fill && nI >= setup_lanes) phase = 0;
			else if(nD) phase = 1;
			else if(nP) phase = 2;
			else phase = 0;
			const int want = phase == 0 ? 0 : phase;
			int swaps = 0;
			for(auto &l : slot)
				if(l[0].state != want && l[1].state == want) { std::swap(l[0], l[1]); ++swaps; }
			if(swaps) t.add(4, c.swap, 32);
			int finished = 0;
			if(phase == 0)
			{
				int takers = 0, missed = 0;
				for(auto &l : slot)
					if(l[0].state == 0 && next < end)
					{
						const Trace &tr = traces[next++];
						++takers;
						++t.rays;
						if(tr.missed_bound) ++missed; else l[0].start(&tr);
					}
				t.add(0, c.setup + c.pool * takers / 256.0, takers);
				if(missed) t.add(3, c.write, missed);
			}
			else if(phase == 2)
			{
				int maxc = 0, pend = 0;
				for(auto &l : slot) if(l[0].state == 2) { maxc = std::max(maxc, l[0].count); ++pend; }
				for(int k = 0; k < maxc; ++k)
				{
					int act = 0;
					for(auto &l : slot) act += (l[0].state == 2 && l[0].count > k);
					t.add(2, c.test, act);
				}
				t.add(2, c.leaf_fixed, pend);
				for(auto &l : slot) if(l[0].state == 2) finished += l[0].afterLeaf();
			}
			else
			{
				bool go[32];
				int going = 0;
				for(int i = 0; i < 32; ++i) { go[i] = slot[i][0].state == 1; going += go[i]; }
				const int started = going;
				for(int s = 0; s < k_steps && going; ++s)
				{
					t.add(1, c.step, going);
					for(int i = 0; i < 32; ++i)
						if(go[i])
						{
							bool fin;
							if(slot[i][0].step(fin)) { go[i] = false; --going; finished += fin; }
						}
				}
				t.add(1, c.descent_fixed, started);
			}
			if(finished) t.add(3, c.write, finished);
		}
	}
	return t;
}

// Policy C: a pool of `n_slots` rays per warp; each round up to 32 rays in the phase's state are processed, whichever they are.
static Tally policyC(const std::vector<Trace> &traces, const Costs &c, int n_slots, int leaf_lanes, int setup_lanes, int k_steps)
{
	Tally t;
	const size_t per_warp = 4096;
	std::vector<RayCursor> slot(n_slots);
	for(size_t base = 0; base + per_warp <= traces.size(); base += per_warp)
	{
		size_t next = base;
		const size_t end = base + per_warp;
		for(auto &s : slot) s = RayCursor();
		for(;;)
		{
			t.add(0, c.round_fixed + 10, 32);
			int nD = 0, nP = 0, nI = 0;
			for(auto &s : slot) { nD += s.state == 1; nP += s.state == 2; nI += s.state == 0; }
			const bool can_fill = next < end;
			if(nD + nP == 0 && !can_fill) break;
			int phase;
			if(nP >= leaf_lanes) phase = 2;
			else if(can_fill && nI >= setup_lanes) phase = 0;
			else if(nD) phase = 1;
			else if(nP) phase = 2;
			else phase = 0;
			std::vector<RayCursor *> pick;
			for(auto &s : slot) if(s.state == phase && int(pick.size()) < 32) pick.push_back(&s);
			int finished = 0;
			if(phase == 0)
			{
				int takers = 0, missed = 0;
				for(auto *s : pick)
					if(next < end)
					{
						const Trace &tr = traces[next++];
						++takers;
						++t.rays;
						if(tr.missed_bound) ++missed; else s->start(&tr);
					}
				t.add(0, c.setup + c.pool * takers / 256.0, takers);
				if(missed) t.add(3, c.write, missed);
			}
			else if(phase == 2)
			{
				int maxc = 0;
				for(auto *s : pick) maxc = std::max(maxc, s->count);
				for(int k = 0; k < maxc; ++k)
				{
					int act = 0;
					for(auto *s : pick) act += s->count > k;
					t.add(2, c.test, act);
				}
				t.add(2, c.leaf_fixed, double(pick.size()));
				t.add(4, c.take, double(pick.size()));
				for(auto *s : pick) finished += s->afterLeaf();
			}
			else
			{
				std::vector<bool> go(pick.size(), true);
				int going = int(pick.size());
				const int started = going;
				t.add(4, c.take, started);
				for(int s = 0; s < k_steps && going; ++s)
				{
					t.add(1, c.step, going);
					for(size_t i = 0; i < pick.size(); ++i)
						if(go[i])
						{
							bool fin;
							if(pick[i]->step(fin)) { go[i] = false; --going; finished += fin; }
						}
				}
				t.add(1, c.descent_fixed, started);
			}
			if(finished) t.add(3, c.write, finished);
		}
	}
	return t;
}

int main(int argc, char **argv)
{
	const std::string dir = argc > 1 ? argv[1] : "/tmp/simt";
	const bool shadow = argc > 2 && std::string(argv[2]) == "shadow";
	na = load<uint32_t>(dir + "/a.bin");
	nb = load<uint32_t>(dir + "/b.bin");
	refs = load<uint32_t>(dir + "/refs.bin");
	xyz = load<float>(dir + "/xyz.bin");
	idx = load<uint32_t>(dir + "/idx.bin");
	rays = load<float>(dir + (shadow ? "/srays.bin" : "/rays.bin"));
	const auto bd = load<float>(dir + "/bound.bin");
	std::copy(bd.begin(), bd.end(), bound);
	const size_t n = rays.size() / 8;
	std::vector<Trace> traces(n);
	double steps = 0, leaves = 0, tests = 0, missed = 0;
	for(size_t i = 0; i < n; ++i)
	{
		traces[i] = traceRay(&rays[8 * i], shadow);
		missed += traces[i].missed_bound;
		for(const Seg &s : traces[i].segs) { steps += s.steps; leaves += s.count != 0; tests += s.count; }
	}
	printf("%zu %s rays: %.2f node steps, %.2f non-empty leaves, %.2f primitive tests per ray; %.1f %% miss the bound\n", n, shadow ? "shadow" : "closest", steps / n, leaves / n, tests / n, 100 * missed / n);
	Costs c;
	policyA(traces, c).print("A shipped (refill 8, leaf 8, 16 steps)");
	policyA(traces, c, 4, 8, 16).print("A refill 4");
	policyA(traces, c, 8, 12, 16).print("A leaf batch 12");
	policyA2(traces, c, 8, 16, 26).print("A2 converged setup, take at 1");
	policyA2(traces, c, 8, 16, 26, 4).print("A2 converged setup, take at 4");
	policyA2(traces, c, 6, 16, 26, 2).print("A2 take at 2, leaf 6");
	policyA2(traces, c, 10, 16, 26, 2).print("A2 take at 2, leaf 10");
	policyA2(traces, c, 8, 8, 26, 2).print("A2 take at 2, 8 steps");
	policyA2(traces, c, 8, 24, 26, 2).print("A2 take at 2, 24 steps");
	for(int k : {8, 16})
		for(int ll : {12, 20, 28})
			for(int sl : {12, 20, 28})
			{
				char name[64];
				snprintf(name, sizeof name, "B 2 rays/lane steps %d leaf %d setup %d", k, ll, sl);
				policyB(traces, c, ll, sl, k).print(name);
			}
	for(int slots : {48, 64, 96})
		for(int k : {4, 8, 16})
		{
			char name[64];
			snprintf(name, sizeof name, "C pool %d steps %d (leaf/setup at 24)", slots, k);
			policyC(traces, c, slots, 24, 24, k).print(name);
		}
	return 0;
}
