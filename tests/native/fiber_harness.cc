// tests/native/fiber_harness.cc -- CPU-only exercise of the wavefront ray queue (integration/src/render/wavefront_b200.cc)
// against a STUB of the five libb200rt entry points it uses (malloc instead of pinned memory; "tracing" a ray returns a
// value computed from the ray).  TEST INFRASTRUCTURE: checks the fiber switch (callee-saved registers, SSE/x87 control
// words, deep recursion on fiber stacks, exceptions inside a fiber), the batching and the mixed-scene grouping on
// several OS threads at once.  Prints "ok <rays> <batches> <calls>" and exits 0 on success.
#include "render/wavefront_b200.h"
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <thread>
#include <vector>

static std::atomic<uint64_t> g_calls{0}, g_rays{0};
static thread_local const char *g_err = "";

extern "C" {
int b200rt_host_alloc(void **ptr, size_t bytes) { *ptr = std::malloc(bytes ? bytes : 1); return *ptr ? B200RT_OK : B200RT_E_MEMORY; }
int b200rt_host_free(void *ptr) { std::free(ptr); return B200RT_OK; }
const char *b200rt_last_error(void) { return g_err; }
// the stub scene pointer is just a tag: results encode it so that a mix-up between groups is caught
int b200rt_trace(b200rt_scene *scene, int query, unsigned flags, const b200rt_ray *rays, size_t n, void *out, int max_depth)
{
	if(!(flags & B200RT_RAYS_TREE_SPACE)) return B200RT_E_INVALID;
	++g_calls;
	g_rays += n;
	const uint32_t tag = uint32_t(reinterpret_cast<uintptr_t>(scene));
	for(size_t i = 0; i < n; ++i)
	{
		const b200rt_ray &r = rays[i];
		if(query == B200RT_QUERY_CLOSEST) static_cast<b200rt_hit *>(out)[i] = b200rt_hit{r.ox * 2.f + 1.f, r.oy, r.oz, tag + uint32_t(r.tmax)};
		else if(query == B200RT_QUERY_SHADOW) static_cast<uint32_t *>(out)[i] = tag * 1000u + uint32_t(r.ox);
		else
		{
			b200rt_tshadow t{};
			t.shadowed = 0; t.n_transparent = uint32_t(max_depth); t.occluder = tag; t.transparent[0].t = r.ox;
			static_cast<b200rt_tshadow *>(out)[i] = t;
		}
	}
	return B200RT_OK;
}
int b200rt_trace_jobs(const b200rt_job *jobs, size_t n_jobs)
{
	for(size_t j = 0; j < n_jobs; ++j)
	{
		if(!(jobs[j].flags & B200RT_BUFFERS_PINNED)) return B200RT_E_INVALID;
		const int rc = b200rt_trace(jobs[j].scene, jobs[j].query, jobs[j].flags, jobs[j].rays, jobs[j].n, jobs[j].out, jobs[j].max_depth);
		if(rc != B200RT_OK) return rc;
	}
	return B200RT_OK;
}
struct b200rt_flight { int rc; };
int b200rt_trace_jobs_begin(const b200rt_job *jobs, size_t n_jobs, b200rt_flight **flight)
{
	*flight = new b200rt_flight{b200rt_trace_jobs(jobs, n_jobs)};
	return (*flight)->rc;
}
int b200rt_trace_jobs_end(b200rt_flight *flight) { const int rc = flight->rc; delete flight; return rc; }
}

using yafaray::b200::RayQueue;

static std::atomic<int> g_failures{0};
#define CHECK(cond) do { if(!(cond)) { ++g_failures; std::fprintf(stderr, "CHECK failed line %d: %s\n", __LINE__, #cond); } } while(0)

// recursive "integrator": one closest query per level, shadow queries in between, floating point live across switches
static double shade(RayQueue &q, b200rt_scene *scene_a, b200rt_scene *scene_b, int job, int depth)
{
	volatile char pad[1500]; // frames of the size the reference's integrate() has, to use the fiber stacks
	pad[0] = char(job); pad[1499] = char(depth);
	const float x = float(job % 1000) + 0.25f * float(depth);
	b200rt_scene *scene = (job % 7 == 0) ? scene_b : scene_a;
	const b200rt_ray ray{x, float(job), float(depth), 0.f, 0.f, 0.f, 1.f, float(depth)};
	const long double keep = std::sqrt(static_cast<long double>(job) + 2.0L); // x87 state across the switch
	const b200rt_hit h = q.closest(scene, ray, 0.f);
	CHECK(h.t == x * 2.f + 1.f);
	CHECK(h.u == float(job) && h.v == float(depth));
	CHECK(h.prim == uint32_t(reinterpret_cast<uintptr_t>(scene)) + uint32_t(depth));
	CHECK(keep == std::sqrt(static_cast<long double>(job) + 2.0L));
	double sum = h.t;
	if(job % 3 == 0)
	{
		const uint32_t occ = q.shadow(scene, ray, 0.f);
		CHECK(occ == uint32_t(reinterpret_cast<uintptr_t>(scene)) * 1000u + uint32_t(x));
		sum += occ;
	}
	if(job % 5 == 0)
	{
		const int md = 1 + job % 4;
		const b200rt_tshadow &t = q.transparentShadow(scene, ray, 0.f, md);
		CHECK(t.n_transparent == uint32_t(md) && t.occluder == uint32_t(reinterpret_cast<uintptr_t>(scene)) && t.transparent[0].t == x);
	}
	if(job % 11 == 0 && depth == 2)
	{
		try { throw std::runtime_error("inside a fiber"); } catch(const std::runtime_error &) { sum += 1.0; } // unwinding on a fiber stack
	}
	CHECK(pad[0] == char(job) && pad[1499] == char(depth));
	return depth > 0 ? sum + shade(q, scene_a, scene_b, job, depth - 1) : sum;
}

int main(int argc, char **argv)
{
	const int n_threads = argc > 1 ? std::atoi(argv[1]) : 4, n_fibers = argc > 2 ? std::atoi(argv[2]) : 256, n_jobs = argc > 3 ? std::atoi(argv[3]) : 20000, depth = argc > 4 ? std::atoi(argv[4]) : 40, n_groups = argc > 5 ? std::atoi(argv[5]) : 2;
	std::atomic<int> next{0};
	std::atomic<uint64_t> batches{0}, calls{0}, rays{0};
	std::vector<std::thread> threads;
	for(int t = 0; t < n_threads; ++t)
		threads.emplace_back([&]() {
			RayQueue q{n_fibers, n_groups, size_t(256) << 10};
			CHECK(q.ok());
			CHECK(RayQueue::current() == nullptr);
			auto *a = reinterpret_cast<b200rt_scene *>(uintptr_t(17)), *b = reinterpret_cast<b200rt_scene *>(uintptr_t(23));
			const bool ok = q.run([&]() {
				CHECK(RayQueue::current() == &q);
				for(int job; (job = next++) < n_jobs;) shade(q, a, b, job, depth);
			});
			CHECK(ok);
			CHECK(RayQueue::current() == nullptr);
			batches += q.stats().batches; calls += q.stats().calls;
			rays += q.stats().rays[0] + q.stats().rays[1] + q.stats().rays[2];
			CHECK(q.stats().peak_fibers <= n_fibers);
		});
	for(auto &t : threads) t.join();
	CHECK(rays.load() == g_rays.load());
	CHECK(calls.load() == g_calls.load());
	// with enough jobs the batches must be close to full: n_fibers rays in flight per thread
	if(n_jobs >= 4 * n_threads * n_fibers) CHECK(double(rays.load()) / double(batches.load()) > 0.5 * n_fibers / n_groups);
	if(g_failures.load()) { std::printf("FAILED %d checks\n", g_failures.load()); return 1; }
	std::printf("ok %llu %llu %llu\n", (unsigned long long) rays.load(), (unsigned long long) batches.load(), (unsigned long long) calls.load());
	return 0;
}
