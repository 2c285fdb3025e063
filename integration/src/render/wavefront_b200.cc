/* integration/src/render/wavefront_b200.cc -- see render/wavefront_b200.h. */
#include "render/wavefront_b200.h"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/mman.h>
#include <unistd.h>

namespace yafaray::b200 {

thread_local RayQueue *RayQueue::current_ = nullptr;

// ---- context switch -----------------------------------------------------------------------------------------------
#if defined(__x86_64__)
// System V x86-64: the callee-saved state is rbx, rbp, r12-r15, the SSE control word and the x87 control word; everything
// else is dead across a call.  b200_fiber_switch(&save, load) pushes that state, stores rsp in *save, adopts `load` as the
// stack pointer and pops the other context's state; its `ret` continues wherever that context called the switch (or, for a
// fresh fiber, in RayQueue::entry).  About 20 instructions, no system call (swapcontext does a sigprocmask per switch).
extern "C" void b200_fiber_switch(void **save_sp, void *load_sp);
asm(R"(
	.text
	.p2align 4
	.globl b200_fiber_switch
	.type b200_fiber_switch, @function
b200_fiber_switch:
	pushq %rbp
	pushq %rbx
	pushq %r12
	pushq %r13
	pushq %r14
	pushq %r15
	subq $8, %rsp
	stmxcsr (%rsp)
	fnstcw 4(%rsp)
	movq %rsp, (%rdi)
	movq %rsi, %rsp
	ldmxcsr (%rsp)
	fldcw 4(%rsp)
	addq $8, %rsp
	popq %r15
	popq %r14
	popq %r13
	popq %r12
	popq %rbx
	popq %rbp
	ret
	.size b200_fiber_switch, .-b200_fiber_switch
	.section .note.GNU-stack,"",@progbits
	.text
)");

namespace {
void *prepareStack(void *stack_top, void (*entry)())
{
	// layout the first switch pops: [mxcsr|fpcw][r15][r14][r13][r12][rbx][rbp][entry][0 = return address of entry]
	auto *sp = reinterpret_cast<uint64_t *>(reinterpret_cast<uintptr_t>(stack_top) & ~uintptr_t(15));
	*--sp = 0;                                  // fake return address of the entry function (it never returns); rsp = 8 (mod 16) at the entry, as after a call
	*--sp = reinterpret_cast<uint64_t>(entry);
	for(int k = 0; k < 6; ++k) *--sp = 0;       // rbp, rbx, r12..r15
	uint32_t mxcsr;
	uint16_t fpcw;
	asm volatile("stmxcsr %0" : "=m"(mxcsr));
	asm volatile("fnstcw %0" : "=m"(fpcw));
	*--sp = uint64_t(mxcsr) | (uint64_t(fpcw) << 32);
	return sp;
}
} //namespace
#define B200_SWITCH(save, load) b200_fiber_switch(save, load)
#else
// Other architectures: not implemented (the GPU hosts this targets are x86-64); RayQueue::run() reports the error and the
// caller falls back to the reference's own per-ray loop, which still traces on the GPU, one ray per call.
namespace {
void *prepareStack(void *, void (*)()) { return nullptr; }
}
#define B200_SWITCH(save, load) ((void) (save), (void) (load))
#endif

// ---- construction ---------------------------------------------------------------------------------------------------
namespace {
constexpr size_t kOutSize[3] = {sizeof(b200rt_hit), sizeof(uint32_t), sizeof(b200rt_tshadow)};

void *pinned(size_t bytes, std::string &error)
{
	void *p = nullptr;
	if(b200rt_host_alloc(&p, bytes) != B200RT_OK) { if(error.empty()) error = std::string("b200rt_host_alloc: ") + b200rt_last_error(); return nullptr; }
	return p;
}
} //namespace

RayQueue::RayQueue(int n_fibers, int n_groups, size_t stack_bytes) : n_fibers_{std::max(1, n_fibers)}
{
	const size_t page = size_t(sysconf(_SC_PAGESIZE));
	stack_bytes = ((std::max(stack_bytes, size_t(64) << 10) + page - 1) / page) * page;
	fibers_.resize(size_t(n_fibers_));
	// ONE mapping for the stacks of all fibers, touched lazily (only the pages a fiber's deepest recursion reached are ever
	// resident), one guard page below the lowest stack.  A mapping and a guard page per fiber cost two system calls each, all of
	// them serialised on the process's address-space lock while sixteen render threads create their queues at the same moment:
	// 12.6 % of the CPU samples of a one-frame SPPM render (profiles/r4e_*).  Between neighbouring stacks a canary takes the
	// guard page's place: checked whenever a fiber parks a ray or ends; an overflow aborts with a message (raise
	// "wavefront_stack_kb") instead of faulting.
	stacks_bytes_ = size_t(n_fibers_) * stack_bytes + page;
	void *m = mmap(nullptr, stacks_bytes_, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE | MAP_STACK, -1, 0);
	if(m == MAP_FAILED) { stacks_ = nullptr; error_ = "mmap of the fiber stacks failed"; return; }
	stacks_ = static_cast<char *>(m);
	mprotect(stacks_, page, PROT_NONE);
	for(size_t i = 0; i < fibers_.size(); ++i)
	{
		fibers_[i].stack = stacks_ + page + i * stack_bytes;
		fibers_[i].stack_bytes = stack_bytes;
		*static_cast<uint64_t *>(fibers_[i].stack) = kStackCanary;
	}
	n_groups = std::max(1, std::min(n_groups, n_fibers_));
	groups_.resize(size_t(n_groups));
	// rounded up to a multiple of 4 slots: every array below then starts 16-byte aligned whatever the fiber / group counts are
	// (32-byte rays, 144- and 16-byte records, 4-byte shadow answers), which the kernels' float4 loads and stores require
	const size_t fibers_per_group = (size_t(n_fibers_) + size_t(n_groups) - 1) / size_t(n_groups);
	const size_t per_group = (fibers_per_group + 3) & ~size_t(3);
	// one pinned slab for the rays and answers of all groups (a pinned allocation costs the driver a fraction of a millisecond and
	// sixteen render threads create their queues at the same moment)
	const size_t per_slot = 3 * sizeof(b200rt_ray) + kOutSize[0] + kOutSize[1] + kOutSize[2] + 3 * sizeof(float); // rays, answers, ray times
	slab_ = static_cast<char *>(pinned(size_t(n_groups) * per_group * per_slot + 64, error_));
	char *cursor = slab_;
	for(size_t g = 0; g < groups_.size(); ++g)
	{
		Group &group = groups_[g];
		for(size_t i = g * fibers_per_group; i < std::min(size_t(n_fibers_), (g + 1) * fibers_per_group); ++i)
		{
			fibers_[i].group = &group;
			group.fibers.push_back(&fibers_[i]);
		}
		group.capacity = uint32_t(per_group);
		group.parked.reserve(group.capacity);
		for(int kind = 0; kind < 3 && slab_; ++kind)
		{
			group.rays[kind] = reinterpret_cast<b200rt_ray *>(cursor);
			cursor += per_group * sizeof(b200rt_ray);
		}
		for(int kind = 2; kind >= 0 && slab_; --kind) // 144-byte records first: keeps every array 16-byte aligned
		{
			group.outs[kind] = cursor;
			cursor += per_group * kOutSize[kind];
		}
		for(int kind = 0; kind < 3 && slab_; ++kind) // Ray::time_ of every parked ray (motion blur), 4-byte values: 16-byte aligned too
		{
			group.times[kind] = reinterpret_cast<float *>(cursor);
			cursor += per_group * sizeof(float);
		}
		for(auto &r : group.requests) r.resize(group.capacity);
	}
	resuming_.reserve(per_group);
}

RayQueue::~RayQueue()
{
	for(Group &group : groups_)
	{
		if(group.flying && group.flight) b200rt_trace_jobs_end(group.flight);
		for(int kind = 0; kind < 3; ++kind)
		{
			b200rt_host_free(group.sorted_rays[kind]);
			b200rt_host_free(group.sorted_times[kind]);
			b200rt_host_free(group.sorted_out[kind]);
		}
	}
	b200rt_host_free(slab_);
	if(stacks_) munmap(stacks_, stacks_bytes_);
}

// ---- fibers ---------------------------------------------------------------------------------------------------------
void RayQueue::entry()
{
	RayQueue *q = current_;
	Fiber *self = q->running_;
	(*q->body_)();
	checkStack(*self);
	self->done = true;
	B200_SWITCH(&self->sp, q->scheduler_sp_); // never resumed
	__builtin_trap();
}

void RayQueue::checkStack(const Fiber &fiber)
{
	if(*static_cast<const uint64_t *>(fiber.stack) == kStackCanary) return;
	std::fprintf(stderr, "b200 wavefront ray queue: a fiber overflowed its %zu KiB stack (accelerator parameter \"wavefront_stack_kb\")\n", fiber.stack_bytes >> 10);
	std::abort();
}

void RayQueue::resume(Fiber &fiber)
{
	running_ = &fiber;
	++stats_.switches;
	if(!fiber.started)
	{
		fiber.started = true;
		fiber.done = false;
		fiber.sp = prepareStack(static_cast<char *>(fiber.stack) + fiber.stack_bytes, &RayQueue::entry);
	}
	B200_SWITCH(&scheduler_sp_, fiber.sp);
	running_ = nullptr;
}

void RayQueue::park(int kind, b200rt_scene *scene, const b200rt_ray &ray, float time, int max_depth)
{
	Fiber *self = running_;
	checkStack(*self);
	Group &group = *self->group;
	const uint32_t slot = group.count[kind]++;
	group.rays[kind][slot] = ray;
	group.times[kind][slot] = time;
	group.requests[kind][slot] = {scene, max_depth};
	self->slot = slot;
	group.parked.push_back(self);
	++stats_.switches;
	B200_SWITCH(&self->sp, scheduler_sp_); // back in run(); returns here once the group's flight has landed
}

b200rt_hit RayQueue::closest(b200rt_scene *scene, const b200rt_ray &ray, float time)
{
	park(B200RT_QUERY_CLOSEST, scene, ray, time, 0);
	return static_cast<const b200rt_hit *>(running_->group->outs[B200RT_QUERY_CLOSEST])[running_->slot];
}

uint32_t RayQueue::shadow(b200rt_scene *scene, const b200rt_ray &ray, float time)
{
	park(B200RT_QUERY_SHADOW, scene, ray, time, 0);
	return static_cast<const uint32_t *>(running_->group->outs[B200RT_QUERY_SHADOW])[running_->slot];
}

const b200rt_tshadow &RayQueue::transparentShadow(b200rt_scene *scene, const b200rt_ray &ray, float time, int max_depth)
{
	park(B200RT_QUERY_TSHADOW, scene, ray, time, max_depth);
	return static_cast<const b200rt_tshadow *>(running_->group->outs[B200RT_QUERY_TSHADOW])[running_->slot];
}

// ---- flights --------------------------------------------------------------------------------------------------------
// Hand the parked rays of a group to libb200rt (b200rt_trace_jobs_begin returns as soon as the kernels are enqueued).
void RayQueue::submit(Group &group)
{
	const auto t0 = std::chrono::steady_clock::now();
	++stats_.batches;
	b200rt_job jobs[3 * kMaxGroupsPerKind];
	size_t n_jobs = 0;
	const unsigned flags = B200RT_RAYS_TREE_SPACE | B200RT_BUFFERS_PINNED;
	for(int kind = 0; kind < 3; ++kind)
	{
		const uint32_t n = group.count[kind];
		group.in_flight[kind] = n;
		group.mixed[kind] = false;
		group.count[kind] = 0;
		if(n == 0) continue;
		stats_.rays[kind] += n;
		const std::vector<Request> &req = group.requests[kind];
		bool uniform = true;
		for(uint32_t i = 1; i < n && uniform; ++i) uniform = req[i].scene == req[0].scene && req[i].max_depth == req[0].max_depth;
		if(uniform)
		{
			jobs[n_jobs++] = {req[0].scene, kind, flags, group.rays[kind], n, group.outs[kind], req[0].max_depth, group.times[kind]};
			continue;
		}
		// several scenes / shadow depths: one job per (scene, depth), rays gathered into the pinned scratch of the kind
		if(!group.sorted_rays[kind])
		{
			group.sorted_rays[kind] = static_cast<b200rt_ray *>(pinned(size_t(group.capacity) * sizeof(b200rt_ray), error_));
			group.sorted_times[kind] = static_cast<float *>(pinned(size_t(group.capacity) * sizeof(float), error_));
			group.sorted_out[kind] = pinned(size_t(group.capacity) * kOutSize[kind], error_);
		}
		std::vector<uint32_t> &order = group.order[kind];
		order.resize(n);
		for(uint32_t i = 0; i < n; ++i) order[i] = i;
		std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return req[a].scene != req[b].scene ? req[a].scene < req[b].scene : req[a].max_depth < req[b].max_depth; });
		size_t pairs = 1;
		for(uint32_t i = 1; i < n; ++i) pairs += (req[order[i]].scene != req[order[i - 1]].scene || req[order[i]].max_depth != req[order[i - 1]].max_depth) ? 1 : 0;
		if(!group.sorted_rays[kind] || !group.sorted_out[kind] || !group.sorted_times[kind] || pairs > kMaxGroupsPerKind)
		{
			// no CPU path: these rays read as misses / unshadowed, and the render reports the error
			if(error_.empty()) error_ = "more distinct (scene, shadow depth) pairs in one flight than the ray queue supports";
			std::memset(group.outs[kind], kind == B200RT_QUERY_TSHADOW ? 0 : 0xFF, size_t(n) * kOutSize[kind]);
			if(kind == B200RT_QUERY_CLOSEST) for(uint32_t i = 0; i < n; ++i) static_cast<b200rt_hit *>(group.outs[kind])[i] = b200rt_hit{0.f, 0.f, 0.f, B200RT_MISS};
			continue;
		}
		group.mixed[kind] = true;
		for(uint32_t i = 0; i < n; ++i) { group.sorted_rays[kind][i] = group.rays[kind][order[i]]; group.sorted_times[kind][i] = group.times[kind][order[i]]; }
		for(uint32_t first = 0; first < n;)
		{
			uint32_t last = first + 1;
			const Request &key = req[order[first]];
			while(last < n && req[order[last]].scene == key.scene && req[order[last]].max_depth == key.max_depth) ++last;
			jobs[n_jobs++] = {key.scene, kind, flags, group.sorted_rays[kind] + first, last - first, static_cast<char *>(group.sorted_out[kind]) + size_t(first) * kOutSize[kind], key.max_depth, group.sorted_times[kind] + first};
			first = last;
		}
	}
	stats_.calls += n_jobs;
	group.flight = nullptr;
	if(b200rt_trace_jobs_begin(jobs, n_jobs, &group.flight) != B200RT_OK && error_.empty()) error_ = std::string("b200rt_trace_jobs_begin: ") + b200rt_last_error();
	group.flying = true;
	stats_.trace_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Wait for a group's flight; afterwards the answers are where its fibers will read them.
void RayQueue::land(Group &group)
{
	const auto t0 = std::chrono::steady_clock::now();
	bool failed = group.flight == nullptr;
	if(group.flight && b200rt_trace_jobs_end(group.flight) != B200RT_OK)
	{
		if(error_.empty()) error_ = std::string("b200rt_trace_jobs_end: ") + b200rt_last_error();
		failed = true;
	}
	group.flight = nullptr;
	group.flying = false;
	for(int kind = 0; kind < 3; ++kind)
	{
		const uint32_t n = group.in_flight[kind];
		if(n == 0) continue;
		if(failed)
		{
			// no CPU path: libb200rt has written nothing reliable, every ray of the flight reads as a miss / unshadowed
			std::memset(group.outs[kind], kind == B200RT_QUERY_TSHADOW ? 0 : 0xFF, size_t(n) * kOutSize[kind]);
			if(kind == B200RT_QUERY_CLOSEST) for(uint32_t i = 0; i < n; ++i) static_cast<b200rt_hit *>(group.outs[kind])[i] = b200rt_hit{0.f, 0.f, 0.f, B200RT_MISS};
		}
		else if(group.mixed[kind])
		{
			const std::vector<uint32_t> &order = group.order[kind];
			for(uint32_t i = 0; i < n; ++i) std::memcpy(static_cast<char *>(group.outs[kind]) + size_t(order[i]) * kOutSize[kind], static_cast<char *>(group.sorted_out[kind]) + size_t(i) * kOutSize[kind], kOutSize[kind]);
		}
	}
	stats_.trace_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

bool RayQueue::run(const std::function<void()> &body)
{
	if(!ok()) return false;
#if !defined(__x86_64__)
	error_ = "the fiber switch of this build is x86-64 only";
	return false;
#endif
	const auto t0 = std::chrono::steady_clock::now();
	RayQueue *outer = current_;
	current_ = this;
	body_ = &body;
	for(Fiber &f : fibers_) f.started = f.done = false;
	// start the groups one after the other: while the first flight is on its way the next group's fibers begin their pixels
	int started = 0;
	bool work_left = true;
	for(Group &group : groups_)
	{
		group.parked.clear();
		for(Fiber *f : group.fibers)
		{
			if(!work_left) break;
			resume(*f);
			++started;
			if(f->done) work_left = false; // its body found no work at all: no further fiber would either
		}
		if(!group.parked.empty()) submit(group);
	}
	stats_.peak_fibers = std::max(stats_.peak_fibers, started);
	// round robin: land a group's flight, resume its fibers until each has parked again or finished, send the new rays off
	for(bool any = true; any;)
	{
		any = false;
		for(Group &group : groups_)
		{
			if(!group.flying) continue;
			any = true;
			land(group);
			resuming_.swap(group.parked);
			group.parked.clear();
			// A resumed fiber finds its stack frames and its answer cold (512 fibers per thread do not fit the core's caches, and the
			// answers were written by the GPU): pull the next fiber's top frames and its answer slot in while this one shades.
			for(size_t i = 0; i < resuming_.size(); ++i)
			{
				if(i + 1 < resuming_.size())
				{
					const Fiber *next = resuming_[i + 1];
					const char *top = static_cast<const char *>(next->sp);
					for(int line = 0; line < 8; ++line) __builtin_prefetch(top + 64 * line);
					for(int kind = 0; kind < 3; ++kind)
						if(next->slot < group.in_flight[kind]) __builtin_prefetch(static_cast<const char *>(group.outs[kind]) + size_t(next->slot) * kOutSize[kind]);
				}
				resume(*resuming_[i]);
			}
			if(!group.parked.empty()) submit(group);
		}
	}
	body_ = nullptr;
	current_ = outer;
	stats_.run_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	return ok();
}

} //namespace yafaray::b200
