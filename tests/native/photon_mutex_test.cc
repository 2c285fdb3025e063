// tests/native/photon_mutex_test.cc -- b200::PhotonMutex (integration/include/render/photon_mutex_b200.h) on the CPU:
// every tuple is drawn under the lock (its four components come from the same step of the four sequences), no tuple is handed
// out twice, in std::mutex mode and in spin mode (batched per OS thread), and the mode can be switched between passes.
#include "render/photon_mutex_b200.h"
#include <algorithm>
#include <cstdio>
#include <thread>
#include <vector>

struct Sequence
{
	int next = 0;
	float getNext() { return float(next++); } // exact up to 2^24 draws
};

static bool pass(yafaray::b200::PhotonMutex &mutex, Sequence *seq, int n_threads, int per_thread, bool spin, int &drawn_before)
{
	mutex.spin(spin);
	std::vector<std::vector<float>> got(n_threads);
	std::vector<std::thread> threads;
	bool consistent = true;
	for(int t = 0; t < n_threads; ++t) threads.emplace_back([&, t]() {
		for(int k = 0; k < per_thread; ++k)
		{
			float a, b, c, d;
			mutex.draw(seq[0], seq[1], seq[2], seq[3], a, b, c, d);
			if(a != b || a != c || a != d) consistent = false;
			got[t].push_back(a);
		}
	});
	for(auto &th : threads) th.join();
	mutex.spin(false);
	std::vector<float> all;
	for(auto &g : got) all.insert(all.end(), g.begin(), g.end());
	std::sort(all.begin(), all.end());
	const bool unique = std::adjacent_find(all.begin(), all.end()) == all.end();
	const bool in_range = !all.empty() && all.front() >= float(drawn_before) && all.back() < float(seq[0].next);
	// std::mutex mode: exactly the tuples [before, before + n) ; spin mode: a subset of what was drawn (a thread may leave part of its last batch unused)
	const bool complete = spin || int(all.size()) == seq[0].next - drawn_before;
	drawn_before = seq[0].next;
	return consistent && unique && in_range && complete && seq[0].next == seq[1].next && seq[0].next == seq[3].next;
}

int main()
{
	yafaray::b200::PhotonMutex mutex;
	Sequence seq[4];
	int before = 0;
	const bool a = pass(mutex, seq, 8, 5000, false, before);
	const bool b = pass(mutex, seq, 8, 5000, true, before);
	const bool c = pass(mutex, seq, 3, 1000, false, before);
	const bool d = pass(mutex, seq, 16, 777, true, before);
	std::printf("%s mutex=%d spin=%d mutex-again=%d spin-again=%d drawn=%d\n", (a && b && c && d) ? "ok" : "FAILED", a, b, c, d, seq[0].next);
	return (a && b && c && d) ? 0 : 1;
}
