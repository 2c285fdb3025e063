// tests/native/pm_host_model.cu -- TEST INFRASTRUCTURE: the photon-map lookup of libyafaray_b200/csrc/pm_kernels.cuh run on the
// CPU.  pmLookupOne is __host__ __device__; this file calls the very same function for one "thread" after the other, with the
// shared-memory heap of MODE 0 laid out as in a block of kPmThreads threads, so that tests/test_pm.py can check the kernel's
// logic (stack handling, the restated libstdc++ heap, result layout) against the reference without a GPU.  Not a product path:
// nothing in libyafaray_b200/ builds or loads it.
#include "../../libyafaray_b200/csrc/pm_kernels.cuh"
#include <vector>

extern "C" int pm_model_lookup(int mode, const uint32_t *nodes4, const float *dirs4, const float *points, const float *normals, uint32_t n_points, uint32_t k,
                               float sq_radius, const float *sq_radii, uint32_t *found2, uint32_t *n_found, float *sq_radius_out, uint32_t *nearest, int round_steps, int single_pop)
{
	using namespace b200pm;
	const uint4 *nodes = reinterpret_cast<const uint4 *>(nodes4);
	const float4 *dirs = reinterpret_cast<const float4 *>(dirs4);
	uint2 *found = reinterpret_cast<uint2 *>(found2);
	std::vector<uint2> smem(size_t(mode == 0 ? k : 1) * kPmThreads);
	for(uint32_t point = 0; point < n_points; ++point)
	{
		const HeapSmem heap_s{smem.data() + point % kPmThreads};
		if(round_steps > 0)
		{
			// the phased state machine (pmStep / pmResolve), one lane
			if(mode == 0 && single_pop) pmLookupPhasedOne<0, true>(nodes, dirs, points, normals, point, k, sq_radius, sq_radii, found, n_found, sq_radius_out, nearest, heap_s, round_steps);
			else if(mode == 1 && single_pop) pmLookupPhasedOne<1, true>(nodes, dirs, points, normals, point, k, sq_radius, sq_radii, found, n_found, sq_radius_out, nearest, heap_s, round_steps);
			else if(mode == 2 && single_pop) pmLookupPhasedOne<2, true>(nodes, dirs, points, normals, point, k, sq_radius, sq_radii, found, n_found, sq_radius_out, nearest, heap_s, round_steps);
			else if(mode == 0) pmLookupPhasedOne<0, false>(nodes, dirs, points, normals, point, k, sq_radius, sq_radii, found, n_found, sq_radius_out, nearest, heap_s, round_steps);
			else if(mode == 1) pmLookupPhasedOne<1, false>(nodes, dirs, points, normals, point, k, sq_radius, sq_radii, found, n_found, sq_radius_out, nearest, heap_s, round_steps);
			else if(mode == 2) pmLookupPhasedOne<2, false>(nodes, dirs, points, normals, point, k, sq_radius, sq_radii, found, n_found, sq_radius_out, nearest, heap_s, round_steps);
			else return -1;
			continue;
		}
		if(mode == 0) pmLookupOne<0>(nodes, dirs, points, normals, point, k, sq_radius, sq_radii, found, n_found, sq_radius_out, nearest, heap_s);
		else if(mode == 1) pmLookupOne<1>(nodes, dirs, points, normals, point, k, sq_radius, sq_radii, found, n_found, sq_radius_out, nearest, heap_s);
		else if(mode == 2) pmLookupOne<2>(nodes, dirs, points, normals, point, k, sq_radius, sq_radii, found, n_found, sq_radius_out, nearest, heap_s);
		else return -1;
	}
	return 0;
}
