// libyafaray_b200/csrc/pm_build.h -- host-side build of the photon map's point kd-tree (include/b200pm.h).
// The tree IS the reference's (include/photon/pkdtree.h:144-218): the lookups' results depend on its shape.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace b200pm {

struct HostTree
{
	// node i in the reference's KdNode terms: b = flags_, a = split bits / photon index (b200pm.h, diagnostics)
	std::vector<uint32_t> a, b;
	uint32_t depth = 0;
};

// pos: 3 floats per photon; n >= 1.  threads <= 0: all hardware threads.
void buildTree(const float *pos, size_t n, int threads, HostTree &out);

} // namespace b200pm
