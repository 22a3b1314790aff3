// rto_internal.h -- declarations shared by the translation units of librto.so (not part of the ABI).
#pragma once
#include "../../include/rto_c.h"
#include "rto_math.h"
#include <cstdint>
#include <vector>

// Records a thread-local message for rto_last_error() and returns `code`.
int rto_fail(int code, const char* fmt, ...);

// ---- host BVH (shape of the reference's BVHNode tree, BVH.h:37-42, stored as a pre-order array) -------------
struct HostBvhNode {
	float    mn[3], mx[3];
	int32_t  left, right;      // node indices; -1/-1 for a leaf
	uint32_t first, count;     // leaf: range in RtoHostBvh::order (count <= 2); internal: count == 0
};

struct RtoHostBvh {
	const RtoTriangle*       tris = nullptr;   // caller-owned (BVH.cpp:21-25 keeps raw pointers too)
	size_t                   numTris = 0;
	std::vector<HostBvhNode> nodes;            // pre-order: left subtree directly after its parent
	std::vector<uint32_t>    order;            // triangle ids in leaf (depth-first, left-to-right) order
};
