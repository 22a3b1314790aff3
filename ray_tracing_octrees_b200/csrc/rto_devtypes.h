// rto_devtypes.h -- device scene descriptors and layout constants shared by the kernels (rto_kernels.cuh) and by the
// translation units that build those layouts on the device (rto_build.cu).  No kernels in here.
#pragma once
#include "rto_internal.h"
#include <cuda_runtime.h>

namespace rto {

// ------------------------------------------------------------------------------------------------
// Device scene descriptors (passed by value as kernel parameters)
// ------------------------------------------------------------------------------------------------
struct BvhDev {
	const float4* nodes;     // inner nodes, 4 x float4 each: [lo0.xyz hi0.x][hi0.yz lo1.xy][lo1.z hi1.xyz][ref0 ref1 - -]
	const float4* tris;      // 4 x float4 (64 B) per triangle in leaf order: [v0.xyz e1.x][e1.yz e2.xy][e2.z id lo.xy][lo.z hi.xyz], e1 = v1 - v0,
	                         // e2 = v2 - v0, lo/hi = the exact box of the REFERENCE leaf the triangle belongs to
	int   rootRef;           // >= 0: inner node index; < 0: ~leafRef, leafRef = (firstPos << 1) | (count - 1)
	int   numTris;
	const float4* exactNodes; // the tree rays take when they are not admitted to the fused tests (rto_kernels.cuh bvh_fused_ok): the reference-
	int   exactRoot;          // shaped tree with exact boxes and the reference's leaves (host route), or this tree itself (device route)
	int   exactLeafBox;       // leafBox of that tree
	float grow;              // > 0: every box of this tree below the root is a conservative one (leaf boxes grown by `grow` on each side), so
	                         // the node tests may use fused multiply-adds (rto_kernels.cuh slab_oct); 0: exact boxes, exact tests only
	int   leafBox;           // 1: the tree's leaves are single triangles under (inflated) boxes of their own, and a triangle is a candidate
	                         // only if the box of its reference leaf (in its record) passes intersectAABB: tested at the leaf.
	                         // 0: the tree's leaves are the reference's leaves and their boxes were tested in the parent node.
	float rootLo[3], rootHi[3];
	int   paired;            // node layout of `nodes`: 0 = child 0's box then child 1's ([lo0.xyz hi0.x][hi0.yz lo1.xy][lo1.z hi1.xyz]); 1 = the two children
	                         // interleaved plane by plane ([lo0.x lo1.x lo0.y lo1.y][lo0.z lo1.z hi0.x hi1.x][hi0.y hi1.y hi0.z hi1.z]), so that the same plane of both
	                         // children sits in one 64-bit register pair and one packed FFMA2 (fma.rn.f32x2, sm_100) tests it for both
	int   exactPaired;       // the same for `exactNodes`
	// the production tree once more, collapsed into 4-wide nodes with 16-bit boxes on one grid (rto_internal.h rto_build_wide_topology):
	// 4 x uint4 per node [x0..x3][y0..y3][z0..z3][ref0..ref3], plane = wideLo[axis] + q * wideStep.  Null: the scene has no wide form.
	const uint4* wide;
	int   wideRoot;
	float wideLo[3], wideStep;
};

struct OctDev {
	const uint32_t* desc;    // compact layout: per node, bit31 = leaf, bit30 = solid, else index of first child (8 contiguous)
	const int32_t*  up;      // compact layout: parent node of sibling group g = (node - 1) >> 3
	const int4*     nodes16; // general layout: RtoGpuNode padded to 16 x int32
	const int4*     inner;   // compact layout, one 16-byte record per INTERNAL node in BFS order (rank): x = index of first child,
	                         // y = rank of the first internal child, z = rank of the parent, w = leafMask | solidMask<<8 | childIdx<<16
	int   numNodes;
	int   rootSize;
	int   compact;
	float gmin[3];
	float voxel;
};

// render kernels: one warp = one 4 x 8 pixel tile, a block = kRenderThreads / 32 tiles side by side (16 x 8 pixels at 128 threads)
#ifndef RTO_RENDER_THREADS
#define RTO_RENDER_THREADS 128
#endif
constexpr int kRenderThreads = RTO_RENDER_THREADS;
constexpr int kRenderBlockW = kRenderThreads / 8;      // pixels per block row
constexpr uint32_t kOctLeaf = 0x80000000u;
constexpr uint32_t kOctSolid = 0x40000000u;
constexpr int kMaxOctDepth = 32;
constexpr int kWideStackDev = 144;  // == kWideStack (rto_internal.h): the builder refuses trees whose walks could need more
constexpr int kBvhStack = 96;       // deepest tree the builders hand out: 92 levels (host SAH falls back to a balanced tree, device LBVH refuses)
constexpr float kMissT = 1e30f;
constexpr float kBelowMissT = 9.99999940e29f;    // the largest float below 1e30f (rto_init checks the bit pattern)
constexpr float kMinPositive = 1.40129846e-45f;  // the smallest positive (denormal) float, bit pattern 0x00000001 (rto_init checks)
// pruning margin of the ordered BVH traversal: a subtree is skipped only if its box entry distance exceeds the
// best hit by more than this relative slack (keeps co-planar / shared-edge candidates)
constexpr float kPruneSlack = 1.00001f;

} // namespace rto
