// rto_internal.h -- declarations shared by the translation units of librto.so (not part of the ABI).
#pragma once
#include "../../include/rto_c.h"
#include "rto_math.h"
#include <cstdint>
#include <exception>
#include <new>
#include <system_error>
#include <thread>
#include <vector>

// Records a thread-local message for rto_last_error() and returns `code`.
int rto_fail(int code, const char* fmt, ...);

// Nothing may throw across the C ABI: every extern "C" entry point that allocates or starts threads is a function-try-block ending in this.
#define RTO_CATCH_ALL(name) \
	catch (const std::bad_alloc&) { return rto_fail(RTO_ERR_ALLOC, name ": out of host memory"); } \
	catch (const std::exception& e_) { return rto_fail(RTO_ERR_ALLOC, name ": %s", e_.what()); } \
	catch (...) { return rto_fail(RTO_ERR_ALLOC, name ": unknown failure"); }

// The two halves of a divide-and-conquer step, the first on a new thread when `parallel` and a thread can be had (otherwise both on
// the calling thread).  An exception on either side surfaces on the calling thread after BOTH halves have finished -- never through
// the destructor of a joinable std::thread.
template <class A, class B> void rto_fork_join(bool parallel, A&& a, B&& b) {
	if (!parallel) { a(); b(); return; }
	std::exception_ptr ea, eb;
	std::thread th;
	bool started = false;
	try { th = std::thread([&] { try { a(); } catch (...) { ea = std::current_exception(); } }); started = true; }
	catch (const std::system_error&) {}
	try { if (!started) a(); b(); } catch (...) { eb = std::current_exception(); }
	if (started) th.join();
	if (ea) std::rethrow_exception(ea);
	if (eb) std::rethrow_exception(eb);
}
// fn(t) for t in [0, n) on n threads (the last one is the caller); threads that cannot be created run on the caller, exceptions are
// rethrown after every thread has been joined.
template <class F> void rto_run_threads(int n, F&& fn) {
	std::vector<std::thread> th;
	std::vector<std::exception_ptr> err((size_t)(n > 0 ? n : 0));
	int started = 0;
	for (; started + 1 < n; started++) {
		const int t = started;
		try { th.emplace_back([&, t] { try { fn(t); } catch (...) { err[t] = std::current_exception(); } }); }
		catch (const std::system_error&) { break; }
	}
	for (int t = started; t < n; t++) { try { fn(t); } catch (...) { err[t] = std::current_exception(); } }
	for (auto& x : th) x.join();
	for (auto& e : err) if (e) std::rethrow_exception(e);
}

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

// ---- device node arrays built from the host BVH (host_builders.cpp) ------------------------------------------
// Node layout (16 floats): child0 lo xyz, hi xyz; child1 lo xyz, hi xyz; ref0, ref1 (int bits); 2 pad.
// A ref >= 0 is an inner node index, < 0 is ~((firstPos << 1) | (count - 1)) for a reference leaf whose triangles sit at
// positions firstPos.. in leaf order.
//   reference topology: the reference's own tree (pre-order) -- needed to replay BVH::query's visit order and box counts;
//   fast topology     : a binned-SAH tree whose leaves are SINGLE triangles under slightly inflated boxes of their own (refs
//                       ~(pos << 1)); the traversal re-tests the exact box of the triangle's reference leaf (stored in the triangle
//                       record) before Moller-Trumbore.  The candidate set of BVH::query depends only on which reference-leaf boxes
//                       pass the slab test (a box that passes makes every enclosing box pass, by monotonic rounding), so a triangle
//                       the reference tests is lost only if the ray misses the triangle's own inflated box -- and then Moller-
//                       Trumbore rejects it too; the visit order changes, and the closest-hit rule is order independent.  Tight
//                       per-triangle boxes cut the Moller-Trumbore tests (the block that runs with the fewest lanes) by a third.
void rto_build_reference_topology(const RtoHostBvh& h, std::vector<float>& nodeBuf, int32_t& rootRef);
void rto_build_fast_topology(const RtoHostBvh& h, std::vector<float>& nodeBuf, int32_t& rootRef, float& grow);
// The fast topology collapsed into 4-wide nodes with boxes quantised to 16 bits per plane on ONE grid for the whole scene:
// 64 bytes per node instead of 64 bytes per PAIR of children, half as many dependent fetches per ray.  Node layout (16 words):
// [x0 x1 x2 x3][y0 y1 y2 y3][z0 z1 z2 z3][ref0 ref1 ref2 ref3], word k of an axis = lo | hi << 16 of child k in grid units
// (plane = lo[axis] + q * step); refs as in the binary arrays, an unused slot holds an inverted box and kWideEmpty.  Every quantised box
// contains the (already grown) box it stands for with two grid steps to spare on every side, which covers the rounding of the
// kernels' dequantising node test (rto_kernels.cuh wide_test).  Returns false (no wide tree) when the tree is a
// single leaf or when a walk could need more postponed entries than the kernels' stack holds.
bool rto_build_wide_topology(const std::vector<float>& fastNodes, int32_t fastRoot, const float rootLo[3], const float rootHi[3], float grow,
	std::vector<uint32_t>& wide, int32_t& wideRoot, float wideLo[3], float& wideStep);
constexpr int32_t kWideEmpty = (int32_t)0x80000000;   // ref of an unused child slot (never a leaf reference: ~it >> 1 is beyond 2^30 triangles)
constexpr int kWideStack = 144;

// ---- device layouts, built on the host (host_layouts.cpp) and uploaded verbatim by rto_device.cu --------------
struct OctLayout {
	bool compact = false;
	bool isTree = false;            // every node but the root is the child of exactly one node (what createOctreeFromVoxelGrid and the frustum
	                                // cull produce); the octreeRaySkip walk, a recursion over a pointer TREE in the reference, is defined only then
	size_t numLeaves = 0;
	std::vector<uint32_t> desc;     // compact: 7 pad words + one word per node (bit31 leaf, bit30 solid, else first child)
	std::vector<int32_t>  up;       // compact: parent node of sibling group (node - 1) >> 3
	std::vector<int32_t>  inner;    // compact: 4 ints per internal node (first child, rank of first internal child, parent rank,
	                                //          leafMask | solidMask << 8 | own octant << 16), ranked in BFS order
	std::vector<int32_t>  padded;   // general: RtoGpuNode padded to 16 ints
};
int rto_build_octree_layout(const RtoGpuNode* nodes, size_t numNodes, OctLayout& out);

struct BvhLayout {
	std::vector<uint32_t> wideNodes;                  // 16 words per 4-wide node (rto_build_wide_topology); empty: no wide tree
	int32_t wideRoot = -1;
	float wideLo[3] = { 0, 0, 0 }, wideStep = 0.0f;
	std::vector<float> refNodes, fastNodes, tris;     // 16 floats per inner node; 16 floats per triangle in leaf order: v0, v1 - v0, v2 - v0, id in [9], box of its reference leaf in [10..15]
	int32_t refRoot = -1, fastRoot = -1;
	float fastGrow = 0.0f;                             // how far the fast topology's leaf boxes were grown (BvhDev::grow)
	float rootLo[3] = { 0, 0, 0 }, rootHi[3] = { 0, 0, 0 };
};
void rto_build_bvh_layout(const RtoHostBvh& h, BvhLayout& out);

// ---- CSV voxeliser (BuildingLoader.cpp:35-290), host part shared by the host and the device fill (host_builders.cpp) ----------
struct CsvScene {
	std::vector<RtoTriangle> tris;      // one per face whose three vertices exist, in file order (vertices cast double -> float)
	int   dims[3] = { 0, 0, 0 };
	float gridMin[3] = { 0, 0, 0 };
	float voxelSize = 0;
};
// Parses both files and derives the grid geometry exactly like loadCSVDataIntoVoxelGrid.  dims == 0 when either file is empty
// or unreadable (the reference returns an empty VoxelGrid then).
int rto_csv_load(const char* vertsCsv, const char* facesCsv, float voxelSize, CsvScene& out);
// cell range of one face and the centre-in-triangle predicate, identical on host and device (rto_voxelize.h)
