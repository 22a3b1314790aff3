// host_builders.cpp -- CPU-side scene construction of the product library (librto.so).
//
// These are the host halves of the reference's hot path, re-implemented for speed but producing the SAME
// data the reference produces (node sets, numbering, triangle order, tree shape), because hit ids are
// indices into those arrays:
//   rto_host_octree_build  == createOctreeFromVoxelGrid (OctreeVoxel.cpp:704-778) + setOctree BFS flatten
//                             (RayTracerBVH.cpp:443-490); built bottom-up from an occupancy pyramid instead
//                             of the reference's top-down full-region rescans.
//   rto_host_mc_mesh       == MarchingCubesRenderer::render over localMC (Renderer.cpp:14-36,
//                             OctreeVoxel.cpp:780-879); only the outer cell layers of a uniform leaf can
//                             straddle a sign change, so interior cells are skipped without changing order.
//   rto_host_bvh_build     == BVH::BVH / BVH::build (BVH.cpp:19-71); index-based, pre-order node array,
//                             subtrees built on worker threads, identical std::sort call sequence per range.
//   rto_host_camera_orbit  == Camera::getView/getPos (Camera.cpp:11-29) + glm::lookAtRH + glm::inverse.
// Compiled with -ffp-contract=off; arithmetic follows rto_math.h (glm operation order).
#include "rto_internal.h"
#include "rto_nvtx.h"
#include "mc_tables.h"
#include "rto_voxelize.h"
#include "rto_frustum.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <cmath>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

using namespace rto;

// =================================================================================================
// Octree: bottom-up uniformity pyramid + BFS emission
// =================================================================================================
namespace {

constexpr uint8_t kMixed = 0xFF;

struct Pyramid {
	// level l holds cells of edge 2^l voxels, clipped to the grid: dims ceil(dim / 2^l). Level 0 aliases the voxels.
	std::vector<std::vector<uint8_t>> store;
	std::vector<const uint8_t*> data;
	std::vector<int> nx, ny, nz;
	// state of the cell of edge 2^l at voxel origin (x, y, z); outside the grid everything is EMPTY (getVoxelSafe, OctreeVoxel.cpp:692-701)
	inline uint8_t at(int l, int x, int y, int z) const {
		int cx = x >> l, cy = y >> l, cz = z >> l;
		if (cx >= nx[l] || cy >= ny[l] || cz >= nz[l]) return 0;
		return data[l][(size_t)cx + (size_t)cy * nx[l] + (size_t)cz * ((size_t)nx[l] * ny[l])];
	}
};

void buildPyramid(const uint8_t* vox, int dx, int dy, int dz, int levels, Pyramid& P) {
	P.store.resize(levels + 1); P.data.resize(levels + 1); P.nx.resize(levels + 1); P.ny.resize(levels + 1); P.nz.resize(levels + 1);
	P.data[0] = vox; P.nx[0] = dx; P.ny[0] = dy; P.nz[0] = dz;
	for (int l = 1; l <= levels; l++) {
		int px = P.nx[l - 1], py = P.ny[l - 1], pz = P.nz[l - 1];
		int cx = (px + 1) / 2, cy = (py + 1) / 2, cz = (pz + 1) / 2;
		P.nx[l] = cx; P.ny[l] = cy; P.nz[l] = cz;
		P.store[l].assign((size_t)cx * cy * cz, 0);
		const uint8_t* src = P.data[l - 1];
		uint8_t* dst = P.store[l].data();
		auto slab = [&](int z0, int z1) {
			for (int z = z0; z < z1; z++)
				for (int y = 0; y < cy; y++)
					for (int x = 0; x < cx; x++) {
						uint8_t first = 0; bool have = false, mixed = false;
						for (int k = 0; k < 8 && !mixed; k++) {
							int sx = 2 * x + (k & 1), sy = 2 * y + ((k >> 1) & 1), sz = 2 * z + (k >> 2);
							uint8_t v = (sx < px && sy < py && sz < pz) ? src[(size_t)sx + (size_t)sy * px + (size_t)sz * ((size_t)px * py)] : 0;
							if (v == kMixed) mixed = true;
							else if (!have) { first = v; have = true; }
							else if (v != first) mixed = true;
						}
						dst[(size_t)x + (size_t)y * cx + (size_t)z * ((size_t)cx * cy)] = mixed ? kMixed : first;
					}
		};
		unsigned hw = std::max(1u, std::thread::hardware_concurrency());
		int nthreads = (int)std::min<size_t>(hw, std::max<size_t>(1, ((size_t)cx * cy * cz) >> 16));
		if (nthreads <= 1) slab(0, cz);
		else {
			rto_run_threads(nthreads, [&](int t) { slab((int)((int64_t)cz * t / nthreads), (int)((int64_t)cz * (t + 1) / nthreads)); });
		}
		P.data[l] = dst;
	}
}

} // namespace

extern "C" int rto_host_octree_build(const uint8_t* voxels, int dimX, int dimY, int dimZ, RtoGpuNode** nodesOut, size_t* numNodes) try {
	RTO_RANGE("rto_host_octree_build");
	if (!nodesOut || !numNodes) return rto_fail(RTO_ERR_INVALID, "rto_host_octree_build: null output");
	*nodesOut = nullptr; *numNodes = 0;
	// createOctreeFromVoxelGrid returns nullptr for an empty grid (OctreeVoxel.cpp:766): zero nodes, still RTO_OK
	if (dimX == 0 || dimY == 0 || dimZ == 0) return RTO_OK;
	if (!voxels || dimX < 0 || dimY < 0 || dimZ < 0) return rto_fail(RTO_ERR_INVALID, "rto_host_octree_build: bad grid");
	for (size_t i = 0, n = (size_t)dimX * dimY * dimZ; i < n; i++)
		if (voxels[i] == kMixed) return rto_fail(RTO_ERR_INVALID, "rto_host_octree_build: voxel value 255 is reserved");
	int maxDim = std::max({ dimX, dimY, dimZ });
	int root = 1, levels = 0;
	while (root < maxDim) { root <<= 1; levels++; }
	Pyramid P;
	buildPyramid(voxels, dimX, dimY, dimZ, levels, P);

	std::vector<RtoGpuNode> nodes;
	struct Pending { int level; };
	std::vector<int8_t> levelOf;              // level of node i (cell edge 2^level)
	nodes.reserve(1 << 16); levelOf.reserve(1 << 16);
	auto emit = [&](int x, int y, int z, int level) {
		RtoGpuNode n;
		n.x = x; n.y = y; n.z = z; n.size = 1 << level;
		uint8_t s = P.at(level, x, y, z);
		bool leaf = (level == 0) || (s != kMixed);
		n.isLeaf = leaf; n.isUniform = leaf; n.isSolid = leaf && (s == 1);
		for (int i = 0; i < 8; i++) n.child[i] = -1;
		nodes.push_back(n); levelOf.push_back((int8_t)level);
	};
	emit(0, 0, 0, levels);
	for (size_t i = 0; i < nodes.size(); i++) {   // the vector itself is the BFS queue
		if (nodes[i].isLeaf) continue;
		int level = levelOf[i] - 1, half = 1 << level;
		int x0 = nodes[i].x, y0 = nodes[i].y, z0 = nodes[i].z;
		for (int c = 0; c < 8; c++) {         // bit0 = x, bit1 = y, bit2 = z (OctreeVoxel.cpp:749-757)
			nodes[i].child[c] = (int32_t)nodes.size();
			emit(x0 + ((c & 1) ? half : 0), y0 + ((c & 2) ? half : 0), z0 + ((c & 4) ? half : 0), level);
		}
		if (nodes.size() > (size_t)0x7fffff00) return rto_fail(RTO_ERR_UNSUPPORTED, "rto_host_octree_build: more than 2^31 nodes");
	}
	RtoGpuNode* out = (RtoGpuNode*)std::malloc(nodes.size() * sizeof(RtoGpuNode));
	if (!out) return rto_fail(RTO_ERR_ALLOC, "rto_host_octree_build: out of memory");
	std::memcpy(out, nodes.data(), nodes.size() * sizeof(RtoGpuNode));
	*nodesOut = out; *numNodes = nodes.size();
	return RTO_OK;
} RTO_CATCH_ALL("rto_host_octree_build")

// =================================================================================================
// Marching cubes in the reference's emission order
// =================================================================================================
namespace {

struct McGrid {
	const uint8_t* vox; int dx, dy, dz; float minX, minY, minZ, vs;
	inline float scalar(int x, int y, int z) const {      // FILLED -> -1, EMPTY or outside -> +1 (OctreeVoxel.cpp:787-792)
		if (x < 0 || y < 0 || z < 0 || x >= dx || y >= dy || z >= dz) return 1.0f;
		return vox[(size_t)x + (size_t)y * dx + (size_t)z * ((size_t)dx * dy)] == 1 ? -1.0f : 1.0f;
	}
};

inline V3 vertexInterp(V3 p1, V3 p2, float v1, float v2) {    // iso level 0 (OctreeVoxel.cpp:633-640)
	if (std::fabs(0.0f - v1) < 0.00001f) return p1;
	if (std::fabs(0.0f - v2) < 0.00001f) return p2;
	if (std::fabs(v1 - v2) < 0.00001f) return p1;
	float mu = (0.0f - v1) / (v2 - v1);
	return p1 + mu * (p2 - p1);
}

inline void mcCell(const McGrid& g, int x, int y, int z, std::vector<RtoTriangle>& out) {
	static const int off[8][3] = { {0,0,0},{1,0,0},{1,1,0},{0,1,0},{0,0,1},{1,0,1},{1,1,1},{0,1,1} };
	float val[8]; int cube = 0;
	for (int c = 0; c < 8; c++) { val[c] = g.scalar(x + off[c][0], y + off[c][1], z + off[c][2]); if (val[c] < 0) cube |= 1 << c; }
	if (cube == 0 || cube == 255) return;
	int flags = mc_edge_flags(cube);
	V3 pos[8];
	for (int c = 0; c < 8; c++)
		pos[c] = mk3(g.minX + (x + off[c][0]) * g.vs, g.minY + (y + off[c][1]) * g.vs, g.minZ + (z + off[c][2]) * g.vs);
	V3 vert[12];
	for (int e = 0; e < 12; e++) if (flags & (1 << e)) {
		int a = kMcEdgeCorner[e][0], b = kMcEdgeCorner[e][1];
		vert[e] = vertexInterp(pos[a], pos[b], val[a], val[b]);
	}
	for (int i = 0; mc_tri_edge(cube, i) != -1; i += 3) {
		RtoTriangle t;
		V3 a = vert[mc_tri_edge(cube, i)], b = vert[mc_tri_edge(cube, i + 1)], c = vert[mc_tri_edge(cube, i + 2)];
		t.v0[0] = a.x; t.v0[1] = a.y; t.v0[2] = a.z; t.v1[0] = b.x; t.v1[1] = b.y; t.v1[2] = b.z; t.v2[0] = c.x; t.v2[1] = c.y; t.v2[2] = c.z;
		out.push_back(t);
	}
}

} // namespace

extern "C" int rto_host_mc_mesh(const uint8_t* voxels, int dimX, int dimY, int dimZ, const float gridMin[3], float voxelSize,
	const RtoGpuNode* nodes, size_t numNodes, RtoTriangle** trisOut, size_t* numTris) try {
	RTO_RANGE("rto_host_mc_mesh");
	if (!trisOut || !numTris) return rto_fail(RTO_ERR_INVALID, "rto_host_mc_mesh: null output");
	*trisOut = nullptr; *numTris = 0;
	if (numNodes == 0) return RTO_OK;
	if (!voxels || !nodes || !gridMin) return rto_fail(RTO_ERR_INVALID, "rto_host_mc_mesh: null input");
	McGrid g{ voxels, dimX, dimY, dimZ, gridMin[0], gridMin[1], gridMin[2], voxelSize };
	std::vector<RtoTriangle> out;
	// depth-first, children 0..7 in order: the recursion of MarchingCubesRenderer::render (Renderer.cpp:26-33)
	std::vector<int32_t> stack; stack.push_back(0);
	while (!stack.empty()) {
		int32_t i = stack.back(); stack.pop_back();
		if (i < 0 || (size_t)i >= numNodes) continue;
		const RtoGpuNode& n = nodes[i];
		if (!n.isLeaf) { for (int c = 7; c >= 0; c--) stack.push_back(n.child[c]); continue; }
		// localMC(grid, x0, y0, z0, size): cells of the leaf region clipped to dim-1 (OctreeVoxel.cpp:795-797).
		int x1 = std::min(n.x + n.size, dimX - 1), y1 = std::min(n.y + n.size, dimY - 1), z1 = std::min(n.z + n.size, dimZ - 1);
		int lastX = n.x + n.size - 1, lastY = n.y + n.size - 1, lastZ = n.z + n.size - 1;
		for (int z = n.z; z < z1; z++)
			for (int y = n.y; y < y1; y++) {
				if (z == lastZ || y == lastY || !n.isUniform) { for (int x = n.x; x < x1; x++) mcCell(g, x, y, z, out); }
				else if (lastX < x1) mcCell(g, lastX, y, z, out);   // interior cells of a uniform leaf see one sign only
			}
	}
	if (out.empty()) return RTO_OK;
	RtoTriangle* buf = (RtoTriangle*)std::malloc(out.size() * sizeof(RtoTriangle));
	if (!buf) return rto_fail(RTO_ERR_ALLOC, "rto_host_mc_mesh: out of memory");
	std::memcpy(buf, out.data(), out.size() * sizeof(RtoTriangle));
	*trisOut = buf; *numTris = out.size();
	return RTO_OK;
} RTO_CATCH_ALL("rto_host_mc_mesh")

// =================================================================================================
// BVH with the reference's shape
// =================================================================================================
namespace {

size_t subtreeNodes(size_t m) {            // nodes of the tree BVH::build makes for m triangles
	if (m <= 2) return 1;
	return 1 + subtreeNodes(m / 2) + subtreeNodes(m - m / 2);
}

struct Builder {
	const RtoTriangle* tris; std::vector<V3> cen; std::vector<uint32_t> idx; std::vector<HostBvhNode>* nodes;

	void build(size_t lo, size_t hi, size_t nodeIdx, int depth) {
		HostBvhNode& nd = (*nodes)[nodeIdx];
		const float M = std::numeric_limits<float>::max();
		V3 mn = mk3(M, M, M), mx = mk3(-M, -M, -M);
		for (size_t i = lo; i < hi; i++) {              // AABB::expand over triangle boxes (BVH.cpp:38-42); min/max are exact
			const RtoTriangle& t = tris[idx[i]];
			V3 tmn = mk3(M, M, M), tmx = mk3(-M, -M, -M);
			const float* v[3] = { t.v0, t.v1, t.v2 };
			for (int k = 0; k < 3; k++) { V3 p = mk3(v[k][0], v[k][1], v[k][2]); tmn = min3(tmn, p); tmx = max3(tmx, p); }
			mn = min3(mn, tmn); mx = max3(mx, tmn); mn = min3(mn, tmx); mx = max3(mx, tmx);
		}
		nd.mn[0] = mn.x; nd.mn[1] = mn.y; nd.mn[2] = mn.z; nd.mx[0] = mx.x; nd.mx[1] = mx.y; nd.mx[2] = mx.z;
		nd.left = nd.right = -1; nd.first = (uint32_t)lo; nd.count = (uint32_t)(hi - lo);
		size_t m = hi - lo;
		if (m <= 2) return;                             // leaf (BVH.cpp:46-49)
		V3 ext = mx - mn;
		int axis = 0;                                   // BVH.cpp:52-55
		if (ext.y > ext.x) axis = 1;
		if (comp(ext, 2) > comp(ext, axis)) axis = 2;
		const V3* c = cen.data();
		std::sort(idx.begin() + lo, idx.begin() + hi, [c, axis](uint32_t a, uint32_t b) { return comp(c[a], axis) < comp(c[b], axis); });   // BVH.cpp:58-60
		size_t mid = m / 2;                             // BVH.cpp:63
		size_t li = nodeIdx + 1, ri = li + subtreeNodes(mid);
		nd.left = (int32_t)li; nd.right = (int32_t)ri; nd.count = 0;
		// up to 32 subtrees in flight
		rto_fork_join(depth < 5 && m > 20000, [=] { build(lo, lo + mid, li, depth + 1); }, [=] { build(lo + mid, hi, ri, depth + 1); });
	}
};

} // namespace

extern "C" int rto_host_bvh_build(const RtoTriangle* tris, size_t numTris, RtoHostBvh** out) try {
	RTO_RANGE("rto_host_bvh_build");
	if (!out) return rto_fail(RTO_ERR_INVALID, "rto_host_bvh_build: null output");
	*out = nullptr;
	if (numTris && !tris) return rto_fail(RTO_ERR_INVALID, "rto_host_bvh_build: null triangles");
	if (numTris >= (size_t)1 << 30) return rto_fail(RTO_ERR_UNSUPPORTED, "rto_host_bvh_build: more than 2^30 triangles");
	RtoHostBvh* h = new (std::nothrow) RtoHostBvh();
	if (!h) return rto_fail(RTO_ERR_ALLOC, "rto_host_bvh_build: out of memory");
	h->tris = tris; h->numTris = numTris;
	h->nodes.resize(subtreeNodes(numTris));
	Builder b; b.tris = tris; b.nodes = &h->nodes;
	b.cen.resize(numTris); b.idx.resize(numTris);
	for (size_t i = 0; i < numTris; i++) {
		const RtoTriangle& t = tris[i];
		V3 s = mk3(t.v0[0], t.v0[1], t.v0[2]) + mk3(t.v1[0], t.v1[1], t.v1[2]) + mk3(t.v2[0], t.v2[1], t.v2[2]);
		b.cen[i] = s / 3.0f;                            // centroid(), BVH.cpp:15-17
		b.idx[i] = (uint32_t)i;
	}
	b.build(0, numTris, 0, 0);
	h->order.swap(b.idx);
	*out = h;
	return RTO_OK;
} RTO_CATCH_ALL("rto_host_bvh_build")

extern "C" void rto_host_bvh_free(RtoHostBvh* bvh) { delete bvh; }
extern "C" size_t rto_host_bvh_num_nodes(const RtoHostBvh* bvh) { return bvh ? bvh->nodes.size() : 0; }

extern "C" int rto_host_bvh_export(const RtoHostBvh* bvh, float* boxes6, int32_t* meta4, size_t capacity) try {
	if (!bvh || !boxes6 || !meta4) return rto_fail(RTO_ERR_INVALID, "rto_host_bvh_export: null argument");
	if (capacity < bvh->nodes.size()) return rto_fail(RTO_ERR_INVALID, "rto_host_bvh_export: capacity too small");
	for (size_t i = 0; i < bvh->nodes.size(); i++) {    // nodes are stored in pre-order already
		const HostBvhNode& n = bvh->nodes[i];
		std::memcpy(boxes6 + 6 * i, n.mn, 12); std::memcpy(boxes6 + 6 * i + 3, n.mx, 12);
		bool leaf = n.left < 0;
		meta4[4 * i] = leaf; meta4[4 * i + 1] = leaf ? (int32_t)n.count : 0;
		meta4[4 * i + 2] = (leaf && n.count > 0) ? (int32_t)bvh->order[n.first] : -1;
		meta4[4 * i + 3] = (leaf && n.count > 1) ? (int32_t)bvh->order[n.first + 1] : -1;
	}
	return RTO_OK;
} RTO_CATCH_ALL("rto_host_bvh_export")

// =================================================================================================
// Device node arrays
// =================================================================================================
static inline int32_t leafRefOf(const HostBvhNode& n) {
	uint32_t cnt = n.count ? n.count : 1;          // count 0 only for the empty tree, which is never referenced
	return ~(int32_t)((n.first << 1) | (cnt - 1));
}

void rto_build_reference_topology(const RtoHostBvh& h, std::vector<float>& nodeBuf, int32_t& rootRef) {
	std::vector<int32_t> innerId(h.nodes.size(), -1);
	int32_t numInner = 0;
	for (size_t i = 0; i < h.nodes.size(); i++) if (h.nodes[i].left >= 0) innerId[i] = numInner++;
	auto refOf = [&](int32_t i) { return h.nodes[i].left >= 0 ? innerId[i] : leafRefOf(h.nodes[i]); };
	nodeBuf.assign((size_t)std::max(numInner, 1) * 16, 0.0f);
	for (size_t i = 0; i < h.nodes.size(); i++) {
		const HostBvhNode& n = h.nodes[i];
		if (n.left < 0) continue;
		float* d = &nodeBuf[(size_t)innerId[i] * 16];
		const HostBvhNode& L = h.nodes[n.left]; const HostBvhNode& R = h.nodes[n.right];
		std::memcpy(d, L.mn, 12); std::memcpy(d + 3, L.mx, 12); std::memcpy(d + 6, R.mn, 12); std::memcpy(d + 9, R.mx, 12);
		int32_t r0 = refOf(n.left), r1 = refOf(n.right);
		std::memcpy(&d[12], &r0, 4); std::memcpy(&d[13], &r1, 4);
	}
	rootRef = h.numTris ? refOf(0) : -1;
}

namespace {
struct SahPrim { float mn[3], mx[3], c[3]; int32_t ref; };
struct SahBuilder {
	std::vector<SahPrim> prims; std::vector<float>* out;
	std::atomic<int> maxDepth{ 0 };
	bool medianOnly = false;        // balanced fallback: split every range in half along its widest centroid axis
	static inline float halfArea(const float* mn, const float* mx) {
		float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
		return dx * dy + dy * dz + dz * dx;
	}
	// Returns the ref of the subtree over prims[lo, hi) and its exact bounds.  A subtree over m leaves has m - 1 inner nodes, so the
	// pre-order position of every node is known in advance (root at `idx`, left subtree at idx + 1, right at idx + (mid - lo)) and
	// the top of the tree can be built by several threads into one preallocated array.
	int32_t build(size_t lo, size_t hi, size_t idx, int depth, float* bmn, float* bmx) {
		const float M = std::numeric_limits<float>::max();
		float cmn[3] = { M, M, M }, cmx[3] = { -M, -M, -M };
		for (int k = 0; k < 3; k++) { bmn[k] = M; bmx[k] = -M; }
		for (size_t i = lo; i < hi; i++)
			for (int k = 0; k < 3; k++) {
				bmn[k] = std::min(bmn[k], prims[i].mn[k]); bmx[k] = std::max(bmx[k], prims[i].mx[k]);
				cmn[k] = std::min(cmn[k], prims[i].c[k]); cmx[k] = std::max(cmx[k], prims[i].c[k]);
			}
		if (hi - lo == 1) return prims[lo].ref;
		{ int seen = maxDepth.load(std::memory_order_relaxed); while (depth + 1 > seen && !maxDepth.compare_exchange_weak(seen, depth + 1)) {} }
		constexpr int NB = 32;
		int bestAxis = -1, bestSplit = 0; float bestCost = M;
		for (int ax = 0; ax < 3 && !medianOnly; ax++) {
			float ext = cmx[ax] - cmn[ax];
			if (!(ext > 0.0f)) continue;
			float scale = NB / ext;
			int cnt[NB] = { 0 }; float bn[NB][3], bx[NB][3];
			for (int b = 0; b < NB; b++) for (int k = 0; k < 3; k++) { bn[b][k] = M; bx[b][k] = -M; }
			for (size_t i = lo; i < hi; i++) {
				int b = std::min(NB - 1, std::max(0, (int)((prims[i].c[ax] - cmn[ax]) * scale)));
				cnt[b]++;
				for (int k = 0; k < 3; k++) { bn[b][k] = std::min(bn[b][k], prims[i].mn[k]); bx[b][k] = std::max(bx[b][k], prims[i].mx[k]); }
			}
			float rightArea[NB]; int rightCnt[NB];
			float an[3] = { M, M, M }, axx[3] = { -M, -M, -M }; int c = 0;
			for (int b = NB - 1; b > 0; b--) {
				for (int k = 0; k < 3; k++) { an[k] = std::min(an[k], bn[b][k]); axx[k] = std::max(axx[k], bx[b][k]); }
				c += cnt[b]; rightCnt[b] = c; rightArea[b] = c ? halfArea(an, axx) : 0.0f;
			}
			for (int k = 0; k < 3; k++) { an[k] = M; axx[k] = -M; }
			c = 0;
			for (int b = 0; b < NB - 1; b++) {
				for (int k = 0; k < 3; k++) { an[k] = std::min(an[k], bn[b][k]); axx[k] = std::max(axx[k], bx[b][k]); }
				c += cnt[b];
				if (c == 0 || rightCnt[b + 1] == 0) continue;
				float cost = halfArea(an, axx) * c + rightArea[b + 1] * rightCnt[b + 1];
				if (cost < bestCost) { bestCost = cost; bestAxis = ax; bestSplit = b + 1; }
			}
		}
		size_t mid;
		if (medianOnly) {
			int ax = 0;
			if (cmx[1] - cmn[1] > cmx[ax] - cmn[ax]) ax = 1;
			if (cmx[2] - cmn[2] > cmx[ax] - cmn[ax]) ax = 2;
			mid = lo + (hi - lo) / 2;
			std::nth_element(prims.begin() + lo, prims.begin() + mid, prims.begin() + hi, [ax](const SahPrim& a, const SahPrim& b) { return a.c[ax] < b.c[ax]; });
		}
		else if (bestAxis < 0) mid = lo + (hi - lo) / 2;   // all centroids coincide: split the list in half
		else {
			float scale = NB / (cmx[bestAxis] - cmn[bestAxis]), base = cmn[bestAxis];
			auto it = std::partition(prims.begin() + lo, prims.begin() + hi, [&](const SahPrim& p) {
				return std::min(NB - 1, std::max(0, (int)((p.c[bestAxis] - base) * scale))) < bestSplit; });
			mid = (size_t)(it - prims.begin());
			if (mid == lo || mid == hi) mid = lo + (hi - lo) / 2;
		}
		float lmn[3], lmx[3], rmn[3], rmx[3];
		int32_t r0, r1;
		rto_fork_join(depth < 4 && hi - lo > 65536, [&] { r0 = build(lo, mid, idx + 1, depth + 1, lmn, lmx); }, [&] { r1 = build(mid, hi, idx + (mid - lo), depth + 1, rmn, rmx); });
		float* d = &(*out)[idx * 16];
		// paired layout (BvhDev::paired): plane k of child 0 and of child 1 side by side, k = lo x, lo y, lo z, hi x, hi y, hi z
		for (int k = 0; k < 3; k++) { d[2 * k] = lmn[k]; d[2 * k + 1] = rmn[k]; d[6 + 2 * k] = lmx[k]; d[6 + 2 * k + 1] = rmx[k]; }
		std::memcpy(&d[12], &r0, 4); std::memcpy(&d[13], &r1, 4);
		return (int32_t)idx;
	}
};
} // namespace

void rto_build_fast_topology(const RtoHostBvh& h, std::vector<float>& nodeBuf, int32_t& rootRef, float& growOut) {
	SahBuilder b; b.out = &nodeBuf;
	nodeBuf.clear();
	// one primitive per triangle, in reference leaf order (position p == tie-break rank of the closest-hit rule); its box is the
	// triangle's own box grown by 2^-18 of the scene's extent on every side (more costs shadow rays, which start 1e-3 voxels above a
	// surface and must not start inside its boxes; less would come too close to the rounding of the Moller-Trumbore test), far more than the rounding of the Moller-Trumbore
	// test can move a hit and than the fused node tests of the kernels can move a plane, far less than a
	// triangle
	float ext = 0.0f;
	if (!h.nodes.empty()) for (int k = 0; k < 3; k++) ext = std::max(ext, std::max(std::fabs(h.nodes[0].mn[k]), std::fabs(h.nodes[0].mx[k])));
	int lg = 18; if (const char* e = getenv("RTO_BVH_GROW_LOG2")) { int v = atoi(e); if (v >= 8 && v <= 22) lg = v; }      // tuning aid
	const float grow = std::ldexp(ext, -lg);
	growOut = grow;
	b.prims.resize(h.numTris);
	for (size_t p = 0; p < h.numTris; p++) {
		const RtoTriangle& t = h.tris[h.order[p]];
		SahPrim& q = b.prims[p];
		for (int k = 0; k < 3; k++) {
			float mn = std::min(t.v0[k], std::min(t.v1[k], t.v2[k])), mx = std::max(t.v0[k], std::max(t.v1[k], t.v2[k]));
			q.mn[k] = mn - grow; q.mx[k] = mx + grow; q.c[k] = 0.5f * mn + 0.5f * mx;
		}
		q.ref = ~(int32_t)((uint32_t)p << 1);
	}
	if (b.prims.empty()) { nodeBuf.assign(16, 0.0f); rootRef = -1; return; }
	nodeBuf.assign(std::max<size_t>(b.prims.size() - 1, 1) * 16, 0.0f);
	float mn[3], mx[3];
	rootRef = b.build(0, b.prims.size(), 0, 0, mn, mx);
	// the kernels keep at most kBvhStack (96) postponed subtrees: a SAH tree deeper than 92 levels (pathological inputs) is replaced
	// by a balanced median-split tree over the same primitives, whose depth is ceil(log2 n) <= 31
	int depthLimit = 92;
	if (const char* e = getenv("RTO_BVH_MAX_DEPTH")) { int v = atoi(e); if (v >= 4 && v < 92) depthLimit = v; }      // (tests force the fallback with it)
	if (b.maxDepth.load() > depthLimit) {
		b.medianOnly = true; b.maxDepth = 0;
		std::fill(nodeBuf.begin(), nodeBuf.end(), 0.0f);
		rootRef = b.build(0, b.prims.size(), 0, 0, mn, mx);
	}
}

// 4-wide collapse of the fast topology with 16-bit quantised boxes (rto_internal.h)
namespace {
struct WideBuilder {
	const float* bin; std::vector<uint32_t>* out;
	double lo[3], step, margin;
	int maxNeed = 0;
	bool offGrid = false;            // a box did not fit on the grid (cannot happen for trees of rto_build_fast_topology; checked, not assumed)
	struct Cand { int32_t ref; float mn[3], mx[3]; };
	static void childBox(const float* d, int k, Cand& c) {        // paired layout: plane p of child k at d[2 * p + k], p = lo xyz, hi xyz
		for (int a = 0; a < 3; a++) { c.mn[a] = d[2 * a + k]; c.mx[a] = d[6 + 2 * a + k]; }
		std::memcpy(&c.ref, &d[12 + k], 4);
	}
	static float area(const Cand& c) { float dx = c.mx[0] - c.mn[0], dy = c.mx[1] - c.mn[1], dz = c.mx[2] - c.mn[2]; return dx * dy + dy * dz + dz * dx; }
	// returns the wide index of binary node b; need = postponed entries a walk below it can hold at once
	int32_t build(int32_t b, int& need) {
		Cand c[4]; int n = 2;
		childBox(bin + (size_t)b * 16, 0, c[0]); childBox(bin + (size_t)b * 16, 1, c[1]);
		while (n < 4) {                                             // open the inner child with the largest box until four children or only leaves
			int pick = -1; float best = -1.0f;
			for (int i = 0; i < n; i++) if (c[i].ref >= 0 && area(c[i]) > best) { best = area(c[i]); pick = i; }
			if (pick < 0) break;
			const float* d = bin + (size_t)c[pick].ref * 16;
			childBox(d, 0, c[pick]); childBox(d, 1, c[n]); n++;
		}
		const size_t idx = out->size() / 16;
		out->resize(out->size() + 16);
		uint32_t w[16];
		int deepest = 0;
		for (int k = 0; k < 4; k++) {
			if (k >= n) { for (int a = 0; a < 3; a++) w[4 * a + k] = 0x0000ffffu; w[12 + k] = (uint32_t)kWideEmpty; continue; }      // lo = 65535, hi = 0: never hit
			for (int a = 0; a < 3; a++) {
				double ql = std::floor(((double)c[k].mn[a] - lo[a]) / step) - margin, qh = std::ceil(((double)c[k].mx[a] - lo[a]) / step) + margin;
				if (ql < 0.0 || qh > 65535.0) { offGrid = true; ql = std::max(0.0, ql); qh = std::min(65535.0, qh); }
				w[4 * a + k] = (uint32_t)ql | ((uint32_t)qh << 16);
			}
			int32_t ref = c[k].ref;
			if (ref >= 0) { int sub = 0; ref = build(ref, sub); deepest = std::max(deepest, sub); }
			w[12 + k] = (uint32_t)ref;
		}
		std::memcpy(&(*out)[idx * 16], w, sizeof(w));
		need = (n - 1) + deepest;
		maxNeed = std::max(maxNeed, need);
		return (int32_t)idx;
	}
};
} // namespace

bool rto_build_wide_topology(const std::vector<float>& fastNodes, int32_t fastRoot, const float rootLo[3], const float rootHi[3], float grow,
	std::vector<uint32_t>& wide, int32_t& wideRoot, float wideLo[3], float& wideStep) {
	wide.clear(); wideRoot = -1;
	if (fastRoot < 0 || fastNodes.empty()) return false;          // empty scene or a single triangle: nothing to collapse
	float ext = 0.0f;
	for (int a = 0; a < 3; a++) ext = std::max(ext, rootHi[a] - rootLo[a]);
	if (!(ext > 0.0f) || !(ext < 1e30f)) return false;
	// One grid for the whole scene: 65 536 positions along the longest side of the root box (whose corners are those of exact boxes)
	// widened by the growth of the leaf boxes and by the margin every quantised box gets on top of rounding outwards.  The margin has
	// to cover what the kernels' dequantising test (rto_kernels.cuh wide_test: t = fma(2^23 + q, step/d, ((lo - o)/d - 2^23 step/d)))
	// and the leaf-level test it must not contradict can be off by, for the rays admitted to the fused tests (|o| <= 16 e, e the largest
	// coordinate of the scene, grow = 64 e 2^-24): half a step from rounding the per-ray constant (2^23 step/d dominates it), and at
	// most (2 x 17 + 17 + 16 + 17) e 2^-24 = 1.31 grow from the roundings of (lo - o), of the products and of the leaf test.
	// (half a `grow` of slack at either end absorbs the rounding of wideLo itself: grow is 32 ulps of the largest coordinate)
	wideStep = (ext + 6.0f * grow) / 65400.0f;
	const double margin = 1.0 + std::ceil(1.5 * (double)grow / (double)wideStep);
	for (int a = 0; a < 3; a++) wideLo[a] = rootLo[a] - 3.0f * grow - 8.0f * wideStep;
	WideBuilder b; b.bin = fastNodes.data(); b.out = &wide; b.step = (double)wideStep; b.margin = margin;
	for (int a = 0; a < 3; a++) b.lo[a] = (double)wideLo[a];
	wide.reserve(fastNodes.size() / 2 + 16);
	int need = 0;
	wideRoot = b.build(fastRoot, need);
	if (getenv("RTO_DEBUG_WIDE")) fprintf(stderr, "[wide] ext %g grow %g step %g margin %g lo %g %g %g nodes %zu need %d offGrid %d\n", ext, grow, wideStep, margin, wideLo[0], wideLo[1], wideLo[2], wide.size() / 16, b.maxNeed, (int)b.offGrid);
	if (b.maxNeed > kWideStack - 4 || b.offGrid) { wide.clear(); wide.shrink_to_fit(); wideRoot = -1; return false; }
	return true;
}

// =================================================================================================
// Camera constants
// =================================================================================================
// glm::inverse(mat4): cofactor expansion, same grouping as glm/detail/func_matrix.inl compute_inverse<4,4>; column-major m[c * 4 + r]
static void glm_inverse4(const float* m, float* outInv) {
	auto M = [&](int c, int r) { return m[c * 4 + r]; };
	float c00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3), c02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3), c03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
	float c04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3), c06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3), c07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
	float c08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2), c10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2), c11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
	float c12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3), c14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3), c15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
	float c16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2), c18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2), c19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
	float c20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1), c22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1), c23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
	const float F0[4] = { c00, c00, c02, c03 }, F1[4] = { c04, c04, c06, c07 }, F2[4] = { c08, c08, c10, c11 };
	const float F3[4] = { c12, c12, c14, c15 }, F4[4] = { c16, c16, c18, c19 }, F5[4] = { c20, c20, c22, c23 };
	const float A0[4] = { M(1, 0), M(0, 0), M(0, 0), M(0, 0) }, A1[4] = { M(1, 1), M(0, 1), M(0, 1), M(0, 1) };
	const float A2[4] = { M(1, 2), M(0, 2), M(0, 2), M(0, 2) }, A3[4] = { M(1, 3), M(0, 3), M(0, 3), M(0, 3) };
	const float sgnA[4] = { 1.0f, -1.0f, 1.0f, -1.0f }, sgnB[4] = { -1.0f, 1.0f, -1.0f, 1.0f };
	float inv[16];
	for (int r = 0; r < 4; r++) {
		inv[0 + r] = ((A1[r] * F0[r] - A2[r] * F1[r]) + A3[r] * F2[r]) * sgnA[r];
		inv[4 + r] = ((A0[r] * F0[r] - A2[r] * F3[r]) + A3[r] * F4[r]) * sgnB[r];
		inv[8 + r] = ((A0[r] * F1[r] - A1[r] * F3[r]) + A3[r] * F5[r]) * sgnA[r];
		inv[12 + r] = ((A0[r] * F2[r] - A1[r] * F4[r]) + A2[r] * F5[r]) * sgnB[r];
	}
	float det = (M(0, 0) * inv[0] + M(0, 1) * inv[4]) + (M(0, 2) * inv[8] + M(0, 3) * inv[12]);
	float ood = 1.0f / det;
	for (int i = 0; i < 16; i++) outInv[i] = inv[i] * ood;
}

// glm::perspective(radians(fovDeg), aspect, zNear, zFar): matrix_clip_space.inl:249-262 (RH, depth -1..1)
static void glm_perspective(float fovDeg, float aspect, float zNear, float zFar, float* P) {
	const float fovy = fovDeg * 0.01745329251994329576923690768489f;      // glm::radians
	const float tanHalf = std::tan(fovy / 2.0f);
	for (int i = 0; i < 16; i++) P[i] = 0.0f;
	P[0] = 1.0f / (aspect * tanHalf);
	P[5] = 1.0f / tanHalf;
	P[10] = -(zFar + zNear) / (zFar - zNear);
	P[11] = -1.0f;
	P[14] = -(2.0f * zFar * zNear) / (zFar - zNear);
}
// glm mat4 * vec4: (m0 * v0 + m1 * v1) + (m2 * v2 + m3 * v3), type_mat4x4.inl:561-572
static void glm_mul_m4v4(const float* m, const float* v, float* out) {
	for (int r = 0; r < 4; r++) out[r] = (m[r] * v[0] + m[4 + r] * v[1]) + (m[8 + r] * v[2] + m[12 + r] * v[3]);
}

extern "C" int rto_host_camera_orbit(float theta, float phi, float radius, const float target[3], float fovDeg, float aspect,
	int width, int height, RtoCamera* out, float* view16) try {
	if (!out || !target) return rto_fail(RTO_ERR_INVALID, "rto_host_camera_orbit: null argument");
	if (width <= 0 || height <= 0) return rto_fail(RTO_ERR_INVALID, "rto_host_camera_orbit: bad image size");
	V3 tgt = mk3(target[0], target[1], target[2]);
	V3 eye = radius * mk3(cosf(theta) * sinf(phi), sinf(theta), cosf(theta) * cosf(phi)) + tgt;    // Camera.cpp:13-17
	// glm::lookAtRH(eye, target, (0,1,0))
	V3 f = normalize3(tgt - eye);
	V3 s = normalize3(cross3(f, mk3(0.0f, 1.0f, 0.0f)));
	V3 u = cross3(s, f);
	float m[16] = { s.x, u.x, -f.x, 0.0f,  s.y, u.y, -f.y, 0.0f,  s.z, u.z, -f.z, 0.0f,  -dot3(s, eye), -dot3(u, eye), dot3(f, eye), 1.0f };
	glm_inverse4(m, out->invView);
	out->camPos[0] = eye.x; out->camPos[1] = eye.y; out->camPos[2] = eye.z;
	float fovRad = fovDeg * 0.01745329251994329576923690768489f;     // glm::radians
	out->tanHalfFov = tanf(fovRad * 0.5f);
	out->aspect = aspect; out->width = width; out->height = height;
	if (view16) std::memcpy(view16, m, 64);
	return RTO_OK;
} RTO_CATCH_ALL("rto_host_camera_orbit")

// =================================================================================================
// sceneCache.bin (CacheUtils.cpp:5-59)
// =================================================================================================
extern "C" int rto_host_grid_load(const char* path, int dims[3], float minAndVoxel[4], uint8_t** voxelsOut) try {
	if (!path || !dims || !minAndVoxel || !voxelsOut) return rto_fail(RTO_ERR_INVALID, "rto_host_grid_load: null argument");
	*voxelsOut = nullptr;
	FILE* f = std::fopen(path, "rb");
	if (!f) return rto_fail(RTO_ERR_IO, "rto_host_grid_load: cannot open file");
	uint64_t n = 0;
	bool ok = std::fread(dims, 4, 3, f) == 3 && std::fread(minAndVoxel, 4, 4, f) == 4 && std::fread(&n, 8, 1, f) == 1;
	uint8_t* buf = nullptr;
	if (ok) {
		if (dims[0] < 0 || dims[1] < 0 || dims[2] < 0 || n != (uint64_t)dims[0] * dims[1] * dims[2]) ok = false;
		else { buf = (uint8_t*)std::malloc(n ? n : 1); ok = buf && std::fread(buf, 1, n, f) == n; }
	}
	std::fclose(f);
	if (!ok) { std::free(buf); return rto_fail(RTO_ERR_IO, "rto_host_grid_load: truncated or inconsistent file"); }
	*voxelsOut = buf;
	return RTO_OK;
} RTO_CATCH_ALL("rto_host_grid_load")

extern "C" int rto_host_grid_save(const char* path, const int dims[3], const float minAndVoxel[4], const uint8_t* voxels) try {
	if (!path || !dims || !minAndVoxel || !voxels) return rto_fail(RTO_ERR_INVALID, "rto_host_grid_save: null argument");
	FILE* f = std::fopen(path, "wb");
	if (!f) return rto_fail(RTO_ERR_IO, "rto_host_grid_save: cannot open file");
	uint64_t n = (uint64_t)dims[0] * dims[1] * dims[2];
	bool ok = std::fwrite(dims, 4, 3, f) == 3 && std::fwrite(minAndVoxel, 4, 4, f) == 4 && std::fwrite(&n, 8, 1, f) == 1 && std::fwrite(voxels, 1, n, f) == n;
	ok = (std::fclose(f) == 0) && ok;
	return ok ? RTO_OK : rto_fail(RTO_ERR_IO, "rto_host_grid_save: write failed");
} RTO_CATCH_ALL("rto_host_grid_save")

extern "C" void rto_host_free(void* p) { std::free(p); }

// =================================================================================================
// CSV voxeliser: loadCSVDataIntoVoxelGrid (BuildingLoader.cpp:153-290) -- the producer of sceneCache.bin
// =================================================================================================
namespace {

// fields of one CSV line, split on ',' and stripped of blanks (BuildingLoader.cpp:28-33, 51-56)
void csvFields(const std::string& line, std::vector<std::string>& out) {
	out.clear();
	std::istringstream ss(line);
	std::string tok;
	while (std::getline(ss, tok, ',')) {
		size_t b = tok.find_first_not_of(" \t\n\r"), e = tok.find_last_not_of(" \t\n\r");
		out.push_back(b == std::string::npos ? std::string() : tok.substr(b, e - b + 1));
	}
}

struct CsvVert { double e, n, h; };

} // namespace

int rto_csv_load(const char* vertsCsv, const char* facesCsv, float voxelSize, CsvScene& out) {
	out = CsvScene();
	out.voxelSize = voxelSize;
	if (!vertsCsv || !facesCsv) return rto_fail(RTO_ERR_INVALID, "csv voxeliser: null path");
	if (!(voxelSize > 0.0f)) return rto_fail(RTO_ERR_INVALID, "csv voxeliser: voxel size must be positive");
	std::unordered_map<long long, CsvVert> verts;          // (mesh, vertex number) -> last definition wins (BuildingLoader.cpp:161-164)
	const double DMAX = std::numeric_limits<double>::max();
	double lo[3] = { DMAX, DMAX, DMAX }, hi[3] = { -DMAX, -DMAX, -DMAX };
	size_t numVerts = 0;
	std::vector<std::string> f;
	std::string line;
	{
		std::ifstream in(vertsCsv);
		if (!in) return RTO_OK;                             // unreadable file: empty grid, like the reference (BuildingLoader.cpp:38-41)
		std::getline(in, line);                             // header
		while (std::getline(in, line)) {
			if (line.empty()) continue;
			csvFields(line, f);
			if (f.size() < 8) continue;
			try {
				int mesh = std::stoi(f[0]), num = std::stoi(f[1]);
				CsvVert v; v.e = std::stod(f[2]); v.n = std::stod(f[3]); v.h = std::stod(f[4]);
				(void)std::stod(f[5]); (void)std::stod(f[6]); (void)std::stod(f[7]);      // latitude, longitude, elevMin must parse too
				verts[((long long)mesh << 32) | (unsigned int)num] = v;
				numVerts++;
				if (std::isfinite(v.e) && std::isfinite(v.n) && std::isfinite(v.h)) {     // bounds over every vertex read (:174-183)
					lo[0] = std::min(lo[0], v.e); lo[1] = std::min(lo[1], v.n); lo[2] = std::min(lo[2], v.h);
					hi[0] = std::max(hi[0], v.e); hi[1] = std::max(hi[1], v.n); hi[2] = std::max(hi[2], v.h);
				}
			}
			catch (const std::exception&) { continue; }
		}
	}
	struct Face { int mesh, a, b, c; };
	std::vector<Face> faces;
	{
		std::ifstream in(facesCsv);
		if (!in) return RTO_OK;
		std::getline(in, line);
		while (std::getline(in, line)) {
			if (line.empty()) continue;
			csvFields(line, f);
			if (f.size() < 4) continue;
			try { Face fc; fc.mesh = std::stoi(f[0]); fc.a = std::stoi(f[1]); fc.b = std::stoi(f[2]); fc.c = std::stoi(f[3]); faces.push_back(fc); }
			catch (const std::exception&) { continue; }
		}
	}
	if (numVerts == 0 || faces.empty()) return RTO_OK;      // :158
	if (!(lo[0] <= hi[0])) return rto_fail(RTO_ERR_INVALID, "csv voxeliser: no vertex with finite coordinates");
	// padding of one voxel, dimensions, cap at 1000 cells per axis by enlarging the voxel (:185-208)
	double vs = voxelSize;
	for (int a = 0; a < 3; a++) { lo[a] -= vs; hi[a] += vs; }
	size_t dim[3];
	for (int a = 0; a < 3; a++) dim[a] = (size_t)std::ceil((hi[a] - lo[a]) / voxelSize);
	const size_t MAX_DIM = 1000;
	if (dim[0] > MAX_DIM || dim[1] > MAX_DIM || dim[2] > MAX_DIM) {
		float scale = (float)std::max({ dim[0] / MAX_DIM, dim[1] / MAX_DIM, dim[2] / MAX_DIM });
		voxelSize *= scale;
		for (int a = 0; a < 3; a++) dim[a] = (size_t)std::ceil((hi[a] - lo[a]) / voxelSize);
	}
	for (int a = 0; a < 3; a++) { out.dims[a] = (int)dim[a]; out.gridMin[a] = (float)lo[a]; }
	out.voxelSize = voxelSize;
	out.tris.reserve(faces.size());
	for (const Face& fc : faces) {                          // faces with a missing vertex are skipped (:232-241)
		auto a = verts.find(((long long)fc.mesh << 32) | (unsigned int)fc.a);
		auto b = verts.find(((long long)fc.mesh << 32) | (unsigned int)fc.b);
		auto c = verts.find(((long long)fc.mesh << 32) | (unsigned int)fc.c);
		if (a == verts.end() || b == verts.end() || c == verts.end()) continue;
		RtoTriangle t;
		t.v0[0] = (float)a->second.e; t.v0[1] = (float)a->second.n; t.v0[2] = (float)a->second.h;
		t.v1[0] = (float)b->second.e; t.v1[1] = (float)b->second.n; t.v1[2] = (float)b->second.h;
		t.v2[0] = (float)c->second.e; t.v2[1] = (float)c->second.n; t.v2[2] = (float)c->second.h;
		out.tris.push_back(t);
	}
	return RTO_OK;
}

extern "C" int rto_host_csv_voxelize(const char* vertsCsv, const char* facesCsv, float voxelSize, int dims[3], float minAndVoxel[4], uint8_t** voxelsOut) try {
	RTO_RANGE("rto_host_csv_voxelize");
	if (!dims || !minAndVoxel || !voxelsOut) return rto_fail(RTO_ERR_INVALID, "rto_host_csv_voxelize: null output");
	*voxelsOut = nullptr; dims[0] = dims[1] = dims[2] = 0;
	CsvScene S;
	int rc = rto_csv_load(vertsCsv, facesCsv, voxelSize, S); if (rc) return rc;
	for (int a = 0; a < 3; a++) { dims[a] = S.dims[a]; minAndVoxel[a] = S.gridMin[a]; }
	minAndVoxel[3] = S.voxelSize;
	const size_t n = (size_t)S.dims[0] * S.dims[1] * S.dims[2];
	if (n == 0) return RTO_OK;
	uint8_t* vox = (uint8_t*)std::calloc(n, 1);
	if (!vox) return rto_fail(RTO_ERR_ALLOC, "rto_host_csv_voxelize: out of memory");
	auto work = [&](size_t f0, size_t f1) {
		for (size_t i = f0; i < f1; i++) {
			const RtoTriangle& t = S.tris[i];
			VoxRange r = vox_face_range(t, S.gridMin, S.voxelSize, S.dims);
			if (r.empty) continue;
			V3 a = mk3(t.v0[0], t.v0[1], t.v0[2]), b = mk3(t.v1[0], t.v1[1], t.v1[2]), c = mk3(t.v2[0], t.v2[1], t.v2[2]);
			for (int z = r.z0; z <= r.z1; z++) for (int y = r.y0; y <= r.y1; y++) for (int x = r.x0; x <= r.x1; x++)
				if (vox_point_in_triangle(vox_center(S.gridMin, S.voxelSize, x, y, z), a, b, c))
					vox[(size_t)x + (size_t)y * S.dims[0] + (size_t)z * ((size_t)S.dims[0] * S.dims[1])] = 1;      // every writer stores the same value
		}
	};
	unsigned hw = std::max(1u, std::thread::hardware_concurrency());
	size_t nt = std::min<size_t>(hw, std::max<size_t>(1, S.tris.size() / 256));
	if (nt <= 1) work(0, S.tris.size());
	else {
		rto_run_threads((int)nt, [&](int k) { work(S.tris.size() * (size_t)k / nt, S.tris.size() * ((size_t)k + 1) / nt); });
	}
	*voxelsOut = vox;
	return RTO_OK;
} RTO_CATCH_ALL("rto_host_csv_voxelize")

// =================================================================================================
// Frustum culling of the flattened octree: RayTracerBVH::renderSceneComputeWithCulling, CPU part (RayTracerBVH.cpp:724-813)
// =================================================================================================
// proj * view with glm::perspective(radians(fovDeg), aspect, zNear, zFar) (matrix_clip_space.inl:249-262, RH, depth -1..1) and
// glm's mat4 * mat4 (type_mat4x4.inl:630-648: ((A0*b0 + A1*b1) + A2*b2) + A3*b3 per column).
extern "C" int rto_host_view_proj(const float view16[16], float fovDeg, float aspect, float zNear, float zFar, float viewProj16[16]) try {
	if (!view16 || !viewProj16) return rto_fail(RTO_ERR_INVALID, "rto_host_view_proj: null argument");
	float P[16];
	glm_perspective(fovDeg, aspect, zNear, zFar, P);
	for (int j = 0; j < 4; j++)
		for (int r = 0; r < 4; r++) {
			float acc = P[0 * 4 + r] * view16[j * 4 + 0];
			acc = acc + P[1 * 4 + r] * view16[j * 4 + 1];
			acc = acc + P[2 * 4 + r] * view16[j * 4 + 2];
			acc = acc + P[3 * 4 + r] * view16[j * 4 + 3];
			viewProj16[j * 4 + r] = acc;
		}
	return RTO_OK;
} RTO_CATCH_ALL("rto_host_view_proj")

extern "C" int rto_host_frustum_cull(const RtoGpuNode* nodes, size_t numNodes, const float gridMin[3], float voxelSize, const float viewProj16[16],
	float margin, RtoGpuNode** culledOut, size_t* numCulled, int32_t** newToOldOut) try {
	if (!culledOut || !numCulled) return rto_fail(RTO_ERR_INVALID, "rto_host_frustum_cull: null output");
	*culledOut = nullptr; *numCulled = 0;
	if (newToOldOut) *newToOldOut = nullptr;
	if (numNodes == 0) return RTO_OK;
	if (!nodes || !gridMin || !viewProj16) return rto_fail(RTO_ERR_INVALID, "rto_host_frustum_cull: null input");
	const FrustumPlanes F = frustum_from_view_proj(viewProj16);
	std::vector<int32_t> newIndex(numNodes, -1);
	size_t visible = 0;
	for (size_t i = 0; i < numNodes; i++) if (frustum_node_visible(F, nodes[i], gridMin, voxelSize, margin)) newIndex[i] = (int32_t)visible++;
	if (visible == 0) return RTO_OK;
	RtoGpuNode* out = (RtoGpuNode*)std::malloc(visible * sizeof(RtoGpuNode));
	int32_t* back = newToOldOut ? (int32_t*)std::malloc(visible * sizeof(int32_t)) : nullptr;
	if (!out || (newToOldOut && !back)) { std::free(out); std::free(back); return rto_fail(RTO_ERR_ALLOC, "rto_host_frustum_cull: out of memory"); }
	for (size_t i = 0; i < numNodes; i++) {
		if (newIndex[i] < 0) continue;
		RtoGpuNode n = nodes[i];
		if (!n.isLeaf)                                       // children of leaves are copied as they are (:788)
			for (int c = 0; c < 8; c++) {
				int32_t oc = n.child[c];
				n.child[c] = (oc >= 0 && (size_t)oc < numNodes && newIndex[oc] >= 0) ? newIndex[oc] : -1;
			}
		out[newIndex[i]] = n;
		if (back) back[newIndex[i]] = (int32_t)i;
	}
	*culledOut = out; *numCulled = visible;
	if (newToOldOut) *newToOldOut = back;
	return RTO_OK;
} RTO_CATCH_ALL("rto_host_frustum_cull")

// =================================================================================================
// The probe rays of VolumeRaycastRenderer's skip-distance estimate (VolumeRaycastRenderer.cpp:1598-1629): a 7 x 7 grid over the
// central +-0.2 of NDC, unprojected through inverse(perspective(45 deg, aspect, 0.1, 5000)) and inverse(view).
// =================================================================================================
extern "C" int rto_host_skip_probe_rays(const float view16[16], const float camPos[3], float aspect, float* origins, float* dirs) try {
	if (!view16 || !camPos || !origins || !dirs) return rto_fail(RTO_ERR_INVALID, "rto_host_skip_probe_rays: null argument");
	float P[16], invP[16], invV[16];
	glm_perspective(45.0f, aspect, 0.1f, 5000.0f, P);
	glm_inverse4(view16, invV);
	glm_inverse4(P, invP);
	const int gridSize = 7; const float sampleOffset = 0.2f;
	V3 ro = mk3(camPos[0], camPos[1], camPos[2]);
	int k = 0;
	for (int y = 0; y < gridSize; y++)
		for (int x = 0; x < gridSize; x++, k++) {
			float ndcX = ((float)x / (gridSize - 1) - 0.5f) * 2.0f * sampleOffset;
			float ndcY = ((float)y / (gridSize - 1) - 0.5f) * 2.0f * sampleOffset;
			float clip[4] = { ndcX, ndcY, 1.0f, 1.0f }, vpos[4], wpos[4];
			glm_mul_m4v4(invP, clip, vpos);
			const float w = vpos[3];
			for (int c = 0; c < 4; c++) vpos[c] = vpos[c] / w;
			glm_mul_m4v4(invV, vpos, wpos);
			V3 rd = normalize3(mk3(wpos[0], wpos[1], wpos[2]) - ro);
			origins[3 * k] = ro.x; origins[3 * k + 1] = ro.y; origins[3 * k + 2] = ro.z;
			dirs[3 * k] = rd.x; dirs[3 * k + 1] = rd.y; dirs[3 * k + 2] = rd.z;
		}
	return RTO_OK;
} RTO_CATCH_ALL("rto_host_skip_probe_rays")

// 15th percentile of the valid probe results, 75 % of it, blended with the previous frame's value (VolumeRaycastRenderer.cpp:1646-1663)
extern "C" float rto_host_skip_distance_from_probes(const float* t, int count, float lastSkipDistance) {
	std::vector<float> valid;
	for (int i = 0; i < count; i++) if (t[i] < 1e30f && t[i] > 0.0f) valid.push_back(t[i]);
	float skip = 0.0f;
	if (!valid.empty()) {
		std::sort(valid.begin(), valid.end());
		int safeIndex = std::max(0, (int)(valid.size() * 0.15f));
		skip = valid[safeIndex];
		skip *= 0.75f;
	}
	const float blendFactor = 0.4f;
	return lastSkipDistance * blendFactor + skip * (1.0f - blendFactor);
}
