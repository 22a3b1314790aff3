// host_dc.cpp -- the Adaptive Dual Contouring mesh of the reference, host side (part of librto.so).
//
//   rto_host_dc_mesh == renderOctree (main.cpp:95-208) over AdaptiveDualContouringRenderer::render -> createTriangles
//                       (AdaptiveDualContouringRenderer.cpp:489-803, 805-1088, 1090-1357, 1367-1530, QEFSolver :46-160):
//                       the triangle soup of configuration C4, same triangles in the same order, bit for bit.
//
// What the reference does.  renderOctree walks the octree depth first (children 0..7), drops subtrees outside the frustum and calls
// createTriangles on every leaf, one after the other on one thread.  A leaf that "contains surface" (a sampling heuristic) gets a
// dual vertex (axis-snapped plane projection or a regularised QEF over the Hermite points of its region); for each of its 12 edges
// whose end voxels differ it collects the dual vertices of up to three neighbour leaves (looked up by origin in g_octreeMap) and
// emits 1-2 triangles whose area exceeds 1e-6; a boundary leaf that produced nothing falls back to bulged face fans.
//
// The order dependence.  Dual vertices go through dualVertexCache: a neighbour's vertex that is not cached yet is computed by the
// VISITING leaf with the visiting leaf's size and cached under the neighbour's key, so the value a key holds is the one its first
// toucher gave it.  That makes the result a function of the visit order, and it is why the reference's own thread pool is not on
// this path.  Two formulations give the reference's result here, both over the per-cell arithmetic of rto_dc.h (shared with the
// device builder rto_dc.cu): an order-free one (rto_host_dc_mesh: first touchers as minima, everything on all host threads) and
// a replay of the cache protocol in visit order (rto_host_dc_mesh_replay).  See the comments above each.
// Limits: the reference's keys are (x << 20 | y << 10 | z), which alias beyond 1024 voxels per axis; larger octrees are refused.
#include "rto_internal.h"
#include "rto_nvtx.h"
#include "rto_dc.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

using namespace rto;
using namespace rto::dc;

namespace {

template <class F> void parallelFor(size_t n, size_t grain, int nthreads, F&& body) {
	if (nthreads <= 1 || n <= grain) { if (n) body(0, n, 0); return; }
	std::atomic<size_t> next{ 0 };
	rto_run_threads(nthreads, [&](int t) { for (;;) { size_t lo = next.fetch_add(grain); if (lo >= n) break; body(lo, std::min(n, lo + grain), t); } });
}

double nowSec() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int hostThreads() { return (int)std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 64); }

int checkArgs(const char* who, const uint8_t* voxels, const float* gridMin, const RtoGpuNode* nodes, size_t numNodes, RtoTriangle** trisOut, size_t* numTris) {
	if (!trisOut || !numTris) return rto_fail(RTO_ERR_INVALID, "%s: null output", who);
	*trisOut = nullptr; *numTris = 0;
	if (numNodes == 0) return RTO_OK;
	if (!voxels || !nodes || !gridMin) return rto_fail(RTO_ERR_INVALID, "%s: null input", who);
	if (nodes[0].size > 1024) return rto_fail(RTO_ERR_UNSUPPORTED, "%s: octree larger than 1024 voxels per axis (the reference's cell keys alias there)", who);
	return RTO_OK;
}

int handOver(const char* who, const std::vector<RtoTriangle>& out, RtoTriangle** trisOut, size_t* numTris) {
	if (out.empty()) return RTO_OK;
	RtoTriangle* buf = (RtoTriangle*)std::malloc(out.size() * sizeof(RtoTriangle));
	if (!buf) return rto_fail(RTO_ERR_ALLOC, "%s: out of memory", who);
	std::memcpy(buf, out.data(), out.size() * sizeof(RtoTriangle));
	*trisOut = buf; *numTris = out.size();
	return RTO_OK;
}

// the visit order of renderOctree's traverse lambda (main.cpp:152-187): depth first, children 0..7, frustum-culled subtrees dropped.
// The subtrees three levels below the root are walked on separate threads and their leaf lists joined in order.
void visitOrder(const RtoGpuNode* nodes, size_t numNodes, const float* viewProj16, const float gridMin[3], float voxelSize, float extraMargin, std::vector<int32_t>& order) {
	FrustumPlanes F;
	if (viewProj16) F = frustum_from_view_proj(viewProj16);
	auto walk = [&](int32_t root, int depthLimit, std::vector<int32_t>& leaves, std::vector<int32_t>* cut) {
		std::vector<std::pair<int32_t, int>> stack; stack.emplace_back(root, 0);
		while (!stack.empty()) {
			auto [i, depth] = stack.back(); stack.pop_back();
			if (i < 0 || (size_t)i >= numNodes) continue;
			const RtoGpuNode& n = nodes[i];
			if (cut && depth == depthLimit) { cut->push_back(i); leaves.push_back(-(int32_t)cut->size()); continue; }     // placeholder -(task + 1)
			if (viewProj16 && !frustum_node_visible(F, n, gridMin, voxelSize, extraMargin)) continue;
			if (n.isLeaf) leaves.push_back(i);
			else for (int c = 7; c >= 0; c--) stack.emplace_back(n.child[c], depth + 1);
		}
	};
	std::vector<int32_t> top, tasks;
	walk(0, 3, top, &tasks);
	std::vector<std::vector<int32_t>> parts(tasks.size());
	parallelFor(tasks.size(), 1, hostThreads(), [&](size_t lo, size_t hi, int) { for (size_t t = lo; t < hi; t++) walk(tasks[t], 0, parts[t], nullptr); });
	size_t total = 0;
	for (int32_t v : top) total += v >= 0 ? 1 : parts[(size_t)(-v - 1)].size();
	order.reserve(total);
	for (int32_t v : top) { if (v >= 0) order.push_back(v); else { const auto& p = parts[(size_t)(-v - 1)]; order.insert(order.end(), p.begin(), p.end()); } }
}

UniformBox boxOf(const RtoGpuNode& n) { UniformBox u; u.x0 = n.x; u.y0 = n.y; u.z0 = n.z; u.size = n.size; return u; }

// One leaf that contains surface
struct LeafRec {
	int32_t  node;
	uint16_t edgeMask;      // bit dir * 4 + edge: end voxels of that edge are both inside the grid and differ (:590-612)
	uint8_t  memoMask;      // (replay only) memo[o] holds the dual vertex of (origin - offset o) for the size of this leaf
	int32_t  tgt[8];        // offset o (bit 0: x - size, bit 1: y - size, bit 2: z - size) -> leaf that starts there and may be joined, or -1 (:646-687)
};

void atomicMinU64(std::atomic<unsigned long long>& a, unsigned long long v) {
	unsigned long long cur = a.load(std::memory_order_relaxed);
	while (v < cur && !a.compare_exchange_weak(cur, v, std::memory_order_relaxed)) {}
}

} // namespace

// =================================================================================================================================
// Order-free formulation (also what rto_dc.cu runs on the device).
//
// The value a cache key holds is the one its FIRST toucher gave it, and a touch is ordered by (position of the visiting leaf in the
// walk, kind): own vertex (0) before the neighbours met along the 12 edges (1) before the face neighbours of the fallback (2).
// The position of a leaf in the depth-first walk is the Morton code of its origin.  Which keys a leaf touches along its edges is a
// pure function of the grid and the octree, so the first toucher of every key is a minimum over independent contributions; the
// vertex follows from the key, the toucher's size and the kind.  Only the fallback of createTriangles (:789-799) depends on
// values: it runs for a boundary leaf none of whose edge triangles survives the area test.  Its touches are added in rounds: with
// the set S of fallback leaves assumed so far, first touchers, vertices and triangle counts are recomputed and S is read off again,
// until it repeats.  A leaf's triangles depend only on touches made before it in the walk, so each round gets a longer prefix of
// the walk right, and a set that repeats is the reference's (it starts empty and normally repeats after one round).
// =================================================================================================================================
static int dcMeshOrderFree(const uint8_t* voxels, int dimX, int dimY, int dimZ, const float gridMin[3], float voxelSize,
	const RtoGpuNode* nodes, size_t numNodes, const float* viewProj16, float extraMargin, RtoTriangle** trisOut, float** normalsOut, size_t* numTris) {
	if (normalsOut) *normalsOut = nullptr;
	int rc = checkArgs("rto_host_dc_mesh", voxels, gridMin, nodes, numNodes, trisOut, numTris);
	if (rc || numNodes == 0) return rc;
	const Grid g{ voxels, dimX, dimY, dimZ, gridMin[0], gridMin[1], gridMin[2], voxelSize };
	const bool verbose = std::getenv("RTO_DC_VERBOSE") != nullptr;
	const int nthreads = hostThreads();
	double t0 = nowSec();

	std::vector<int32_t> order;
	visitOrder(nodes, numNodes, viewProj16, gridMin, voxelSize, extraMargin, order);
	// leaves that contain surface, in visit order
	std::vector<uint8_t> surf(order.size(), 0);
	parallelFor(order.size(), 4096, nthreads, [&](size_t lo, size_t hi, int) {
		for (size_t i = lo; i < hi; i++) { const RtoGpuNode& n = nodes[order[i]]; surf[i] = cellContainsSurface(g, n.x, n.y, n.z, n.size) ? 1 : 0; }
	});
	std::vector<LeafRec> recs;
	for (size_t i = 0; i < order.size(); i++) if (surf[i]) { LeafRec R; R.node = order[i]; R.edgeMask = 0; R.memoMask = 0; recs.push_back(R); }
	std::vector<int32_t>().swap(order); std::vector<uint8_t>().swap(surf);
	double t1 = nowSec();

	// first touchers from the edge walks (kinds 0 and 1)
	std::vector<std::atomic<unsigned long long>> firstAB(numNodes);
	parallelFor(numNodes, 1 << 16, nthreads, [&](size_t lo, size_t hi, int) { for (size_t i = lo; i < hi; i++) firstAB[i].store(kNoTouch, std::memory_order_relaxed); });
	parallelFor(recs.size(), 1024, nthreads, [&](size_t lo, size_t hi, int) {
		for (size_t r = lo; r < hi; r++) {
			LeafRec& R = recs[r];
			const RtoGpuNode& n = nodes[R.node];
			R.edgeMask = (uint16_t)edgeSignMask(g, n.x, n.y, n.z, n.size);
			const uint32_t asked = askedOffsets(R.edgeMask);
			R.tgt[0] = R.node; R.tgt[7] = -1;
			atomicMinU64(firstAB[(size_t)R.node], touchKey(n, 0));
			for (int o = 1; o < 7; o++) {
				R.tgt[o] = (asked & (1u << o)) ? joinTarget(g, nodes, n.x, n.y, n.z, n.size, o) : -1;
				if (R.tgt[o] >= 0) atomicMinU64(firstAB[(size_t)R.tgt[o]], touchKey(n, 1));
			}
		}
	});
	double t2 = nowSec();

	std::vector<unsigned long long> first(numNodes), firstPrev(numNodes, kNoTouch);
	std::vector<V3> val(numNodes);
	std::vector<uint8_t> fallback(recs.size(), 0), fallbackNew(recs.size(), 0);
	std::vector<int32_t> edgeCount(recs.size(), 0);
	int rounds = 0;
	size_t numFallback = 0, vertices = 0;
	for (;; rounds++) {
		if (rounds > 64) {
			if (normalsOut) return rto_fail(RTO_ERR_UNSUPPORTED, "rto_host_dc_mesh_normals: fallback rounds did not settle");
			return rto_host_dc_mesh_replay(voxels, dimX, dimY, dimZ, gridMin, voxelSize, nodes, numNodes, viewProj16, extraMargin, trisOut, numTris);
		}
		parallelFor(numNodes, 1 << 16, nthreads, [&](size_t lo, size_t hi, int) { for (size_t i = lo; i < hi; i++) first[i] = firstAB[i].load(std::memory_order_relaxed); });
		// touches of the assumed fallback leaves (kind 2), sequentially: they are few
		for (size_t r = 0; r < recs.size(); r++) if (fallback[r]) {
			const RtoGpuNode& n = nodes[recs[r].node];
			for (int face = 0; face < 6; face++) {
				int nx, ny, nz; int32_t k;
				if (fallbackFace(g, nodes, n, face, nx, ny, nz, k) && k >= 0) first[(size_t)k] = std::min(first[(size_t)k], touchKey(n, 2));
			}
		}
		// vertices of keys whose first toucher changed
		std::atomic<size_t> computed{ 0 };
		parallelFor(numNodes, 4096, nthreads, [&](size_t lo, size_t hi, int) {
			size_t c = 0;
			for (size_t i = lo; i < hi; i++) if (first[i] != firstPrev[i]) { if (first[i] != kNoTouch) { val[i] = vertexFromTouch(g, nodes[i], first[i]); c++; } firstPrev[i] = first[i]; }
			computed += c;
		});
		vertices += computed.load();
		// edge triangles of every leaf, and who falls back
		std::atomic<int> changed{ 0 };
		parallelFor(recs.size(), 4096, nthreads, [&](size_t lo, size_t hi, int) {
			int ch = 0;
			for (size_t r = lo; r < hi; r++) {
				const LeafRec& R = recs[r];
				const RtoGpuNode& n = nodes[R.node];
				V3 vtx[8];
				for (int o = 0; o < 7; o++) if (R.tgt[o] >= 0) vtx[o] = val[(size_t)R.tgt[o]];
				edgeCount[r] = edgeTriangles(R.edgeMask, R.tgt, vtx, nullptr);
				fallbackNew[r] = (edgeCount[r] == 0 && touchesBoundary(g, n.x, n.y, n.z, n.size)) ? 1 : 0;
				ch |= fallbackNew[r] != fallback[r];
			}
			if (ch) changed.store(1);
		});
		if (!changed.load()) break;
		fallback.swap(fallbackNew);
	}
	double t3 = nowSec();

	// emission in visit order
	std::vector<size_t> offset(recs.size() + 1, 0);
	for (size_t r = 0; r < recs.size(); r++) {
		size_t c = (size_t)edgeCount[r];
		if (fallback[r]) {
			numFallback++;
			const RtoGpuNode& n = nodes[recs[r].node];
			for (int face = 0; face < 6; face++) { int nx, ny, nz; int32_t k; if (fallbackFace(g, nodes, n, face, nx, ny, nz, k)) c += 32; }
		}
		offset[r + 1] = offset[r] + c;
	}
	const size_t total = offset[recs.size()];
	if (verbose) std::fprintf(stderr, "rto_host_dc_mesh: %zu leaves with surface, %zu triangles, %zu fallback leaves, %d extra rounds, %zu vertices; walk+surface %.2f s, touches %.2f s, vertices+counts %.2f s (%d threads)\n",
		recs.size(), total, numFallback, rounds, vertices, t1 - t0, t2 - t1, t3 - t2, nthreads);
	if (total == 0) return RTO_OK;
	RtoTriangle* buf = (RtoTriangle*)std::malloc(total * sizeof(RtoTriangle));
	V3* nbuf = normalsOut ? (V3*)std::malloc(total * sizeof(V3)) : nullptr;
	if (!buf || (normalsOut && !nbuf)) { std::free(buf); std::free(nbuf); return rto_fail(RTO_ERR_ALLOC, "rto_host_dc_mesh: out of memory"); }
	parallelFor(recs.size(), 4096, nthreads, [&](size_t lo, size_t hi, int) {
		for (size_t r = lo; r < hi; r++) {
			if (offset[r + 1] == offset[r]) continue;
			const LeafRec& R = recs[r];
			const RtoGpuNode& n = nodes[R.node];
			RtoTriangle* out = buf + offset[r];
			V3* nrm = nbuf ? nbuf + offset[r] : nullptr;
			V3 vtx[8];
			for (int o = 0; o < 7; o++) if (R.tgt[o] >= 0) vtx[o] = val[(size_t)R.tgt[o]];
			const int ne = edgeTriangles(R.edgeMask, R.tgt, vtx, out, nrm, n.isSolid != 0);
			out += ne; if (nrm) nrm += ne;
			if (fallback[r])
				for (int face = 0; face < 6; face++) {
					int nx, ny, nz; int32_t k;
					if (!fallbackFace(g, nodes, n, face, nx, ny, nz, k)) continue;
					const V3 neighborVertex = k >= 0 ? val[(size_t)k] : g.centre(nx, ny, nz, n.size);
					faceFan(vtx[0], neighborVertex, face, n.size, voxelSize, out, nrm, n.isSolid != 0);
					out += 32; if (nrm) nrm += 32;
				}
		}
	});
	*trisOut = buf; *numTris = total;
	if (normalsOut) *normalsOut = reinterpret_cast<float*>(nbuf);
	return RTO_OK;
}

extern "C" int rto_host_dc_mesh(const uint8_t* voxels, int dimX, int dimY, int dimZ, const float gridMin[3], float voxelSize,
	const RtoGpuNode* nodes, size_t numNodes, const float* viewProj16, float extraMargin, RtoTriangle** trisOut, size_t* numTris) try {
	RTO_RANGE("rto_host_dc_mesh");
	return dcMeshOrderFree(voxels, dimX, dimY, dimZ, gridMin, voxelSize, nodes, numNodes, viewProj16, extraMargin, trisOut, nullptr, numTris);
} RTO_CATCH_ALL("rto_host_dc_mesh")

extern "C" int rto_host_dc_mesh_normals(const uint8_t* voxels, int dimX, int dimY, int dimZ, const float gridMin[3], float voxelSize,
	const RtoGpuNode* nodes, size_t numNodes, const float* viewProj16, float extraMargin, RtoTriangle** trisOut, float** normalsOut, size_t* numTris) try {
	if (!normalsOut) return rto_fail(RTO_ERR_INVALID, "rto_host_dc_mesh_normals: null output");
	return dcMeshOrderFree(voxels, dimX, dimY, dimZ, gridMin, voxelSize, nodes, numNodes, viewProj16, extraMargin, trisOut, normalsOut, numTris);
} RTO_CATCH_ALL("rto_host_dc_mesh_normals")

// =================================================================================================================================
// The triangle cache of the application (main.cpp:27-67, written after every Dual-Contouring run and read back instead of meshing):
// size_t count, then count x MCTriangle { vec3 v[3]; vec3 normal[3]; } = 72 bytes each.
// =================================================================================================================================
extern "C" int rto_host_tricache_save(const char* path, const RtoTriangle* tris, const float* normals3, size_t numTris) try {
	if (!path || (numTris && !tris)) return rto_fail(RTO_ERR_INVALID, "rto_host_tricache_save: null argument");
	FILE* f = std::fopen(path, "wb");
	if (!f) return rto_fail(RTO_ERR_IO, "rto_host_tricache_save: cannot open %s", path);
	const size_t n = numTris;
	bool ok = std::fwrite(&n, sizeof(n), 1, f) == 1;
	std::vector<float> rec((size_t)18 * 4096);
	for (size_t i0 = 0; i0 < n && ok; i0 += 4096) {
		const size_t m = std::min<size_t>(4096, n - i0);
		for (size_t i = 0; i < m; i++) {
			float* r = rec.data() + 18 * i;
			std::memcpy(r, &tris[i0 + i], 36);
			V3 nm;
			if (normals3) nm = mk3(normals3[3 * (i0 + i)], normals3[3 * (i0 + i) + 1], normals3[3 * (i0 + i) + 2]);
			else {       // flat normal of the geometry, what localMC stores (OctreeVoxel.cpp:863-870)
				const float* v = tris[i0 + i].v0;
				nm = normalize3(cross3(mk3(v[3], v[4], v[5]) - mk3(v[0], v[1], v[2]), mk3(v[6], v[7], v[8]) - mk3(v[0], v[1], v[2])));
			}
			for (int k = 0; k < 3; k++) { r[9 + 3 * k] = nm.x; r[10 + 3 * k] = nm.y; r[11 + 3 * k] = nm.z; }
		}
		ok = std::fwrite(rec.data(), 72, m, f) == m;
	}
	ok = (std::fclose(f) == 0) && ok;
	return ok ? RTO_OK : rto_fail(RTO_ERR_IO, "rto_host_tricache_save: write to %s failed", path);
} RTO_CATCH_ALL("rto_host_tricache_save")

extern "C" int rto_host_tricache_load(const char* path, RtoTriangle** trisOut, float** normals9Out, size_t* numTris) try {
	if (!path || !trisOut || !numTris) return rto_fail(RTO_ERR_INVALID, "rto_host_tricache_load: null argument");
	*trisOut = nullptr; *numTris = 0; if (normals9Out) *normals9Out = nullptr;
	FILE* f = std::fopen(path, "rb");
	if (!f) return rto_fail(RTO_ERR_IO, "rto_host_tricache_load: cannot open %s", path);
	size_t n = 0;
	if (std::fread(&n, sizeof(n), 1, f) != 1) { std::fclose(f); return rto_fail(RTO_ERR_IO, "rto_host_tricache_load: %s is truncated", path); }
	std::fseek(f, 0, SEEK_END);
	const long bytes = std::ftell(f);
	if (bytes < 0 || (size_t)bytes < 8 || ((size_t)bytes - 8) / 72 < n) { std::fclose(f); return rto_fail(RTO_ERR_IO, "rto_host_tricache_load: %s holds fewer than the %zu triangles it announces", path, n); }
	std::fseek(f, 8, SEEK_SET);
	if (n == 0) { std::fclose(f); return RTO_OK; }
	RtoTriangle* t = (RtoTriangle*)std::malloc(n * sizeof(RtoTriangle));
	float* nm = normals9Out ? (float*)std::malloc(n * 36) : nullptr;
	if (!t || (normals9Out && !nm)) { std::free(t); std::free(nm); std::fclose(f); return rto_fail(RTO_ERR_ALLOC, "rto_host_tricache_load: out of memory"); }
	std::vector<float> rec((size_t)18 * 4096);
	bool ok = true;
	for (size_t i0 = 0; i0 < n && ok; i0 += 4096) {
		const size_t m = std::min<size_t>(4096, n - i0);
		ok = std::fread(rec.data(), 72, m, f) == m;
		for (size_t i = 0; i < m && ok; i++) { std::memcpy(&t[i0 + i], rec.data() + 18 * i, 36); if (nm) std::memcpy(nm + 9 * (i0 + i), rec.data() + 18 * i + 9, 36); }
	}
	std::fclose(f);
	if (!ok) { std::free(t); std::free(nm); return rto_fail(RTO_ERR_IO, "rto_host_tricache_load: read from %s failed", path); }
	*trisOut = t; *numTris = n; if (normals9Out) *normals9Out = nm;
	return RTO_OK;
} RTO_CATCH_ALL("rto_host_tricache_load")

// =================================================================================================================================
// Replay formulation: the reference's cache protocol executed leaf by leaf in visit order, with the pure per-cell work prepared
// on all host threads for a block of leaves at a time.  Same result; kept as the cross-check of the order-free formulation and as
// its way out should the fallback rounds not settle.
//   * a neighbour at offset {-s, 0}^3 always precedes the visiting leaf in depth-first order, so a neighbour that contains surface
//     has cached its own vertex by then; a speculative vertex is prepared only for the others;
//   * the sequential pass computes on the spot whatever the parallel pass did not foresee: correctness never depends on the
//     speculation, only speed does.
// =================================================================================================================================
extern "C" int rto_host_dc_mesh_replay(const uint8_t* voxels, int dimX, int dimY, int dimZ, const float gridMin[3], float voxelSize,
	const RtoGpuNode* nodes, size_t numNodes, const float* viewProj16, float extraMargin, RtoTriangle** trisOut, size_t* numTris) try {
	RTO_RANGE("rto_host_dc_mesh_replay");
	int rc = checkArgs("rto_host_dc_mesh_replay", voxels, gridMin, nodes, numNodes, trisOut, numTris);
	if (rc || numNodes == 0) return rc;
	const Grid g{ voxels, dimX, dimY, dimZ, gridMin[0], gridMin[1], gridMin[2], voxelSize };
	std::vector<int32_t> order;
	visitOrder(nodes, numNodes, viewProj16, gridMin, voxelSize, extraMargin, order);

	// dualVertexCache, keyed by leaf: slotOf[node] >= 0 -> cacheVal[slot]; -1 not cached; -2 (during a block) will cache itself in this block
	std::vector<int32_t> slotOf(numNodes, -1);
	std::vector<V3> cacheVal;
	std::vector<RtoTriangle> out;
	const int nthreads = hostThreads();
	const size_t kBlock = (size_t)1 << 18;
	std::vector<uint8_t> surf;
	std::vector<LeafRec> recs;
	std::vector<V3> memo;            // 8 per record
	std::vector<uint32_t> recIndex;
	for (size_t b0 = 0; b0 < order.size(); b0 += kBlock) {
		const size_t bn = std::min(kBlock, order.size() - b0);
		surf.assign(bn, 0);
		parallelFor(bn, 2048, nthreads, [&](size_t lo, size_t hi, int) {
			for (size_t i = lo; i < hi; i++) { const RtoGpuNode& n = nodes[order[b0 + i]]; surf[i] = cellContainsSurface(g, n.x, n.y, n.z, n.size) ? 1 : 0; }
		});
		recIndex.clear();
		for (size_t i = 0; i < bn; i++) if (surf[i]) { recIndex.push_back((uint32_t)i); int32_t nd = order[b0 + i]; if (slotOf[nd] == -1) slotOf[nd] = -2; }
		recs.resize(recIndex.size()); memo.resize(recIndex.size() * 8);
		parallelFor(recIndex.size(), 256, nthreads, [&](size_t lo, size_t hi, int) {
			for (size_t r = lo; r < hi; r++) {
				LeafRec& R = recs[r];
				R.node = order[b0 + recIndex[r]];
				const RtoGpuNode& n = nodes[R.node];
				R.edgeMask = (uint16_t)edgeSignMask(g, n.x, n.y, n.z, n.size); R.memoMask = 0;
				const uint32_t asked = askedOffsets(R.edgeMask);
				R.tgt[0] = R.node; R.tgt[7] = -1;
				for (int o = 1; o < 7; o++) R.tgt[o] = (asked & (1u << o)) ? joinTarget(g, nodes, n.x, n.y, n.z, n.size, o) : -1;
				if (slotOf[R.node] < 0) { memo[r * 8] = dualVertex(g, n.x, n.y, n.z, n.size, boxOf(n)); R.memoMask |= 1; }
				for (int o = 1; o < 7; o++) {
					if (R.tgt[o] < 0 || slotOf[R.tgt[o]] != -1) continue;
					const RtoGpuNode& k = nodes[R.tgt[o]];
					memo[r * 8 + o] = dualVertex(g, k.x, k.y, k.z, n.size, boxOf(k)); R.memoMask |= (uint8_t)(1u << o);
				}
			}
		});
		auto cached = [&](int32_t node, V3& v) { int32_t s = slotOf[node]; if (s < 0) return false; v = cacheVal[(size_t)s]; return true; };
		auto store = [&](int32_t node, V3 v) { slotOf[node] = (int32_t)cacheVal.size(); cacheVal.push_back(v); };
		for (size_t r = 0; r < recs.size(); r++) {
			const LeafRec& R = recs[r];
			const RtoGpuNode& n = nodes[R.node];
			V3 vtx[8];
			if (!cached(R.node, vtx[0])) { vtx[0] = (R.memoMask & 1) ? memo[r * 8] : dualVertex(g, n.x, n.y, n.z, n.size, boxOf(n)); store(R.node, vtx[0]); }
			// the look-ups of the edge walk, in its order (:586-726); an offset met again finds what the first meeting cached
			for (int dir = 0; dir < 3; dir++) for (int edge = 0; edge < 4; edge++) if (R.edgeMask & (1u << (dir * 4 + edge)))
				for (int adjIdx = 1; adjIdx < 4; adjIdx++) {
					const int o = adjOffsetBits(dir, edge, adjIdx);
					const int32_t k = R.tgt[o];
					if (k < 0 || o == 0) continue;
					if (!cached(k, vtx[o])) { const RtoGpuNode& kn = nodes[k]; vtx[o] = (R.memoMask & (1u << o)) ? memo[r * 8 + o] : dualVertex(g, kn.x, kn.y, kn.z, n.size, boxOf(kn)); store(k, vtx[o]); }
				}
			const size_t before = out.size();
			out.resize(before + 24);
			out.resize(before + (size_t)edgeTriangles(R.edgeMask, R.tgt, vtx, out.data() + before));
			// fallback for boundary cells that produced nothing (:789-799 -> createFaceTriangles)
			if (out.size() == before && touchesBoundary(g, n.x, n.y, n.z, n.size))
				for (int face = 0; face < 6; face++) {
					int nx, ny, nz; int32_t k;
					if (!fallbackFace(g, nodes, n, face, nx, ny, nz, k)) continue;
					V3 neighborVertex;
					if (!(k >= 0 && cached(k, neighborVertex))) {
						neighborVertex = g.centre(nx, ny, nz, n.size);
						if (k >= 0) store(k, neighborVertex);          // a key no leaf starts at is never read back
					}
					const size_t at = out.size();
					out.resize(at + 32);
					faceFan(vtx[0], neighborVertex, face, n.size, voxelSize, out.data() + at);
				}
		}
	}
	return handOver("rto_host_dc_mesh_replay", out, trisOut, numTris);
} RTO_CATCH_ALL("rto_host_dc_mesh_replay")
