// rto_dc.cu -- the Adaptive Dual Contouring mesh on the GPU (second half): the triangle soup of renderOctree
// (main.cpp:95-208) over AdaptiveDualContouringRenderer::createTriangles, same triangles in the same order as the reference and as
// rto_host_dc_mesh, bit for bit.
//
// The reference's mesher is sequential by construction: its dual-vertex cache gives every key the value of its first toucher in
// visit order.  The order-free formulation (host_dc.cpp, comment above rto_host_dc_mesh) turns that into data-parallel steps:
//   1. flag the leaves the walk reaches that contain surface; compact; sort by the Morton code of the origin (= visit order)
//   2. per leaf: edge sign mask, joinable neighbours; atomicMin of the touch key (Morton of the visitor, kind, visitor's size)
//      into every key it touches
//   3. rounds: add the touches of the fallback leaves assumed so far, compute the vertex of every key whose first toucher
//      changed, count the edge triangles of every leaf, read off who falls back; stop when the set repeats
//   4. exclusive scan of the triangle counts, emission
// Per-cell arithmetic comes from rto_dc.h (shared with the host), compiled -fmad=false -prec-div=true -prec-sqrt=true.
// One thread per cell: Hermite sums are sequential float sums in the reference's order and cannot be split across threads.
#include "rto_scene.cuh"
#include "rto_nvtx.h"
#include "rto_dc.h"

#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace rto;
using namespace rto::dc;

namespace {

#define DC_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return rto_fail(RTO_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); } while (0)

struct Pool {
	std::vector<void*> blocks;
	~Pool() { for (void* p : blocks) if (p) cudaFree(p); }
	template <typename T> cudaError_t alloc(T** p, size_t count) {
		*p = nullptr;
		void* q = nullptr;
		cudaError_t e = cudaMalloc(&q, (count ? count : 1) * sizeof(T));
		if (e == cudaSuccess) { blocks.push_back(q); *p = (T*)q; }
		return e;
	}
	void free_now(void* p) { for (void*& b : blocks) if (b == p) { cudaFree(p); b = nullptr; } }
};

struct WalkArgs { int cull; FrustumPlanes F; float gridMin[3]; float voxel, margin; };

__global__ void k_dc_flags(Grid g, const RtoGpuNode* __restrict__ nodes, size_t numNodes, WalkArgs W, uint8_t* __restrict__ flag) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= numNodes) return;
	const RtoGpuNode n = nodes[i];
	bool f = n.isLeaf != 0;
	if (f && W.cull) f = leafReached(nodes, n, W.F, W.gridMin, W.voxel, W.margin);
	if (f) f = cellContainsSurface(g, n.x, n.y, n.z, n.size);
	flag[i] = f ? 1 : 0;
}

__global__ void k_dc_keys(const RtoGpuNode* __restrict__ nodes, const int32_t* __restrict__ recNode, size_t numRecs, uint32_t* __restrict__ keys) {
	const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= numRecs) return;
	const RtoGpuNode& n = nodes[recNode[r]];
	keys[r] = morton30(n.x, n.y, n.z);
}

__global__ void k_dc_fill(unsigned long long* __restrict__ a, size_t n, unsigned long long v) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) a[i] = v;
}

// step 2: what the edge walk of each leaf touches
__global__ void k_dc_touch_edges(Grid g, const RtoGpuNode* __restrict__ nodes, const int32_t* __restrict__ recNode, size_t numRecs,
	uint16_t* __restrict__ edgeMask, int32_t* __restrict__ tgt /* 8 per record */, unsigned long long* __restrict__ firstAB) {
	const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= numRecs) return;
	const int32_t node = recNode[r];
	const RtoGpuNode n = nodes[node];
	const uint32_t mask = edgeSignMask(g, n.x, n.y, n.z, n.size);
	const uint32_t asked = askedOffsets(mask);
	edgeMask[r] = (uint16_t)mask;
	int32_t* T = tgt + r * 8;
	T[0] = node; T[7] = -1;
	atomicMin(firstAB + node, touchKey(n, 0));
	for (int o = 1; o < 7; o++) {
		const int32_t k = (asked & (1u << o)) ? joinTarget(g, nodes, n.x, n.y, n.z, n.size, o) : -1;
		T[o] = k;
		if (k >= 0) atomicMin(firstAB + k, touchKey(n, 1));
	}
}

// step 3a: touches of the leaves assumed to fall back
__global__ void k_dc_touch_fallback(Grid g, const RtoGpuNode* __restrict__ nodes, const int32_t* __restrict__ recNode, size_t numRecs,
	const uint8_t* __restrict__ fallback, unsigned long long* __restrict__ first) {
	const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= numRecs || !fallback[r]) return;
	const RtoGpuNode n = nodes[recNode[r]];
	for (int face = 0; face < 6; face++) {
		int nx, ny, nz; int32_t k;
		if (fallbackFace(g, nodes, n, face, nx, ny, nz, k) && k >= 0) atomicMin(first + k, touchKey(n, 2));
	}
}

// step 3b: the vertex of every key whose first toucher changed
__global__ void k_dc_vertices(Grid g, const RtoGpuNode* __restrict__ nodes, size_t numNodes, const unsigned long long* __restrict__ first,
	unsigned long long* __restrict__ firstPrev, V3* __restrict__ val) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= numNodes) return;
	const unsigned long long f = first[i];
	if (f == firstPrev[i]) return;
	firstPrev[i] = f;
	if (f != kNoTouch) val[i] = vertexFromTouch(g, nodes[i], f);
}

// step 3c: edge triangles of every leaf, and who falls back
__global__ void k_dc_count(Grid g, const RtoGpuNode* __restrict__ nodes, const int32_t* __restrict__ recNode, size_t numRecs,
	const uint16_t* __restrict__ edgeMask, const int32_t* __restrict__ tgt, const V3* __restrict__ val,
	int32_t* __restrict__ edgeCount, const uint8_t* __restrict__ fallback, uint8_t* __restrict__ fallbackNew, int* __restrict__ changed) {
	const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= numRecs) return;
	int32_t T[8]; V3 vtx[8];
	for (int o = 0; o < 8; o++) { T[o] = tgt[r * 8 + o]; if (T[o] >= 0) vtx[o] = val[T[o]]; }
	const int c = edgeTriangles(edgeMask[r], T, vtx, nullptr);
	edgeCount[r] = c;
	uint8_t fb = 0;
	if (c == 0) { const RtoGpuNode& n = nodes[recNode[r]]; fb = touchesBoundary(g, n.x, n.y, n.z, n.size) ? 1 : 0; }
	fallbackNew[r] = fb;
	if (fb != fallback[r]) *changed = 1;
}

// step 4a: triangles per leaf
__global__ void k_dc_totals(Grid g, const RtoGpuNode* __restrict__ nodes, const int32_t* __restrict__ recNode, size_t numRecs,
	const int32_t* __restrict__ edgeCount, const uint8_t* __restrict__ fallback, unsigned long long* __restrict__ counts) {
	const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= numRecs) return;
	unsigned long long c = (unsigned long long)edgeCount[r];
	if (fallback[r]) {
		const RtoGpuNode n = nodes[recNode[r]];
		for (int face = 0; face < 6; face++) { int nx, ny, nz; int32_t k; if (fallbackFace(g, nodes, n, face, nx, ny, nz, k)) c += 32; }
	}
	counts[r] = c;
}

// step 4b: emission in visit order
__global__ void k_dc_emit(Grid g, const RtoGpuNode* __restrict__ nodes, const int32_t* __restrict__ recNode, size_t numRecs,
	const uint16_t* __restrict__ edgeMask, const int32_t* __restrict__ tgt, const V3* __restrict__ val, const uint8_t* __restrict__ fallback,
	const unsigned long long* __restrict__ offsets, const unsigned long long* __restrict__ counts, RtoTriangle* __restrict__ tris) {
	const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= numRecs || counts[r] == 0) return;
	int32_t T[8]; V3 vtx[8];
	for (int o = 0; o < 8; o++) { T[o] = tgt[r * 8 + o]; if (T[o] >= 0) vtx[o] = val[T[o]]; }
	RtoTriangle* out = tris + offsets[r];
	out += edgeTriangles(edgeMask[r], T, vtx, out);
	if (fallback[r]) {
		const RtoGpuNode n = nodes[recNode[r]];
		for (int face = 0; face < 6; face++) {
			int nx, ny, nz; int32_t k;
			if (!fallbackFace(g, nodes, n, face, nx, ny, nz, k)) continue;
			const V3 neighborVertex = k >= 0 ? val[k] : g.centre(nx, ny, nz, n.size);
			faceFan(vtx[0], neighborVertex, face, n.size, g.vs, out);
			out += 32;
		}
	}
}

inline unsigned blocksFor(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

} // namespace

// The mesh of a grid and node array that already live on the device; *dTrisOut is cudaMalloc'ed and belongs to the caller.
int rto_dc_extract_device(const uint8_t* dVox, int dimX, int dimY, int dimZ, const float gridMin[3], float voxelSize,
	const RtoGpuNode* dNodes, size_t numNodes, const float* viewProj16, float extraMargin, cudaStream_t st,
	RtoTriangle** dTrisOut, size_t* numTris, size_t* numLeavesOut, int* roundsOut) {
	*dTrisOut = nullptr; *numTris = 0;
	if (numLeavesOut) *numLeavesOut = 0;
	if (roundsOut) *roundsOut = 0;
	if (numNodes >= ((size_t)1 << 31)) return rto_fail(RTO_ERR_UNSUPPORTED, "Dual Contouring on the device: too many nodes");
	Pool pool;
	const Grid g{ dVox, dimX, dimY, dimZ, gridMin[0], gridMin[1], gridMin[2], voxelSize };
	WalkArgs W; std::memset(&W, 0, sizeof(W));
	W.cull = viewProj16 ? 1 : 0;
	if (viewProj16) W.F = frustum_from_view_proj(viewProj16);
	W.gridMin[0] = gridMin[0]; W.gridMin[1] = gridMin[1]; W.gridMin[2] = gridMin[2]; W.voxel = voxelSize; W.margin = extraMargin;

	// 1. leaves with surface, in visit order
	uint8_t* flag = nullptr; int32_t* recNodeUnsorted = nullptr; int* dNumRecs = nullptr;
	DC_TRY(pool.alloc(&flag, numNodes)); DC_TRY(pool.alloc(&recNodeUnsorted, numNodes)); DC_TRY(pool.alloc(&dNumRecs, 1));
	k_dc_flags<<<blocksFor(numNodes, 256), 256, 0, st>>>(g, dNodes, numNodes, W, flag);
	DC_TRY(cudaGetLastError());
	size_t tmpBytes = 0; void* tmp = nullptr;
	cub::CountingInputIterator<int32_t> ids(0);
	DC_TRY(cub::DeviceSelect::Flagged(nullptr, tmpBytes, ids, flag, recNodeUnsorted, dNumRecs, (int)numNodes, st));
	DC_TRY(pool.alloc((uint8_t**)&tmp, tmpBytes));
	DC_TRY(cub::DeviceSelect::Flagged(tmp, tmpBytes, ids, flag, recNodeUnsorted, dNumRecs, (int)numNodes, st));
	int numRecsI = 0;
	DC_TRY(cudaMemcpyAsync(&numRecsI, dNumRecs, sizeof(int), cudaMemcpyDeviceToHost, st));
	DC_TRY(cudaStreamSynchronize(st));
	const size_t numRecs = (size_t)numRecsI;
	pool.free_now(tmp); pool.free_now(flag);
	if (numLeavesOut) *numLeavesOut = numRecs;
	if (numRecs == 0) return RTO_OK;
	uint32_t* keys = nullptr; uint32_t* keysSorted = nullptr; int32_t* recNode = nullptr;
	DC_TRY(pool.alloc(&keys, numRecs)); DC_TRY(pool.alloc(&keysSorted, numRecs)); DC_TRY(pool.alloc(&recNode, numRecs));
	k_dc_keys<<<blocksFor(numRecs, 256), 256, 0, st>>>(dNodes, recNodeUnsorted, numRecs, keys);
	DC_TRY(cudaGetLastError());
	tmpBytes = 0;
	DC_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys, keysSorted, recNodeUnsorted, recNode, (int)numRecs, 0, 30, st));
	DC_TRY(pool.alloc((uint8_t**)&tmp, tmpBytes));
	DC_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, keys, keysSorted, recNodeUnsorted, recNode, (int)numRecs, 0, 30, st));
	DC_TRY(cudaStreamSynchronize(st));
	pool.free_now(tmp); pool.free_now(keys); pool.free_now(keysSorted); pool.free_now(recNodeUnsorted);

	// 2. touches of the edge walks
	uint16_t* edgeMask = nullptr; int32_t* tgt = nullptr; unsigned long long* firstAB = nullptr; unsigned long long* first = nullptr; unsigned long long* firstPrev = nullptr;
	V3* val = nullptr; int32_t* edgeCount = nullptr; uint8_t* fallback = nullptr; uint8_t* fallbackNew = nullptr; int* dChanged = nullptr;
	DC_TRY(pool.alloc(&edgeMask, numRecs)); DC_TRY(pool.alloc(&tgt, numRecs * 8));
	DC_TRY(pool.alloc(&firstAB, numNodes)); DC_TRY(pool.alloc(&first, numNodes)); DC_TRY(pool.alloc(&firstPrev, numNodes));
	DC_TRY(pool.alloc(&val, numNodes)); DC_TRY(pool.alloc(&edgeCount, numRecs));
	DC_TRY(pool.alloc(&fallback, numRecs)); DC_TRY(pool.alloc(&fallbackNew, numRecs)); DC_TRY(pool.alloc(&dChanged, 1));
	k_dc_fill<<<blocksFor(numNodes, 256), 256, 0, st>>>(firstAB, numNodes, kNoTouch);
	k_dc_fill<<<blocksFor(numNodes, 256), 256, 0, st>>>(firstPrev, numNodes, kNoTouch);
	DC_TRY(cudaMemsetAsync(fallback, 0, numRecs, st));
	k_dc_touch_edges<<<blocksFor(numRecs, 128), 128, 0, st>>>(g, dNodes, recNode, numRecs, edgeMask, tgt, firstAB);
	DC_TRY(cudaGetLastError());

	// 3. rounds
	int rounds = 0;
	for (;; rounds++) {
		if (rounds > 64) return rto_fail(RTO_ERR_UNSUPPORTED, "Dual Contouring on the device: fallback rounds did not settle (use rto_host_dc_mesh_replay)");
		DC_TRY(cudaMemcpyAsync(first, firstAB, numNodes * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
		if (rounds > 0) k_dc_touch_fallback<<<blocksFor(numRecs, 128), 128, 0, st>>>(g, dNodes, recNode, numRecs, fallback, first);
		k_dc_vertices<<<blocksFor(numNodes, 64), 64, 0, st>>>(g, dNodes, numNodes, first, firstPrev, val);
		DC_TRY(cudaMemsetAsync(dChanged, 0, sizeof(int), st));
		k_dc_count<<<blocksFor(numRecs, 128), 128, 0, st>>>(g, dNodes, recNode, numRecs, edgeMask, tgt, val, edgeCount, fallback, fallbackNew, dChanged);
		DC_TRY(cudaGetLastError());
		int changed = 0;
		DC_TRY(cudaMemcpyAsync(&changed, dChanged, sizeof(int), cudaMemcpyDeviceToHost, st));
		DC_TRY(cudaStreamSynchronize(st));
		if (!changed) break;
		uint8_t* t = fallback; fallback = fallbackNew; fallbackNew = t;
	}
	if (roundsOut) *roundsOut = rounds;

	// 4. counts, scan, emission
	unsigned long long* counts = nullptr; unsigned long long* offsets = nullptr;
	DC_TRY(pool.alloc(&counts, numRecs)); DC_TRY(pool.alloc(&offsets, numRecs));
	k_dc_totals<<<blocksFor(numRecs, 128), 128, 0, st>>>(g, dNodes, recNode, numRecs, edgeCount, fallback, counts);
	DC_TRY(cudaGetLastError());
	tmpBytes = 0;
	DC_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, counts, offsets, (int)numRecs, st));
	DC_TRY(pool.alloc((uint8_t**)&tmp, tmpBytes));
	DC_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, counts, offsets, (int)numRecs, st));
	unsigned long long lastOff = 0, lastCnt = 0;
	DC_TRY(cudaMemcpyAsync(&lastOff, offsets + (numRecs - 1), sizeof(lastOff), cudaMemcpyDeviceToHost, st));
	DC_TRY(cudaMemcpyAsync(&lastCnt, counts + (numRecs - 1), sizeof(lastCnt), cudaMemcpyDeviceToHost, st));
	DC_TRY(cudaStreamSynchronize(st));
	const size_t total = (size_t)(lastOff + lastCnt);
	pool.free_now(tmp); pool.free_now(firstAB); pool.free_now(first); pool.free_now(firstPrev);
	if (total == 0) return RTO_OK;
	RtoTriangle* dTris = nullptr;
	DC_TRY(cudaMalloc((void**)&dTris, total * sizeof(RtoTriangle)));
	k_dc_emit<<<blocksFor(numRecs, 128), 128, 0, st>>>(g, dNodes, recNode, numRecs, edgeMask, tgt, val, fallback, offsets, counts, dTris);
	cudaError_t e = cudaGetLastError();
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	if (e != cudaSuccess) { cudaFree(dTris); return rto_fail(RTO_ERR_CUDA, "Dual Contouring on the device: %s", cudaGetErrorString(e)); }
	*dTrisOut = dTris; *numTris = total;
	return RTO_OK;
}

extern "C" int rto_device_dc_mesh(const uint8_t* voxels, int dimX, int dimY, int dimZ, const float gridMin[3], float voxelSize,
	const RtoGpuNode* nodes, size_t numNodes, const float* viewProj16, float extraMargin, RtoTriangle** trisOut, size_t* numTris) try {
	RTO_RANGE("rto_device_dc_mesh");
	if (!trisOut || !numTris) return rto_fail(RTO_ERR_INVALID, "rto_device_dc_mesh: null output");
	*trisOut = nullptr; *numTris = 0;
	if (numNodes == 0) return RTO_OK;
	if (!voxels || !nodes || !gridMin || dimX <= 0 || dimY <= 0 || dimZ <= 0) return rto_fail(RTO_ERR_INVALID, "rto_device_dc_mesh: bad input");
	if (nodes[0].size > 1024) return rto_fail(RTO_ERR_UNSUPPORTED, "rto_device_dc_mesh: octree larger than 1024 voxels per axis (the reference's cell keys alias there)");
	int rc = rto_require_device(); if (rc) return rc;
	const bool verbose = std::getenv("RTO_DC_VERBOSE") != nullptr;
	cudaStream_t st = nullptr;
	DC_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
	struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamSynchronize(s); cudaStreamDestroy(s); } } guard{ st };
	Pool pool;
	cudaEvent_t ev0, ev1;
	DC_TRY(cudaEventCreate(&ev0)); DC_TRY(cudaEventCreate(&ev1));
	struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } evGuard{ ev0, ev1 };
	const size_t nvox = (size_t)dimX * dimY * dimZ;
	uint8_t* dVox = nullptr; RtoGpuNode* dNodes = nullptr;
	DC_TRY(pool.alloc(&dVox, nvox)); DC_TRY(pool.alloc(&dNodes, numNodes));
	DC_TRY(cudaMemcpyAsync(dVox, voxels, nvox, cudaMemcpyHostToDevice, st));
	DC_TRY(cudaMemcpyAsync(dNodes, nodes, numNodes * sizeof(RtoGpuNode), cudaMemcpyHostToDevice, st));
	DC_TRY(cudaEventRecord(ev0, st));
	RtoTriangle* dTris = nullptr; size_t total = 0, leaves = 0; int rounds = 0;
	rc = rto_dc_extract_device(dVox, dimX, dimY, dimZ, gridMin, voxelSize, dNodes, numNodes, viewProj16, extraMargin, st, &dTris, &total, &leaves, &rounds);
	if (rc || total == 0) return rc;
	pool.blocks.push_back(dTris);
	DC_TRY(cudaEventRecord(ev1, st));
	RtoTriangle* host = (RtoTriangle*)std::malloc(total * sizeof(RtoTriangle));
	if (!host) return rto_fail(RTO_ERR_ALLOC, "rto_device_dc_mesh: out of host memory");
	cudaError_t e = cudaMemcpyAsync(host, dTris, total * sizeof(RtoTriangle), cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	if (e != cudaSuccess) { std::free(host); return rto_fail(RTO_ERR_CUDA, "rto_device_dc_mesh: %s", cudaGetErrorString(e)); }
	if (verbose) {
		float ms = 0; cudaEventElapsedTime(&ms, ev0, ev1);
		std::fprintf(stderr, "rto_device_dc_mesh: %zu leaves with surface, %zu triangles, %d extra rounds, %.1f ms on the device between upload and read-back\n", leaves, total, rounds, ms);
	}
	*trisOut = host; *numTris = total;
	return RTO_OK;
} RTO_CATCH_ALL("rto_device_dc_mesh")
