// rto_build.cu -- scene construction on the GPU (SURVEY.md 8f "next" rows 2 and 3): the steps in front of the ray path.
//
//   octree   createOctreeFromVoxelGrid (OctreeVoxel.cpp:704-778) + RayTracerBVH::setOctree's BFS numbering
//            (RayTracerBVH.cpp:443-490): the same node set, the same indices, built bottom-up from a uniformity pyramid and
//            emitted level by level with prefix sums (a BFS level of that tree is the Morton-ordered list of its cells).
//            Integer work only: results are bit-identical to rto_host_octree_build, which tests pin to the reference.
//   mesh     MarchingCubesRenderer::render over localMC (Renderer.cpp:14-36, OctreeVoxel.cpp:780-879): every grid cell that
//            straddles the surface is found in one pass, put into the reference's emission order by a radix sort on
//            (leaf in depth-first order, z, y, x inside the leaf) and triangulated with the reference's arithmetic
//            (-fmad=false), so the triangle array equals rto_host_mc_mesh's bit for bit.
// Memory traffic, not arithmetic, bounds all of it (bytes in, bytes out, a few passes); see DESIGN.md section 5.
#include "rto_scene.cuh"
#include "rto_nvtx.h"
#include "rto_sahchunk.h"
#include "mc_tables.h"
#include "rto_voxelize.h"
#include "rto_frustum.h"

#include <cub/cub.cuh>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace rto;

namespace {

constexpr uint8_t kMixedCell = 0xFF;
constexpr int kMaxLevels = 17;          // root edge up to 65536 voxels

// frees every cudaMalloc'ed block it was given unless released
struct DevPool {
	std::vector<void*> blocks;
	~DevPool() { for (void* p : blocks) if (p) cudaFree(p); }
	template <typename T> cudaError_t alloc(T** p, size_t count) {
		*p = nullptr;
		void* q = nullptr;
		cudaError_t e = cudaMalloc(&q, (count ? count : 1) * sizeof(T));
		if (e == cudaSuccess) { blocks.push_back(q); *p = (T*)q; }
		return e;
	}
	void release(void* p) { for (void*& b : blocks) if (b == p) b = nullptr; }
	void free_now(void* p) { for (void*& b : blocks) if (b == p) { cudaFree(p); b = nullptr; } }
};

struct PyramidDev {
	const uint8_t* data[kMaxLevels + 1];
	int nx[kMaxLevels + 1], ny[kMaxLevels + 1], nz[kMaxLevels + 1];
	int levels;
};

// state of the cell of edge 2^l at voxel origin (x, y, z); outside the grid everything is EMPTY (getVoxelSafe, OctreeVoxel.cpp:692-701)
__device__ __forceinline__ uint8_t pyr_at(const PyramidDev& P, int l, int x, int y, int z) {
	int cx = x >> l, cy = y >> l, cz = z >> l;
	if (cx >= P.nx[l] || cy >= P.ny[l] || cz >= P.nz[l]) return 0;
	return __ldg(P.data[l] + (size_t)cx + (size_t)cy * P.nx[l] + (size_t)cz * ((size_t)P.nx[l] * P.ny[l]));
}

__global__ void k_check_voxels(const uint8_t* __restrict__ v, size_t n, int* bad) {
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
	bool b = false;
	for (; i < n; i += stride) b |= (v[i] == kMixedCell);
	if (__any_sync(0xffffffffu, b) && (threadIdx.x & 31) == 0) atomicExch(bad, 1);
}

// one pyramid level: a cell is the common state of its 8 children, or "mixed"
__global__ void k_pyramid_level(const uint8_t* __restrict__ src, int px, int py, int pz, uint8_t* __restrict__ dst, int cx, int cy, int cz,
	unsigned long long* mixedCount) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
	bool isMixed = false;
	if (x < cx) {
		uint8_t first = 0; bool have = false;
#pragma unroll
		for (int k = 0; k < 8; k++) {
			int sx = 2 * x + (k & 1), sy = 2 * y + ((k >> 1) & 1), sz = 2 * z + (k >> 2);
			uint8_t v = (sx < px && sy < py && sz < pz) ? __ldg(src + (size_t)sx + (size_t)sy * px + (size_t)sz * ((size_t)px * py)) : (uint8_t)0;
			if (v == kMixedCell) isMixed = true;
			else if (!have) { first = v; have = true; }
			else if (v != first) isMixed = true;
		}
		dst[(size_t)x + (size_t)y * cx + (size_t)z * ((size_t)cx * cy)] = isMixed ? kMixedCell : first;
	}
	unsigned m = __ballot_sync(0xffffffffu, isMixed);
	if (m && (threadIdx.x & 31) == 0) atomicAdd(mixedCount, (unsigned long long)__popc(m));
}

// ---- BFS emission ------------------------------------------------------------------------------------------------
// F = internal nodes of depth d in BFS order: origin, node index, rank of the parent among all internal nodes.
struct FrontierDev { ushort4* xyz; int* gidx; int* prank; };

// children of F (8 per node): state and "is internal" flag
__global__ void k_oct_classify(PyramidDev P, FrontierDev F, int count, int childLevel, uint8_t* __restrict__ state, int* __restrict__ flag) {
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= 8 * count) return;
	const int i = j >> 3, c = j & 7, half = 1 << childLevel;
	const ushort4 p = F.xyz[i];
	const uint8_t s = pyr_at(P, childLevel, p.x + ((c & 1) ? half : 0), p.y + ((c & 2) ? half : 0), p.z + ((c & 4) ? half : 0));
	state[j] = s;
	flag[j] = (childLevel > 0 && s == kMixedCell) ? 1 : 0;
}

struct OctOut {
	uint32_t* desc;        // compact layout (offset by 7 words, see host_layouts.cpp), may be null
	RtoGpuNode* nodes;     // reference layout, may be null
};

// writes the nodes of depth d+1 (index base1 + j), appends the internal ones to the next frontier
__global__ void k_oct_emit(FrontierDev F, int count, int innerBase, int childLevel, const uint8_t* __restrict__ state, const int* __restrict__ flag,
	const int* __restrict__ ex, int base1, int base2, FrontierDev next, OctOut out) {
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= 8 * count) return;
	const int i = j >> 3, c = j & 7, half = 1 << childLevel;
	const ushort4 p = F.xyz[i];
	const int x = p.x + ((c & 1) ? half : 0), y = p.y + ((c & 2) ? half : 0), z = p.z + ((c & 4) ? half : 0);
	const int gidx = base1 + j;
	const bool internal = flag[j] != 0;
	const uint8_t s = state[j];
	const int firstChild = base2 + 8 * ex[j];
	if (internal) {
		const int r = ex[j];
		next.xyz[r] = make_ushort4((unsigned short)x, (unsigned short)y, (unsigned short)z, 0);
		next.gidx[r] = gidx;
		next.prank[r] = innerBase + i;
	}
	if (out.desc) out.desc[7 + gidx] = internal ? (uint32_t)firstChild : (kOctLeaf | (s == 1 ? kOctSolid : 0u));
	if (out.nodes) {
		RtoGpuNode n;
		n.x = x; n.y = y; n.z = z; n.size = half;
		n.isLeaf = internal ? 0 : 1; n.isUniform = n.isLeaf; n.isSolid = (!internal && s == 1) ? 1 : 0;
#pragma unroll
		for (int k = 0; k < 8; k++) n.child[k] = internal ? firstChild + k : -1;
		out.nodes[gidx] = n;
	}
}

// the 16-byte record of every internal node of depth d (host_layouts.cpp: rto_build_octree_layout) and the parent link of its children
__global__ void k_oct_inner(FrontierDev F, int count, int innerBase, int innerBaseNext, const uint8_t* __restrict__ state, const int* __restrict__ flag,
	const int* __restrict__ ex, int base1, int4* __restrict__ inner, int32_t* __restrict__ up) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	unsigned leafMask = 0, solidMask = 0;
#pragma unroll
	for (int c = 0; c < 8; c++) {
		if (!flag[8 * i + c]) { leafMask |= 1u << c; if (state[8 * i + c] == 1) solidMask |= 1u << c; }
	}
	const int gidx = F.gidx[i];
	const unsigned oct = gidx > 0 ? (unsigned)((gidx - 1) & 7) : 0u;
	int4 r;
	r.x = base1 + 8 * i;
	r.y = (leafMask != 0xffu) ? innerBaseNext + ex[8 * i] : 0;
	r.z = F.prank[i];
	r.w = (int)(leafMask | (solidMask << 8) | (oct << 16));
	inner[innerBase + i] = r;
	up[(base1 + 8 * i - 1) >> 3] = gidx;
}

__global__ void k_oct_root(PyramidDev P, OctOut out, int rootLevel, int numNodes) {
	const uint8_t s = pyr_at(P, rootLevel, 0, 0, 0);
	const bool internal = rootLevel > 0 && s == kMixedCell;
	if (out.desc) out.desc[7] = internal ? 1u : (kOctLeaf | (s == 1 ? kOctSolid : 0u));
	if (out.nodes) {
		RtoGpuNode n;
		n.x = 0; n.y = 0; n.z = 0; n.size = 1 << rootLevel;
		n.isLeaf = internal ? 0 : 1; n.isUniform = n.isLeaf; n.isSolid = (!internal && s == 1) ? 1 : 0;
		for (int k = 0; k < 8; k++) n.child[k] = internal ? 1 + k : -1;
		out.nodes[0] = n;
	}
}

struct OctBuild {
	DevPool pool;
	PyramidDev P{};
	const uint8_t* dVoxels = nullptr;   // device copy of the grid (level 0)
	int dims[3] = { 0, 0, 0 };
	int rootLevel = 0;
	size_t numNodes = 0, numInner = 0, numLeaves = 0;
	std::vector<size_t> innerAtDepth;    // internal nodes per depth (depth 0 = root)
	// outputs (device), allocated by build()
	uint32_t* desc = nullptr; int32_t* up = nullptr; int4* inner = nullptr; RtoGpuNode* nodes = nullptr;
	size_t descWords = 0, upWords = 0, innerRecs = 0;
};

#define BUILD_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return rto_fail(RTO_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); } while (0)

// Uploads the grid (unless it already is on the device) and builds the uniformity pyramid; fills the node counts.
int oct_pyramid(OctBuild& B, const uint8_t* voxels, bool onDevice, int dimX, int dimY, int dimZ, cudaStream_t st) {
	const size_t nvox = (size_t)dimX * dimY * dimZ;
	int maxDim = dimX > dimY ? dimX : dimY; if (dimZ > maxDim) maxDim = dimZ;
	int root = 1, levels = 0;
	while (root < maxDim) { root <<= 1; levels++; }
	if (levels > kMaxLevels - 1) return rto_fail(RTO_ERR_UNSUPPORTED, "octree build on the device: grid edge above 65536 voxels");
	B.rootLevel = levels; B.dims[0] = dimX; B.dims[1] = dimY; B.dims[2] = dimZ;
	if (onDevice) B.dVoxels = voxels;
	else {
		uint8_t* d = nullptr;
		BUILD_TRY(B.pool.alloc(&d, nvox));
		BUILD_TRY(cudaMemcpyAsync(d, voxels, nvox, cudaMemcpyHostToDevice, st));
		B.dVoxels = d;
	}
	int* dBad = nullptr; unsigned long long* dMixed = nullptr;
	BUILD_TRY(B.pool.alloc(&dBad, 1));
	BUILD_TRY(B.pool.alloc(&dMixed, kMaxLevels + 1));
	BUILD_TRY(cudaMemsetAsync(dBad, 0, sizeof(int), st));
	BUILD_TRY(cudaMemsetAsync(dMixed, 0, sizeof(unsigned long long) * (kMaxLevels + 1), st));
	k_check_voxels<<<1184, 256, 0, st>>>(B.dVoxels, nvox, dBad);
	B.P.levels = levels;
	B.P.data[0] = B.dVoxels; B.P.nx[0] = dimX; B.P.ny[0] = dimY; B.P.nz[0] = dimZ;
	for (int l = 1; l <= levels; l++) {
		const int px = B.P.nx[l - 1], py = B.P.ny[l - 1], pz = B.P.nz[l - 1];
		const int cx = (px + 1) / 2, cy = (py + 1) / 2, cz = (pz + 1) / 2;
		uint8_t* d = nullptr;
		BUILD_TRY(B.pool.alloc(&d, (size_t)cx * cy * cz));
		if (cy > 65535 || cz > 65535) return rto_fail(RTO_ERR_UNSUPPORTED, "octree build on the device: pyramid level too large");
		dim3 block(128), grid((cx + 127) / 128, cy, cz);
		k_pyramid_level<<<grid, block, 0, st>>>(B.P.data[l - 1], px, py, pz, d, cx, cy, cz, dMixed + l);
		B.P.data[l] = d; B.P.nx[l] = cx; B.P.ny[l] = cy; B.P.nz[l] = cz;
	}
	BUILD_TRY(cudaGetLastError());
	int bad = 0; unsigned long long mixed[kMaxLevels + 1];
	BUILD_TRY(cudaMemcpyAsync(&bad, dBad, sizeof(int), cudaMemcpyDeviceToHost, st));
	BUILD_TRY(cudaMemcpyAsync(mixed, dMixed, sizeof(mixed), cudaMemcpyDeviceToHost, st));
	BUILD_TRY(cudaStreamSynchronize(st));
	if (bad) return rto_fail(RTO_ERR_INVALID, "octree build: voxel value 255 is reserved");
	// every mixed cell of level l is an internal node of depth (levels - l): its ancestors are mixed too
	B.innerAtDepth.assign(levels + 1, 0);
	B.numInner = 0;
	for (int l = levels; l >= 1; l--) { B.innerAtDepth[levels - l] = (size_t)mixed[l]; B.numInner += (size_t)mixed[l]; }
	B.numNodes = 1 + 8 * B.numInner;
	B.numLeaves = B.numNodes - B.numInner;
	if (B.numNodes > (size_t)0x3ffffff0) return rto_fail(RTO_ERR_UNSUPPORTED, "octree build on the device: more than 2^30 nodes");
	return RTO_OK;
}

// Emits the tree level by level.  wantCompact: desc/up/inner (device layout of rto_scene_create_octree); wantNodes: GPUNodes array.
int oct_emit(OctBuild& B, bool wantCompact, bool wantNodes, cudaStream_t st) {
	if (wantCompact) {
		B.descWords = B.numNodes + 8; B.upWords = (B.numNodes + 7) / 8 + 1; B.innerRecs = B.numInner ? B.numInner : 1;
		BUILD_TRY(B.pool.alloc(&B.desc, B.descWords));
		BUILD_TRY(B.pool.alloc(&B.up, B.upWords));
		BUILD_TRY(B.pool.alloc(&B.inner, B.innerRecs));
		BUILD_TRY(cudaMemsetAsync(B.desc, 0, B.descWords * 4, st));
		BUILD_TRY(cudaMemsetAsync(B.up, 0, B.upWords * 4, st));
		BUILD_TRY(cudaMemsetAsync(B.inner, 0, B.innerRecs * 16, st));
	}
	if (wantNodes) BUILD_TRY(B.pool.alloc(&B.nodes, B.numNodes));
	OctOut out; out.desc = B.desc; out.nodes = B.nodes;
	k_oct_root<<<1, 1, 0, st>>>(B.P, out, B.rootLevel, (int)B.numNodes);
	size_t maxF = 1;
	for (size_t c : B.innerAtDepth) if (c > maxF) maxF = c;
	if (B.numInner == 0) { BUILD_TRY(cudaGetLastError()); return RTO_OK; }
	FrontierDev F[2];
	for (int k = 0; k < 2; k++) {
		BUILD_TRY(B.pool.alloc(&F[k].xyz, maxF)); BUILD_TRY(B.pool.alloc(&F[k].gidx, maxF)); BUILD_TRY(B.pool.alloc(&F[k].prank, maxF));
	}
	uint8_t* state = nullptr; int *flag = nullptr, *ex = nullptr;
	BUILD_TRY(B.pool.alloc(&state, 8 * maxF)); BUILD_TRY(B.pool.alloc(&flag, 8 * maxF)); BUILD_TRY(B.pool.alloc(&ex, 8 * maxF));
	size_t tmpBytes = 0;
	BUILD_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, flag, ex, (int)(8 * maxF), st));
	void* tmp = nullptr;
	{ uint8_t* t8 = nullptr; BUILD_TRY(B.pool.alloc(&t8, tmpBytes)); tmp = t8; }
	{	// depth 0: the root
		ushort4 r0 = make_ushort4(0, 0, 0, 0); int zero = 0;
		BUILD_TRY(cudaMemcpyAsync(F[0].xyz, &r0, sizeof(r0), cudaMemcpyHostToDevice, st));
		BUILD_TRY(cudaMemcpyAsync(F[0].gidx, &zero, 4, cudaMemcpyHostToDevice, st));
		BUILD_TRY(cudaMemcpyAsync(F[0].prank, &zero, 4, cudaMemcpyHostToDevice, st));
		BUILD_TRY(cudaStreamSynchronize(st));      // (r0, zero are stack variables)
	}
	size_t base1 = 1, innerBase = 0;
	int cur = 0;
	for (int d = 0; d < B.rootLevel; d++) {
		const size_t count = B.innerAtDepth[d];
		if (count == 0) break;
		const int childLevel = B.rootLevel - d - 1;
		const size_t nChildren = 8 * count, base2 = base1 + nChildren, innerBaseNext = innerBase + count;
		const unsigned blocksC = (unsigned)((nChildren + 255) / 256), blocksP = (unsigned)((count + 255) / 256);
		k_oct_classify<<<blocksC, 256, 0, st>>>(B.P, F[cur], (int)count, childLevel, state, flag);
		BUILD_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, flag, ex, (int)nChildren, st));
		k_oct_emit<<<blocksC, 256, 0, st>>>(F[cur], (int)count, (int)innerBase, childLevel, state, flag, ex, (int)base1, (int)base2, F[cur ^ 1], out);
		if (wantCompact) k_oct_inner<<<blocksP, 256, 0, st>>>(F[cur], (int)count, (int)innerBase, (int)innerBaseNext, state, flag, ex, (int)base1, B.inner, B.up);
		base1 = base2; innerBase = innerBaseNext; cur ^= 1;
	}
	BUILD_TRY(cudaGetLastError());
	if (base1 != B.numNodes || innerBase != B.numInner) return rto_fail(RTO_ERR_CUDA, "octree build on the device: node count mismatch (%zu vs %zu)", base1, B.numNodes);
	return RTO_OK;
}

} // namespace

// ------------------------------------------------------------------------------------------------
// C ABI: octree
// ------------------------------------------------------------------------------------------------
extern "C" int rto_device_octree_build(const uint8_t* voxels, int dimX, int dimY, int dimZ, RtoGpuNode** nodesOut, size_t* numNodes) try {
	RTO_RANGE("rto_device_octree_build");
	if (!nodesOut || !numNodes) return rto_fail(RTO_ERR_INVALID, "rto_device_octree_build: null output");
	*nodesOut = nullptr; *numNodes = 0;
	if (dimX == 0 || dimY == 0 || dimZ == 0) return RTO_OK;        // createOctreeFromVoxelGrid returns nullptr (OctreeVoxel.cpp:766)
	if (!voxels || dimX < 0 || dimY < 0 || dimZ < 0) return rto_fail(RTO_ERR_INVALID, "rto_device_octree_build: bad grid");
	int rc = rto_require_device(); if (rc) return rc;
	cudaStream_t st = nullptr;
	BUILD_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
	OctBuild B;
	rc = oct_pyramid(B, voxels, false, dimX, dimY, dimZ, st);
	if (!rc) rc = oct_emit(B, false, true, st);
	RtoGpuNode* host = nullptr;
	if (!rc) {
		host = (RtoGpuNode*)std::malloc(B.numNodes * sizeof(RtoGpuNode));
		if (!host) rc = rto_fail(RTO_ERR_ALLOC, "rto_device_octree_build: out of host memory");
	}
	if (!rc) {
		cudaError_t e = cudaMemcpyAsync(host, B.nodes, B.numNodes * sizeof(RtoGpuNode), cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess) e = cudaStreamSynchronize(st);
		if (e != cudaSuccess) { std::free(host); host = nullptr; rc = rto_fail(RTO_ERR_CUDA, "rto_device_octree_build: %s", cudaGetErrorString(e)); }
	}
	cudaStreamSynchronize(st);
	cudaStreamDestroy(st);
	if (rc) return rc;
	*nodesOut = host; *numNodes = B.numNodes;
	return RTO_OK;
} RTO_CATCH_ALL("rto_device_octree_build")

extern "C" int rto_scene_create_octree_from_grid(const uint8_t* voxels, int dimX, int dimY, int dimZ, const float gridMin[3], float voxelSize,
	RtoScene** out) try {
	RTO_RANGE("rto_scene_create_octree_from_grid");
	if (!out) return rto_fail(RTO_ERR_INVALID, "rto_scene_create_octree_from_grid: null output");
	*out = nullptr;
	if (!voxels || !gridMin || dimX <= 0 || dimY <= 0 || dimZ <= 0) return rto_fail(RTO_ERR_INVALID, "rto_scene_create_octree_from_grid: empty grid (nothing to trace)");
	RtoScene* s = nullptr;
	int rc = rto_scene_new(&s); if (rc) return rc;
	OctBuild B;
	rc = oct_pyramid(B, voxels, false, dimX, dimY, dimZ, s->stream);
	if (!rc) rc = oct_emit(B, true, false, s->stream);
	if (!rc) { cudaError_t e = cudaStreamSynchronize(s->stream); if (e != cudaSuccess) rc = rto_fail(RTO_ERR_CUDA, "octree build on the device failed: %s", cudaGetErrorString(e)); }
	if (rc) { rto_scene_destroy(s); return rc; }
	s->kind = RTO_MODE_OCTREE_GLSL;
	s->numNodes = B.numNodes; s->numPrims = B.numLeaves;
	OctDev& D = s->oct;
	D.numNodes = (int)B.numNodes; D.rootSize = 1 << B.rootLevel;
	D.gmin[0] = gridMin[0]; D.gmin[1] = gridMin[1]; D.gmin[2] = gridMin[2]; D.voxel = voxelSize;
	D.compact = 1;
	D.desc = B.desc + 7; D.up = B.up; D.inner = B.inner;
	B.pool.release(B.desc); B.pool.release(B.up); B.pool.release(B.inner);
	rto_scene_adopt(s, B.desc, B.descWords * 4); rto_scene_adopt(s, B.up, B.upWords * 4); rto_scene_adopt(s, B.inner, B.innerRecs * 16);
	*out = s;
	return RTO_OK;
} RTO_CATCH_ALL("rto_scene_create_octree_from_grid")

// Diagnostic: copies the compact device layout of an octree scene to the host (desc: numNodes + 8 words, up: (numNodes + 7) / 8 + 1
// words, inner: 4 words per internal node); any pointer may be null.  Used by the tests to compare both construction routes.
extern "C" int rto_scene_octree_layout_read(RtoScene* s, uint32_t* desc, int32_t* up, int32_t* inner4, size_t* numInnerOut) try {
	if (!s || s->kind == RTO_MODE_BVH || !s->oct.compact) return rto_fail(RTO_ERR_INVALID, "rto_scene_octree_layout_read: not a compact octree scene");
	BUILD_TRY(cudaSetDevice(s->device));
	const size_t n = s->numNodes, numInner = n > 1 ? (n - 1) / 8 : 0;
	if (numInnerOut) *numInnerOut = numInner;
	if (desc) BUILD_TRY(cudaMemcpyAsync(desc, s->oct.desc - 7, (n + 8) * 4, cudaMemcpyDeviceToHost, s->stream));
	if (up) BUILD_TRY(cudaMemcpyAsync(up, s->oct.up, ((n + 7) / 8 + 1) * 4, cudaMemcpyDeviceToHost, s->stream));
	if (inner4) BUILD_TRY(cudaMemcpyAsync(inner4, s->oct.inner, (numInner ? numInner : 1) * 16, cudaMemcpyDeviceToHost, s->stream));
	BUILD_TRY(cudaStreamSynchronize(s->stream));
	return RTO_OK;
} RTO_CATCH_ALL("rto_scene_octree_layout_read")

// ------------------------------------------------------------------------------------------------
// Marching cubes
// ------------------------------------------------------------------------------------------------
namespace {

struct McDev {
	const uint8_t* vox; int dx, dy, dz; float minX, minY, minZ, vs;
};

// sign pattern of the 8 cell corners: bit c set <=> corner c FILLED (scalar -1 < iso 0); outside the grid is EMPTY
// (OctreeVoxel.cpp:787-792); corner numbering of the reference's cornerOffset table
__device__ __forceinline__ int mc_cube_index(const McDev& g, int x, int y, int z) {
	const int off[8][3] = { {0,0,0},{1,0,0},{1,1,0},{0,1,0},{0,0,1},{1,0,1},{1,1,1},{0,1,1} };
	int cube = 0;
#pragma unroll
	for (int c = 0; c < 8; c++) {
		const int vx = x + off[c][0], vy = y + off[c][1], vz = z + off[c][2];
		const bool in = vx < g.dx && vy < g.dy && vz < g.dz;
		if (in && __ldg(g.vox + (size_t)vx + (size_t)vy * g.dx + (size_t)vz * ((size_t)g.dx * g.dy)) == 1) cube |= 1 << c;
	}
	return cube;
}

__device__ __forceinline__ uint64_t morton3(uint32_t x, uint32_t y, uint32_t z) {     // bit0 = x, bit1 = y, bit2 = z per level (child index of OctreeVoxel.cpp:749-757)
	auto spread = [](uint64_t v) {
		v &= 0x1fffffull;
		v = (v | (v << 32)) & 0x1f00000000ffffull;
		v = (v | (v << 16)) & 0x1f0000ff0000ffull;
		v = (v | (v << 8)) & 0x100f00f00f00f00full;
		v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
		v = (v | (v << 2)) & 0x1249249249249249ull;
		return v;
	};
	return spread(x) | (spread(y) << 1) | (spread(z) << 2);
}

// pass 1 (cells == nullptr): count surface cells; pass 2: append them (unordered; the sort below fixes the order)
__global__ void k_mc_find(McDev g, unsigned long long* counter, uint32_t* __restrict__ cellXY, uint16_t* __restrict__ cellZ, uint8_t* __restrict__ cellCube,
	unsigned long long capacity) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
	int cube = 0;
	if (x < g.dx - 1) cube = mc_cube_index(g, x, y, z);
	const bool active = cube != 0 && cube != 255;
	const unsigned m = __ballot_sync(0xffffffffu, active);
	if (!m) return;
	const int lane = threadIdx.x & 31, leader = __ffs((int)m) - 1;
	unsigned long long base = 0;
	if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
	base = __shfl_sync(0xffffffffu, base, leader);
	if (active && cellXY) {
		const unsigned long long slot = base + (unsigned)__popc(m & ((1u << lane) - 1u));
		if (slot < capacity) { cellXY[slot] = (uint32_t)x | ((uint32_t)y << 16); cellZ[slot] = (uint16_t)z; cellCube[slot] = (uint8_t)cube; }
	}
}

// sort key of a surface cell = its place in the reference's emission order: leaves in depth-first order (children 0..7,
// Renderer.cpp:26-33) == Morton order of the leaf; inside a leaf z, then y, then x (OctreeVoxel.cpp:799-801)
__global__ void k_mc_keys(const uint32_t* __restrict__ desc, int rootSize, const uint32_t* __restrict__ cellXY, const uint16_t* __restrict__ cellZ, size_t n,
	uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t xy = cellXY[i];
	const uint32_t x = xy & 0xffffu, y = xy >> 16, z = cellZ[i];
	// leaf containing voxel (x, y, z)
	int node = 0, size = rootSize;
	for (;;) {
		const uint32_t d = __ldg(desc + node);
		if (d & kOctLeaf) break;
		size >>= 1;
		node = (int)d + ((x & size) ? 1 : 0) + ((y & size) ? 2 : 0) + ((z & size) ? 4 : 0);
	}
	const int k = 31 - __clz(size);                       // leaf edge = 2^k
	const uint64_t lowMask = (k >= 21) ? ~0ull : ((1ull << (3 * k)) - 1ull);
	const uint32_t m = (uint32_t)size - 1u;
	const uint64_t local = ((uint64_t)(z & m) << (2 * k)) | ((uint64_t)(y & m) << k) | (uint64_t)(x & m);
	keys[i] = (morton3(x, y, z) & ~lowMask) | local;
	vals[i] = (uint32_t)i;
}

__device__ __forceinline__ int mc_tri_count(const uint64_t* __restrict__ triTable, int cube) {
	const uint64_t t = __ldg(triTable + cube);
	int n = 0;
	while (n < 15 && ((t >> (4 * n)) & 0xF) != 0xF) n += 3;
	return n / 3;
}

__global__ void k_mc_count(const uint64_t* __restrict__ triTable, const uint32_t* __restrict__ order, const uint8_t* __restrict__ cellCube, size_t n, int* __restrict__ counts) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	counts[i] = mc_tri_count(triTable, cellCube[order[i]]);
}

__device__ __forceinline__ V3 mc_vertex_interp(V3 p1, V3 p2, float v1, float v2) {    // iso level 0 (OctreeVoxel.cpp:633-640)
	if (fabsf(0.0f - v1) < 0.00001f) return p1;
	if (fabsf(0.0f - v2) < 0.00001f) return p2;
	if (fabsf(v1 - v2) < 0.00001f) return p1;
	float mu = (0.0f - v1) / (v2 - v1);
	return p1 + mu * (p2 - p1);
}

__global__ void k_mc_emit(McDev g, const uint64_t* __restrict__ triTable, const uint32_t* __restrict__ order, const uint32_t* __restrict__ cellXY,
	const uint16_t* __restrict__ cellZ, const uint8_t* __restrict__ cellCube, const long long* __restrict__ offsets, size_t n, RtoTriangle* __restrict__ out) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t c = order[i];
	const uint32_t xy = cellXY[c];
	const int x = (int)(xy & 0xffffu), y = (int)(xy >> 16), z = (int)cellZ[c];
	const int cube = cellCube[c];
	const int off[8][3] = { {0,0,0},{1,0,0},{1,1,0},{0,1,0},{0,0,1},{1,0,1},{1,1,1},{0,1,1} };
	const int edgeCorner[12][2] = { {0,1},{1,2},{2,3},{3,0},{4,5},{5,6},{6,7},{7,4},{0,4},{1,5},{2,6},{3,7} };
	const uint64_t tt = __ldg(triTable + cube);
	RtoTriangle* dst = out + offsets[i];
	for (int t = 0; t < 5; t++) {
		if (((tt >> (12 * t)) & 0xF) == 0xF) break;
		float v[9];
#pragma unroll
		for (int k = 0; k < 3; k++) {
			const int e = (int)((tt >> (12 * t + 4 * k)) & 0xF);
			const int a = edgeCorner[e][0], b = edgeCorner[e][1];
			const float va = ((cube >> a) & 1) ? -1.0f : 1.0f, vb = ((cube >> b) & 1) ? -1.0f : 1.0f;
			// corner positions: grid.min + (x + dx) * voxelSize (OctreeVoxel.cpp:806-811; int -> float, one product, one sum)
			V3 pa = mk3(g.minX + float(x + off[a][0]) * g.vs, g.minY + float(y + off[a][1]) * g.vs, g.minZ + float(z + off[a][2]) * g.vs);
			V3 pb = mk3(g.minX + float(x + off[b][0]) * g.vs, g.minY + float(y + off[b][1]) * g.vs, g.minZ + float(z + off[b][2]) * g.vs);
			V3 p = mc_vertex_interp(pa, pb, va, vb);
			v[3 * k] = p.x; v[3 * k + 1] = p.y; v[3 * k + 2] = p.z;
		}
		RtoTriangle tri;
		tri.v0[0] = v[0]; tri.v0[1] = v[1]; tri.v0[2] = v[2]; tri.v1[0] = v[3]; tri.v1[1] = v[4]; tri.v1[2] = v[5]; tri.v2[0] = v[6]; tri.v2[1] = v[7]; tri.v2[2] = v[8];
		dst[t] = tri;
	}
}

struct McBuild {
	DevPool pool;
	RtoTriangle* tris = nullptr;     // device, reference emission order
	size_t numTris = 0;
};

// needs B.P (pyramid not used), B.dVoxels and the compact desc[] of the octree (for the leaf of every surface cell)
int mc_extract(const OctBuild& O, const float gridMin[3], float voxelSize, McBuild& M, cudaStream_t st) {
	McDev g{ O.dVoxels, O.dims[0], O.dims[1], O.dims[2], gridMin[0], gridMin[1], gridMin[2], voxelSize };
	if (g.dx < 2 || g.dy < 2 || g.dz < 2) return RTO_OK;            // no cell has all 8 corners: localMC's loops are empty (OctreeVoxel.cpp:795-801)
	if (g.dy - 1 > 65535 || g.dz - 1 > 65535) return rto_fail(RTO_ERR_UNSUPPORTED, "mesh extraction on the device: grid too large");
	unsigned long long* dCounter = nullptr;
	BUILD_TRY(M.pool.alloc(&dCounter, 1));
	BUILD_TRY(cudaMemsetAsync(dCounter, 0, 8, st));
	dim3 block(128), grid((g.dx - 1 + 127) / 128, g.dy - 1, g.dz - 1);
	k_mc_find<<<grid, block, 0, st>>>(g, dCounter, nullptr, nullptr, nullptr, 0);
	unsigned long long nCells = 0;
	BUILD_TRY(cudaMemcpyAsync(&nCells, dCounter, 8, cudaMemcpyDeviceToHost, st));
	BUILD_TRY(cudaStreamSynchronize(st));
	if (nCells == 0) return RTO_OK;
	if (nCells >= 0xffffffffull) return rto_fail(RTO_ERR_UNSUPPORTED, "mesh extraction on the device: more than 2^32 surface cells");
	uint32_t* cellXY = nullptr; uint16_t* cellZ = nullptr; uint8_t* cellCube = nullptr;
	BUILD_TRY(M.pool.alloc(&cellXY, nCells)); BUILD_TRY(M.pool.alloc(&cellZ, nCells)); BUILD_TRY(M.pool.alloc(&cellCube, nCells));
	BUILD_TRY(cudaMemsetAsync(dCounter, 0, 8, st));
	k_mc_find<<<grid, block, 0, st>>>(g, dCounter, cellXY, cellZ, cellCube, nCells);
	uint64_t *keys = nullptr, *keysOut = nullptr; uint32_t *vals = nullptr, *order = nullptr;
	BUILD_TRY(M.pool.alloc(&keys, nCells)); BUILD_TRY(M.pool.alloc(&keysOut, nCells)); BUILD_TRY(M.pool.alloc(&vals, nCells)); BUILD_TRY(M.pool.alloc(&order, nCells));
	const unsigned blocksN = (unsigned)((nCells + 255) / 256);
	k_mc_keys<<<blocksN, 256, 0, st>>>(O.desc + 7, 1 << O.rootLevel, cellXY, cellZ, nCells, keys, vals);
	size_t sortBytes = 0, scanBytes = 0;
	const int endBit = 3 * O.rootLevel > 0 ? 3 * O.rootLevel : 1;
	BUILD_TRY(cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, keys, keysOut, vals, order, (int)nCells, 0, endBit, st));
	int* counts = nullptr; long long* offsets = nullptr;
	BUILD_TRY(M.pool.alloc(&counts, nCells)); BUILD_TRY(M.pool.alloc(&offsets, nCells + 1));
	BUILD_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, counts, offsets, (int)nCells, st));
	uint8_t* tmp = nullptr;
	BUILD_TRY(M.pool.alloc(&tmp, sortBytes > scanBytes ? sortBytes : scanBytes));
	BUILD_TRY(cub::DeviceRadixSort::SortPairs(tmp, sortBytes, keys, keysOut, vals, order, (int)nCells, 0, endBit, st));
	uint64_t* dTable = nullptr;
	BUILD_TRY(M.pool.alloc(&dTable, 256));
	BUILD_TRY(cudaMemcpyAsync(dTable, kMcTriPacked, 256 * 8, cudaMemcpyHostToDevice, st));
	k_mc_count<<<blocksN, 256, 0, st>>>(dTable, order, cellCube, nCells, counts);
	BUILD_TRY(cub::DeviceScan::ExclusiveSum(tmp, scanBytes, counts, offsets, (int)nCells, st));
	long long lastOff = 0; int lastCnt = 0;
	BUILD_TRY(cudaMemcpyAsync(&lastOff, offsets + (nCells - 1), 8, cudaMemcpyDeviceToHost, st));
	BUILD_TRY(cudaMemcpyAsync(&lastCnt, counts + (nCells - 1), 4, cudaMemcpyDeviceToHost, st));
	BUILD_TRY(cudaStreamSynchronize(st));
	M.numTris = (size_t)(lastOff + lastCnt);
	if (M.numTris == 0) return RTO_OK;
	M.pool.free_now(keys); M.pool.free_now(keysOut); M.pool.free_now(vals);
	BUILD_TRY(M.pool.alloc(&M.tris, M.numTris));
	k_mc_emit<<<blocksN, 256, 0, st>>>(g, dTable, order, cellXY, cellZ, cellCube, offsets, nCells, M.tris);
	BUILD_TRY(cudaGetLastError());
	return RTO_OK;
}

} // namespace

extern "C" int rto_device_mc_mesh(const uint8_t* voxels, int dimX, int dimY, int dimZ, const float gridMin[3], float voxelSize,
	RtoTriangle** trisOut, size_t* numTris) try {
	RTO_RANGE("rto_device_mc_mesh");
	if (!trisOut || !numTris) return rto_fail(RTO_ERR_INVALID, "rto_device_mc_mesh: null output");
	*trisOut = nullptr; *numTris = 0;
	if (dimX == 0 || dimY == 0 || dimZ == 0) return RTO_OK;
	if (!voxels || !gridMin || dimX < 0 || dimY < 0 || dimZ < 0) return rto_fail(RTO_ERR_INVALID, "rto_device_mc_mesh: bad grid");
	int rc = rto_require_device(); if (rc) return rc;
	cudaStream_t st = nullptr;
	BUILD_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
	OctBuild B; McBuild M;
	rc = oct_pyramid(B, voxels, false, dimX, dimY, dimZ, st);
	if (!rc) rc = oct_emit(B, true, false, st);
	if (!rc) rc = mc_extract(B, gridMin, voxelSize, M, st);
	RtoTriangle* host = nullptr;
	if (!rc && M.numTris) {
		host = (RtoTriangle*)std::malloc(M.numTris * sizeof(RtoTriangle));
		if (!host) rc = rto_fail(RTO_ERR_ALLOC, "rto_device_mc_mesh: out of host memory");
		else {
			cudaError_t e = cudaMemcpyAsync(host, M.tris, M.numTris * sizeof(RtoTriangle), cudaMemcpyDeviceToHost, st);
			if (e == cudaSuccess) e = cudaStreamSynchronize(st);
			if (e != cudaSuccess) { std::free(host); host = nullptr; rc = rto_fail(RTO_ERR_CUDA, "rto_device_mc_mesh: %s", cudaGetErrorString(e)); }
		}
	}
	cudaStreamSynchronize(st);
	cudaStreamDestroy(st);
	if (rc) return rc;
	*trisOut = host; *numTris = M.numTris;
	return RTO_OK;
} RTO_CATCH_ALL("rto_device_mc_mesh")

// ------------------------------------------------------------------------------------------------
// BVH build on the device: the fast, NON-reference-shaped mode.
//
// BVH::build (BVH.cpp:33-71) sorts with std::sort, whose order on equal centroids is a property of one libstdc++ version on one
// input sequence; no parallel algorithm can reproduce that tree, and rto_host_bvh_build stays the route whose leaves (hence
// candidate sets, hence hit ids) equal the reference's exactly.  This route builds a linear BVH instead (Morton sort, leaves of
// two Morton-neighbours, Karras' radix tree, bottom-up exact union boxes) in milliseconds where the host route takes seconds to
// a minute.  The traversal kernels and the closest-hit rule (min t, then position) are the same; hit ids can differ from the
// reference only where two triangles are hit at the same t to the last bit (shared edges) or a ray grazes the edge of a leaf
// box -- the "documented near-tie" class of the north star; tests/test_gpu_builders.py measures the rate against the oracle.
// ------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ int float_as_ordered(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ordered_as_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void k_lbvh_bounds(const RtoTriangle* __restrict__ tris, size_t n, int* bounds /* lo xyz, hi xyz as ordered ints */) {
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
	float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
	for (; i < n; i += stride) {
		const float* v = reinterpret_cast<const float*>(tris + i);
#pragma unroll
		for (int k = 0; k < 9; k++) { float f = v[k]; lo[k % 3] = fminf(lo[k % 3], f); hi[k % 3] = fmaxf(hi[k % 3], f); }
	}
#pragma unroll
	for (int a = 0; a < 3; a++) {
		for (int o = 16; o > 0; o >>= 1) { lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o)); hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o)); }
		if ((threadIdx.x & 31) == 0) { atomicMin(bounds + a, float_as_ordered(lo[a])); atomicMax(bounds + 3 + a, float_as_ordered(hi[a])); }
	}
}

__global__ void k_lbvh_codes(const RtoTriangle* __restrict__ tris, size_t n, const int* __restrict__ bounds, uint64_t* __restrict__ codes, uint32_t* __restrict__ ids) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float* v = reinterpret_cast<const float*>(tris + i);
	// one scale for all three axes (the largest extent): Morton cells stay cubes, which matters for flat scenes such as the DT
	// grid (425 x 243 x 29 voxels), where per-axis scaling made the tree 3x slower to traverse
	float ext = 0.0f;
#pragma unroll
	for (int a = 0; a < 3; a++) ext = fmaxf(ext, ordered_as_float(bounds[3 + a]) - ordered_as_float(bounds[a]));
	uint32_t q[3];
#pragma unroll
	for (int a = 0; a < 3; a++) {
		const float lo = ordered_as_float(bounds[a]);
		const float c = (v[a] + v[3 + a] + v[6 + a]) * (1.0f / 3.0f);
		float u = ext > 0.0f ? (c - lo) / ext : 0.0f;
		u = fminf(fmaxf(u, 0.0f), 1.0f);
		q[a] = (uint32_t)fminf(u * 2097152.0f, 2097151.0f);
	}
	codes[i] = morton3(q[0], q[1], q[2]);
	ids[i] = (uint32_t)i;
}

// triangle records in sorted order (v0, e1, e2, id: rto_kernels.cuh TriV) and the exact box of every leaf (two neighbours)
__global__ void k_lbvh_leaves(const RtoTriangle* __restrict__ tris, const uint32_t* __restrict__ order, size_t n, int perLeaf, float grow, float4* __restrict__ rec, float* __restrict__ leafBox /* 6 per leaf */) {
	const size_t leaf = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const size_t numLeaves = (n + perLeaf - 1) / perLeaf;
	if (leaf >= numLeaves) return;
	float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
	for (size_t p = perLeaf * leaf; p < perLeaf * leaf + perLeaf && p < n; p++) {
		const uint32_t id = order[p];
		const float* v = reinterpret_cast<const float*>(tris + id);
		float f[9];
#pragma unroll
		for (int k = 0; k < 9; k++) { f[k] = v[k]; lo[k % 3] = fminf(lo[k % 3], f[k]); hi[k % 3] = fmaxf(hi[k % 3], f[k]); }
		rec[4 * p] = make_float4(f[0], f[1], f[2], f[3] - f[0]);
		rec[4 * p + 1] = make_float4(f[4] - f[1], f[5] - f[2], f[6] - f[0], f[7] - f[1]);
		rec[4 * p + 2] = make_float4(f[8] - f[2], __int_as_float((int)id), 0.0f, 0.0f);      // (reference-leaf box unused: BvhDev::leafBox == 0)
		rec[4 * p + 3] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	}
#pragma unroll
	for (int a = 0; a < 3; a++) { leafBox[6 * leaf + a] = lo[a] - grow; leafBox[6 * leaf + 3 + a] = hi[a] + grow; }     // conservative boxes: BvhDev::grow
}

// common-prefix length of the keys of leaves i and j (key = Morton code of the leaf's first triangle, ties broken by index)
__device__ __forceinline__ int lbvh_delta(const uint64_t* __restrict__ codes, int perLeaf, int numLeaves, int i, int j) {
	if (j < 0 || j >= numLeaves) return -1;
	const uint64_t a = codes[perLeaf * (size_t)i], b = codes[perLeaf * (size_t)j];
	if (a == b) return 64 + __clz(i ^ j);
	return __clzll((long long)(a ^ b));
}

// Karras 2012: internal node i of the binary radix tree over the sorted leaves; child refs into the node, parent links for the fit pass
__global__ void k_lbvh_tree(const uint64_t* __restrict__ codes, int perLeaf, int numLeaves, size_t numTris, float4* __restrict__ nodes, int* __restrict__ parentOfInner, int* __restrict__ parentOfLeaf,
	int2* __restrict__ rangeOf /* may be null: [first, last] leaf of every internal node */) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= numLeaves - 1) return;
	const int d = (lbvh_delta(codes, perLeaf, numLeaves, i, i + 1) - lbvh_delta(codes, perLeaf, numLeaves, i, i - 1)) >= 0 ? 1 : -1;
	const int dmin = lbvh_delta(codes, perLeaf, numLeaves, i, i - d);
	int lmax = 2;
	while (lbvh_delta(codes, perLeaf, numLeaves, i, i + lmax * d) > dmin) lmax <<= 1;
	int l = 0;
	for (int t = lmax >> 1; t >= 1; t >>= 1) if (lbvh_delta(codes, perLeaf, numLeaves, i, i + (l + t) * d) > dmin) l += t;
	const int j = i + l * d;
	const int dnode = lbvh_delta(codes, perLeaf, numLeaves, i, j);
	int s = 0;
	for (int t = (l + 1) >> 1; ; t = (t + 1) >> 1) {
		if (lbvh_delta(codes, perLeaf, numLeaves, i, i + (s + t) * d) > dnode) s += t;
		if (t == 1) break;
	}
	const int gamma = i + s * d + min(d, 0);
	const int first = min(i, j), last = max(i, j);
	if (rangeOf) rangeOf[i] = make_int2(first, last);
	auto leafRef = [&](int leaf) {
		const size_t p = perLeaf * (size_t)leaf;
		const int cnt = (perLeaf == 2 && p + 1 < numTris) ? 2 : 1;
		return ~(int)((p << 1) | (size_t)(cnt - 1));
	};
	int r0, r1;
	if (gamma == first) { r0 = leafRef(gamma); parentOfLeaf[gamma] = 2 * i; } else { r0 = gamma; parentOfInner[gamma] = 2 * i; }
	if (gamma + 1 == last) { r1 = leafRef(gamma + 1); parentOfLeaf[gamma + 1] = 2 * i + 1; } else { r1 = gamma + 1; parentOfInner[gamma + 1] = 2 * i + 1; }
	nodes[4 * (size_t)i + 3] = make_float4(__int_as_float(r0), __int_as_float(r1), 0.0f, 0.0f);
	if (i == 0) parentOfInner[0] = -1;
}

// the subtrees the surface-area rebuild takes (rto_sahchunk.h): at most kSahChunk leaves, and the parent has more
__global__ void k_sah_select(int numInner, const int2* __restrict__ rangeOf, const int* __restrict__ parentOfInner, int* __restrict__ list, int* __restrict__ count) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= numInner) return;
	const int2 r = rangeOf[i];
	const int m = r.y - r.x + 1;
	if (m < 3 || m > kSahChunk) return;
	const int link = parentOfInner[i];
	if (link >= 0) { const int2 pr = rangeOf[link >> 1]; if (pr.y - pr.x + 1 <= kSahChunk) return; }
	list[atomicAdd(count, 1)] = i;
}
// one block per subtree (rto_sahchunk.h, sah_rebuild_chunk_block)
__global__ void __launch_bounds__(kSahBlock) k_sah_rebuild(const int* __restrict__ list, const int2* __restrict__ rangeOf, const float* __restrict__ leafBox, float4* __restrict__ nodes,
	int* __restrict__ parentOfInner, int* __restrict__ parentOfLeaf) {
	const int i = list[blockIdx.x];
	const int2 r = rangeOf[i];
	sah_rebuild_chunk_block(leafBox, r.x, r.y, i, nodes, parentOfInner, parentOfLeaf);
}
// one thread per subtree: the form the CPU emulation checks, kept as a cross-check of the block form (RTO_SAH_SERIAL=1; tests/test_gpu_builders.py)
__global__ void __launch_bounds__(64) k_sah_rebuild_serial(int count, const int* __restrict__ list, const int2* __restrict__ rangeOf, const float* __restrict__ leafBox, float4* __restrict__ nodes,
	int* __restrict__ parentOfInner, int* __restrict__ parentOfLeaf) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= count) return;
	const int i = list[t];
	const int2 r = rangeOf[i];
	sah_rebuild_chunk(leafBox, r.x, r.y, i, nodes, parentOfInner, parentOfLeaf);
}

__device__ __forceinline__ void lbvh_store_child_box(float* node16, int slot, const float lo[3], const float hi[3]) {
	// paired layout (BvhDev::paired): plane k of both children side by side
	for (int k = 0; k < 3; k++) { node16[2 * k + slot] = lo[k]; node16[6 + 2 * k + slot] = hi[k]; }
}

// bottom-up: every leaf writes its box into its parent's slot; the second child to arrive at a node unites the two and climbs
__global__ void k_lbvh_fit(int numLeaves, const float* __restrict__ leafBox, float* nodes /* 16 floats per inner node */, const int* __restrict__ parentOfInner,
	const int* __restrict__ parentOfLeaf, int* arrived, float* rootBox, int* maxDepth) {
	const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
	if (leaf >= numLeaves) return;
	{	// depth of this leaf = upper bound of the traversal stack a ray can need on its way here
		int depth = 1, up = parentOfLeaf[leaf] >> 1;
		while ((up = parentOfInner[up]) >= 0) { up >>= 1; depth++; }
		if (depth > 92) atomicMax(maxDepth, depth);
	}
	float lo[3], hi[3];
#pragma unroll
	for (int a = 0; a < 3; a++) { lo[a] = leafBox[6 * (size_t)leaf + a]; hi[a] = leafBox[6 * (size_t)leaf + 3 + a]; }
	int link = parentOfLeaf[leaf];
	for (;;) {
		const int node = link >> 1, slot = link & 1;
		float* n16 = nodes + 16 * (size_t)node;
		lbvh_store_child_box(n16, slot, lo, hi);
		__threadfence();
		if (atomicAdd(arrived + node, 1) == 0) return;          // the sibling subtree is not finished yet: its thread will continue
		__threadfence();                                        // acquire side: the sibling's box (stored before ITS fence + atomic) is read after our atomic
		volatile float* other = n16 + (slot ^ 1);
#pragma unroll
		for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], other[2 * a]); hi[a] = fmaxf(hi[a], other[6 + 2 * a]); }
		link = parentOfInner[node];
		if (link < 0) {
#pragma unroll
			for (int a = 0; a < 3; a++) { rootBox[a] = lo[a]; rootBox[3 + a] = hi[a]; }
			return;
		}
	}
}

// tris: device array of numTris RtoTriangle (reference emission order, index == hit id).  Fills s->bvh / s->bvhFast.
int lbvh_build(RtoScene* s, const RtoTriangle* dTris, size_t numTris) {
	cudaStream_t st = s->stream;
	BvhDev D{};
	D.numTris = (int)numTris; D.rootRef = -1; D.leafBox = 0; D.grow = 0.0f; D.paired = 1; D.exactPaired = 1;
	s->numPrims = numTris;
	if (numTris == 0) { s->bvh = D; s->bvhFast = D; s->numNodes = 0; return RTO_OK; }
	if (numTris >= (size_t)1 << 30) return rto_fail(RTO_ERR_UNSUPPORTED, "BVH build on the device: more than 2^30 triangles");
	DevPool tmp;
	// triangles per leaf: 1 (measured 1.2-1.7x faster to traverse than leaves of two Morton neighbours: tight leaf boxes save
	// Moller-Trumbore tests, the block that runs with the fewest lanes) or 2 (RTO_LBVH_LEAF=2, tuning aid)
	static const int perLeaf = [] { const char* e = getenv("RTO_LBVH_LEAF"); return (e && e[0] == '2') ? 2 : 1; }();
	const int numLeaves = (int)((numTris + perLeaf - 1) / perLeaf), numInner = numLeaves - 1;
	int* dBounds = nullptr;
	BUILD_TRY(tmp.alloc(&dBounds, 6));
	{
		int init[6]; const float big = FLT_MAX;
		int hiInit, loInit; std::memcpy(&loInit, &big, 4); float nb = -FLT_MAX; std::memcpy(&hiInit, &nb, 4);
		hiInit = hiInit ^ 0x7fffffff;         // ordered form of -FLT_MAX
		for (int a = 0; a < 3; a++) { init[a] = loInit; init[3 + a] = hiInit; }
		BUILD_TRY(cudaMemcpyAsync(dBounds, init, sizeof(init), cudaMemcpyHostToDevice, st));
		BUILD_TRY(cudaStreamSynchronize(st));
	}
	k_lbvh_bounds<<<1184, 256, 0, st>>>(dTris, numTris, dBounds);
	uint64_t *codes = nullptr, *codesSorted = nullptr; uint32_t *ids = nullptr, *order = nullptr;
	BUILD_TRY(tmp.alloc(&codes, numTris)); BUILD_TRY(tmp.alloc(&codesSorted, numTris)); BUILD_TRY(tmp.alloc(&ids, numTris)); BUILD_TRY(tmp.alloc(&order, numTris));
	const unsigned blocksT = (unsigned)((numTris + 255) / 256), blocksL = (unsigned)((numLeaves + 255) / 256);
	k_lbvh_codes<<<blocksT, 256, 0, st>>>(dTris, numTris, dBounds, codes, ids);
	size_t sortBytes = 0;
	BUILD_TRY(cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, codes, codesSorted, ids, order, (int)numTris, 0, 63, st));
	uint8_t* sortTmp = nullptr;
	BUILD_TRY(tmp.alloc(&sortTmp, sortBytes));
	BUILD_TRY(cub::DeviceRadixSort::SortPairs(sortTmp, sortBytes, codes, codesSorted, ids, order, (int)numTris, 0, 63, st));
	tmp.free_now(codes); tmp.free_now(ids); tmp.free_now(sortTmp);
	// scene-owned outputs
	void *dRec = nullptr, *dNodes = nullptr;
	int rc;
	if ((rc = rto_scene_alloc(s, &dRec, numTris * 64))) return rc;
	if ((rc = rto_scene_alloc(s, &dNodes, (size_t)(numInner > 0 ? numInner : 1) * 64))) return rc;
	float* leafBox = nullptr; float* dRoot = nullptr;
	BUILD_TRY(tmp.alloc(&leafBox, 6 * (size_t)numLeaves)); BUILD_TRY(tmp.alloc(&dRoot, 6));
	// leaf boxes grown by 2^-18 of the scene extent: the kernels' fused node tests need conservative boxes (rto_kernels.cuh slab_oct)
	float grow = 0.0f;
	{
		int hb[6];
		BUILD_TRY(cudaMemcpyAsync(hb, dBounds, sizeof(hb), cudaMemcpyDeviceToHost, st));
		BUILD_TRY(cudaStreamSynchronize(st));
		for (int a = 0; a < 6; a++) { int i = hb[a] >= 0 ? hb[a] : hb[a] ^ 0x7fffffff; float f; std::memcpy(&f, &i, 4); grow = std::max(grow, std::fabs(f)); }
		int lg = 18; if (const char* e = getenv("RTO_BVH_GROW_LOG2")) { int v = atoi(e); if (v >= 8 && v <= 22) lg = v; }      // tuning aid
		grow = std::ldexp(grow, -lg);
	}
	D.grow = grow;
	k_lbvh_leaves<<<blocksL, 256, 0, st>>>(dTris, order, numTris, perLeaf, grow, (float4*)dRec, leafBox);
	float root[6];
	if (numInner == 0) {
		BUILD_TRY(cudaMemcpyAsync(root, leafBox, sizeof(root), cudaMemcpyDeviceToHost, st));
		BUILD_TRY(cudaStreamSynchronize(st));
		D.rootRef = ~(int)(numTris - 1);          // leafRef = (0 << 1) | (count - 1)
	}
	else {
		int *pInner = nullptr, *pLeaf = nullptr, *arrived = nullptr, *dDepth = nullptr;
		BUILD_TRY(tmp.alloc(&pInner, numInner)); BUILD_TRY(tmp.alloc(&pLeaf, numLeaves)); BUILD_TRY(tmp.alloc(&arrived, numInner)); BUILD_TRY(tmp.alloc(&dDepth, 1));
		BUILD_TRY(cudaMemsetAsync(arrived, 0, sizeof(int) * (size_t)numInner, st));
		BUILD_TRY(cudaMemsetAsync(dDepth, 0, sizeof(int), st));
		// the radix tree's top, surface-area splits below (rto_sahchunk.h); RTO_LBVH_SAH=0 keeps the plain radix tree (tuning aid)
		static const bool sahBottom = [] { const char* e = getenv("RTO_LBVH_SAH"); return !(e && e[0] == '0'); }();
		int2* rangeOf = nullptr; int* chunkList = nullptr; int* chunkCount = nullptr;
		if (sahBottom && perLeaf == 1) {
			BUILD_TRY(tmp.alloc(&rangeOf, numInner)); BUILD_TRY(tmp.alloc(&chunkList, numInner)); BUILD_TRY(tmp.alloc(&chunkCount, 1));
			BUILD_TRY(cudaMemsetAsync(chunkCount, 0, sizeof(int), st));
		}
		k_lbvh_tree<<<(unsigned)((numInner + 255) / 256), 256, 0, st>>>(codesSorted, perLeaf, numLeaves, numTris, (float4*)dNodes, pInner, pLeaf, rangeOf);
		if (rangeOf) {
			k_sah_select<<<(unsigned)((numInner + 255) / 256), 256, 0, st>>>(numInner, rangeOf, pInner, chunkList, chunkCount);
			int chunks = 0;
			BUILD_TRY(cudaMemcpyAsync(&chunks, chunkCount, sizeof(int), cudaMemcpyDeviceToHost, st));
			BUILD_TRY(cudaStreamSynchronize(st));
			const char* serial = getenv("RTO_SAH_SERIAL");
			if (chunks > 0 && serial && serial[0] == '1') k_sah_rebuild_serial<<<(unsigned)((chunks + 63) / 64), 64, 0, st>>>(chunks, chunkList, rangeOf, leafBox, (float4*)dNodes, pInner, pLeaf);
			else if (chunks > 0) k_sah_rebuild<<<(unsigned)chunks, kSahBlock, 0, st>>>(chunkList, rangeOf, leafBox, (float4*)dNodes, pInner, pLeaf);
			BUILD_TRY(cudaGetLastError());
		}
		k_lbvh_fit<<<blocksL, 256, 0, st>>>(numLeaves, leafBox, (float*)dNodes, pInner, pLeaf, arrived, dRoot, dDepth);
		BUILD_TRY(cudaGetLastError());
		int depth = 0;
		BUILD_TRY(cudaMemcpyAsync(root, dRoot, sizeof(root), cudaMemcpyDeviceToHost, st));
		BUILD_TRY(cudaMemcpyAsync(&depth, dDepth, sizeof(int), cudaMemcpyDeviceToHost, st));
		BUILD_TRY(cudaStreamSynchronize(st));
		// the kernels keep at most kBvhStack (96) postponed subtrees; a radix tree deeper than 92 levels needs > 2^29 triangles sharing one Morton
		// cell and is refused rather than traversed wrongly (rto_scene_create_bvh builds a balanced tree for such input)
		if (depth > 92) return rto_fail(RTO_ERR_UNSUPPORTED, "BVH build on the device: tree depth %d exceeds the traversal stack; use rto_scene_create_bvh", depth);
		D.rootRef = 0;
	}
	for (int a = 0; a < 3; a++) { D.rootLo[a] = root[a]; D.rootHi[a] = root[3 + a]; }
	D.nodes = (const float4*)dNodes; D.tris = (const float4*)dRec;
	D.exactNodes = D.nodes; D.exactRoot = D.rootRef; D.exactLeafBox = 0;
	s->bvh = D; s->bvhFast = D;
	s->numNodes = (size_t)numLeaves + (size_t)numInner;
	s->deviceBuiltBvh = true;
	return RTO_OK;
}

} // namespace

extern "C" int rto_scene_create_bvh_device(const RtoTriangle* tris, size_t numTris, RtoScene** out) try {
	RTO_RANGE("rto_scene_create_bvh_device");
	if (!out) return rto_fail(RTO_ERR_INVALID, "rto_scene_create_bvh_device: null output");
	*out = nullptr;
	if (numTris && !tris) return rto_fail(RTO_ERR_INVALID, "rto_scene_create_bvh_device: null triangles");
	RtoScene* s = nullptr;
	int rc = rto_scene_new(&s); if (rc) return rc;
	s->kind = RTO_MODE_BVH;
	DevPool pool;
	RtoTriangle* d = nullptr;
	cudaError_t e = pool.alloc(&d, numTris);
	if (e == cudaSuccess && numTris) e = cudaMemcpyAsync(d, tris, numTris * sizeof(RtoTriangle), cudaMemcpyHostToDevice, s->stream);
	if (e != cudaSuccess) { rto_scene_destroy(s); return rto_fail(RTO_ERR_CUDA, "rto_scene_create_bvh_device: %s", cudaGetErrorString(e)); }
	rc = lbvh_build(s, d, numTris);
	if (rc) { rto_scene_destroy(s); return rc; }
	*out = s;
	return RTO_OK;
} RTO_CATCH_ALL("rto_scene_create_bvh_device")

extern "C" int rto_scene_create_bvh_from_grid(const uint8_t* voxels, int dimX, int dimY, int dimZ, const float gridMin[3], float voxelSize,
	RtoScene** out) try {
	RTO_RANGE("rto_scene_create_bvh_from_grid");
	if (!out) return rto_fail(RTO_ERR_INVALID, "rto_scene_create_bvh_from_grid: null output");
	*out = nullptr;
	if (!voxels || !gridMin || dimX <= 0 || dimY <= 0 || dimZ <= 0) return rto_fail(RTO_ERR_INVALID, "rto_scene_create_bvh_from_grid: empty grid");
	RtoScene* s = nullptr;
	int rc = rto_scene_new(&s); if (rc) return rc;
	s->kind = RTO_MODE_BVH;
	{
		OctBuild B; McBuild M;
		rc = oct_pyramid(B, voxels, false, dimX, dimY, dimZ, s->stream);
		if (!rc) rc = oct_emit(B, true, false, s->stream);
		if (!rc) rc = mc_extract(B, gridMin, voxelSize, M, s->stream);
		if (!rc) rc = lbvh_build(s, M.tris, M.numTris);
		if (!rc) { cudaError_t e = cudaStreamSynchronize(s->stream); if (e != cudaSuccess) rc = rto_fail(RTO_ERR_CUDA, "scene build on the device failed: %s", cudaGetErrorString(e)); }
	}
	if (rc) { rto_scene_destroy(s); return rc; }
	*out = s;
	return RTO_OK;
} RTO_CATCH_ALL("rto_scene_create_bvh_from_grid")

// The pipeline configuration C4 names, entirely on the device: grid -> octree -> Adaptive Dual Contouring mesh (rto_dc.cu: the reference's
// triangles in the reference's order) -> linear BVH.  Hit ids index the soup rto_device_dc_mesh / rto_host_dc_mesh return for the
// same arguments.
extern "C" int rto_scene_create_bvh_from_grid_dc(const uint8_t* voxels, int dimX, int dimY, int dimZ, const float gridMin[3], float voxelSize,
	const float* viewProj16, float extraMargin, RtoScene** out) try {
	RTO_RANGE("rto_scene_create_bvh_from_grid_dc");
	if (!out) return rto_fail(RTO_ERR_INVALID, "rto_scene_create_bvh_from_grid_dc: null output");
	*out = nullptr;
	if (!voxels || !gridMin || dimX <= 0 || dimY <= 0 || dimZ <= 0) return rto_fail(RTO_ERR_INVALID, "rto_scene_create_bvh_from_grid_dc: empty grid");
	RtoScene* s = nullptr;
	int rc = rto_scene_new(&s); if (rc) return rc;
	s->kind = RTO_MODE_BVH;
	{
		OctBuild B;
		RtoTriangle* dTris = nullptr; size_t numTris = 0;
		rc = oct_pyramid(B, voxels, false, dimX, dimY, dimZ, s->stream);
		if (!rc && (1 << B.rootLevel) > 1024) rc = rto_fail(RTO_ERR_UNSUPPORTED, "rto_scene_create_bvh_from_grid_dc: grid larger than 1024 voxels per axis (the reference's cell keys alias there)");
		if (!rc) rc = oct_emit(B, false, true, s->stream);
		if (!rc) rc = rto_dc_extract_device(B.dVoxels, dimX, dimY, dimZ, gridMin, voxelSize, B.nodes, B.numNodes, viewProj16, extraMargin, s->stream, &dTris, &numTris, nullptr, nullptr);
		if (!rc) rc = lbvh_build(s, dTris, numTris);
		if (!rc) { cudaError_t e = cudaStreamSynchronize(s->stream); if (e != cudaSuccess) rc = rto_fail(RTO_ERR_CUDA, "scene build on the device failed: %s", cudaGetErrorString(e)); }
		else cudaStreamSynchronize(s->stream);
		if (dTris) cudaFree(dTris);
	}
	if (rc) { rto_scene_destroy(s); return rc; }
	*out = s;
	return RTO_OK;
} RTO_CATCH_ALL("rto_scene_create_bvh_from_grid_dc")


// ------------------------------------------------------------------------------------------------
// CSV voxeliser, fill on the device: the files are parsed on the host (rto_csv_load), one warp rasterises one
// face into the grid with the reference's arithmetic (rto_voxelize.h).  Every writer stores FILLED, so the order does not matter
// (the reference's OpenMP loop relies on the same fact, BuildingLoader.cpp:229, 279).
// ------------------------------------------------------------------------------------------------
namespace {
struct VoxGridDev { int dims[3]; float gmin[3]; float voxel; };

__global__ void k_voxelize(const RtoTriangle* __restrict__ tris, size_t numFaces, VoxGridDev G, uint8_t* __restrict__ vox) {
	const size_t face = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (face >= numFaces) return;
	const RtoTriangle t = tris[face];
	const VoxRange r = vox_face_range(t, G.gmin, G.voxel, G.dims);
	if (r.empty) return;
	const V3 a = mk3(t.v0[0], t.v0[1], t.v0[2]), b = mk3(t.v1[0], t.v1[1], t.v1[2]), c = mk3(t.v2[0], t.v2[1], t.v2[2]);
	const long long nx = r.x1 - r.x0 + 1, ny = r.y1 - r.y0 + 1, nz = r.z1 - r.z0 + 1, n = nx * ny * nz;
	for (long long i = lane; i < n; i += 32) {
		const int x = r.x0 + (int)(i % nx), y = r.y0 + (int)((i / nx) % ny), z = r.z0 + (int)(i / (nx * ny));
		if (vox_point_in_triangle(vox_center(G.gmin, G.voxel, x, y, z), a, b, c))
			vox[(size_t)x + (size_t)y * G.dims[0] + (size_t)z * ((size_t)G.dims[0] * G.dims[1])] = 1;
	}
}
} // namespace

extern "C" int rto_device_csv_voxelize(const char* vertsCsv, const char* facesCsv, float voxelSize, int dims[3], float minAndVoxel[4], uint8_t** voxelsOut) try {
	RTO_RANGE("rto_device_csv_voxelize");
	if (!dims || !minAndVoxel || !voxelsOut) return rto_fail(RTO_ERR_INVALID, "rto_device_csv_voxelize: null output");
	*voxelsOut = nullptr; dims[0] = dims[1] = dims[2] = 0;
	int rc = rto_require_device(); if (rc) return rc;
	CsvScene S;
	rc = rto_csv_load(vertsCsv, facesCsv, voxelSize, S); if (rc) return rc;
	for (int a = 0; a < 3; a++) { dims[a] = S.dims[a]; minAndVoxel[a] = S.gridMin[a]; }
	minAndVoxel[3] = S.voxelSize;
	const size_t n = (size_t)S.dims[0] * S.dims[1] * S.dims[2];
	if (n == 0) return RTO_OK;
	uint8_t* host = (uint8_t*)std::malloc(n);
	if (!host) return rto_fail(RTO_ERR_ALLOC, "rto_device_csv_voxelize: out of host memory");
	DevPool pool;
	cudaStream_t st = nullptr;
	uint8_t* dVox = nullptr; RtoTriangle* dTris = nullptr;
	cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = pool.alloc(&dVox, n);
	if (e == cudaSuccess) e = pool.alloc(&dTris, S.tris.size());
	if (e == cudaSuccess) e = cudaMemsetAsync(dVox, 0, n, st);
	if (e == cudaSuccess && !S.tris.empty()) e = cudaMemcpyAsync(dTris, S.tris.data(), S.tris.size() * sizeof(RtoTriangle), cudaMemcpyHostToDevice, st);
	if (e == cudaSuccess && !S.tris.empty()) {
		VoxGridDev G; for (int a = 0; a < 3; a++) { G.dims[a] = S.dims[a]; G.gmin[a] = S.gridMin[a]; } G.voxel = S.voxelSize;
		const size_t threads = S.tris.size() * 32;
		k_voxelize<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(dTris, S.tris.size(), G, dVox);
		e = cudaGetLastError();
	}
	if (e == cudaSuccess) e = cudaMemcpyAsync(host, dVox, n, cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	if (st) cudaStreamDestroy(st);
	if (e != cudaSuccess) { std::free(host); return rto_fail(RTO_ERR_CUDA, "rto_device_csv_voxelize: %s", cudaGetErrorString(e)); }
	*voxelsOut = host;
	return RTO_OK;
} RTO_CATCH_ALL("rto_device_csv_voxelize")


// ------------------------------------------------------------------------------------------------
// Frustum culling on the device: the per-node test, the compaction and the child remap of RayTracerBVH.cpp:724-813 (the reference
// runs them on one CPU thread every time the frustum is refreshed).  Same arrays as rto_host_frustum_cull.
// ------------------------------------------------------------------------------------------------
namespace {
struct CullArgs { FrustumPlanes F; float gmin[3]; float voxel; float margin; };

__global__ void k_cull_flags(const RtoGpuNode* __restrict__ nodes, size_t n, CullArgs A, int* __restrict__ flag) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	flag[i] = frustum_node_visible(A.F, nodes[i], A.gmin, A.voxel, A.margin) ? 1 : 0;
}
__global__ void k_cull_emit(const RtoGpuNode* __restrict__ nodes, size_t n, const int* __restrict__ flag, const int* __restrict__ ex,
	RtoGpuNode* __restrict__ out, int32_t* __restrict__ newToOld) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || !flag[i]) return;
	RtoGpuNode nd = nodes[i];
	if (!nd.isLeaf)
		for (int c = 0; c < 8; c++) {
			const int oc = nd.child[c];
			nd.child[c] = (oc >= 0 && (size_t)oc < n && flag[oc]) ? ex[oc] : -1;
		}
	out[ex[i]] = nd;
	newToOld[ex[i]] = (int32_t)i;
}
} // namespace

extern "C" int rto_device_frustum_cull(const RtoGpuNode* nodes, size_t numNodes, const float gridMin[3], float voxelSize, const float viewProj16[16],
	float margin, RtoGpuNode** culledOut, size_t* numCulled, int32_t** newToOldOut) try {
	RTO_RANGE("rto_device_frustum_cull");
	if (!culledOut || !numCulled) return rto_fail(RTO_ERR_INVALID, "rto_device_frustum_cull: null output");
	*culledOut = nullptr; *numCulled = 0;
	if (newToOldOut) *newToOldOut = nullptr;
	if (numNodes == 0) return RTO_OK;
	if (!nodes || !gridMin || !viewProj16) return rto_fail(RTO_ERR_INVALID, "rto_device_frustum_cull: null input");
	if (numNodes >= (size_t)0x7fffff00) return rto_fail(RTO_ERR_UNSUPPORTED, "rto_device_frustum_cull: too many nodes");
	int rc = rto_require_device(); if (rc) return rc;
	CullArgs A; A.F = frustum_from_view_proj(viewProj16);
	for (int k = 0; k < 3; k++) A.gmin[k] = gridMin[k];
	A.voxel = voxelSize; A.margin = margin;
	DevPool pool;
	cudaStream_t st = nullptr;
	BUILD_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
	RtoGpuNode *dIn = nullptr, *dOut = nullptr; int *flag = nullptr, *ex = nullptr; int32_t* dBack = nullptr; uint8_t* tmp = nullptr;
	RtoGpuNode* host = nullptr; int32_t* back = nullptr;
	auto fail = [&](int code) { cudaStreamSynchronize(st); cudaStreamDestroy(st); std::free(host); std::free(back); return code; };
#define CULL_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return fail(rto_fail(RTO_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_))); } while (0)
	CULL_TRY(pool.alloc(&dIn, numNodes)); CULL_TRY(pool.alloc(&flag, numNodes)); CULL_TRY(pool.alloc(&ex, numNodes));
	CULL_TRY(cudaMemcpyAsync(dIn, nodes, numNodes * sizeof(RtoGpuNode), cudaMemcpyHostToDevice, st));
	const unsigned blocks = (unsigned)((numNodes + 255) / 256);
	k_cull_flags<<<blocks, 256, 0, st>>>(dIn, numNodes, A, flag);
	size_t tmpBytes = 0;
	CULL_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, flag, ex, (int)numNodes, st));
	CULL_TRY(pool.alloc(&tmp, tmpBytes));
	CULL_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, flag, ex, (int)numNodes, st));
	int lastEx = 0, lastFlag = 0;
	CULL_TRY(cudaMemcpyAsync(&lastEx, ex + (numNodes - 1), 4, cudaMemcpyDeviceToHost, st));
	CULL_TRY(cudaMemcpyAsync(&lastFlag, flag + (numNodes - 1), 4, cudaMemcpyDeviceToHost, st));
	CULL_TRY(cudaStreamSynchronize(st));
	const size_t visible = (size_t)lastEx + (size_t)lastFlag;
	if (visible == 0) { cudaStreamDestroy(st); return RTO_OK; }
	CULL_TRY(pool.alloc(&dOut, visible)); CULL_TRY(pool.alloc(&dBack, visible));
	k_cull_emit<<<blocks, 256, 0, st>>>(dIn, numNodes, flag, ex, dOut, dBack);
	CULL_TRY(cudaGetLastError());
	host = (RtoGpuNode*)std::malloc(visible * sizeof(RtoGpuNode));
	back = (int32_t*)std::malloc(visible * sizeof(int32_t));
	if (!host || !back) return fail(rto_fail(RTO_ERR_ALLOC, "rto_device_frustum_cull: out of host memory"));
	CULL_TRY(cudaMemcpyAsync(host, dOut, visible * sizeof(RtoGpuNode), cudaMemcpyDeviceToHost, st));
	CULL_TRY(cudaMemcpyAsync(back, dBack, visible * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
	CULL_TRY(cudaStreamSynchronize(st));
#undef CULL_TRY
	cudaStreamDestroy(st);
	*culledOut = host; *numCulled = visible;
	if (newToOldOut) *newToOldOut = back; else std::free(back);
	return RTO_OK;
} RTO_CATCH_ALL("rto_device_frustum_cull")
