// rto_device.cu -- device scenes, kernel launches and the C ABI of librto.so (see include/rto_c.h).
//
// Built for sm_100a only with -fmad=false (exact parity with the CPU oracle needs unfused multiply/add).
// There is no CPU fallback in this file: every entry point that traces rays requires a CUDA device.
#include "rto_scene.cuh"
#include "rto_nvtx.h"
#include "rto_kernels.cuh"
#include "rto_sort.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <vector>

using namespace rto;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int rto_fail(int code, const char* fmt, ...) {
	va_list ap; va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return code;
}
extern "C" const char* rto_last_error(void) { return g_err; }
extern "C" const char* rto_version(void) { return "rto-b200 0.1 (sm_100a)"; }

static int require_device() { return rto_require_device(); }
int rto_require_device() {
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0) {
		cudaGetLastError();
		return rto_fail(RTO_ERR_NO_DEVICE, "no CUDA device available (%s); librto has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
	}
	return RTO_OK;
}

extern "C" int rto_init(int device) try {
	{	// constants whose exact bit patterns the traversal code relies on
		float a = kMissT, b = kBelowMissT, c = kMinPositive; uint32_t ua, ub, uc;
		std::memcpy(&ua, &a, 4); std::memcpy(&ub, &b, 4); std::memcpy(&uc, &c, 4);
		if (ua != 0x7149f2cau || ub != 0x7149f2c9u || uc != 1u) return rto_fail(RTO_ERR_UNSUPPORTED, "librto was built with a compiler that rounds 1e30f differently");
	}
	int rc = require_device(); if (rc) return rc;
	CUDA_TRY(cudaSetDevice(device));
	cudaDeviceProp p;
	CUDA_TRY(cudaGetDeviceProperties(&p, device));
	if (p.major != 10) return rto_fail(RTO_ERR_NO_DEVICE, "device %d is sm_%d%d; librto is built for sm_100a only", device, p.major, p.minor);
	return RTO_OK;
} RTO_CATCH_ALL("rto_init")

extern "C" int rto_device_info(int* smCount, int* ccMajor, int* ccMinor, size_t* l2Bytes, size_t* totalMem) try {
	int rc = require_device(); if (rc) return rc;
	int dev = 0; CUDA_TRY(cudaGetDevice(&dev));
	cudaDeviceProp p; CUDA_TRY(cudaGetDeviceProperties(&p, dev));
	if (smCount) *smCount = p.multiProcessorCount;
	if (ccMajor) *ccMajor = p.major;
	if (ccMinor) *ccMinor = p.minor;
	if (l2Bytes) *l2Bytes = (size_t)p.l2CacheSize;
	if (totalMem) *totalMem = p.totalGlobalMem;
	return RTO_OK;
} RTO_CATCH_ALL("rto_device_info")

// ------------------------------------------------------------------------------------------------
// scene (struct in rto_scene.cuh)
// ------------------------------------------------------------------------------------------------
int rto_scene_alloc(RtoScene* s, void** p, size_t bytes) {
	*p = nullptr;
	cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
	if (e != cudaSuccess) return rto_fail(RTO_ERR_ALLOC, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
	s->owned.push_back(*p); s->ownedBytes.push_back(bytes ? bytes : 16);
	s->deviceBytes += bytes;
	return RTO_OK;
}
int rto_scene_adopt(RtoScene* s, void* p, size_t bytes) {
	s->owned.push_back(p); s->ownedBytes.push_back(bytes);
	s->deviceBytes += bytes;
	return RTO_OK;
}

int rto_scene_scratch(RtoScene* s, int slot, size_t bytes, void** p) {
	if (s->scratchBytes[slot] < bytes) {
		if (s->scratch[slot]) cudaFree(s->scratch[slot]);
		s->scratch[slot] = nullptr; s->scratchBytes[slot] = 0;
		size_t want = bytes + bytes / 8 + 256;
		cudaError_t e = cudaMalloc(&s->scratch[slot], want);
		if (e != cudaSuccess) return rto_fail(RTO_ERR_ALLOC, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
		s->scratchBytes[slot] = want;
	}
	*p = s->scratch[slot];
	return RTO_OK;
}

int rto_scene_new(RtoScene** out) {
	int rc = require_device(); if (rc) return rc;
	RtoScene* s = new (std::nothrow) RtoScene();
	if (!s) return rto_fail(RTO_ERR_ALLOC, "out of host memory");
	cudaError_t e = cudaGetDevice(&s->device);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->copyStream, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->evFrame, cudaEventDisableTiming);
	if (e == cudaSuccess) e = cudaEventCreate(&s->evStart);
	if (e == cudaSuccess) e = cudaEventCreate(&s->evStop);
	if (e == cudaSuccess) e = cudaDeviceGetAttribute(&s->smCount, cudaDevAttrMultiProcessorCount, s->device);
	if (e != cudaSuccess) { delete s; return rto_fail(RTO_ERR_CUDA, "scene setup failed: %s", cudaGetErrorString(e)); }
	*out = s;
	return RTO_OK;
}
static int scene_alloc(RtoScene* s, void** p, size_t bytes) { return rto_scene_alloc(s, p, bytes); }
static int scene_scratch(RtoScene* s, int slot, size_t bytes, void** p) { return rto_scene_scratch(s, slot, bytes, p); }
static int scene_new(RtoScene** out) { return rto_scene_new(out); }

extern "C" void rto_scene_destroy(RtoScene* s) {
	if (!s) return;
	cudaSetDevice(s->device);
	if (s->stream) cudaStreamSynchronize(s->stream);
	for (void* p : s->owned) cudaFree(p);
	for (void* p : s->scratch) if (p) cudaFree(p);
	for (auto& c : s->camRing) {
		if (c.host) cudaFreeHost(c.host);
		if (c.dev) cudaFree(c.dev);
		if (c.uploaded) cudaEventDestroy(c.uploaded);
		for (cudaEvent_t e : c.readerDone) if (e) cudaEventDestroy(e);
	}
	if (s->evStart) cudaEventDestroy(s->evStart);
	if (s->evStop) cudaEventDestroy(s->evStop);
	if (s->evFrame) cudaEventDestroy(s->evFrame);
	if (s->evTable) cudaEventDestroy(s->evTable);
	if (s->copyStream) { cudaStreamSynchronize(s->copyStream); cudaStreamDestroy(s->copyStream); }
	if (s->stream) cudaStreamDestroy(s->stream);
	delete s;
}

extern "C" int rto_scene_info(const RtoScene* s, int* kind, size_t* numPrims, size_t* numNodes, size_t* deviceBytes, int* compactLayout) try {
	if (!s) return rto_fail(RTO_ERR_INVALID, "rto_scene_info: null scene");
	if (kind) *kind = s->kind;
	if (numPrims) *numPrims = s->numPrims;
	if (numNodes) *numNodes = s->numNodes;
	if (deviceBytes) *deviceBytes = s->deviceBytes;
	if (compactLayout) *compactLayout = (s->kind == RTO_MODE_BVH) ? 0 : s->oct.compact;
	return RTO_OK;
} RTO_CATCH_ALL("rto_scene_info")
extern "C" void* rto_scene_stream(const RtoScene* s) { return s ? (void*)s->stream : nullptr; }
extern "C" uint64_t rto_scene_launch_count(const RtoScene* s) { return s ? s->launches : 0; }
extern "C" int rto_scene_sync(RtoScene* s) try {
	if (!s) return rto_fail(RTO_ERR_INVALID, "rto_scene_sync: null scene");
	CUDA_TRY(cudaStreamSynchronize(s->stream));
	return RTO_OK;
} RTO_CATCH_ALL("rto_scene_sync")
extern "C" int rto_scene_last_kernel_ms(RtoScene* s, float* ms) try {
	if (!s || !ms) return rto_fail(RTO_ERR_INVALID, "rto_scene_last_kernel_ms: null argument");
	if (!s->timed) return rto_fail(RTO_ERR_INVALID, "rto_scene_last_kernel_ms: nothing rendered yet");
	CUDA_TRY(cudaEventSynchronize(s->evStop));
	CUDA_TRY(cudaEventElapsedTime(ms, s->evStart, s->evStop));
	return RTO_OK;
} RTO_CATCH_ALL("rto_scene_last_kernel_ms")

// ------------------------------------------------------------------------------------------------
// octree upload: RayTracerBVH::setOctree's SSBO (RayTracerBVH.cpp:492-504) -> pointer-free device arrays
// ------------------------------------------------------------------------------------------------
extern "C" int rto_scene_create_octree(const RtoGpuNode* nodes, size_t numNodes, const float gridMin[3], float voxelSize, RtoScene** out) try {
	RTO_RANGE("rto_scene_create_octree");
	if (!out) return rto_fail(RTO_ERR_INVALID, "rto_scene_create_octree: null output");
	*out = nullptr;
	if (!nodes || numNodes == 0 || !gridMin) return rto_fail(RTO_ERR_INVALID, "rto_scene_create_octree: empty octree (the reference's setOctree(nullptr) clears the scene; nothing to trace)");
	OctLayout L;
	int rc = rto_build_octree_layout(nodes, numNodes, L); if (rc) return rc;
	RtoScene* s = nullptr;
	rc = scene_new(&s); if (rc) return rc;
	s->kind = RTO_MODE_OCTREE_GLSL;
	s->numNodes = numNodes; s->numPrims = L.numLeaves;
	OctDev& D = s->oct;
	D.numNodes = (int)numNodes; D.rootSize = nodes[0].size;
	D.gmin[0] = gridMin[0]; D.gmin[1] = gridMin[1]; D.gmin[2] = gridMin[2]; D.voxel = voxelSize;
	D.compact = L.compact ? 1 : 0;
	s->octIsTree = L.isTree;
	auto up = [&](const void* src, size_t bytes, void** dst) -> int {
		int r = scene_alloc(s, dst, bytes); if (r) return r;
		cudaError_t e = cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, s->stream);
		return e == cudaSuccess ? RTO_OK : rto_fail(RTO_ERR_CUDA, "octree upload failed: %s", cudaGetErrorString(e));
	};
	void *dDesc = nullptr, *dUp = nullptr, *dInner = nullptr, *dN = nullptr;
	if (L.compact) {
		if ((rc = up(L.desc.data(), L.desc.size() * 4, &dDesc)) || (rc = up(L.up.data(), L.up.size() * 4, &dUp)) || (rc = up(L.inner.data(), L.inner.size() * 4, &dInner))) { rto_scene_destroy(s); return rc; }
		D.desc = (const uint32_t*)dDesc + 7; D.up = (const int32_t*)dUp; D.inner = (const int4*)dInner;
	}
	else {
		if ((rc = up(L.padded.data(), L.padded.size() * 4, &dN))) { rto_scene_destroy(s); return rc; }
		D.nodes16 = (const int4*)dN;
	}
	cudaError_t e = cudaStreamSynchronize(s->stream);
	if (e != cudaSuccess) { rto_scene_destroy(s); return rto_fail(RTO_ERR_CUDA, "octree upload failed: %s", cudaGetErrorString(e)); }
	*out = s;
	return RTO_OK;
} RTO_CATCH_ALL("rto_scene_create_octree")

// ------------------------------------------------------------------------------------------------
// BVH upload: reference-shaped host tree -> child-boxes-in-parent nodes + leaf-ordered triangles
// ------------------------------------------------------------------------------------------------
// the device scene of a host-built layout, on the current device
int rto_scene_from_bvh_layout(const BvhLayout& L, size_t numTris, size_t numRefNodes, RtoScene** out) {
	RtoScene* s = nullptr;
	int rc = scene_new(&s); if (rc) return rc;
	s->kind = RTO_MODE_BVH; s->numPrims = numTris; s->numNodes = numRefNodes;
	BvhDev& D = s->bvh;
	D.numTris = (int)numTris;
	for (int k = 0; k < 3; k++) { D.rootLo[k] = L.rootLo[k]; D.rootHi[k] = L.rootHi[k]; }
	D.rootRef = L.refRoot;
	void *dN = nullptr, *dT = nullptr, *dF = nullptr;
	if ((rc = scene_alloc(s, &dN, L.refNodes.size() * 4)) || (rc = scene_alloc(s, &dT, L.tris.size() * 4)) ||
		(!L.fastNodes.empty() && (rc = scene_alloc(s, &dF, L.fastNodes.size() * 4)))) { rto_scene_destroy(s); return rc; }
	cudaError_t e = cudaMemcpyAsync(dN, L.refNodes.data(), L.refNodes.size() * 4, cudaMemcpyHostToDevice, s->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(dT, L.tris.data(), L.tris.size() * 4, cudaMemcpyHostToDevice, s->stream);
	if (e == cudaSuccess && dF) e = cudaMemcpyAsync(dF, L.fastNodes.data(), L.fastNodes.size() * 4, cudaMemcpyHostToDevice, s->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
	if (e != cudaSuccess) { rto_scene_destroy(s); return rto_fail(RTO_ERR_CUDA, "BVH upload failed: %s", cudaGetErrorString(e)); }
	D.nodes = (const float4*)dN; D.tris = (const float4*)dT;
	D.leafBox = 0; D.grow = 0.0f; D.paired = 0; D.exactPaired = 0;
	D.exactNodes = D.nodes; D.exactRoot = D.rootRef; D.exactLeafBox = 0;
	s->bvh = D;
	s->bvhFast = D;
	if (dF) { s->bvhFast.nodes = (const float4*)dF; s->bvhFast.rootRef = L.fastRoot; s->bvhFast.leafBox = 1; s->bvhFast.grow = L.fastGrow; s->bvhFast.paired = 1; }
	if (dF && !L.wideNodes.empty() && L.wideRoot >= 0) {
		void* dW = nullptr;
		if ((rc = scene_alloc(s, &dW, L.wideNodes.size() * 4))) { rto_scene_destroy(s); return rc; }
		e = cudaMemcpyAsync(dW, L.wideNodes.data(), L.wideNodes.size() * 4, cudaMemcpyHostToDevice, s->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
		if (e != cudaSuccess) { rto_scene_destroy(s); return rto_fail(RTO_ERR_CUDA, "BVH upload failed: %s", cudaGetErrorString(e)); }
		s->bvhFast.wide = (const uint4*)dW; s->bvhFast.wideRoot = L.wideRoot; s->bvhFast.wideStep = L.wideStep;
		for (int k = 0; k < 3; k++) s->bvhFast.wideLo[k] = L.wideLo[k];
	}
	*out = s;
	return RTO_OK;
}

// reference-shaped host tree (built here unless given) -> the two device topologies and the triangle records, on the host
int rto_bvh_layout_from_tris(const RtoTriangle* tris, size_t numTris, const RtoHostBvh* prebuilt, BvhLayout& L, size_t* numRefNodes) {
	RtoHostBvh* ownedBvh = nullptr;
	const RtoHostBvh* h = prebuilt;
	if (!h) { int rc = rto_host_bvh_build(tris, numTris, &ownedBvh); if (rc) return rc; h = ownedBvh; }
	else if (h->numTris != numTris || h->tris != tris) return rto_fail(RTO_ERR_INVALID, "rto_scene_create_bvh: prebuilt BVH belongs to another triangle array");
	int rc = RTO_OK;
	try { rto_build_bvh_layout(*h, L); *numRefNodes = h->nodes.size(); }
	catch (...) { rc = rto_fail(RTO_ERR_ALLOC, "rto_scene_create_bvh: out of host memory while building the device layout"); }
	rto_host_bvh_free(ownedBvh);
	return rc;
}

extern "C" int rto_scene_create_bvh(const RtoTriangle* tris, size_t numTris, const RtoHostBvh* prebuilt, RtoScene** out) try {
	RTO_RANGE("rto_scene_create_bvh");
	if (!out) return rto_fail(RTO_ERR_INVALID, "rto_scene_create_bvh: null output");
	*out = nullptr;
	if (numTris && !tris) return rto_fail(RTO_ERR_INVALID, "rto_scene_create_bvh: null triangles");
	int rc = require_device(); if (rc) return rc;
	BvhLayout L; size_t numRefNodes = 0;
	if ((rc = rto_bvh_layout_from_tris(tris, numTris, prebuilt, L, &numRefNodes))) return rc;
	return rto_scene_from_bvh_layout(L, numTris, numRefNodes, out);
} RTO_CATCH_ALL("rto_scene_create_bvh")

// ------------------------------------------------------------------------------------------------
// rendering
// ------------------------------------------------------------------------------------------------
static int check_mode(const RtoScene* s, int mode) {
	if (s->kind == RTO_MODE_BVH) { if (mode != RTO_MODE_BVH) return rto_fail(RTO_ERR_INVALID, "scene is a BVH; mode must be RTO_MODE_BVH"); }
	else if (mode != RTO_MODE_OCTREE_SKIP && mode != RTO_MODE_OCTREE_GLSL) return rto_fail(RTO_ERR_INVALID, "scene is an octree; mode must be RTO_MODE_OCTREE_SKIP or RTO_MODE_OCTREE_GLSL");
	else if (mode == RTO_MODE_OCTREE_SKIP && !s->octIsTree)
		return rto_fail(RTO_ERR_INVALID, "RTO_MODE_OCTREE_SKIP: the node array's child graph is not a tree (a node is shared or lies on a cycle); octreeRaySkip is a recursion over a pointer tree");
	return RTO_OK;
}

// BVH scenes: per-thread while-loop kernel; default flags trace the SAH topology with t-pruning, RTO_FLAG_NO_PRUNE
// visits every box the reference's queryNode visits, on the reference's own topology (verification path).
// Octree scenes: thread-per-pixel kernel.  (Other schedulings were built, measured and removed in round 1 -- for the BVH loop a
// persistent per-lane refill kernel, a warp-phased while-while kernel and a per-step warp vote; for the octree walks a persistent
// kernel refilling idle lanes from a pixel counter; see profiles/README.md.)
static int launch_render(RtoScene* s, const RenderArgs& A, int width, int numCams, int mode, cudaStream_t st) {
	dim3 block(kRenderThreads), grid((width + kRenderBlockW - 1) / kRenderBlockW, (A.y1 - A.y0 + 7) / 8, numCams);
	if (s->kind == RTO_MODE_BVH) {
		bool sh = (A.flags & RTO_FLAG_SHADOWS) != 0, prune = (A.flags & RTO_FLAG_NO_PRUNE) == 0;
		const bool wide = s->bvhFast.wide != nullptr;
		if (sh && prune && wide) k_render_bvh<true, true, true><<<grid, block, 0, st>>>(s->bvhFast, A);
		else if (prune && wide) k_render_bvh<false, true, true><<<grid, block, 0, st>>>(s->bvhFast, A);
		else if (sh && prune) k_render_bvh<true, true><<<grid, block, 0, st>>>(s->bvhFast, A);
		else if (prune) k_render_bvh<false, true><<<grid, block, 0, st>>>(s->bvhFast, A);
		else if (sh) k_render_bvh<true, false><<<grid, block, 0, st>>>(s->bvh, A);
		else k_render_bvh<false, false><<<grid, block, 0, st>>>(s->bvh, A);
	}
	else if (mode == RTO_MODE_OCTREE_SKIP) k_render_octree<RTO_MODE_OCTREE_SKIP><<<grid, block, 0, st>>>(s->oct, A);
	else k_render_octree<RTO_MODE_OCTREE_GLSL><<<grid, block, 0, st>>>(s->oct, A);
	s->launches++;
	return RTO_OK;
}

// Camera array of a batched call: copied into a pinned slot of the scene's ring and from there to the device on `st`, so the call
// returns without waiting for anything.  A slot is reused only after the kernel that read it has finished (release_cameras).
static int stage_cameras(RtoScene* s, const RtoCamera* cams, int n, cudaStream_t st, int* slotOut, const RtoCamera** devOut) {
	// the same camera array again (an orbit traced batch after batch, the sub-ranges of a sharded batch): the copy that is already on the
	// device is read-only and still valid -- no copy, no wait
	for (int k = 0; k < RtoScene::kCamSlots; k++) {
		RtoScene::CamSlot& c = s->camRing[k];
		if (c.count == n && c.host && std::memcmp(c.host, cams, sizeof(RtoCamera) * n) == 0) {
			CUDA_TRY(cudaStreamWaitEvent(st, c.uploaded, 0));          // (another stream may have staged it)
			*slotOut = k; *devOut = c.dev;
			return RTO_OK;
		}
	}
	const int slot = s->camNext;
	s->camNext = (s->camNext + 1) % RtoScene::kCamSlots;
	RtoScene::CamSlot& c = s->camRing[slot];
	if (!c.uploaded) CUDA_TRY(cudaEventCreateWithFlags(&c.uploaded, cudaEventDisableTiming));
	for (int k = 0; k < 4; k++) if (c.readerDone[k]) CUDA_TRY(cudaEventSynchronize(c.readerDone[k]));      // every kernel that read this slot has finished
	c.count = 0;
	if (c.cap < n) {
		if (c.host) cudaFreeHost(c.host);
		if (c.dev) cudaFree(c.dev);
		c.host = nullptr; c.dev = nullptr; c.cap = 0;
		const int want = n < 64 ? 64 : n;
		CUDA_TRY(cudaHostAlloc((void**)&c.host, sizeof(RtoCamera) * want, cudaHostAllocDefault));
		CUDA_TRY(cudaMalloc((void**)&c.dev, sizeof(RtoCamera) * want));
		c.cap = want;
	}
	std::memcpy(c.host, cams, sizeof(RtoCamera) * n);
	CUDA_TRY(cudaMemcpyAsync(c.dev, c.host, sizeof(RtoCamera) * n, cudaMemcpyHostToDevice, st));
	CUDA_TRY(cudaEventRecord(c.uploaded, st));
	c.count = n;
	*slotOut = slot; *devOut = c.dev;
	return RTO_OK;
}
static int release_cameras(RtoScene* s, int slot, cudaStream_t st) {
	if (slot < 0) return RTO_OK;
	RtoScene::CamSlot& c = s->camRing[slot];
	int k = 0;
	while (k < 4 && c.readerDone[k] && c.readerStream[k] != st) k++;
	if (k == 4) { k = 0; CUDA_TRY(cudaEventSynchronize(c.readerDone[0])); }      // a fifth stream: take over the first entry once its reader is done
	if (!c.readerDone[k]) CUDA_TRY(cudaEventCreateWithFlags(&c.readerDone[k], cudaEventDisableTiming));
	c.readerStream[k] = st;
	CUDA_TRY(cudaEventRecord(c.readerDone[k], st));
	return RTO_OK;
}

extern "C" int rto_render_batch(RtoScene* s, const RtoCamera* cams, int numCams, int mode, uint32_t flags, float shadowBias,
	int y0, int y1, const RtoFrame* frame) try {
	RTO_RANGE("rto_render_batch");
	if (!s || !cams || !frame || numCams <= 0) return rto_fail(RTO_ERR_INVALID, "rto_render: null argument");
	int rc = check_mode(s, mode); if (rc) return rc;
	const int W = cams[0].width, H = cams[0].height;
	if (W <= 0 || H <= 0 || y0 < 0 || y1 > H || y0 >= y1) return rto_fail(RTO_ERR_INVALID, "rto_render: bad image size or row range [%d,%d) of %dx%d", y0, y1, W, H);
	for (int c = 1; c < numCams; c++) if (cams[c].width != W || cams[c].height != H) return rto_fail(RTO_ERR_INVALID, "rto_render_batch: all cameras must share one image size");
	if (numCams > 65535) return rto_fail(RTO_ERR_INVALID, "rto_render_batch: at most 65535 cameras per call");
	if (frame->memory != RTO_MEM_HOST && frame->memory != RTO_MEM_DEVICE) return rto_fail(RTO_ERR_INVALID, "rto_render: bad RtoFrame.memory");
	CUDA_TRY(cudaSetDevice(s->device));
	const size_t npix = (size_t)numCams * (size_t)(y1 - y0) * W;
	RenderArgs A{};
	A.cam0 = cams[0]; A.cams = nullptr; A.y0 = y0; A.y1 = y1; A.shadowBias = shadowBias; A.flags = flags;
	const bool host = frame->memory == RTO_MEM_HOST;
	int camSlot = -1;
	if (numCams > 1 && !host) { if ((rc = stage_cameras(s, cams, numCams, s->stream, &camSlot, &A.cams))) return rc; }
	if (host) {
		void* p = nullptr;
		if (frame->rgba) { if ((rc = scene_scratch(s, 0, npix * 16, &p))) return rc; A.rgba = (float4*)p; }
		if (frame->hitId) { if ((rc = scene_scratch(s, 1, npix * 4, &p))) return rc; A.hitId = (int*)p; }
		if (frame->t) { if ((rc = scene_scratch(s, 2, npix * 4, &p))) return rc; A.t = (float*)p; }
	}
	else { A.rgba = (float4*)frame->rgba; A.hitId = frame->hitId; A.t = frame->t; }
	if ((unsigned long long)npix >= 0xffffffffull) return rto_fail(RTO_ERR_UNSUPPORTED, "rto_render: more than 2^32 pixels in one call");
	CUDA_TRY(cudaEventRecord(s->evStart, s->stream));
	// a single frame into host memory is cut into row bands for the same reason (the call renderSceneCompute maps to): the copy of band b
	// runs while band b + 1 is traced, so the call takes one band's trace plus the copies instead of the whole trace plus the copies
	const int rows = y1 - y0;
	int bands = (host && numCams == 1 && rows >= 256) ? 4 : 1;
	if (bands > 1) if (const char* e = getenv("RTO_HOST_BANDS")) { const int v = atoi(e); if (v >= 1 && v <= 64) bands = v; }      // tuning aid
	if (host && (numCams > 1 || bands > 1)) {
		// one launch per frame (or band); the planes of unit i travel to the host on the copy stream while unit i + 1 is traced
		const size_t fpix = (size_t)rows * W;
		const int bandRows = ((rows + bands - 1) / bands + 7) & ~7;              // multiples of the kernels' 8-row tiles
		for (int c = 0; c < numCams; c++) {
			for (int r0 = 0; r0 < rows; r0 += bandRows) {
				const int r1 = r0 + bandRows < rows ? r0 + bandRows : rows;
				const size_t off = c * fpix + (size_t)r0 * W, bpix = (size_t)(r1 - r0) * W;
				RenderArgs Ac = A;
				Ac.cam0 = cams[c]; Ac.cams = nullptr; Ac.y0 = y0 + r0; Ac.y1 = y0 + r1;
				Ac.rgba = A.rgba ? A.rgba + off : nullptr; Ac.hitId = A.hitId ? A.hitId + off : nullptr; Ac.t = A.t ? A.t + off : nullptr;
				if ((rc = launch_render(s, Ac, W, 1, mode, s->stream))) return rc;
				CUDA_TRY(cudaGetLastError());
				CUDA_TRY(cudaEventRecord(s->evFrame, s->stream));
				CUDA_TRY(cudaStreamWaitEvent(s->copyStream, s->evFrame, 0));
				if (frame->rgba) CUDA_TRY(cudaMemcpyAsync(frame->rgba + 4 * off, Ac.rgba, bpix * 16, cudaMemcpyDeviceToHost, s->copyStream));
				if (frame->hitId) CUDA_TRY(cudaMemcpyAsync(frame->hitId + off, Ac.hitId, bpix * 4, cudaMemcpyDeviceToHost, s->copyStream));
				if (frame->t) CUDA_TRY(cudaMemcpyAsync(frame->t + off, Ac.t, bpix * 4, cudaMemcpyDeviceToHost, s->copyStream));
			}
		}
		CUDA_TRY(cudaEventRecord(s->evStop, s->stream));
		s->timed = true;
		CUDA_TRY(cudaStreamSynchronize(s->copyStream));
		CUDA_TRY(cudaStreamSynchronize(s->stream));
		return RTO_OK;
	}
	if ((rc = launch_render(s, A, W, numCams, mode, s->stream))) return rc;
	CUDA_TRY(cudaEventRecord(s->evStop, s->stream));
	if ((rc = release_cameras(s, camSlot, s->stream))) return rc;
	s->timed = true;
	CUDA_TRY(cudaGetLastError());
	if (host) {
		if (frame->rgba) CUDA_TRY(cudaMemcpyAsync(frame->rgba, A.rgba, npix * 16, cudaMemcpyDeviceToHost, s->stream));
		if (frame->hitId) CUDA_TRY(cudaMemcpyAsync(frame->hitId, A.hitId, npix * 4, cudaMemcpyDeviceToHost, s->stream));
		if (frame->t) CUDA_TRY(cudaMemcpyAsync(frame->t, A.t, npix * 4, cudaMemcpyDeviceToHost, s->stream));
		CUDA_TRY(cudaStreamSynchronize(s->stream));
	}
	return RTO_OK;
} RTO_CATCH_ALL("rto_render_batch")

// ------------------------------------------------------------------------------------------------
// hit codes: 4 bytes per pixel out of the trace kernel, planes rebuilt where they are wanted (rto_c.h "compact frames")
// ------------------------------------------------------------------------------------------------
static int check_rows(const char* who, const RtoCamera* cams, int numCams, int y0, int y1, bool tileAligned) {
	if (!cams || numCams <= 0) return rto_fail(RTO_ERR_INVALID, "%s: no cameras", who);
	const int W = cams[0].width, H = cams[0].height;
	if (W <= 0 || H <= 0 || y0 < 0 || y1 > H || y0 >= y1) return rto_fail(RTO_ERR_INVALID, "%s: bad image size or row range [%d,%d) of %dx%d", who, y0, y1, W, H);
	for (int c = 1; c < numCams; c++) if (cams[c].width != W || cams[c].height != H) return rto_fail(RTO_ERR_INVALID, "%s: all cameras must share one image size", who);
	if (numCams > 65535) return rto_fail(RTO_ERR_INVALID, "%s: at most 65535 cameras per call", who);
	if ((unsigned long long)numCams * (unsigned long long)(y1 - y0) * W >= 0xffffffffull) return rto_fail(RTO_ERR_UNSUPPORTED, "%s: more than 2^32 pixels in one call", who);
	if (tileAligned && ((y0 & 7) || ((y1 & 7) && y1 != H))) return rto_fail(RTO_ERR_INVALID, "%s: with hit codes y0 must be a multiple of 8 and y1 a multiple of 8 or the image height", who);
	return RTO_OK;
}

int rto_enqueue_render(RtoScene* s, const RtoCamera* cams, int numCams, int mode, uint32_t flags, float shadowBias, int y0, int y1,
	float4* rgba, int32_t* hitId, float* t, uint32_t* codes, size_t codeFrame0, cudaStream_t st) {
	RenderArgs A{};
	A.cam0 = cams[0]; A.cams = nullptr; A.y0 = y0; A.y1 = y1; A.shadowBias = shadowBias; A.flags = flags;
	A.rgba = rgba; A.hitId = hitId; A.t = t;
	A.codes = codes; A.codeFrame0 = (unsigned)codeFrame0; A.codeTilesY = (unsigned)((cams[0].height + 7) / 8);
	int slot = -1, rc;
	if (numCams > 1 && (rc = stage_cameras(s, cams, numCams, st, &slot, &A.cams))) return rc;
	if ((rc = launch_render(s, A, cams[0].width, numCams, mode, st))) return rc;
	CUDA_TRY(cudaGetLastError());
	return release_cameras(s, slot, st);
}

int rto_enqueue_resolve(RtoScene* s, const RtoCamera* cams, int numCams, int y0, int y1, const uint32_t* codes, size_t codeFrame0,
	float4* rgba, int32_t* hitId, float* t, cudaStream_t st) {
	RenderArgs A{};
	A.cam0 = cams[0]; A.cams = nullptr; A.y0 = y0; A.y1 = y1;
	A.rgba = rgba; A.hitId = hitId; A.t = t;
	A.codes = const_cast<uint32_t*>(codes); A.codeFrame0 = (unsigned)codeFrame0; A.codeTilesY = (unsigned)((cams[0].height + 7) / 8);
	int slot = -1, rc;
	if (numCams > 1 && (rc = stage_cameras(s, cams, numCams, st, &slot, &A.cams))) return rc;
	if (!s->shadeTable) {
		// once per scene, on the stream this first expansion runs on (later ones on other streams come after it in host order and
		// are ordered behind it by the event below)
		void* p = nullptr;
		if ((rc = scene_alloc(s, &p, (s->numPrims ? s->numPrims : 1) * sizeof(float)))) return rc;
		if (s->numPrims) k_shade_table<<<(unsigned)((s->numPrims + 255) / 256), 256, 0, st>>>(s->bvhFast.tris, (int)s->numPrims, (float*)p);
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaEventCreateWithFlags(&s->evTable, cudaEventDisableTiming));
		CUDA_TRY(cudaEventRecord(s->evTable, st));
		s->shadeTable = (float*)p;
		s->launches++;
	}
	CUDA_TRY(cudaStreamWaitEvent(st, s->evTable, 0));
	dim3 block(kRenderThreads), grid((cams[0].width + kRenderBlockW - 1) / kRenderBlockW, (y1 - y0 + 7) / 8, numCams);
	k_resolve_bvh<<<grid, block, 0, st>>>(s->bvhFast, A, s->shadeTable);
	s->launches++;
	CUDA_TRY(cudaGetLastError());
	return release_cameras(s, slot, st);
}

extern "C" size_t rto_codes_frame_words(int width, int height) {
	if (width <= 0 || height <= 0) return 0;
	return (size_t)((width + kRenderBlockW - 1) / kRenderBlockW) * (size_t)((height + 7) / 8) * kRenderThreads;
}

extern "C" int rto_render_codes(RtoScene* s, const RtoCamera* cams, int numCams, uint32_t flags, float shadowBias, int y0, int y1,
	uint32_t* codes, size_t firstFrame, void* stream) try {
	RTO_RANGE("rto_render_codes");
	if (!s || !codes) return rto_fail(RTO_ERR_INVALID, "rto_render_codes: null argument");
	if (s->kind != RTO_MODE_BVH) return rto_fail(RTO_ERR_UNSUPPORTED, "rto_render_codes: hit codes exist for BVH scenes only");
	int rc = check_rows("rto_render_codes", cams, numCams, y0, y1, true); if (rc) return rc;
	if (firstFrame + (size_t)numCams > 0xffffffffull) return rto_fail(RTO_ERR_INVALID, "rto_render_codes: frame index out of range");
	CUDA_TRY(cudaSetDevice(s->device));
	cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
	if (!stream) CUDA_TRY(cudaEventRecord(s->evStart, st));
	if ((rc = rto_enqueue_render(s, cams, numCams, RTO_MODE_BVH, flags, shadowBias, y0, y1, nullptr, nullptr, nullptr, codes, firstFrame, st))) return rc;
	if (!stream) { CUDA_TRY(cudaEventRecord(s->evStop, st)); s->timed = true; }
	return RTO_OK;
} RTO_CATCH_ALL("rto_render_codes")

extern "C" int rto_resolve_codes(RtoScene* s, const RtoCamera* cams, int numCams, int y0, int y1, const uint32_t* codes, size_t firstFrame,
	const RtoFrame* frame, void* stream) try {
	RTO_RANGE("rto_resolve_codes");
	if (!s || !codes || !frame) return rto_fail(RTO_ERR_INVALID, "rto_resolve_codes: null argument");
	if (s->kind != RTO_MODE_BVH) return rto_fail(RTO_ERR_UNSUPPORTED, "rto_resolve_codes: hit codes exist for BVH scenes only");
	int rc = check_rows("rto_resolve_codes", cams, numCams, y0, y1, true); if (rc) return rc;
	if (frame->memory != RTO_MEM_HOST && frame->memory != RTO_MEM_DEVICE) return rto_fail(RTO_ERR_INVALID, "rto_resolve_codes: bad RtoFrame.memory");
	CUDA_TRY(cudaSetDevice(s->device));
	cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
	const bool host = frame->memory == RTO_MEM_HOST;
	const size_t npix = (size_t)numCams * (size_t)(y1 - y0) * cams[0].width;
	float4* rgba = (float4*)frame->rgba; int32_t* hid = frame->hitId; float* t = frame->t;
	if (host) {
		void* p = nullptr;
		if (frame->rgba) { if ((rc = scene_scratch(s, 0, npix * 16, &p))) return rc; rgba = (float4*)p; }
		if (frame->hitId) { if ((rc = scene_scratch(s, 1, npix * 4, &p))) return rc; hid = (int32_t*)p; }
		if (frame->t) { if ((rc = scene_scratch(s, 2, npix * 4, &p))) return rc; t = (float*)p; }
	}
	if (!stream) CUDA_TRY(cudaEventRecord(s->evStart, st));
	if ((rc = rto_enqueue_resolve(s, cams, numCams, y0, y1, codes, firstFrame, rgba, hid, t, st))) return rc;
	if (!stream) { CUDA_TRY(cudaEventRecord(s->evStop, st)); s->timed = true; }
	if (host) {
		if (frame->rgba) CUDA_TRY(cudaMemcpyAsync(frame->rgba, rgba, npix * 16, cudaMemcpyDeviceToHost, st));
		if (frame->hitId) CUDA_TRY(cudaMemcpyAsync(frame->hitId, hid, npix * 4, cudaMemcpyDeviceToHost, st));
		if (frame->t) CUDA_TRY(cudaMemcpyAsync(frame->t, t, npix * 4, cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
	}
	return RTO_OK;
} RTO_CATCH_ALL("rto_resolve_codes")

// ------------------------------------------------------------------------------------------------
// exchange memory: device buffers another GPU of the box writes hit codes into (same process: peer access; other process: CUDA IPC)
// ------------------------------------------------------------------------------------------------
static_assert(sizeof(RtoIpcHandle) == sizeof(cudaIpcMemHandle_t), "RtoIpcHandle must hold a cudaIpcMemHandle_t");

extern "C" int rto_exchange_alloc(size_t bytes, void** devPtr, RtoIpcHandle* handleOut) try {
	if (!devPtr) return rto_fail(RTO_ERR_INVALID, "rto_exchange_alloc: null output");
	*devPtr = nullptr;
	int rc = require_device(); if (rc) return rc;
	void* p = nullptr;
	cudaError_t e = cudaMalloc(&p, bytes ? bytes : 256);
	if (e != cudaSuccess) return rto_fail(RTO_ERR_ALLOC, "rto_exchange_alloc: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
	e = cudaMemset(p, 0, bytes ? bytes : 256);                       // signal words start at 0; hit codes at "miss"
	if (e != cudaSuccess) { cudaFree(p); return rto_fail(RTO_ERR_CUDA, "rto_exchange_alloc: cudaMemset failed: %s", cudaGetErrorString(e)); }
	if (handleOut) {
		cudaIpcMemHandle_t h;
		e = cudaIpcGetMemHandle(&h, p);
		if (e != cudaSuccess) { cudaFree(p); cudaGetLastError(); return rto_fail(RTO_ERR_CUDA, "rto_exchange_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); }
		std::memcpy(handleOut, &h, sizeof(h));
	}
	*devPtr = p;
	return RTO_OK;
} RTO_CATCH_ALL("rto_exchange_alloc")
extern "C" int rto_exchange_free(void* devPtr) try {
	if (devPtr) CUDA_TRY(cudaFree(devPtr));
	return RTO_OK;
} RTO_CATCH_ALL("rto_exchange_free")
extern "C" int rto_exchange_open(const RtoIpcHandle* handle, void** devPtr) try {
	if (!handle || !devPtr) return rto_fail(RTO_ERR_INVALID, "rto_exchange_open: null argument");
	*devPtr = nullptr;
	int rc = require_device(); if (rc) return rc;
	cudaIpcMemHandle_t h;
	std::memcpy(&h, handle, sizeof(h));
	cudaError_t e = cudaIpcOpenMemHandle(devPtr, h, cudaIpcMemLazyEnablePeerAccess);
	if (e != cudaSuccess) { cudaGetLastError(); *devPtr = nullptr; return rto_fail(RTO_ERR_CUDA, "rto_exchange_open: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e)); }
	return RTO_OK;
} RTO_CATCH_ALL("rto_exchange_open")
extern "C" int rto_exchange_close(void* devPtr) try {
	if (devPtr) CUDA_TRY(cudaIpcCloseMemHandle(devPtr));
	return RTO_OK;
} RTO_CATCH_ALL("rto_exchange_close")

extern "C" int rto_render(RtoScene* s, const RtoCamera* cam, int mode, uint32_t flags, float shadowBias, int y0, int y1, const RtoFrame* frame) try {
	return rto_render_batch(s, cam, 1, mode, flags, shadowBias, y0, y1, frame);
} RTO_CATCH_ALL("rto_render")

extern "C" int rto_render_stats(RtoScene* s, const RtoCamera* cam, int mode, uint32_t flags, float shadowBias, int y0, int y1, uint64_t stats[5]) try {
	RTO_RANGE("rto_render_stats");
	if (!s || !cam || !stats) return rto_fail(RTO_ERR_INVALID, "rto_render_stats: null argument");
	int rc = check_mode(s, mode); if (rc) return rc;
	if (s->deviceBuiltBvh) return rto_fail(RTO_ERR_UNSUPPORTED, "rto_render_stats: the scene's BVH was built on the device; the reference's work counters need the reference-shaped tree (rto_scene_create_bvh)");
	if (cam->width <= 0 || y0 < 0 || y1 > cam->height || y0 >= y1) return rto_fail(RTO_ERR_INVALID, "rto_render_stats: bad row range");
	CUDA_TRY(cudaSetDevice(s->device));
	void* d = nullptr;
	if ((rc = scene_scratch(s, 4, 5 * 8, &d))) return rc;
	CUDA_TRY(cudaMemsetAsync(d, 0, 5 * 8, s->stream));
	RenderArgs A{};
	A.cam0 = *cam; A.y0 = y0; A.y1 = y1; A.shadowBias = shadowBias; A.flags = flags;
	dim3 block(kRenderThreads), grid((cam->width + kRenderBlockW - 1) / kRenderBlockW, (y1 - y0 + 7) / 8, 1);
	if (s->kind == RTO_MODE_BVH) k_stats_bvh<<<grid, block, 0, s->stream>>>(s->bvh, A, (unsigned long long*)d);
	else k_stats_octree<<<grid, block, 0, s->stream>>>(s->oct, A, mode, (unsigned long long*)d);
	s->launches++;
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaMemcpyAsync(stats, d, 5 * 8, cudaMemcpyDeviceToHost, s->stream));
	CUDA_TRY(cudaStreamSynchronize(s->stream));
	return RTO_OK;
} RTO_CATCH_ALL("rto_render_stats")

// ------------------------------------------------------------------------------------------------
// explicit ray lists
// ------------------------------------------------------------------------------------------------
extern "C" int rto_trace_rays(RtoScene* s, int mode, uint32_t flags, const float* origins, const float* dirs, size_t numRays,
	float tMin, float tMax, float* tOut, int32_t* idOut, int memory) try {
	RTO_RANGE("rto_trace_rays");
	if (!s || !origins || !dirs) return rto_fail(RTO_ERR_INVALID, "rto_trace_rays: null argument");
	int rc = check_mode(s, mode); if (rc) return rc;
	if (memory != RTO_MEM_HOST && memory != RTO_MEM_DEVICE) return rto_fail(RTO_ERR_INVALID, "rto_trace_rays: bad memory selector %d", memory);
	// the closest-hit rule of BVH scenes has no interval (accept t > 1e-4, strict minimum); octree scenes take the caller's
	if (s->kind == RTO_MODE_BVH && !(tMin == 0.0f && tMax >= 1e30f))
		return rto_fail(RTO_ERR_INVALID, "rto_trace_rays: BVH scenes take tMin = 0 and tMax >= 1e30 only (the hit rule has no interval)");
	if (numRays == 0) return RTO_OK;
	CUDA_TRY(cudaSetDevice(s->device));
	const bool host = memory == RTO_MEM_HOST;
	const float *dO = origins, *dD = dirs; float* dT = tOut; int32_t* dI = idOut;
	if (host) {
		void* p = nullptr;
		if ((rc = scene_scratch(s, 0, numRays * 24, &p))) return rc;
		dO = (const float*)p; dD = dO + 3 * numRays;
		CUDA_TRY(cudaMemcpyAsync(p, origins, numRays * 12, cudaMemcpyHostToDevice, s->stream));
		CUDA_TRY(cudaMemcpyAsync((float*)p + 3 * numRays, dirs, numRays * 12, cudaMemcpyHostToDevice, s->stream));
		if (tOut) { if ((rc = scene_scratch(s, 1, numRays * 4, &p))) return rc; dT = (float*)p; }
		if (idOut) { if ((rc = scene_scratch(s, 2, numRays * 4, &p))) return rc; dI = (int32_t*)p; }
	}
	unsigned blocks = (unsigned)((numRays + 127) / 128);
	const uint32_t* perm = nullptr;
	CUDA_TRY(cudaEventRecord(s->evStart, s->stream));          // rto_scene_last_kernel_ms: key generation + sort + trace
	if ((flags & RTO_FLAG_SORT_RAYS) && numRays > 1) {
		// ray coherence sorting: trace the rays in the order of their sort keys (rto_kernels.cuh k_ray_sort_keys); results land at each ray's own index
		if (numRays >= ((size_t)1 << 31)) return rto_fail(RTO_ERR_UNSUPPORTED, "rto_trace_rays: RTO_FLAG_SORT_RAYS takes fewer than 2^31 rays");
		float lo[3], hi[3];
		if (s->kind == RTO_MODE_BVH) for (int a = 0; a < 3; a++) { lo[a] = s->bvhFast.rootLo[a]; hi[a] = s->bvhFast.rootHi[a]; }
		else for (int a = 0; a < 3; a++) { lo[a] = s->oct.gmin[a]; hi[a] = s->oct.gmin[a] + (float)s->oct.rootSize * s->oct.voxel; }
		void *k0 = nullptr, *k1 = nullptr, *i0 = nullptr, *i1 = nullptr, *tmp = nullptr;
		if ((rc = scene_scratch(s, 4, numRays * 8, &k0))) return rc;
		if ((rc = scene_scratch(s, 5, numRays * 8, &i0))) return rc;
		k1 = (uint32_t*)k0 + numRays; i1 = (uint32_t*)i0 + numRays;
		if ((rc = scene_scratch(s, 6, sort_scratch_bytes(numRays), &tmp))) return rc;
		k_ray_sort_keys<<<(unsigned)((numRays + 255) / 256), 256, 0, s->stream>>>(dO, dD, numRays, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], (uint32_t*)k0, (uint32_t*)i0);
		s->launches++;
		CUDA_TRY(sort_pairs_u32((uint32_t*)k0, (uint32_t*)i0, (uint32_t*)k1, (uint32_t*)i1, numRays, tmp, s->stream, &s->launches));      // rto_sort.cuh
		perm = (const uint32_t*)i0;
	}
	if (s->kind == RTO_MODE_BVH) k_trace_bvh<<<blocks, 128, 0, s->stream>>>((flags & RTO_FLAG_NO_PRUNE) ? s->bvh : s->bvhFast, flags, dO, dD, numRays, dT, dI, perm);
	else k_trace_octree<<<blocks, 128, 0, s->stream>>>(s->oct, mode, dO, dD, numRays, tMin, tMax, dT, dI, perm);
	CUDA_TRY(cudaEventRecord(s->evStop, s->stream));
	s->timed = true;
	s->launches++;
	CUDA_TRY(cudaGetLastError());
	if (host) {
		if (tOut) CUDA_TRY(cudaMemcpyAsync(tOut, dT, numRays * 4, cudaMemcpyDeviceToHost, s->stream));
		if (idOut) CUDA_TRY(cudaMemcpyAsync(idOut, dI, numRays * 4, cudaMemcpyDeviceToHost, s->stream));
		CUDA_TRY(cudaStreamSynchronize(s->stream));
	}
	return RTO_OK;
} RTO_CATCH_ALL("rto_trace_rays")

// Page-locked host memory for frame planes: a device -> host copy into ordinary (pageable) memory is staged by the driver and runs at a
// fraction of the link's speed; the reference never reads its frame back (a GL texture), so this has no counterpart there.
extern "C" int rto_host_alloc_pinned(size_t bytes, void** out) try {
	if (!out) return rto_fail(RTO_ERR_INVALID, "rto_host_alloc_pinned: null output");
	*out = nullptr;
	int rc = rto_require_device(); if (rc) return rc;
	if (bytes == 0) bytes = 1;
	cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
	if (e != cudaSuccess) { *out = nullptr; return rto_fail(e == cudaErrorMemoryAllocation ? RTO_ERR_ALLOC : RTO_ERR_CUDA, "rto_host_alloc_pinned(%zu bytes): %s", bytes, cudaGetErrorString(e)); }
	return RTO_OK;
} RTO_CATCH_ALL("rto_host_alloc_pinned")
extern "C" void rto_host_free_pinned(void* p) { if (p) cudaFreeHost(p); }

// The device sort behind RTO_FLAG_SORT_RAYS on its own (rto_sort.cuh): n key / value pairs in host memory, sorted by key in place, stable.
extern "C" int rto_device_sort_pairs(uint32_t* keys, uint32_t* vals, size_t n) try {
	RTO_RANGE("rto_device_sort_pairs");
	if (n && (!keys || !vals)) return rto_fail(RTO_ERR_INVALID, "rto_device_sort_pairs: null argument");
	if (n >= ((size_t)1 << 31)) return rto_fail(RTO_ERR_UNSUPPORTED, "rto_device_sort_pairs: fewer than 2^31 pairs");
	int rc = rto_require_device(); if (rc) return rc;
	if (n < 2) return RTO_OK;
	uint32_t* d = nullptr; void* tmp = nullptr;
	cudaError_t e = cudaMalloc((void**)&d, 4 * n * sizeof(uint32_t));
	if (e == cudaSuccess) e = cudaMalloc(&tmp, sort_scratch_bytes(n));
	if (e == cudaSuccess) e = cudaMemcpy(d, keys, n * 4, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMemcpy(d + n, vals, n * 4, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = sort_pairs_u32(d, d + n, d + 2 * n, d + 3 * n, n, tmp, 0);
	if (e == cudaSuccess) e = cudaMemcpy(keys, d, n * 4, cudaMemcpyDeviceToHost);
	if (e == cudaSuccess) e = cudaMemcpy(vals, d + n, n * 4, cudaMemcpyDeviceToHost);
	cudaFree(d); cudaFree(tmp);
	if (e != cudaSuccess) return rto_fail(RTO_ERR_CUDA, "rto_device_sort_pairs: %s", cudaGetErrorString(e));
	return RTO_OK;
} RTO_CATCH_ALL("rto_device_sort_pairs")

extern "C" int rto_bvh_query(RtoScene* s, const float* origins, const float* dirs, size_t numRays,
	int64_t* offsets, int32_t* ids, size_t idsCapacity, size_t* totalOut) try {
	RTO_RANGE("rto_bvh_query");
	if (!s || !origins || !dirs || !offsets) return rto_fail(RTO_ERR_INVALID, "rto_bvh_query: null argument");
	if (s->kind != RTO_MODE_BVH) return rto_fail(RTO_ERR_INVALID, "rto_bvh_query: scene is not a BVH");
	if (s->deviceBuiltBvh) return rto_fail(RTO_ERR_UNSUPPORTED, "rto_bvh_query: the scene's BVH was built on the device; BVH::query's candidate order needs the reference-shaped tree (rto_scene_create_bvh)");
	offsets[0] = 0;
	if (totalOut) *totalOut = 0;
	if (numRays == 0) return RTO_OK;
	CUDA_TRY(cudaSetDevice(s->device));
	int rc; void* p = nullptr;
	if ((rc = scene_scratch(s, 0, numRays * 24, &p))) return rc;
	float* dO = (float*)p; float* dD = dO + 3 * numRays;
	CUDA_TRY(cudaMemcpyAsync(dO, origins, numRays * 12, cudaMemcpyHostToDevice, s->stream));
	CUDA_TRY(cudaMemcpyAsync(dD, dirs, numRays * 12, cudaMemcpyHostToDevice, s->stream));
	if ((rc = scene_scratch(s, 1, numRays * 4, &p))) return rc;
	int* dCounts = (int*)p;
	unsigned blocks = (unsigned)((numRays + 127) / 128);
	k_bvh_query<<<blocks, 128, 0, s->stream>>>(s->bvh, dO, dD, numRays, nullptr, dCounts, nullptr);
	s->launches++;
	CUDA_TRY(cudaGetLastError());
	std::vector<int> counts(numRays);
	CUDA_TRY(cudaMemcpyAsync(counts.data(), dCounts, numRays * 4, cudaMemcpyDeviceToHost, s->stream));
	CUDA_TRY(cudaStreamSynchronize(s->stream));
	int64_t total = 0;
	for (size_t i = 0; i < numRays; i++) { offsets[i] = total; total += counts[i]; }
	offsets[numRays] = total;
	if (totalOut) *totalOut = (size_t)total;
	if (!ids || total == 0) return RTO_OK;
	if ((size_t)total > idsCapacity) return rto_fail(RTO_ERR_INVALID, "rto_bvh_query: ids capacity %zu < %lld candidates", idsCapacity, (long long)total);
	if ((rc = scene_scratch(s, 2, (numRays + 1) * 8, &p))) return rc;
	long long* dOff = (long long*)p;
	CUDA_TRY(cudaMemcpyAsync(dOff, offsets, (numRays + 1) * 8, cudaMemcpyHostToDevice, s->stream));
	if ((rc = scene_scratch(s, 5, (size_t)total * 4, &p))) return rc;
	int* dIds = (int*)p;
	k_bvh_query<<<blocks, 128, 0, s->stream>>>(s->bvh, dO, dD, numRays, dOff, nullptr, dIds);
	s->launches++;
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaMemcpyAsync(ids, dIds, (size_t)total * 4, cudaMemcpyDeviceToHost, s->stream));
	CUDA_TRY(cudaStreamSynchronize(s->stream));
	return RTO_OK;
} RTO_CATCH_ALL("rto_bvh_query")

// ------------------------------------------------------------------------------------------------
// VolumeRaycastRenderer's per-frame use of octreeRaySkip (VolumeRaycastRenderer.cpp:1598-1664): 49 probe rays -> one start distance
// ------------------------------------------------------------------------------------------------
extern "C" int rto_octree_skip_distance(RtoScene* s, const float view16[16], const float camPos[3], float aspect, float lastSkipDistance,
	float* skipDistanceOut, float* probeT /* 49 floats, may be NULL */) try {
	RTO_RANGE("rto_octree_skip_distance");
	if (!s || !skipDistanceOut) return rto_fail(RTO_ERR_INVALID, "rto_octree_skip_distance: null argument");
	if (s->kind == RTO_MODE_BVH) return rto_fail(RTO_ERR_INVALID, "rto_octree_skip_distance: scene is not an octree");
	float o[49 * 3], d[49 * 3], t[49];
	int rc = rto_host_skip_probe_rays(view16, camPos, aspect, o, d); if (rc) return rc;
	rc = rto_trace_rays(s, RTO_MODE_OCTREE_SKIP, 0, o, d, 49, 0.0f, 1e30f, t, nullptr, RTO_MEM_HOST); if (rc) return rc;
	if (probeT) std::memcpy(probeT, t, sizeof(t));
	*skipDistanceOut = rto_host_skip_distance_from_probes(t, 49, lastSkipDistance);
	return RTO_OK;
} RTO_CATCH_ALL("rto_octree_skip_distance")
