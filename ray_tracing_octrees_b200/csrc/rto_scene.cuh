// rto_scene.cuh -- the device scene object and its allocation helpers, shared by the translation units that enqueue GPU work
// (rto_device.cu: uploads, rendering; rto_build.cu: scene construction on the GPU).  Not part of the ABI.
#pragma once
#include "rto_devtypes.h"
#include <vector>

#define CUDA_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return rto_fail(RTO_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); } while (0)

// RTO_OK, or RTO_ERR_NO_DEVICE with a message: there is no CPU fallback for anything that traces or builds on the device
int rto_require_device();

struct RtoScene {
	int kind = RTO_MODE_BVH;          // RTO_MODE_BVH or RTO_MODE_OCTREE_GLSL (any octree)
	int device = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t copyStream = nullptr;            // device->host plane copies of RTO_MEM_HOST batches overlap the next frame's kernel
	cudaEvent_t evStart = nullptr, evStop = nullptr, evFrame = nullptr;
	bool timed = false;
	uint64_t launches = 0;
	size_t deviceBytes = 0, numPrims = 0, numNodes = 0;
	rto::BvhDev bvh{};                // reference topology (BVH::query replay, stats, RTO_FLAG_NO_PRUNE)
	rto::BvhDev bvhFast{};            // SAH topology over the same leaves (production closest-hit / shadow rays)
	rto::OctDev oct{};
	std::vector<void*> owned;         // device allocations of the scene
	std::vector<size_t> ownedBytes;   // and their sizes (rto_scene_save writes the ones the descriptors point into)
	// growable scratch (device outputs for RTO_MEM_HOST calls, cameras, ray lists)
	void* scratch[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
	size_t scratchBytes[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
	int smCount = 0;
	// camera arrays of batched calls: pinned host slot -> device slot, a ring so that a call never waits for the previous one's kernel
	struct CamSlot {
		RtoCamera* host = nullptr; RtoCamera* dev = nullptr; int cap = 0, count = 0;
		cudaEvent_t uploaded = nullptr;
		cudaStream_t readerStream[4] = { nullptr, nullptr, nullptr, nullptr };      // last kernel that read the slot, per stream that ever did
		cudaEvent_t readerDone[4] = { nullptr, nullptr, nullptr, nullptr };
		bool used = false;
	};
	static constexpr int kCamSlots = 64;
	CamSlot camRing[kCamSlots];
	int camNext = 0;
	cudaEvent_t evTable = nullptr;    // "the table below is filled"
	float* shadeTable = nullptr;      // BVH scenes: Lambert term of every triangle's normal in leaf order (rto_resolve_codes; built on first use)
	bool octIsTree = true;            // octree scenes: the uploaded child graph is a tree (OctLayout::isTree)
	bool deviceBuiltBvh = false;      // linear BVH built on the device: the reference-shaped tree (BVH::query replay, work counters) does not exist
};

int rto_scene_new(RtoScene** out);                             // stream, events, device id
int rto_scene_alloc(RtoScene* s, void** p, size_t bytes);      // device memory owned by the scene
int rto_scene_adopt(RtoScene* s, void* p, size_t bytes);       // take ownership of an existing cudaMalloc'ed block
int rto_scene_scratch(RtoScene* s, int slot, size_t bytes, void** p);

// Enqueue a render of rows [y0, y1) of `numCams` cameras on `st` (a stream of the scene's device; the planes and the code buffer are
// device-accessible addresses, any of them may be null; codes: BVH scenes only, see include/rto_c.h rto_render_codes), and the
// expansion of hit codes into planes.  Camera arrays are staged through the scene's pinned ring.  Nothing here synchronises.
int rto_enqueue_render(RtoScene* s, const RtoCamera* cams, int numCams, int mode, uint32_t flags, float shadowBias, int y0, int y1,
	float4* rgba, int32_t* hitId, float* t, uint32_t* codes, size_t codeFrame0, cudaStream_t st);
int rto_enqueue_resolve(RtoScene* s, const RtoCamera* cams, int numCams, int y0, int y1, const uint32_t* codes, size_t codeFrame0,
	float4* rgba, int32_t* hitId, float* t, cudaStream_t st);
// the device scene of a triangle soup on the CURRENT device from a layout built once on the host (rto_group.cu replicates it)
struct BvhLayout;
int rto_bvh_layout_from_tris(const RtoTriangle* tris, size_t numTris, const RtoHostBvh* prebuilt, BvhLayout& L, size_t* numRefNodes);
int rto_scene_from_bvh_layout(const BvhLayout& L, size_t numTris, size_t numRefNodes, RtoScene** out);

// Dual Contouring on the device (rto_dc.cu) for a grid and a node array that already live there; *dTrisOut is cudaMalloc'ed and
// belongs to the caller.  Synchronises the stream.
int rto_dc_extract_device(const uint8_t* dVox, int dimX, int dimY, int dimZ, const float gridMin[3], float voxelSize,
	const RtoGpuNode* dNodes, size_t numNodes, const float* viewProj16, float extraMargin, cudaStream_t st,
	RtoTriangle** dTrisOut, size_t* numTris, size_t* numLeavesOut, int* roundsOut);
