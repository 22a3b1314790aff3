// rto_blob.cu -- a device scene as one file: the flattened arrays the kernels walk, written once and uploaded as they are.
//
// The reference keeps two caches next to the path: sceneCache.bin (the voxel grid, CacheUtils.cpp:5-59) and the Dual-Contouring triangle
// cache (main.cpp:27-92); everything after them -- octree, flatten, BVH -- is rebuilt at every start (main.cpp:1127-1131).  Both formats
// are read and written by rto_host_grid_* / rto_host_tricache_*.  This is the third cache SURVEY.md section 5 asks for: the DEVICE
// layout itself (production and reference-shaped BVH nodes, triangle records, wide form; compact or general octree arrays), whatever
// route built it (host builders, device builders, a frustum-culled array), so that a start costs one read and one upload.
//
// File: header | segment sizes | descriptors with every device pointer replaced by (segment, byte offset) | segment bytes | checksum.
// A file written by another build of the library (different descriptor sizes or version) is refused, not guessed at.
#include "rto_scene.cuh"
#include "rto_nvtx.h"
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

using namespace rto;

namespace {

constexpr char kMagic[8] = { 'R', 'T', 'O', 'S', 'C', 'N', '1', 0 };
constexpr uint32_t kBlobVersion = 1;
constexpr int kNumPtrs = 8;      // BvhDev (bvh): nodes, tris; BvhDev (bvhFast): nodes, exactNodes, wide; OctDev: desc, up, inner / nodes16 (see ptr_table)

struct BlobHeader {
	char magic[8];
	uint32_t version, kind;
	uint32_t sizeofBvhDev, sizeofOctDev;
	uint64_t numPrims, numNodes;
	uint32_t octIsTree, deviceBuiltBvh;
	uint32_t numSegments, numPtrs;
};
struct BlobPtr { int32_t segment; uint32_t pad; uint64_t offset; };      // segment < 0: null pointer

// Every device pointer of the three descriptors, in a fixed order.  bvhFast.tris always equals bvh.tris and bvh.exactNodes / bvh.wide
// are copies of fields listed here, so they are re-derived on load the way the builders set them.
struct PtrSlot { const void** field; };
static void ptr_table(RtoScene* s, const void** out[kNumPtrs]) {
	out[0] = reinterpret_cast<const void**>(&s->bvh.nodes);
	out[1] = reinterpret_cast<const void**>(&s->bvh.tris);
	out[2] = reinterpret_cast<const void**>(&s->bvhFast.nodes);
	out[3] = reinterpret_cast<const void**>(&s->bvhFast.wide);
	out[4] = reinterpret_cast<const void**>(&s->oct.desc);
	out[5] = reinterpret_cast<const void**>(&s->oct.up);
	out[6] = reinterpret_cast<const void**>(&s->oct.inner);
	out[7] = reinterpret_cast<const void**>(&s->oct.nodes16);
}

// sum of the 64-bit words (tail bytes zero-extended): cheap enough for multi-gigabyte scenes, enough to catch truncation and bit rot
struct Checksum {
	uint64_t sum = 0x9E3779B97F4A7C15ull;
	void add(const void* data, size_t n) {
		const unsigned char* p = static_cast<const unsigned char*>(data);
		size_t i = 0;
		for (; i + 8 <= n; i += 8) { uint64_t w; std::memcpy(&w, p + i, 8); sum = (sum ^ w) * 0x100000001B3ull + (sum >> 29); }
		if (i < n) { uint64_t w = 0; std::memcpy(&w, p + i, n - i); sum = (sum ^ w) * 0x100000001B3ull + (sum >> 29); }
	}
};

struct File {
	std::FILE* f = nullptr;
	~File() { if (f) std::fclose(f); }
};

}  // namespace

extern "C" int rto_scene_save(RtoScene* s, const char* path) try {
	RTO_RANGE("rto_scene_save");
	if (!s || !path) return rto_fail(RTO_ERR_INVALID, "rto_scene_save: null argument");
	CUDA_TRY(cudaSetDevice(s->device));
	CUDA_TRY(cudaStreamSynchronize(s->stream));
	const void** ptrs[kNumPtrs];
	ptr_table(s, ptrs);
	// the allocations the descriptors point into, in the order they are first met
	std::vector<int> segOf(s->owned.size(), -1);
	std::vector<int> segments;
	BlobPtr table[kNumPtrs];
	for (int k = 0; k < kNumPtrs; k++) {
		table[k].segment = -1; table[k].pad = 0; table[k].offset = 0;
		const char* p = static_cast<const char*>(*ptrs[k]);
		if (!p) continue;
		int found = -1;
		for (size_t a = 0; a < s->owned.size(); a++) {
			const char* base = static_cast<const char*>(s->owned[a]);
			if (p >= base && p < base + s->ownedBytes[a]) { found = (int)a; break; }
		}
		if (found < 0) return rto_fail(RTO_ERR_UNSUPPORTED, "rto_scene_save: a scene array does not lie in memory the scene owns");
		if (segOf[found] < 0) { segOf[found] = (int)segments.size(); segments.push_back(found); }
		table[k].segment = segOf[found];
		table[k].offset = (uint64_t)(p - static_cast<const char*>(s->owned[found]));
	}
	BlobHeader H{};
	std::memcpy(H.magic, kMagic, 8);
	H.version = kBlobVersion; H.kind = (uint32_t)s->kind;
	H.sizeofBvhDev = (uint32_t)sizeof(BvhDev); H.sizeofOctDev = (uint32_t)sizeof(OctDev);
	H.numPrims = s->numPrims; H.numNodes = s->numNodes;
	H.octIsTree = s->octIsTree ? 1u : 0u; H.deviceBuiltBvh = s->deviceBuiltBvh ? 1u : 0u;
	H.numSegments = (uint32_t)segments.size(); H.numPtrs = kNumPtrs;
	// descriptors with their pointers blanked (the table carries them)
	RtoScene blank;
	blank.bvh = s->bvh; blank.bvhFast = s->bvhFast; blank.oct = s->oct;
	{
		const void** bp[kNumPtrs];
		ptr_table(&blank, bp);
		for (int k = 0; k < kNumPtrs; k++) *bp[k] = nullptr;
		blank.bvh.exactNodes = nullptr; blank.bvh.wide = nullptr; blank.bvhFast.tris = nullptr; blank.bvhFast.exactNodes = nullptr;
	}
	// which tree the production descriptor falls back to for rays outside the fused tests: the reference-shaped one or itself
	const uint32_t exactIsOwn = (s->bvhFast.exactNodes == s->bvhFast.nodes) ? 1u : 0u;

	File out; out.f = std::fopen(path, "wb");
	if (!out.f) return rto_fail(RTO_ERR_IO, "rto_scene_save: cannot open %s for writing", path);
	Checksum ck;
	auto put = [&](const void* data, size_t n) -> bool { ck.add(data, n); return std::fwrite(data, 1, n, out.f) == n; };
	bool ok = put(&H, sizeof H);
	for (int a : segments) { const uint64_t n = s->ownedBytes[a]; ok = ok && put(&n, 8); }
	ok = ok && put(table, sizeof table) && put(&exactIsOwn, 4) && put(&blank.bvh, sizeof(BvhDev)) && put(&blank.bvhFast, sizeof(BvhDev)) && put(&blank.oct, sizeof(OctDev));
	// segments through a pinned staging buffer, 64 MiB at a time
	const size_t chunk = (size_t)64 << 20;
	void* stage = nullptr;
	if (ok && !segments.empty()) CUDA_TRY(cudaHostAlloc(&stage, chunk, cudaHostAllocDefault));
	std::unique_ptr<void, void (*)(void*)> stageGuard(stage, [](void* p) { if (p) cudaFreeHost(p); });
	for (size_t k = 0; ok && k < segments.size(); k++) {
		const char* base = static_cast<const char*>(s->owned[segments[k]]);
		const size_t n = s->ownedBytes[segments[k]];
		for (size_t off = 0; ok && off < n; off += chunk) {
			const size_t m = n - off < chunk ? n - off : chunk;
			CUDA_TRY(cudaMemcpy(stage, base + off, m, cudaMemcpyDeviceToHost));
			ok = put(stage, m);
		}
	}
	const uint64_t sum = ck.sum;
	ok = ok && std::fwrite(&sum, 1, 8, out.f) == 8;
	ok = ok && std::fflush(out.f) == 0;
	if (!ok) return rto_fail(RTO_ERR_IO, "rto_scene_save: write to %s failed", path);
	return RTO_OK;
} RTO_CATCH_ALL("rto_scene_save")

extern "C" int rto_scene_load(const char* path, RtoScene** outScene) try {
	RTO_RANGE("rto_scene_load");
	if (!outScene) return rto_fail(RTO_ERR_INVALID, "rto_scene_load: null output");
	*outScene = nullptr;
	if (!path) return rto_fail(RTO_ERR_INVALID, "rto_scene_load: null path");
	int rc = rto_require_device(); if (rc) return rc;
	File in; in.f = std::fopen(path, "rb");
	if (!in.f) return rto_fail(RTO_ERR_IO, "rto_scene_load: cannot open %s", path);
	Checksum ck;
	auto get = [&](void* data, size_t n) -> bool { if (std::fread(data, 1, n, in.f) != n) return false; ck.add(data, n); return true; };
	BlobHeader H{};
	if (!get(&H, sizeof H) || std::memcmp(H.magic, kMagic, 8) != 0) return rto_fail(RTO_ERR_IO, "rto_scene_load: %s is not a scene file", path);
	if (H.version != kBlobVersion || H.sizeofBvhDev != sizeof(BvhDev) || H.sizeofOctDev != sizeof(OctDev) || H.numPtrs != kNumPtrs)
		return rto_fail(RTO_ERR_UNSUPPORTED, "rto_scene_load: %s was written by another version of the library (rebuild the scene)", path);
	if (H.numSegments > kNumPtrs || (H.kind != RTO_MODE_BVH && H.kind != RTO_MODE_OCTREE_GLSL)) return rto_fail(RTO_ERR_IO, "rto_scene_load: %s: bad header", path);
	uint64_t segBytes[kNumPtrs] = {};
	for (uint32_t k = 0; k < H.numSegments; k++) if (!get(&segBytes[k], 8)) return rto_fail(RTO_ERR_IO, "rto_scene_load: %s is truncated", path);
	BlobPtr table[kNumPtrs]; uint32_t exactIsOwn = 0;
	RtoScene* s = nullptr;
	rc = rto_scene_new(&s); if (rc) return rc;
	std::unique_ptr<RtoScene, void (*)(RtoScene*)> guard(s, [](RtoScene* p) { rto_scene_destroy(p); });
	if (!get(table, sizeof table) || !get(&exactIsOwn, 4) || !get(&s->bvh, sizeof(BvhDev)) || !get(&s->bvhFast, sizeof(BvhDev)) || !get(&s->oct, sizeof(OctDev)))
		return rto_fail(RTO_ERR_IO, "rto_scene_load: %s is truncated", path);
	for (int k = 0; k < kNumPtrs; k++)
		if (table[k].segment >= (int32_t)H.numSegments || (table[k].segment >= 0 && table[k].offset >= segBytes[table[k].segment]))
			return rto_fail(RTO_ERR_IO, "rto_scene_load: %s: bad pointer table", path);
	s->kind = (int)H.kind; s->numPrims = (size_t)H.numPrims; s->numNodes = (size_t)H.numNodes;
	s->octIsTree = H.octIsTree != 0; s->deviceBuiltBvh = H.deviceBuiltBvh != 0;
	// segments: file -> pinned staging -> device
	const size_t chunk = (size_t)64 << 20;
	void* stage = nullptr;
	if (H.numSegments) CUDA_TRY(cudaHostAlloc(&stage, chunk, cudaHostAllocDefault));
	std::unique_ptr<void, void (*)(void*)> stageGuard(stage, [](void* p) { if (p) cudaFreeHost(p); });
	void* segDev[kNumPtrs] = {};
	for (uint32_t k = 0; k < H.numSegments; k++) {
		if ((rc = rto_scene_alloc(s, &segDev[k], (size_t)segBytes[k]))) return rc;      // (a size the file lies about fails here or at the read below)
		for (size_t off = 0; off < segBytes[k]; off += chunk) {
			const size_t m = segBytes[k] - off < chunk ? (size_t)(segBytes[k] - off) : chunk;
			if (!get(stage, m)) return rto_fail(RTO_ERR_IO, "rto_scene_load: %s is truncated", path);
			CUDA_TRY(cudaMemcpy(static_cast<char*>(segDev[k]) + off, stage, m, cudaMemcpyHostToDevice));
		}
	}
	uint64_t sum = 0;
	if (std::fread(&sum, 1, 8, in.f) != 8 || sum != ck.sum) return rto_fail(RTO_ERR_IO, "rto_scene_load: %s is damaged (checksum)", path);
	const void** ptrs[kNumPtrs];
	ptr_table(s, ptrs);
	for (int k = 0; k < kNumPtrs; k++) *ptrs[k] = table[k].segment < 0 ? nullptr : static_cast<const char*>(segDev[table[k].segment]) + table[k].offset;
	// the derived pointers, as the builders set them (rto_scene_from_bvh_layout, rto_build.cu)
	s->bvhFast.tris = s->bvh.tris;
	s->bvh.exactNodes = s->bvh.nodes; s->bvh.wide = nullptr;
	s->bvhFast.exactNodes = exactIsOwn ? s->bvhFast.nodes : s->bvh.nodes;
	if (s->kind == RTO_MODE_BVH) {
		if (s->bvh.numTris < 0 || (size_t)s->bvh.numTris != s->numPrims || (s->numPrims && (!s->bvh.tris || !s->bvhFast.nodes)))
			return rto_fail(RTO_ERR_IO, "rto_scene_load: %s: inconsistent BVH scene", path);
	}
	else if (s->oct.numNodes <= 0 || (size_t)s->oct.numNodes != s->numNodes || (s->oct.compact ? (!s->oct.desc || !s->oct.up || !s->oct.inner) : !s->oct.nodes16))
		return rto_fail(RTO_ERR_IO, "rto_scene_load: %s: inconsistent octree scene", path);
	*outScene = guard.release();
	return RTO_OK;
} RTO_CATCH_ALL("rto_scene_load")
