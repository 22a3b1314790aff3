// VolumeRaycastRenderer.h (shim) -- the one piece of VolumeRaycastRenderer that is a traversal: octreeRaySkip
// (453-skeleton/VolumeRaycastRenderer.cpp:50-155), the recursive ray / octree walk its per-frame skip-distance estimate calls for 49
// probe rays (:1598-1664).  Same signature for a single ray (the visibility map the reference may pass holds only `true` entries --
// :1123 is its only writer -- and is not modelled); the walk itself runs on the GPU (RTO_MODE_OCTREE_SKIP).  Batches of rays and the
// whole skip-distance estimate should use rto_trace_rays / rto_octree_skip_distance directly: one launch instead of one per ray.
#pragma once
#include "RayTracerBVH.h"
#include <unordered_map>

namespace rto_shim {
// scenes of the (sub)trees octreeRaySkip has been called on, keyed by node; rto_shim_forget_octree(node) before freeOctree
inline std::unordered_map<const OctreeNode*, RtoScene*>& raySkipScenes() { static std::unordered_map<const OctreeNode*, RtoScene*> m; return m; }
}

inline void rto_shim_forget_octree(const OctreeNode* node) {
	auto& m = rto_shim::raySkipScenes();
	auto it = m.find(node);
	if (it != m.end()) { rto_scene_destroy(it->second); m.erase(it); }
}

// returns the entry distance of the first solid leaf the walk accepts, or 1e30f (VolumeRaycastRenderer.cpp:50-155)
inline float octreeRaySkip(const OctreeNode* node, const rto_shim::vec3& ro, const rto_shim::vec3& rd, float tMin, float tMax, const VoxelGrid& grid) {
	if (!node) return 1e30f;
	auto& m = rto_shim::raySkipScenes();
	auto it = m.find(node);
	if (it == m.end()) {
		std::vector<RtoGpuNode> flat = rto_shim_flatten(node);               // the subtree under `node`, node coordinates stay absolute
		float gmin[3] = { grid.minX, grid.minY, grid.minZ };
		RtoScene* sc = nullptr;
		if (rto_scene_create_octree(flat.data(), flat.size(), gmin, grid.voxelSize, &sc) != RTO_OK) { std::fprintf(stderr, "[octreeRaySkip] %s\n", rto_last_error()); return 1e30f; }
		it = m.emplace(node, sc).first;
	}
	const float o[3] = { ro.x, ro.y, ro.z }, d[3] = { rd.x, rd.y, rd.z };
	float t = 1e30f;
	if (rto_trace_rays(it->second, RTO_MODE_OCTREE_SKIP, 0, o, d, 1, tMin, tMax, &t, nullptr, RTO_MEM_HOST) != RTO_OK) { std::fprintf(stderr, "[octreeRaySkip] %s\n", rto_last_error()); return 1e30f; }
	return t;
}
