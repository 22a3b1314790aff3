// OctreeVoxel.h (shim) -- the reference's voxel / octree types and builders (453-skeleton/OctreeVoxel.h:10-69) on top of librto.
// createOctreeFromVoxelGrid returns the same pointer tree the reference builds (same nodes, same child order); the tree is
// materialised from librto's flat BFS array; rto_shim_flat_index(node) gives a node's GPUNodes index (== leaf id of RayTracerBVH)
// from a side table, so that OctreeNode itself has exactly the reference's members.
#pragma once
#include "../../../include/rto_c.h"
#include "rto_shim_math.h"
#include <cstdint>
#include <unordered_map>
#include <vector>

enum class VoxelState : uint8_t { EMPTY = 0, FILLED = 1 };

struct MCTriangle { rto_shim::vec3 v[3]; rto_shim::vec3 normal[3]; };      // OctreeVoxel.h:22-25 (72 bytes)

struct VoxelGrid {                                                         // OctreeVoxel.h:28-42
	int dimX = 0, dimY = 0, dimZ = 0;
	float minX = 0.f, minY = 0.f, minZ = 0.f;
	float voxelSize = 1.f;
	std::vector<VoxelState> data;
	int index(int x, int y, int z) const { return x + y * dimX + z * (dimX * dimY); }
};

struct OctreeNode {                                                        // OctreeVoxel.h:45-62
	int x, y, z, size;
	bool isLeaf, isSolid, isUniform;
	OctreeNode* parent;
	OctreeNode* children[8];
	OctreeNode(int _x, int _y, int _z, int _size) : x(_x), y(_y), z(_z), size(_size), isLeaf(false), isSolid(false), isUniform(false), parent(nullptr) {
		for (int i = 0; i < 8; i++) children[i] = nullptr;
	}
};

// node -> index in the BFS GPUNodes array of the octree it was created in (-1 for nodes this library did not create).  One table per
// process, refilled by every createOctreeFromVoxelGrid like the reference's own global g_octreeMap (OctreeVoxel.cpp:773-776).
namespace rto_shim { inline std::unordered_map<const OctreeNode*, int>& flatIndexTable() { static std::unordered_map<const OctreeNode*, int> t; return t; } }
inline int rto_shim_flat_index(const OctreeNode* node) {
	auto& t = rto_shim::flatIndexTable();
	auto it = t.find(node);
	return it == t.end() ? -1 : it->second;
}

inline VoxelState getVoxelSafe(const VoxelGrid& g, int x, int y, int z) {   // OctreeVoxel.cpp:692-701
	if (x < 0 || y < 0 || z < 0 || x >= g.dimX || y >= g.dimY || z >= g.dimZ) return VoxelState::EMPTY;
	return g.data[g.index(x, y, z)];
}

inline OctreeNode* createOctreeFromVoxelGrid(const VoxelGrid& grid) {      // OctreeVoxel.cpp:765-778
	RtoGpuNode* flat = nullptr; size_t n = 0;
	if (rto_host_octree_build(reinterpret_cast<const uint8_t*>(grid.data.data()), grid.dimX, grid.dimY, grid.dimZ, &flat, &n) != RTO_OK || n == 0) return nullptr;
	std::vector<OctreeNode*> nodes(n, nullptr);
	auto& table = rto_shim::flatIndexTable();
	table.clear();
	for (size_t i = 0; i < n; i++) {
		nodes[i] = new OctreeNode(flat[i].x, flat[i].y, flat[i].z, flat[i].size);
		nodes[i]->isLeaf = flat[i].isLeaf != 0; nodes[i]->isSolid = flat[i].isSolid != 0; nodes[i]->isUniform = flat[i].isUniform != 0;
		table[nodes[i]] = (int)i;
	}
	for (size_t i = 0; i < n; i++)
		for (int c = 0; c < 8; c++)
			if (flat[i].child[c] >= 0) { nodes[i]->children[c] = nodes[flat[i].child[c]]; nodes[flat[i].child[c]]->parent = nodes[i]; }
	OctreeNode* root = nodes[0];
	rto_host_free(flat);
	return root;
}

inline void freeOctree(OctreeNode* node) {                                 // OctreeVoxel.cpp:881-888
	if (!node) return;
	for (int i = 0; i < 8; i++) freeOctree(node->children[i]);
	delete node;
}

// MarchingCubesRenderer::render(root, grid, 0,0,0, root->size) (Renderer.cpp:14-36): same triangles in the same order.
// Normals are the flat face normals localMC stores (normalize(cross(v1-v0, v2-v0)), OctreeVoxel.cpp:863-870).
std::vector<MCTriangle> rto_shim_marching_cubes(const OctreeNode* root, const VoxelGrid& grid);
