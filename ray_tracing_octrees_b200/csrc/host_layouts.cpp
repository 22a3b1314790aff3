// host_layouts.cpp -- device memory layouts of the scenes, built on the host (no CUDA calls here).
// rto_device.cu uploads these arrays verbatim; tests/emu links this file to run the same traversal code on the CPU.
#include "rto_internal.h"
#include <algorithm>
#include <cstdlib>
#include <cstring>

#ifndef RTO_BVH_WIDE_DEFAULT
#define RTO_BVH_WIDE_DEFAULT false
#endif

// The compact layout needs the shape the reference builder always produces: every internal node has 8 children
// with consecutive indices, child boxes are the 8 octants of the parent, leaf <=> uniform.
static bool octree_is_compactable(const RtoGpuNode* n, size_t count) {
	if (count == 0 || n[0].x != 0 || n[0].y != 0 || n[0].z != 0 || n[0].size <= 0 || (n[0].size & (n[0].size - 1))) return false;
	if (count >= 0x3fffffffu) return false;
	std::vector<uint8_t> seen(count, 0);
	seen[0] = 1;
	for (size_t i = 0; i < count; i++) {
		const RtoGpuNode& p = n[i];
		bool leafLike = (p.isLeaf == 1) || (p.isUniform == 1);
		if (p.isLeaf != 0 && p.isLeaf != 1) return false;
		if (p.isUniform != p.isLeaf) return false;
		if (p.isSolid != 0 && p.isSolid != 1) return false;
		if (leafLike) continue;
		int first = p.child[0];
		if (first <= 0 || ((first - 1) & 7) != 0 || (size_t)first + 7 >= count || p.size < 2) return false;
		int half = p.size / 2;
		for (int c = 0; c < 8; c++) {
			if (p.child[c] != first + c) return false;
			const RtoGpuNode& q = n[first + c];
			if (q.size != half || q.x != p.x + ((c & 1) ? half : 0) || q.y != p.y + ((c & 2) ? half : 0) || q.z != p.z + ((c & 4) ? half : 0)) return false;
			if (seen[first + c]) return false;
			seen[first + c] = 1;
		}
	}
	for (size_t i = 0; i < count; i++) if (!seen[i]) return false;
	return true;
}


int rto_build_octree_layout(const RtoGpuNode* nodes, size_t numNodes, OctLayout& L) {
	if (numNodes > (size_t)0x7fffff00) return rto_fail(RTO_ERR_UNSUPPORTED, "rto_scene_create_octree: too many nodes");
	for (size_t i = 0; i < numNodes; i++)
		for (int c = 0; c < 8; c++)
			if (nodes[i].child[c] >= (int64_t)numNodes) return rto_fail(RTO_ERR_INVALID, "rto_scene_create_octree: node %zu child %d out of range", i, c);
	L.numLeaves = 0;
	for (size_t i = 0; i < numNodes; i++) L.numLeaves += nodes[i].isLeaf ? 1 : 0;
	{	// in-degree <= 1 everywhere and 0 at the root  <=>  what can be reached from the root is a tree (no shared children, no cycles)
		std::vector<uint8_t> refs(numNodes, 0);
		L.isTree = true;
		for (size_t i = 0; i < numNodes && L.isTree; i++)
			for (int c = 0; c < 8; c++) {
				const int32_t k = nodes[i].child[c];
				if (k < 0) continue;
				if (k == 0 || refs[k]) { L.isTree = false; break; }
				refs[k] = 1;
			}
	}
	L.compact = octree_is_compactable(nodes, numNodes);
	if (!L.compact) {
		L.padded.assign(numNodes * 16, -1);
		for (size_t i = 0; i < numNodes; i++) std::memcpy(&L.padded[16 * i], &nodes[i], sizeof(RtoGpuNode));
		return RTO_OK;
	}
	// desc[] is offset by 7 words so that every sibling group (indices 1+8g .. 8+8g) is one aligned 32-byte sector
	L.desc.assign(numNodes + 8, 0);
	L.up.assign((numNodes + 7) / 8 + 1, 0);
	for (size_t i = 0; i < numNodes; i++) {
		const RtoGpuNode& p = nodes[i];
		if (p.isLeaf) L.desc[7 + i] = 0x80000000u | (p.isSolid ? 0x40000000u : 0u);
		else { L.desc[7 + i] = (uint32_t)p.child[0]; L.up[(p.child[0] - 1) >> 3] = (int32_t)i; }
	}
	// one 16-byte record per internal node, ranked in BFS order (the children of a node are 8 consecutive BFS indices, so its
	// internal children have consecutive ranks starting at `internalBase`)
	std::vector<int32_t> rankOf(numNodes, -1);
	int32_t numInner = 0;
	for (size_t i = 0; i < numNodes; i++) if (!nodes[i].isLeaf) rankOf[i] = numInner++;
	L.inner.assign((size_t)std::max(numInner, 1) * 4, 0);
	for (size_t i = 0; i < numNodes; i++) {
		const RtoGpuNode& p = nodes[i];
		if (p.isLeaf) continue;
		int32_t* r = &L.inner[(size_t)rankOf[i] * 4];
		uint32_t leafMask = 0, solidMask = 0; int32_t base = -1;
		for (int c = 0; c < 8; c++) {
			const RtoGpuNode& q = nodes[p.child[c]];
			if (q.isLeaf) { leafMask |= 1u << c; if (q.isSolid) solidMask |= 1u << c; }
			else if (base < 0) base = rankOf[p.child[c]];
		}
		r[0] = p.child[0]; r[1] = base < 0 ? 0 : base;
		r[3] |= (int32_t)(leafMask | (solidMask << 8));            // (bits 16-18 were set by the parent, which comes earlier in BFS order)
		for (int c = 0; c < 8; c++) {                              // tell each internal child who its parent is and which octant it is
			int32_t cr = rankOf[p.child[c]];
			if (cr >= 0) { L.inner[(size_t)cr * 4 + 2] = rankOf[i]; L.inner[(size_t)cr * 4 + 3] |= (int32_t)((uint32_t)c << 16); }
		}
	}
	return RTO_OK;
}

void rto_build_bvh_layout(const RtoHostBvh& h, BvhLayout& L) {
	// two node arrays over the same reference leaves (rto_internal.h): the reference's own topology for replaying
	// BVH::query, and a binned-SAH topology for the production closest-hit / shadow traversal
	rto_build_reference_topology(h, L.refNodes, L.refRoot);
	static const bool refTopology = getenv("RTO_BVH_REFERENCE_TOPOLOGY") != nullptr;     // tuning aid: trace through the reference's tree
	if (!refTopology) rto_build_fast_topology(h, L.fastNodes, L.fastRoot, L.fastGrow);
	L.tris.assign(std::max<size_t>(h.numTris, 1) * 16, 0.0f);
	for (const HostBvhNode& n : h.nodes) {              // every triangle carries the exact box of its reference leaf (floats 10..15)
		if (n.left >= 0) continue;
		for (uint32_t p = n.first; p < n.first + n.count; p++) { std::memcpy(&L.tris[(size_t)p * 16 + 10], n.mn, 12); std::memcpy(&L.tris[(size_t)p * 16 + 13], n.mx, 12); }
	}
	for (size_t p = 0; p < h.numTris; p++) {
		uint32_t id = h.order[p];
		float* d = &L.tris[p * 16];
		const RtoTriangle& T = h.tris[id];
		for (int k = 0; k < 3; k++) { d[k] = T.v0[k]; d[3 + k] = T.v1[k] - T.v0[k]; d[6 + k] = T.v2[k] - T.v0[k]; }     // v0, e1, e2 (rto_kernels.cuh TriV)
		int32_t iid = (int32_t)id;
		std::memcpy(&d[9], &iid, 4);
	}
	const HostBvhNode& root = h.nodes[0];
	for (int k = 0; k < 3; k++) { L.rootLo[k] = root.mn[k]; L.rootHi[k] = root.mx[k]; }
	// the 4-wide quantised form of the production tree (RTO_BVH_WIDE=0 keeps the binary form only)
	const char* we = getenv("RTO_BVH_WIDE");
	const bool wantWide = we ? (atoi(we) != 0) : RTO_BVH_WIDE_DEFAULT;
	if (wantWide && !L.fastNodes.empty() && L.fastRoot >= 0) {
		rto_build_wide_topology(L.fastNodes, L.fastRoot, L.rootLo, L.rootHi, L.fastGrow, L.wideNodes, L.wideRoot, L.wideLo, L.wideStep);
	}
}
