// BVH.h (shim) -- the reference's BVH class (453-skeleton/BVH.h:7-63) on top of librto: the tree is built on the host with
// the reference's exact shape, uploaded once, and query() runs on the GPU (rto_bvh_query).  closestHit()/render are the
// north-star extensions.  Like the reference, the triangle vector must outlive the BVH (BVH.cpp:21-25).
#pragma once
#include "../../../include/rto_c.h"
#include "rto_shim_math.h"
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

struct Triangle { rto_shim::vec3 v0, v1, v2; };                            // BVH.h:7-11 (36 bytes)
static_assert(sizeof(Triangle) == sizeof(RtoTriangle), "Triangle must be 9 packed floats");

struct AABB {                                                              // BVH.h:14-34
	rto_shim::vec3 min, max;
	AABB() : min(std::numeric_limits<float>::max()), max(-std::numeric_limits<float>::max()) {}
	void expand(const rto_shim::vec3& p) {                                 // glm::min / glm::max: (b < a) ? b : a, (a < b) ? b : a
		min = rto_shim::vec3(p.x < min.x ? p.x : min.x, p.y < min.y ? p.y : min.y, p.z < min.z ? p.z : min.z);
		max = rto_shim::vec3(max.x < p.x ? p.x : max.x, max.y < p.y ? p.y : max.y, max.z < p.z ? p.z : max.z);
	}
	void expand(const AABB& other) { expand(other.min); expand(other.max); }
};

struct BVHNode {                                                           // BVH.h:37-42
	AABB bounds;
	BVHNode* left = nullptr;
	BVHNode* right = nullptr;
	std::vector<const Triangle*> triangles;                                // non-empty for leaf nodes
};

class BVH {
public:
	explicit BVH(const std::vector<Triangle>& triangles) : m_tris(&triangles) {
		if (rto_host_bvh_build(reinterpret_cast<const RtoTriangle*>(triangles.data()), triangles.size(), &m_host) != RTO_OK)
			throw std::runtime_error(std::string("BVH build: ") + rto_last_error());
	}
	~BVH() { rto_scene_destroy(m_scene); rto_host_bvh_free(m_host); }
	BVH(const BVH&) = delete;
	BVH& operator=(const BVH&) = delete;

	// BVH.cpp:107-113: candidate triangles of every leaf whose box chain the ray hits, in the reference's order.
	void query(const rto_shim::vec3& origin, const rto_shim::vec3& direction, std::vector<const Triangle*>& outCandidates) const {
		ensureScene();
		float o[3] = { origin.x, origin.y, origin.z }, d[3] = { direction.x, direction.y, direction.z };
		int64_t off[2]; size_t total = 0;
		check(rto_bvh_query(m_scene, o, d, 1, off, nullptr, 0, &total));
		std::vector<int32_t> ids(total);
		if (total) check(rto_bvh_query(m_scene, o, d, 1, off, ids.data(), total, &total));
		for (int32_t id : ids) outCandidates.push_back(&(*m_tris)[id]);
	}
	// extension: closest Moller-Trumbore hit over the candidates (SURVEY.md 8c rule); returns triangle index or -1
	int closestHit(const rto_shim::vec3& origin, const rto_shim::vec3& direction, float* tOut = nullptr) const {
		ensureScene();
		float o[3] = { origin.x, origin.y, origin.z }, d[3] = { direction.x, direction.y, direction.z };
		float t; int32_t id;
		check(rto_trace_rays(m_scene, RTO_MODE_BVH, 0, o, d, 1, 0.0f, 1e30f, &t, &id, RTO_MEM_HOST));
		if (tOut) *tOut = t;
		return id;
	}
	RtoScene* scene() const { ensureScene(); return m_scene; }
	const RtoHostBvh* host() const { return m_host; }
	const std::vector<Triangle>& triangles() const { return *m_tris; }
	// the pointer tree the reference keeps in its private `root` (BVH.h:54), for callers that walk it themselves: same shape, same
	// boxes, same leaves, materialised on first use from the pre-order export of the host tree
	const BVHNode* root() const {
		if (m_nodes) return m_nodes->empty() ? nullptr : m_nodes->data();
		const size_t n = rto_host_bvh_num_nodes(m_host);
		std::vector<float> boxes(n * 6); std::vector<int32_t> meta(n * 4);
		check(rto_host_bvh_export(m_host, boxes.data(), meta.data(), n));
		m_nodes.reset(new std::vector<BVHNode>(n));
		// pre-order, left subtree first: the left child of node i is i + 1, the right child follows the left subtree
		std::vector<size_t> stack;                                          // inner nodes whose right child is still to come
		for (size_t i = 0; i < n; i++) {
			BVHNode& nd = (*m_nodes)[i];
			nd.bounds.min = rto_shim::vec3(boxes[6 * i], boxes[6 * i + 1], boxes[6 * i + 2]);
			nd.bounds.max = rto_shim::vec3(boxes[6 * i + 3], boxes[6 * i + 4], boxes[6 * i + 5]);
			if (i > 0) {
				BVHNode& parent = (*m_nodes)[stack.back()];
				if (!parent.left) parent.left = &nd; else { parent.right = &nd; stack.pop_back(); }
			}
			if (meta[4 * i]) for (int k = 0; k < meta[4 * i + 1]; k++) nd.triangles.push_back(&(*m_tris)[meta[4 * i + 2 + k]]);
			else stack.push_back(i);
		}
		return m_nodes->empty() ? nullptr : m_nodes->data();
	}

private:
	void ensureScene() const {
		if (m_scene) return;
		check(rto_scene_create_bvh(reinterpret_cast<const RtoTriangle*>(m_tris->data()), m_tris->size(), m_host, &m_scene));
	}
	static void check(int rc) { if (rc != RTO_OK) throw std::runtime_error(std::string("librto: ") + rto_last_error()); }
	const std::vector<Triangle>* m_tris;
	RtoHostBvh* m_host = nullptr;
	mutable RtoScene* m_scene = nullptr;
	mutable std::unique_ptr<std::vector<BVHNode>> m_nodes;
};
