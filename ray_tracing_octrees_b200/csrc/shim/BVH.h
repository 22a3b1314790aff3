// BVH.h (shim) -- the reference's BVH class (453-skeleton/BVH.h:7-63) on top of librto: the tree is built on the host with
// the reference's exact shape, uploaded once, and query() runs on the GPU (rto_bvh_query).  closestHit()/render are the
// north-star extensions.  Like the reference, the triangle vector must outlive the BVH (BVH.cpp:21-25).
#pragma once
#include "../../../include/rto_c.h"
#include "rto_shim_math.h"
#include <stdexcept>
#include <string>
#include <vector>

struct Triangle { rto_shim::vec3 v0, v1, v2; };                            // BVH.h:7-11 (36 bytes)
static_assert(sizeof(Triangle) == sizeof(RtoTriangle), "Triangle must be 9 packed floats");

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

private:
	void ensureScene() const {
		if (m_scene) return;
		check(rto_scene_create_bvh(reinterpret_cast<const RtoTriangle*>(m_tris->data()), m_tris->size(), m_host, &m_scene));
	}
	static void check(int rc) { if (rc != RTO_OK) throw std::runtime_error(std::string("librto: ") + rto_last_error()); }
	const std::vector<Triangle>* m_tris;
	RtoHostBvh* m_host = nullptr;
	mutable RtoScene* m_scene = nullptr;
};
