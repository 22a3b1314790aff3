// Frustum.h (shim) -- the reference's frustum class (453-skeleton/Frustum.h:6-24, Frustum.cpp:5-93): Gribb-Hartmann planes from a
// view-projection matrix, normalised, and the p-vertex / n-vertex box test with a margin.  Host arithmetic in glm's operation order;
// the bulk test over a whole node array is rto_host_frustum_cull / rto_device_frustum_cull (what RayTracerBVH.h uses).
#pragma once
#include "rto_shim_math.h"
#include <array>
#include <cmath>

class Frustum {
public:
	enum Planes { LEFT = 0, RIGHT, TOP, BOTTOM, NEAR, FAR, COUNT };

	explicit Frustum(const rto_shim::mat4& viewProj) {                     // Frustum.cpp:5-48
		auto set = [&](int plane, int row, float sign) {
			for (int c = 0; c < 4; c++) m_planes[plane][c] = sign > 0 ? viewProj[c][3] + viewProj[c][row] : viewProj[c][3] - viewProj[c][row];
		};
		set(LEFT, 0, 1.f); set(RIGHT, 0, -1.f); set(BOTTOM, 1, 1.f); set(TOP, 1, -1.f); set(NEAR, 2, 1.f); set(FAR, 2, -1.f);
		for (int i = 0; i < COUNT; i++) {
			const float len = std::sqrt((m_planes[i][0] * m_planes[i][0] + m_planes[i][1] * m_planes[i][1]) + m_planes[i][2] * m_planes[i][2]);
			for (int c = 0; c < 4; c++) m_planes[i][c] /= len;
		}
	}
	// 1 = fully inside, 0 = intersecting, -1 = fully outside (Frustum.cpp:52-93)
	int testAABB(const rto_shim::vec3& min, const rto_shim::vec3& max, float extraMargin) const {
		const float emin[3] = { min.x - extraMargin, min.y - extraMargin, min.z - extraMargin };
		const float emax[3] = { max.x + extraMargin, max.y + extraMargin, max.z + extraMargin };
		int result = 1;
		for (int i = 0; i < COUNT; i++) {
			const float* pl = m_planes[i].data();
			const float p[3] = { pl[0] > 0 ? emax[0] : emin[0], pl[1] > 0 ? emax[1] : emin[1], pl[2] > 0 ? emax[2] : emin[2] };
			if (((pl[0] * p[0] + pl[1] * p[1]) + pl[2] * p[2]) + pl[3] < 0) return -1;
			const float n[3] = { pl[0] < 0 ? emax[0] : emin[0], pl[1] < 0 ? emax[1] : emin[1], pl[2] < 0 ? emax[2] : emin[2] };
			if (((pl[0] * n[0] + pl[1] * n[1]) + pl[2] * n[2]) + pl[3] < 0) result = 0;
		}
		return result;
	}

private:
	std::array<std::array<float, 4>, COUNT> m_planes;
};
