// rto_frustum.h -- the frustum test of RayTracerBVH::renderSceneComputeWithCulling (RayTracerBVH.cpp:724-757) and of Frustum
// (Frustum.cpp:5-93), written once for the host cull (host_builders.cpp) and the device cull (rto_build.cu).  glm's operation
// order (rto_math.h), one IEEE operation per operator.
#pragma once
#include "rto_math.h"
#include "../../include/rto_c.h"

namespace rto {

struct FrustumPlanes { float p[6][4]; };      // normalised planes: LEFT, RIGHT, BOTTOM, TOP, NEAR, FAR (the verdict does not depend on the order)

// Gribb-Hartmann extraction from a column-major view-projection matrix (m[col * 4 + row] == viewProj[col][row]), Frustum.cpp:5-48
RTO_HD FrustumPlanes frustum_from_view_proj(const float* m) {
	FrustumPlanes F;
	const int rowOf[6] = { 0, 0, 1, 1, 2, 2 };
	const float sgn[6] = { 1.0f, -1.0f, 1.0f, -1.0f, 1.0f, -1.0f };
	for (int i = 0; i < 6; i++) {
		for (int c = 0; c < 4; c++) {
			float a = m[c * 4 + 3], b = m[c * 4 + rowOf[i]];
			F.p[i][c] = sgn[i] > 0.0f ? a + b : a - b;
		}
		float len = sqrtf(dot3(mk3(F.p[i][0], F.p[i][1], F.p[i][2]), mk3(F.p[i][0], F.p[i][1], F.p[i][2])));      // glm::length(vec3)
		for (int c = 0; c < 4; c++) F.p[i][c] = F.p[i][c] / len;
	}
	return F;
}

// Frustum::testAABB(min, max, margin) != -1 for the world box of an octree node (RayTracerBVH.cpp:741-756, Frustum.cpp:52-93):
// false only if the box, grown by the margin, lies entirely behind one of the planes
RTO_HD bool frustum_node_visible(const FrustumPlanes& F, const RtoGpuNode& n, const float gridMin[3], float voxel, float margin) {
	V3 mn = mk3(gridMin[0] + float(n.x) * voxel, gridMin[1] + float(n.y) * voxel, gridMin[2] + float(n.z) * voxel);
	float w = float(n.size) * voxel;
	V3 mx = mk3(mn.x + w, mn.y + w, mn.z + w);
	V3 emn = mk3(mn.x - margin, mn.y - margin, mn.z - margin), emx = mk3(mx.x + margin, mx.y + margin, mx.z + margin);
	for (int i = 0; i < 6; i++) {
		V3 p = mk3(F.p[i][0] > 0 ? emx.x : emn.x, F.p[i][1] > 0 ? emx.y : emn.y, F.p[i][2] > 0 ? emx.z : emn.z);
		if (dot3(mk3(F.p[i][0], F.p[i][1], F.p[i][2]), p) + F.p[i][3] < 0) return false;
	}
	return true;
}

} // namespace rto
