// rto_voxelize.h -- the per-face part of the CSV voxeliser (BuildingLoader.cpp:131-150, 229-287), written once for the host
// fill (host_builders.cpp) and the device fill (rto_build.cu).  One IEEE operation per operator, glm's dot order (rto_math.h).
#pragma once
#include "rto_math.h"
#include "../../include/rto_c.h"

namespace rto {

// isPointInTriangle (BuildingLoader.cpp:131-150): barycentric coordinates of p's projection onto the triangle's plane
RTO_HD bool vox_point_in_triangle(V3 p, V3 a, V3 b, V3 c) {
	V3 v0 = c - a, v1 = b - a, v2 = p - a;
	float dot00 = dot3(v0, v0), dot01 = dot3(v0, v1), dot02 = dot3(v0, v2), dot11 = dot3(v1, v1), dot12 = dot3(v1, v2);
	float invDenom = dot00 * dot11 - dot01 * dot01;
	if (fabsf(invDenom) < 1e-7f) return false;
	invDenom = 1.0f / invDenom;
	float u = (dot11 * dot02 - dot01 * dot12) * invDenom;
	float v = (dot00 * dot12 - dot01 * dot02) * invDenom;
	return (u >= 0) && (v >= 0) && (u + v <= 1);
}

struct VoxRange { int x0, y0, z0, x1, y1, z1; bool empty; };

RTO_HD float vox_min3(float a, float b, float c) { float m = a; if (b < m) m = b; if (c < m) m = c; return m; }      // std::min({a, b, c})
RTO_HD float vox_max3(float a, float b, float c) { float m = a; if (m < b) m = b; if (m < c) m = c; return m; }      // std::max({a, b, c})

// voxel range tested for one face (BuildingLoader.cpp:247-261): bounding box in cells, one cell of margin on the far side
RTO_HD VoxRange vox_face_range(const RtoTriangle& t, const float gmin[3], float voxel, const int dims[3]) {
	VoxRange r;
	int lo[3], hi[3];
	for (int a = 0; a < 3; a++) {
		float mn = vox_min3(t.v0[a], t.v1[a], t.v2[a]), mx = vox_max3(t.v0[a], t.v1[a], t.v2[a]);
		int s = (int)((mn - gmin[a]) / voxel), e = (int)((mx - gmin[a]) / voxel) + 1;
		lo[a] = s > 0 ? s : 0;
		hi[a] = e < dims[a] - 1 ? e : dims[a] - 1;
	}
	r.x0 = lo[0]; r.y0 = lo[1]; r.z0 = lo[2]; r.x1 = hi[0]; r.y1 = hi[1]; r.z1 = hi[2];
	r.empty = r.x1 < r.x0 || r.y1 < r.y0 || r.z1 < r.z0;
	return r;
}

// centre of voxel (x, y, z) (BuildingLoader.cpp:266-270)
RTO_HD V3 vox_center(const float gmin[3], float voxel, int x, int y, int z) {
	return mk3(gmin[0] + (float(x) + 0.5f) * voxel, gmin[1] + (float(y) + 0.5f) * voxel, gmin[2] + (float(z) + 0.5f) * voxel);
}

} // namespace rto
