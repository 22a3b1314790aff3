// rto_dc.h -- the per-cell arithmetic of the reference's Adaptive Dual Contouring mesher (AdaptiveDualContouringRenderer.cpp), written
// once for the host builder (host_dc.cpp) and the device builder (rto_dc.cu).  Every function is a pure function of the grid, the
// octree array and its arguments; the order-dependent part of the reference (its dual-vertex cache) is resolved by the callers.
// One IEEE operation per operator in glm 0.9.9.7's order (rto_math.h); host code is compiled with -ffp-contract=off, device code
// with -fmad=false -prec-div=true -prec-sqrt=true, so both round like the reference.
#pragma once
#include "rto_math.h"
#include "rto_frustum.h"
#include "../../include/rto_c.h"

namespace rto {
namespace dc {

RTO_HD int imin(int a, int b) { return a < b ? a : b; }
RTO_HD int imax(int a, int b) { return a > b ? a : b; }

struct Grid {
	const uint8_t* v; int dx, dy, dz; float mnx, mny, mnz, vs;
	RTO_HD bool inb(int x, int y, int z) const { return x >= 0 && y >= 0 && z >= 0 && x < dx && y < dy && z < dz; }
	RTO_HD bool filled(int x, int y, int z) const { return v[(size_t)x + (size_t)y * dx + (size_t)z * ((size_t)dx * dy)] == 1; }
	// gridToWorld, AdaptiveDualContouringRenderer.cpp:1359-1365
	RTO_HD V3 toWorld(int x, int y, int z) const { return mk3(mnx + float(x) * vs, mny + float(y) * vs, mnz + float(z) * vs); }
	// cell centre as createTriangles forms it (:549-550): gridToWorld + vec3(size * 0.5f * voxelSize)
	RTO_HD V3 centre(int x, int y, int z, int size) const { float h = float(size) * 0.5f * vs; return toWorld(x, y, z) + mk3(h, h, h); }
};

struct Hermite { V3 p, n; };

RTO_HD V3 clamp3(V3 x, V3 lo, V3 hi) { return min3(max3(x, lo), hi); }                 // glm::clamp = min(max(x, lo), hi)
RTO_HD V3 mix3(V3 x, V3 y, float a) { return x * (1.0f - a) + y * a; }                  // glm::mix, func_common.inl:104-112
RTO_HD bool isNan(float x) { return x != x; }
RTO_HD bool isInf(float x) { return fabsf(x) > 3.402823466e+38f; }

// calculateIntersection (:1236-1357) for the only case gatherHermiteData produces: (x2, y2, z2) = (x1, y1, z1) + unit axis d, end
// voxels of different state.  The edge cache of the reference holds pure values and is not needed.
RTO_HD Hermite intersection(const Grid& g, int x1, int y1, int z1, int d) {
	const int dx = d == 0, dy = d == 1, dz = d == 2;
	const int x2 = x1 + dx, y2 = y1 + dy, z2 = z1 + dz;
	const bool isFilled1 = g.filled(x1, y1, z1), isFilled2 = g.filled(x2, y2, z2);
	const float v1 = isFilled1 ? -1.0f : 1.0f, v2 = isFilled2 ? -1.0f : 1.0f;
	const V3 p1 = g.toWorld(x1, y1, z1), p2 = g.toWorld(x2, y2, z2);
	float t = v1 / (v1 - v2);
	t = minf(maxf(t, 0.0f), 1.0f);
	Hermite hp;
	hp.p = p1 + t * (p2 - p1);
#define RTO_DC_SCALAR(X, Y, Z) (!g.inb((X), (Y), (Z)) ? 1.0f : (g.filled((X), (Y), (Z)) ? -1.0f : 1.0f))
	V3 normal;
	if (dx != 0) {
		float gy = RTO_DC_SCALAR(x1, y1 + 1, z1) - RTO_DC_SCALAR(x1, y1 - 1, z1);
		float gz = RTO_DC_SCALAR(x1, y1, z1 + 1) - RTO_DC_SCALAR(x1, y1, z1 - 1);
		normal = mk3(0.0f, gy, gz);
	}
	else if (dy != 0) {
		float gx = RTO_DC_SCALAR(x1 + 1, y1, z1) - RTO_DC_SCALAR(x1 - 1, y1, z1);
		float gz = RTO_DC_SCALAR(x1, y1, z1 + 1) - RTO_DC_SCALAR(x1, y1, z1 - 1);
		normal = mk3(gx, 0.0f, gz);
	}
	else {
		float gx = RTO_DC_SCALAR(x1 + 1, y1, z1) - RTO_DC_SCALAR(x1 - 1, y1, z1);
		float gy = RTO_DC_SCALAR(x1, y1 + 1, z1) - RTO_DC_SCALAR(x1, y1 - 1, z1);
		normal = mk3(gx, gy, 0.0f);
	}
#undef RTO_DC_SCALAR
	if ((double)dot3(normal, normal) < 1e-10) normal = mk3(float(dx), float(dy), float(dz));
	else normal = normalize3(normal);
	float dotProduct = normal.x * float(dx) + normal.y * float(dy) + normal.z * float(dz);
	bool normalPointsWithEdge = dotProduct > 0;
	bool edgePointsToFilled = isFilled2;
	if (normalPointsWithEdge == edgePointsToFilled) normal = -normal;
	hp.n = normal;
	return hp;
}

// The uniform leaf a gather starts in: every voxel of [x0, x0 + size)^3 has one state (outside the grid counts as EMPTY, and a leaf
// that sticks out of the grid is EMPTY throughout), so a sample whose +1 neighbours all lie inside it sees no sign change.
struct UniformBox { int x0, y0, z0, size; };

// gatherHermiteData (:1090-1144): samples [x0, min(x0 + size, dim - 1)] inclusive on every axis with stride 2 for size > 8, and
// hands every Hermite point to acc in the reference's order (z, y, x, direction).  Samples that cannot see a sign change because
// they and their +1 neighbours lie inside the uniform leaf `u` are stepped over (u.size == 0: no such knowledge).
template <class Acc> RTO_HD void forEachHermite(const Grid& g, int x0, int y0, int z0, int size, UniformBox u, Acc& acc) {
	const int maxX = imin(x0 + size, g.dx - 1), maxY = imin(y0 + size, g.dy - 1), maxZ = imin(z0 + size, g.dz - 1);
	const int minX = imax(x0, 0), minY = imax(y0, 0), minZ = imax(z0, 0);
	int stride = (size > 8) ? 2 : 1;
	if (size <= 4) stride = 1;
	const int ux1 = u.x0 + u.size, uy1 = u.y0 + u.size, uz1 = u.z0 + u.size;
	for (int z = minZ; z <= maxZ; z += stride)
		for (int y = minY; y <= maxY; y += stride) {
			const bool rowInside = u.size > 0 && y >= u.y0 && y + 1 < uy1 && z >= u.z0 && z + 1 < uz1;
			int x = minX;
			while (x <= maxX) {
				if (rowInside && x >= u.x0 && x + 1 < ux1) { x += ((ux1 - 1 - x + stride - 1) / stride) * stride; continue; }
				const bool currentFilled = g.filled(x, y, z);
				for (int d = 0; d < 3; d++) {
					const int nx = x + (d == 0), ny = y + (d == 1), nz = z + (d == 2);
					if (nx >= g.dx || ny >= g.dy || nz >= g.dz) continue;
					if (currentFilled != g.filled(nx, ny, nz)) acc(intersection(g, x, y, z, d));
				}
				x += stride;
			}
		}
}

// QEFSolver (:46-160)
struct Qef {
	float ata[3][3]; V3 atb, pointSum; int numPoints;
	RTO_HD void clear() { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) ata[i][j] = 0.0f; atb = mk3(0, 0, 0); pointSum = mk3(0, 0, 0); numPoints = 0; }
	RTO_HD void addPoint(V3 point, V3 normal) {
		V3 n = normalize3(normal);
		ata[0][0] += n.x * n.x; ata[0][1] += n.x * n.y; ata[0][2] += n.x * n.z;
		ata[1][0] += n.y * n.x; ata[1][1] += n.y * n.y; ata[1][2] += n.y * n.z;
		ata[2][0] += n.z * n.x; ata[2][1] += n.z * n.y; ata[2][2] += n.z * n.z;
		float d = -dot3(n, point);
		atb.x += n.x * d; atb.y += n.y * d; atb.z += n.z * d;
		pointSum = pointSum + point;
		numPoints++;
	}
	RTO_HD V3 solve(V3 cellCenter, float cellSize) const {
		V3 masspoint = (numPoints > 0) ? pointSum / float(numPoints) : cellCenter;
		if (numPoints <= 2) return masspoint;
		float m[3][3];
		for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = ata[i][j];
		const float reg = 0.3f;
		m[0][0] += reg; m[1][1] += reg; m[2][2] += reg;
		bool invertible = true;
		float inv[3][3];
		// glm::determinant(mat3), func_matrix.inl:211-220
		float det = m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2]) + m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]);
		if ((double)fabsf(det) < 1e-10) invertible = false;
		else {
			// glm::inverse(mat3), func_matrix.inl:269-291
			float ood = 1.0f / (m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2]) + m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]));
			inv[0][0] = +(m[1][1] * m[2][2] - m[2][1] * m[1][2]) * ood;
			inv[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]) * ood;
			inv[2][0] = +(m[1][0] * m[2][1] - m[2][0] * m[1][1]) * ood;
			inv[0][1] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]) * ood;
			inv[1][1] = +(m[0][0] * m[2][2] - m[2][0] * m[0][2]) * ood;
			inv[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]) * ood;
			inv[0][2] = +(m[0][1] * m[1][2] - m[1][1] * m[0][2]) * ood;
			inv[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]) * ood;
			inv[2][2] = +(m[0][0] * m[1][1] - m[1][0] * m[0][1]) * ood;
			for (int i = 0; i < 3 && invertible; i++)
				for (int j = 0; j < 3 && invertible; j++)
					if (isNan(inv[i][j]) || isInf(inv[i][j]) || (double)fabsf(inv[i][j]) > 1e6) invertible = false;
		}
		if (invertible) {
			// mat3 * vec3, type_mat3x3.inl:468-474
			V3 solution = mk3(inv[0][0] * atb.x + inv[1][0] * atb.y + inv[2][0] * atb.z,
				inv[0][1] * atb.x + inv[1][1] * atb.y + inv[2][1] * atb.z,
				inv[0][2] * atb.x + inv[1][2] * atb.y + inv[2][2] * atb.z);
			const float relaxation = 0.7f;
			solution = masspoint + relaxation * (solution - masspoint);
			if (!isNan(solution.x) && !isNan(solution.y) && !isNan(solution.z)) {
				V3 dlt = masspoint - solution;                       // distance2(p0, p1) = length2(p1 - p0), gtx/norm.inl:39-44
				float distSq = dot3(dlt, dlt);
				const float MAX_DIST_SQ = cellSize * cellSize;
				if (distSq < MAX_DIST_SQ) return mix3(solution, masspoint, 0.2f);
			}
		}
		return masspoint;
	}
	RTO_HD V3 solveConstrained(V3 minBound, V3 maxBound) const {
		V3 cellCenter = (minBound + maxBound) * 0.5f;
		float cellSize = maxBound.x - minBound.x;
		V3 solution = solve(cellCenter, cellSize);
		return clamp3(solution, minBound, maxBound);
	}
};

// The three sweeps generateDualVertex (:1146-1234) makes over the Hermite points: sums, points near the snapped plane, the QEF.
// The reference stores the points in a vector; a sweep re-derives them instead, in the same order, so no storage is needed.
struct SumAcc { V3 mass, normal; int n; RTO_HD void operator()(const Hermite& hp) { mass = mass + hp.p; normal = normal + hp.n; n++; } };
struct PlaneAcc { V3 avgNormal, planePoint; int count; RTO_HD void operator()(const Hermite& hp) { float alignment = dot3(normalize3(hp.n), avgNormal); if (alignment > 0.7f) { planePoint = planePoint + hp.p; count++; } } };
struct QefAcc { Qef q; RTO_HD void operator()(const Hermite& hp) { q.addPoint(hp.p, hp.n); } };

// The dual vertex createTriangles gives the cell at (x, y, z) when it computes it with cell size `size` (:571-577 for the visited
// leaf itself, :706-718 for a neighbour, where `size` is the VISITING leaf's): the cell centre if the region holds no Hermite
// point, else generateDualVertex.  u: the uniform leaf that starts at (x, y, z).
RTO_HD V3 dualVertex(const Grid& g, int x, int y, int z, int size, UniformBox u) {
	const V3 cellCenter = g.centre(x, y, z, size);
	const float cellSize = float(size) * g.vs;
	SumAcc s; s.mass = mk3(0.0f, 0.0f, 0.0f); s.normal = mk3(0.0f, 0.0f, 0.0f); s.n = 0;
	forEachHermite(g, x, y, z, size, u, s);
	if (s.n == 0) return cellCenter;
	float hs = cellSize * 0.5f;
	V3 halfSize = mk3(hs, hs, hs);
	V3 minBound = cellCenter - halfSize, maxBound = cellCenter + halfSize;
	float inset = cellSize * 0.001f;
	minBound = minBound + mk3(inset, inset, inset);
	maxBound = maxBound - mk3(inset, inset, inset);
	V3 massPoint = s.mass / float(s.n);
	V3 avgNormal = s.normal;
	if (sqrtf(dot3(avgNormal, avgNormal)) > 0.0001f) {
		avgNormal = normalize3(avgNormal);
		V3 absNormal = mk3(fabsf(avgNormal.x), fabsf(avgNormal.y), fabsf(avgNormal.z));
		float maxComp = maxf(maxf(absNormal.x, absNormal.y), absNormal.z);
		if (maxComp > 0.85f) {
			if (absNormal.x == maxComp) avgNormal = mk3(avgNormal.x > 0 ? 1.0f : -1.0f, 0.0f, 0.0f);
			else if (absNormal.y == maxComp) avgNormal = mk3(0.0f, avgNormal.y > 0 ? 1.0f : -1.0f, 0.0f);
			else avgNormal = mk3(0.0f, 0.0f, avgNormal.z > 0 ? 1.0f : -1.0f);
			PlaneAcc p; p.avgNormal = avgNormal; p.planePoint = mk3(0.0f, 0.0f, 0.0f); p.count = 0;
			forEachHermite(g, x, y, z, size, u, p);
			if (p.count > 0) {
				V3 planePoint = p.planePoint / float(p.count);
				float d = -dot3(avgNormal, planePoint);
				float t = -(dot3(avgNormal, cellCenter) + d);
				V3 projectedVertex = cellCenter + t * avgNormal;
				return clamp3(projectedVertex, minBound, maxBound);
			}
		}
	}
	QefAcc q; q.q.clear();
	forEachHermite(g, x, y, z, size, u, q);
	V3 qefSolution = q.q.solveConstrained(minBound, maxBound);
	return mix3(qefSolution, massPoint, 0.1f);
}

// cellContainsSurface (:1367-1530)
RTO_HD bool cellContainsSurface(const Grid& g, int x0, int y0, int z0, int size) {
	const int maxX = imin(x0 + size, g.dx), maxY = imin(y0 + size, g.dy), maxZ = imin(z0 + size, g.dz);
	const int minX = imax(x0, 0), minY = imax(y0, 0), minZ = imax(z0, 0);
	if (minX >= maxX || minY >= maxY || minZ >= maxZ) return false;
	bool anyFilled = false, anyEmpty = false;
	for (int i = 0; i < 8; i++) {      // corner order of :1389-1392
		const int cx = (i == 1 || i == 2 || i == 5 || i == 6) ? maxX - 1 : minX, cy = (i == 2 || i == 3 || i == 6 || i == 7) ? maxY - 1 : minY, cz = (i >= 4) ? maxZ - 1 : minZ;
		if (g.filled(cx, cy, cz)) anyFilled = true; else anyEmpty = true;
		if (anyFilled && anyEmpty) return true;
	}
	for (int dir = 0; dir < 3; dir++) {
		const int stride = imax(1, size / 4);
		for (int offset = 0; offset < size; offset += stride) {
			if (dir == 0) {
				const int y1 = minY + offset, z1 = minZ + offset;
				if (y1 >= maxY || z1 >= maxZ) continue;
				int x1 = minX - 1, x2 = minX;
				if (x1 >= 0 && x2 < g.dx && g.filled(x1, y1, z1) != g.filled(x2, y1, z1)) return true;
				x1 = maxX - 1; x2 = maxX;
				if (x1 >= 0 && x2 < g.dx && g.filled(x1, y1, z1) != g.filled(x2, y1, z1)) return true;
			}
			else if (dir == 1) {
				const int x1 = minX + offset, z1 = minZ + offset;
				if (x1 >= maxX || z1 >= maxZ) continue;
				int y1 = minY - 1, y2 = minY;
				if (y1 >= 0 && y2 < g.dy && g.filled(x1, y1, z1) != g.filled(x1, y2, z1)) return true;
				y1 = maxY - 1; y2 = maxY;
				if (y1 >= 0 && y2 < g.dy && g.filled(x1, y1, z1) != g.filled(x1, y2, z1)) return true;
			}
			else {
				const int x1 = minX + offset, y1 = minY + offset;
				if (x1 >= maxX || y1 >= maxY) continue;
				int z1 = minZ - 1, z2 = minZ;
				if (z1 >= 0 && z2 < g.dz && g.filled(x1, y1, z1) != g.filled(x1, y1, z2)) return true;
				z1 = maxZ - 1; z2 = maxZ;
				if (z1 >= 0 && z2 < g.dz && g.filled(x1, y1, z1) != g.filled(x1, y1, z2)) return true;
			}
		}
	}
	if (size <= 4) {
		for (int z = minZ; z < maxZ - 1; z++)
			for (int y = minY; y < maxY - 1; y++)
				for (int x = minX; x < maxX - 1; x++) {
					const bool s = g.filled(x, y, z);
					if (s != g.filled(x + 1, y, z) || s != g.filled(x, y + 1, z) || s != g.filled(x, y, z + 1)) return true;
				}
	}
	return false;
}

// What g_octreeMap answers for a key: buildOctreeRec (OctreeVoxel.cpp:704-762) stores every node under its origin, a child 0 after its
// parent, so the entry that survives under an origin is the leaf that starts there.  -1: no node starts at (x, y, z).
RTO_HD int32_t leafAtOrigin(const RtoGpuNode* nodes, int x, int y, int z) {
	int32_t i = 0;
	for (;;) {
		const RtoGpuNode& n = nodes[i];
		if (n.isLeaf) return (n.x == x && n.y == y && n.z == z) ? i : -1;
		const int half = n.size / 2;
		const int ci = (x >= n.x + half ? 1 : 0) | (y >= n.y + half ? 2 : 0) | (z >= n.z + half ? 4 : 0);
		i = n.child[ci];
		if (i < 0) return -1;
	}
}

// bit dir * 4 + edge: both end voxels of that cell edge are inside the grid and differ (:590-612)
RTO_HD uint32_t edgeSignMask(const Grid& g, int x0, int y0, int z0, int size) {
	uint32_t mask = 0;
	for (int dir = 0; dir < 3; dir++)
		for (int edge = 0; edge < 4; edge++) {
			int ex1 = x0, ey1 = y0, ez1 = z0;
			if (dir == 0) { ey1 += (edge & 1) ? size : 0; ez1 += (edge & 2) ? size : 0; }
			else if (dir == 1) { ex1 += (edge & 1) ? size : 0; ez1 += (edge & 2) ? size : 0; }
			else { ex1 += (edge & 1) ? size : 0; ey1 += (edge & 2) ? size : 0; }
			const int ex2 = ex1 + (dir == 0 ? size : 0), ey2 = ey1 + (dir == 1 ? size : 0), ez2 = ez1 + (dir == 2 ? size : 0);
			if (!g.inb(ex1, ey1, ez1) || !g.inb(ex2, ey2, ez2)) continue;
			if (g.filled(ex1, ey1, ez1) == g.filled(ex2, ey2, ez2)) continue;
			mask |= 1u << (dir * 4 + edge);
		}
	return mask;
}

// offset bits of neighbour adjIdx (1..3) of edge `edge` in direction `dir` (:630-645): bit 0 / 1 / 2 = the neighbour starts one
// cell size lower in x / y / z.  Its coordinate on an axis is e1 - size: origin - size where the edge sits on the low side of
// the cell, the origin itself where it sits on the high side.
RTO_HD int adjOffsetBits(int dir, int edge, int adjIdx) {
	const int a = (dir == 0) ? 1 : 0, b = (dir == 2) ? 1 : 2;          // the two axes the edge index steps along: bit 0 -> a, bit 1 -> b
	const bool lowA = !(edge & 1), lowB = !(edge & 2);
	int bits = 0;
	if ((adjIdx == 1 || adjIdx == 3) && lowA) bits |= 1 << a;
	if ((adjIdx == 2 || adjIdx == 3) && lowB) bits |= 1 << b;
	return bits;
}
// the offsets the flagged edges of a cell ask for (bit o set: offset o)
RTO_HD uint32_t askedOffsets(uint32_t edgeMask) {
	uint32_t asked = 0;
	for (int dir = 0; dir < 3; dir++) for (int edge = 0; edge < 4; edge++) if (edgeMask & (1u << (dir * 4 + edge)))
		for (int adjIdx = 1; adjIdx < 4; adjIdx++) asked |= 1u << adjOffsetBits(dir, edge, adjIdx);
	return asked;
}
// the leaf a cell may join at offset o (:646-687): inside the grid, a leaf starts exactly there, sizes within a factor of two
RTO_HD int32_t joinTarget(const Grid& g, const RtoGpuNode* nodes, int x0, int y0, int z0, int size, int o) {
	const int ax = x0 - ((o & 1) ? size : 0), ay = y0 - ((o & 2) ? size : 0), az = z0 - ((o & 4) ? size : 0);
	if (!g.inb(ax, ay, az)) return -1;
	const int32_t k = leafAtOrigin(nodes, ax, ay, az);
	if (k < 0) return -1;
	const int adjSize = nodes[k].size;
	if (imax(size, adjSize) > imin(size, adjSize) * 2) return -1;      // MAX_SIZE_DIFFERENCE, :681-685
	return k;
}

RTO_HD void putTri(RtoTriangle* t, V3 a, V3 b, V3 c) {
	t->v0[0] = a.x; t->v0[1] = a.y; t->v0[2] = a.z; t->v1[0] = b.x; t->v1[1] = b.y; t->v1[2] = b.z; t->v2[0] = c.x; t->v2[1] = c.y; t->v2[2] = c.z;
}
// a triangle is kept if 0.5 * |cross(e1, e2)| > 1e-6 (:735-738)
RTO_HD bool hasArea(V3 a, V3 b, V3 c) {
	V3 e1 = b - a, e2 = c - a;
	V3 cr = cross3(e1, e2);
	float area = 0.5f * sqrtf(dot3(cr, cr));
	return area > 1e-6f;
}

// The triangles createTriangles emits for the 12 edges of one leaf (:586-787) given the dual vertices its cache look-ups return:
// vtx[0] the leaf's own, vtx[o] the joinable neighbour's at offset o (tgt[o] >= 0).  out == nullptr: count only.
// nrm (may be null, only with out): the flat normal the reference stores three times per MCTriangle, normalize(cross(e1, e2)), negated
// when the emitting leaf is solid (:730-734).
RTO_HD V3 flatNormal(V3 a, V3 b, V3 c, bool flip) { V3 n = normalize3(cross3(b - a, c - a)); return flip ? -n : n; }
RTO_HD int edgeTriangles(uint32_t edgeMask, const int32_t tgt[8], const V3 vtx[8], RtoTriangle* out, V3* nrm = nullptr, bool flip = false) {
	int n = 0;
	for (int dir = 0; dir < 3; dir++)
		for (int edge = 0; edge < 4; edge++) {
			if (!(edgeMask & (1u << (dir * 4 + edge)))) continue;
			V3 adj[4]; int cnt = 0;
			adj[cnt++] = vtx[0];
			for (int adjIdx = 1; adjIdx < 4; adjIdx++) {
				const int o = adjOffsetBits(dir, edge, adjIdx);
				if (tgt[o] < 0) continue;
				adj[cnt++] = vtx[o];
			}
			if (cnt >= 3) {
				if (hasArea(adj[0], adj[1], adj[2])) { if (out) putTri(out + n, adj[0], adj[1], adj[2]); if (nrm) nrm[n] = flatNormal(adj[0], adj[1], adj[2], flip); n++; }
				if (cnt >= 4 && hasArea(adj[0], adj[2], adj[3])) { if (out) putTri(out + n, adj[0], adj[2], adj[3]); if (nrm) nrm[n] = flatNormal(adj[0], adj[2], adj[3], flip); n++; }
			}
		}
	return n;
}

RTO_HD bool touchesBoundary(const Grid& g, int x0, int y0, int z0, int size) {      // :791-794
	return x0 == 0 || y0 == 0 || z0 == 0 || (x0 + size) >= g.dx || (y0 + size) >= g.dy || (z0 + size) >= g.dz;
}

// One face of createFaceTriangles (:843-878): does the fallback of the leaf `n` produce fans through face `face`, and towards which
// leaf (k >= 0) or loose position (k < 0)?
RTO_HD bool fallbackFace(const Grid& g, const RtoGpuNode* nodes, const RtoGpuNode& n, int face, int& nx, int& ny, int& nz, int32_t& k) {
	const int fx = face == 0 ? 1 : (face == 1 ? -1 : 0), fy = face == 2 ? 1 : (face == 3 ? -1 : 0), fz = face == 4 ? 1 : (face == 5 ? -1 : 0);
	const int size = n.size;
	nx = n.x + fx * size; ny = n.y + fy * size; nz = n.z + fz * size;
	if (!g.inb(nx, ny, nz)) return false;
	const bool currentSolid = n.isSolid != 0;
	bool neighborSolid;
	k = leafAtOrigin(nodes, nx, ny, nz);
	if (k >= 0) {
		const int adjSize = nodes[k].size;
		if (imax(size, adjSize) > imin(size, adjSize) * 2) return false;
		neighborSolid = nodes[k].isSolid != 0;
	}
	else {
		int cx = nx + size / 2, cy = ny + size / 2, cz = nz + size / 2;
		cx = imin(imax(cx, 0), g.dx - 1); cy = imin(imax(cy, 0), g.dy - 1); cz = imin(imax(cz, 0), g.dz - 1);
		neighborSolid = g.filled(cx, cy, cz);
	}
	return currentSolid != neighborSolid;
}

// The 32 triangles of one face of createFaceTriangles (:906-1084): two bulged fans of 16 over a 3 x 3 point grid
// nrm (may be null): the face normal, pointing away from the solid side for the cell's fan and the other way for the neighbour's (:936-940)
RTO_HD void faceFan(V3 cellVertex, V3 neighborVertex, int face, int size, float vs, RtoTriangle* out, V3* nrm = nullptr, bool currentSolid = false) {
	const int fx = face == 0 ? 1 : (face == 1 ? -1 : 0), fy = face == 2 ? 1 : (face == 3 ? -1 : 0), fz = face == 4 ? 1 : (face == 5 ? -1 : 0);
	const float halfSize = float(size) * vs * 0.5f;
	const V3 faceNormal = mk3(float(fx), float(fy), float(fz));
	const V3 faceCenter = (cellVertex + neighborVertex) * 0.5f;
	V3 tangent1, tangent2;
	if (fabsf(faceNormal.x) > 0.5f) { tangent1 = mk3(0, 1, 0); tangent2 = mk3(0, 0, 1); }
	else if (fabsf(faceNormal.y) > 0.5f) { tangent1 = mk3(1, 0, 0); tangent2 = mk3(0, 0, 1); }
	else { tangent1 = mk3(1, 0, 0); tangent2 = mk3(0, 1, 0); }
	const int divisions = 2;
	V3 gridPoints[9];
	int k = 0;
	for (int i = 0; i <= divisions; i++) {
		float u = float(i) / float(divisions);
		for (int j = 0; j <= divisions; j++) {
			float v = float(j) / float(divisions);
			float mappedU = 2.0f * u - 1.0f, mappedV = 2.0f * v - 1.0f;
			V3 point = faceCenter + tangent1 * (mappedU * halfSize) + tangent2 * (mappedV * halfSize);
			float distFromCenter = sqrtf(mappedU * mappedU + mappedV * mappedV);      // glm::length(vec2)
			float bulge = 0.05f * halfSize * (1.0f - distFromCenter * distFromCenter);
			point = point + faceNormal * bulge;
			gridPoints[k++] = point;
		}
	}
	int t = 0;
	for (int i = 0; i < divisions; i++)
		for (int j = 0; j < divisions; j++) {
			const int idx00 = i * (divisions + 1) + j, idx10 = (i + 1) * (divisions + 1) + j, idx01 = i * (divisions + 1) + (j + 1), idx11 = (i + 1) * (divisions + 1) + (j + 1);
			putTri(out + t++, cellVertex, gridPoints[idx00], gridPoints[idx10]);
			putTri(out + t++, cellVertex, gridPoints[idx10], gridPoints[idx11]);
			putTri(out + t++, cellVertex, gridPoints[idx11], gridPoints[idx01]);
			putTri(out + t++, cellVertex, gridPoints[idx01], gridPoints[idx00]);
		}
	for (int i = 0; i < divisions; i++)
		for (int j = 0; j < divisions; j++) {
			const int idx00 = i * (divisions + 1) + j, idx10 = (i + 1) * (divisions + 1) + j, idx01 = i * (divisions + 1) + (j + 1), idx11 = (i + 1) * (divisions + 1) + (j + 1);
			putTri(out + t++, neighborVertex, gridPoints[idx10], gridPoints[idx00]);
			putTri(out + t++, neighborVertex, gridPoints[idx11], gridPoints[idx10]);
			putTri(out + t++, neighborVertex, gridPoints[idx01], gridPoints[idx11]);
			putTri(out + t++, neighborVertex, gridPoints[idx00], gridPoints[idx01]);
		}
	if (nrm) {
		V3 normal = faceNormal;
		if (!currentSolid) normal = -normal;
		for (int i = 0; i < 16; i++) { nrm[i] = normal; nrm[16 + i] = -normal; }
	}
}

// ---- the order of renderOctree's walk, as a number ------------------------------------------------------------------------
// Leaves are visited depth first with children 0..7 (bit 0 = x, 1 = y, 2 = z): leaf A comes before leaf B iff the Morton code of
// A's origin (x in the lowest bit) is smaller.  A touch of a cache key is ordered by (visiting leaf, kind): a leaf looks up its own
// vertex first (0), then its neighbours' while it walks its edges (1), then, in the fallback, its face neighbours' (2).
RTO_HD uint32_t spread10(uint32_t v) { v &= 0x3ff; v = (v | (v << 16)) & 0x030000ff; v = (v | (v << 8)) & 0x0300f00f; v = (v | (v << 4)) & 0x030c30c3; v = (v | (v << 2)) & 0x09249249; return v; }
RTO_HD uint32_t morton30(int x, int y, int z) { return spread10((uint32_t)x) | (spread10((uint32_t)y) << 1) | (spread10((uint32_t)z) << 2); }
RTO_HD int log2i(int v) { int l = 0; while ((1 << l) < v) l++; return l; }
constexpr unsigned long long kNoTouch = ~0ull;
// touch key: Morton code of the visiting leaf << 8 | kind << 4 | log2(size of the visiting leaf)
RTO_HD unsigned long long touchKey(const RtoGpuNode& visitor, int kind) { return ((unsigned long long)morton30(visitor.x, visitor.y, visitor.z) << 8) | ((unsigned long long)kind << 4) | (unsigned long long)log2i(visitor.size); }
RTO_HD int touchKind(unsigned long long key) { return (int)((key >> 4) & 3); }
RTO_HD int touchSize(unsigned long long key) { return 1 << (int)(key & 15); }

// The value a cache key holds, from the touch that came first: the visiting leaf's size decides the region and the cell size
RTO_HD V3 vertexFromTouch(const Grid& g, const RtoGpuNode& leaf, unsigned long long key) {
	const int s = touchSize(key);
	if (touchKind(key) == 2) return g.centre(leaf.x, leaf.y, leaf.z, s);                 // createFaceTriangles caches the bare centre (:899-906)
	UniformBox u; u.x0 = leaf.x; u.y0 = leaf.y; u.z0 = leaf.z; u.size = leaf.size;
	return dualVertex(g, leaf.x, leaf.y, leaf.z, s, u);
}

// is the leaf reached by renderOctree's walk (main.cpp:152-187)?  Every node on the way down must pass the frustum test.
RTO_HD bool leafReached(const RtoGpuNode* nodes, const RtoGpuNode& leaf, const FrustumPlanes& F, const float gridMin[3], float voxel, float margin) {
	int32_t i = 0;
	for (;;) {
		const RtoGpuNode& n = nodes[i];
		if (!frustum_node_visible(F, n, gridMin, voxel, margin)) return false;
		if (n.isLeaf) return true;
		const int half = n.size / 2;
		const int ci = (leaf.x >= n.x + half ? 1 : 0) | (leaf.y >= n.y + half ? 2 : 0) | (leaf.z >= n.z + half ? 4 : 0);
		i = n.child[ci];
		if (i < 0) return false;
	}
}

} // namespace dc
} // namespace rto
