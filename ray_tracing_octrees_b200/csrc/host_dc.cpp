// host_dc.cpp -- the Adaptive Dual Contouring mesh of the reference, host side (part of librto.so).
//
//   rto_host_dc_mesh == renderOctree (main.cpp:95-208) over AdaptiveDualContouringRenderer::render -> createTriangles
//                       (AdaptiveDualContouringRenderer.cpp:489-803, 805-1088, 1090-1357, 1367-1530, QEFSolver :46-160):
//                       the triangle soup of configuration C4, same triangles in the same order, bit for bit.
//
// What the reference does.  renderOctree walks the octree depth first (children 0..7), drops subtrees outside the frustum and calls
// createTriangles on every leaf, one after the other on one thread.  A leaf that "contains surface" (a sampling heuristic) gets a
// dual vertex (axis-snapped plane projection or a regularised QEF over the Hermite points of its region); for each of its 12 edges
// whose end voxels differ it collects the dual vertices of up to three neighbour leaves (looked up by origin in g_octreeMap) and
// emits 1-2 triangles whose area exceeds 1e-6; a boundary leaf that produced nothing falls back to bulged face fans.
//
// The order dependence.  Dual vertices go through dualVertexCache: a neighbour's vertex that is not cached yet is computed by the
// VISITING leaf with the visiting leaf's size and cached under the neighbour's key, so the value a key holds is the one its first
// toucher gave it.  That makes the result a function of the visit order, and it is why the reference's own thread pool is not on
// this path.  Here the order is kept and the work is not:
//   * everything expensive is a pure function of (cell origin, size): cellContainsSurface, the edge sign tests, the neighbour
//     look-ups, and the dual vertex F(origin, size) itself.  These run on all host threads for a block of leaves at a time.
//     A neighbour at offset {-s, 0}^3 always precedes the visiting leaf in depth-first order, so a neighbour that contains surface
//     has cached its own vertex by then; F(neighbour, visiting size) is needed only for the others.
//   * a sequential pass then replays the cache protocol over the precomputed records (look-ups, area tests, emission, the rare
//     face fallback), computing on the spot whatever the parallel pass did not foresee.  Correctness never depends on the
//     speculation, only speed does.
//
// Arithmetic: one IEEE operation per operator in glm 0.9.9.7's order (rto_math.h; mat3 inverse/determinant: func_matrix.inl:211-291,
// mix: func_common.inl:104-112, clamp = min(max(x, lo), hi)); compiled with -ffp-contract=off.
// Limits: the reference's keys are (x << 20 | y << 10 | z), which alias beyond 1024 voxels per axis; larger octrees are refused.
#include "rto_internal.h"
#include "rto_frustum.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

using namespace rto;

namespace {

struct DcGrid {
	const uint8_t* v; int dx, dy, dz; float mnx, mny, mnz, vs;
	inline bool inb(int x, int y, int z) const { return x >= 0 && y >= 0 && z >= 0 && x < dx && y < dy && z < dz; }
	inline bool filled(int x, int y, int z) const { return v[(size_t)x + (size_t)y * dx + (size_t)z * ((size_t)dx * dy)] == 1; }
	// gridToWorld, AdaptiveDualContouringRenderer.cpp:1359-1365
	inline V3 toWorld(int x, int y, int z) const { return mk3(mnx + float(x) * vs, mny + float(y) * vs, mnz + float(z) * vs); }
	// cell centre as createTriangles forms it (:549-550): gridToWorld + vec3(size * 0.5f * voxelSize)
	inline V3 centre(int x, int y, int z, int size) const { float h = float(size) * 0.5f * vs; return toWorld(x, y, z) + mk3(h, h, h); }
};

struct Hermite { V3 p, n; };

inline V3 clamp3(V3 x, V3 lo, V3 hi) { return min3(max3(x, lo), hi); }
inline V3 mix3(V3 x, V3 y, float a) { return x * (1.0f - a) + y * a; }

// calculateIntersection (:1236-1357) for the only case gatherHermiteData produces: (x2, y2, z2) = (x1, y1, z1) + unit axis d, end
// voxels of different state.  The edge cache of the reference holds pure values and is not needed.
Hermite intersection(const DcGrid& g, int x1, int y1, int z1, int d) {
	const int dx = d == 0, dy = d == 1, dz = d == 2;
	const int x2 = x1 + dx, y2 = y1 + dy, z2 = z1 + dz;
	const bool isFilled1 = g.filled(x1, y1, z1), isFilled2 = g.filled(x2, y2, z2);
	const float v1 = isFilled1 ? -1.0f : 1.0f, v2 = isFilled2 ? -1.0f : 1.0f;
	const V3 p1 = g.toWorld(x1, y1, z1), p2 = g.toWorld(x2, y2, z2);
	float t = v1 / (v1 - v2);
	t = minf(maxf(t, 0.0f), 1.0f);
	Hermite hp;
	hp.p = p1 + t * (p2 - p1);
	auto getScalar = [&](int x, int y, int z) -> float {
		if (!g.inb(x, y, z)) return 1.0f;
		return g.filled(x, y, z) ? -1.0f : 1.0f;
	};
	V3 normal;
	if (dx != 0) {
		float gy = getScalar(x1, y1 + 1, z1) - getScalar(x1, y1 - 1, z1);
		float gz = getScalar(x1, y1, z1 + 1) - getScalar(x1, y1, z1 - 1);
		normal = mk3(0.0f, gy, gz);
	}
	else if (dy != 0) {
		float gx = getScalar(x1 + 1, y1, z1) - getScalar(x1 - 1, y1, z1);
		float gz = getScalar(x1, y1, z1 + 1) - getScalar(x1, y1, z1 - 1);
		normal = mk3(gx, 0.0f, gz);
	}
	else {
		float gx = getScalar(x1 + 1, y1, z1) - getScalar(x1 - 1, y1, z1);
		float gy = getScalar(x1, y1 + 1, z1) - getScalar(x1, y1 - 1, z1);
		normal = mk3(gx, gy, 0.0f);
	}
	if ((double)dot3(normal, normal) < 1e-10) normal = mk3(float(dx), float(dy), float(dz));
	else normal = normalize3(normal);
	float dotProduct = normal.x * float(dx) + normal.y * float(dy) + normal.z * float(dz);
	bool normalPointsWithEdge = dotProduct > 0;
	bool edgePointsToFilled = isFilled2;
	if (normalPointsWithEdge == edgePointsToFilled) normal = -normal;
	hp.n = normal;
	return hp;
}

// gatherHermiteData (:1090-1144): region [x0, min(x0 + size, dim - 1)] inclusive on every axis, stride 2 for size > 8
void gatherHermite(const DcGrid& g, int x0, int y0, int z0, int size, std::vector<Hermite>& points) {
	points.clear();
	const int maxX = std::min(x0 + size, g.dx - 1), maxY = std::min(y0 + size, g.dy - 1), maxZ = std::min(z0 + size, g.dz - 1);
	const int minX = std::max(x0, 0), minY = std::max(y0, 0), minZ = std::max(z0, 0);
	int stride = (size > 8) ? 2 : 1;
	if (size <= 4) stride = 1;
	for (int z = minZ; z <= maxZ; z += stride)
		for (int y = minY; y <= maxY; y += stride)
			for (int x = minX; x <= maxX; x += stride) {
				const bool currentFilled = g.filled(x, y, z);
				for (int d = 0; d < 3; d++) {
					const int nx = x + (d == 0), ny = y + (d == 1), nz = z + (d == 2);
					if (nx >= g.dx || ny >= g.dy || nz >= g.dz) continue;
					if (currentFilled != g.filled(nx, ny, nz)) points.push_back(intersection(g, x, y, z, d));
				}
			}
}

// QEFSolver (:46-160)
struct Qef {
	float ata[3][3]; V3 atb, pointSum; int numPoints;
	Qef() { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) ata[i][j] = 0.0f; atb = mk3(0, 0, 0); pointSum = mk3(0, 0, 0); numPoints = 0; }
	void addPoint(V3 point, V3 normal) {
		V3 n = normalize3(normal);
		ata[0][0] += n.x * n.x; ata[0][1] += n.x * n.y; ata[0][2] += n.x * n.z;
		ata[1][0] += n.y * n.x; ata[1][1] += n.y * n.y; ata[1][2] += n.y * n.z;
		ata[2][0] += n.z * n.x; ata[2][1] += n.z * n.y; ata[2][2] += n.z * n.z;
		float d = -dot3(n, point);
		atb.x += n.x * d; atb.y += n.y * d; atb.z += n.z * d;
		pointSum = pointSum + point;
		numPoints++;
	}
	V3 solve(V3 cellCenter, float cellSize) const {
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
		if ((double)std::fabs(det) < 1e-10) invertible = false;
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
					if (std::isnan(inv[i][j]) || std::isinf(inv[i][j]) || (double)std::fabs(inv[i][j]) > 1e6) invertible = false;
		}
		if (invertible) {
			// mat3 * vec3, type_mat3x3.inl:468-474
			V3 solution = mk3(inv[0][0] * atb.x + inv[1][0] * atb.y + inv[2][0] * atb.z,
				inv[0][1] * atb.x + inv[1][1] * atb.y + inv[2][1] * atb.z,
				inv[0][2] * atb.x + inv[1][2] * atb.y + inv[2][2] * atb.z);
			const float relaxation = 0.7f;
			solution = masspoint + relaxation * (solution - masspoint);
			if (!std::isnan(solution.x) && !std::isnan(solution.y) && !std::isnan(solution.z)) {
				V3 dlt = masspoint - solution;                       // distance2(p0, p1) = length2(p1 - p0), gtx/norm.inl:39-44
				float distSq = dot3(dlt, dlt);
				const float MAX_DIST_SQ = cellSize * cellSize;
				if (distSq < MAX_DIST_SQ) return mix3(solution, masspoint, 0.2f);
			}
		}
		return masspoint;
	}
	V3 solveConstrained(V3 minBound, V3 maxBound) const {
		V3 cellCenter = (minBound + maxBound) * 0.5f;
		float cellSize = maxBound.x - minBound.x;
		V3 solution = solve(cellCenter, cellSize);
		return clamp3(solution, minBound, maxBound);
	}
};

// generateDualVertex (:1146-1234); hermiteData is not empty
V3 generateDualVertex(const std::vector<Hermite>& hermiteData, V3 cellCenter, float cellSize) {
	float hs = cellSize * 0.5f;
	V3 halfSize = mk3(hs, hs, hs);
	V3 minBound = cellCenter - halfSize, maxBound = cellCenter + halfSize;
	float inset = cellSize * 0.001f;
	minBound = minBound + mk3(inset, inset, inset);
	maxBound = maxBound - mk3(inset, inset, inset);
	V3 massPoint = mk3(0.0f, 0.0f, 0.0f);
	for (const Hermite& hp : hermiteData) massPoint = massPoint + hp.p;
	massPoint = massPoint / float(hermiteData.size());
	V3 avgNormal = mk3(0.0f, 0.0f, 0.0f);
	for (const Hermite& hp : hermiteData) avgNormal = avgNormal + hp.n;
	if (sqrtf(dot3(avgNormal, avgNormal)) > 0.0001f) {
		avgNormal = normalize3(avgNormal);
		V3 absNormal = mk3(std::fabs(avgNormal.x), std::fabs(avgNormal.y), std::fabs(avgNormal.z));
		float maxComp = maxf(maxf(absNormal.x, absNormal.y), absNormal.z);
		if (maxComp > 0.85f) {
			if (absNormal.x == maxComp) avgNormal = mk3(avgNormal.x > 0 ? 1.0f : -1.0f, 0.0f, 0.0f);
			else if (absNormal.y == maxComp) avgNormal = mk3(0.0f, avgNormal.y > 0 ? 1.0f : -1.0f, 0.0f);
			else avgNormal = mk3(0.0f, 0.0f, avgNormal.z > 0 ? 1.0f : -1.0f);
			V3 planePoint = mk3(0.0f, 0.0f, 0.0f);
			int planePointCount = 0;
			for (const Hermite& hp : hermiteData) {
				float alignment = dot3(normalize3(hp.n), avgNormal);
				if (alignment > 0.7f) { planePoint = planePoint + hp.p; planePointCount++; }
			}
			if (planePointCount > 0) {
				planePoint = planePoint / float(planePointCount);
				float d = -dot3(avgNormal, planePoint);
				float t = -(dot3(avgNormal, cellCenter) + d);
				V3 projectedVertex = cellCenter + t * avgNormal;
				return clamp3(projectedVertex, minBound, maxBound);
			}
		}
	}
	Qef qef;
	for (const Hermite& hp : hermiteData) qef.addPoint(hp.p, hp.n);
	V3 qefSolution = qef.solveConstrained(minBound, maxBound);
	return mix3(qefSolution, massPoint, 0.1f);
}

// The dual vertex createTriangles gives a cell at (x, y, z) when it computes it with cell size `size` (:571-577 for the cell itself,
// :706-718 for a neighbour, where `size` is the VISITING cell's): a pure function of its arguments.
V3 dualVertex(const DcGrid& g, int x, int y, int z, int size, std::vector<Hermite>& scratch) {
	V3 c = g.centre(x, y, z, size);
	gatherHermite(g, x, y, z, size, scratch);
	if (scratch.empty()) return c;
	return generateDualVertex(scratch, c, float(size) * g.vs);
}

// cellContainsSurface (:1367-1530)
bool cellContainsSurface(const DcGrid& g, int x0, int y0, int z0, int size) {
	const int maxX = std::min(x0 + size, g.dx), maxY = std::min(y0 + size, g.dy), maxZ = std::min(z0 + size, g.dz);
	const int minX = std::max(x0, 0), minY = std::max(y0, 0), minZ = std::max(z0, 0);
	if (minX >= maxX || minY >= maxY || minZ >= maxZ) return false;
	bool anyFilled = false, anyEmpty = false;
	const int corners[8][3] = {
		{minX, minY, minZ}, {maxX - 1, minY, minZ}, {maxX - 1, maxY - 1, minZ}, {minX, maxY - 1, minZ},
		{minX, minY, maxZ - 1}, {maxX - 1, minY, maxZ - 1}, {maxX - 1, maxY - 1, maxZ - 1}, {minX, maxY - 1, maxZ - 1} };
	for (int i = 0; i < 8; i++) {
		if (!g.inb(corners[i][0], corners[i][1], corners[i][2])) continue;
		if (g.filled(corners[i][0], corners[i][1], corners[i][2])) anyFilled = true; else anyEmpty = true;
		if (anyFilled && anyEmpty) return true;
	}
	for (int dir = 0; dir < 3; dir++) {
		const int stride = std::max(1, size / 4);
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
int32_t leafAtOrigin(const RtoGpuNode* nodes, int x, int y, int z) {
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

// One leaf that contains surface, prepared by the parallel pass
struct LeafRec {
	int32_t  node;
	uint16_t edgeMask;      // bit dir * 4 + edge: end voxels of that edge are both inside the grid and differ (:590-612)
	uint8_t  memoMask;      // memo[o] holds F(origin - offset o, size of this leaf)
	int32_t  tgt[8];        // offset o (bit 0: x - size, bit 1: y - size, bit 2: z - size) -> leaf that starts there and may be joined, or -1 (:646-687)
	V3       memo[8];
};

// offset bits of neighbour adjIdx (1..3) of edge `edge` in direction `dir` (:630-645); the neighbour's coordinate on an axis is
// e1 - size, i.e. origin - size where the edge sits on the low side of the cell and origin where it sits on the high side
inline int adjOffsetBits(int dir, int edge, int adjIdx) {
	const int a = (dir == 0) ? 1 : 0, b = (dir == 2) ? 1 : 2;          // the two axes the edge index steps along: bit 0 -> a, bit 1 -> b
	const bool lowA = !(edge & 1), lowB = !(edge & 2);
	int bits = 0;
	if ((adjIdx == 1 || adjIdx == 3) && lowA) bits |= 1 << a;
	if ((adjIdx == 2 || adjIdx == 3) && lowB) bits |= 1 << b;
	return bits;
}

template <class F> void parallelFor(size_t n, size_t grain, int nthreads, F&& body) {
	if (nthreads <= 1 || n <= grain) { body(0, n, 0); return; }
	std::atomic<size_t> next{ 0 };
	std::vector<std::thread> th;
	for (int t = 0; t < nthreads; t++)
		th.emplace_back([&, t] { for (;;) { size_t lo = next.fetch_add(grain); if (lo >= n) break; body(lo, std::min(n, lo + grain), t); } });
	for (auto& x : th) x.join();
}

// The bulged face fans of createFaceTriangles (:805-1088) for one face with a sign change
void emitFaceFan(V3 cellVertex, V3 neighborVertex, const int fd[3], int size, float vs, std::vector<RtoTriangle>& out) {
	auto put = [&](V3 a, V3 b, V3 c) {
		RtoTriangle t; t.v0[0] = a.x; t.v0[1] = a.y; t.v0[2] = a.z; t.v1[0] = b.x; t.v1[1] = b.y; t.v1[2] = b.z; t.v2[0] = c.x; t.v2[1] = c.y; t.v2[2] = c.z;
		out.push_back(t);
	};
	const float halfSize = float(size) * vs * 0.5f;
	const V3 faceNormal = mk3(float(fd[0]), float(fd[1]), float(fd[2]));
	const V3 faceCenter = (cellVertex + neighborVertex) * 0.5f;
	V3 tangent1, tangent2;
	if (std::fabs(faceNormal.x) > 0.5f) { tangent1 = mk3(0, 1, 0); tangent2 = mk3(0, 0, 1); }
	else if (std::fabs(faceNormal.y) > 0.5f) { tangent1 = mk3(1, 0, 0); tangent2 = mk3(0, 0, 1); }
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
	for (int i = 0; i < divisions; i++)
		for (int j = 0; j < divisions; j++) {
			const int idx00 = i * (divisions + 1) + j, idx10 = (i + 1) * (divisions + 1) + j, idx01 = i * (divisions + 1) + (j + 1), idx11 = (i + 1) * (divisions + 1) + (j + 1);
			put(cellVertex, gridPoints[idx00], gridPoints[idx10]);
			put(cellVertex, gridPoints[idx10], gridPoints[idx11]);
			put(cellVertex, gridPoints[idx11], gridPoints[idx01]);
			put(cellVertex, gridPoints[idx01], gridPoints[idx00]);
		}
	for (int i = 0; i < divisions; i++)
		for (int j = 0; j < divisions; j++) {
			const int idx00 = i * (divisions + 1) + j, idx10 = (i + 1) * (divisions + 1) + j, idx01 = i * (divisions + 1) + (j + 1), idx11 = (i + 1) * (divisions + 1) + (j + 1);
			put(neighborVertex, gridPoints[idx10], gridPoints[idx00]);
			put(neighborVertex, gridPoints[idx11], gridPoints[idx10]);
			put(neighborVertex, gridPoints[idx01], gridPoints[idx11]);
			put(neighborVertex, gridPoints[idx00], gridPoints[idx01]);
		}
}

} // namespace

extern "C" int rto_host_dc_mesh(const uint8_t* voxels, int dimX, int dimY, int dimZ, const float gridMin[3], float voxelSize,
	const RtoGpuNode* nodes, size_t numNodes, const float* viewProj16, float extraMargin, RtoTriangle** trisOut, size_t* numTris) {
	if (!trisOut || !numTris) return rto_fail(RTO_ERR_INVALID, "rto_host_dc_mesh: null output");
	*trisOut = nullptr; *numTris = 0;
	if (numNodes == 0) return RTO_OK;
	if (!voxels || !nodes || !gridMin) return rto_fail(RTO_ERR_INVALID, "rto_host_dc_mesh: null input");
	if (nodes[0].size > 1024) return rto_fail(RTO_ERR_UNSUPPORTED, "rto_host_dc_mesh: octree larger than 1024 voxels per axis (the reference's cell keys alias there)");
	const DcGrid g{ voxels, dimX, dimY, dimZ, gridMin[0], gridMin[1], gridMin[2], voxelSize };

	// 1. the visit order of renderOctree's traverse lambda (main.cpp:152-187): depth first, children 0..7, frustum-culled subtrees dropped
	std::vector<int32_t> order;
	{
		FrustumPlanes F;
		if (viewProj16) F = frustum_from_view_proj(viewProj16);
		std::vector<int32_t> stack; stack.push_back(0);
		while (!stack.empty()) {
			int32_t i = stack.back(); stack.pop_back();
			if (i < 0 || (size_t)i >= numNodes) continue;
			const RtoGpuNode& n = nodes[i];
			if (viewProj16 && !frustum_node_visible(F, n, gridMin, voxelSize, extraMargin)) continue;
			if (n.isLeaf) order.push_back(i);
			else for (int c = 7; c >= 0; c--) stack.push_back(n.child[c]);
		}
	}

	// dualVertexCache, keyed by leaf: slotOf[node] >= 0 -> cacheVal[slot]; -1 not cached; -2 (during a block) will cache itself in this block
	std::vector<int32_t> slotOf(numNodes, -1);
	std::vector<V3> cacheVal;
	std::vector<RtoTriangle> out;
	const int nthreads = (int)std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 64);
	std::vector<std::vector<Hermite>> scratch((size_t)nthreads);
	std::vector<Hermite> seqScratch;

	const bool verbose = std::getenv("RTO_DC_VERBOSE") != nullptr;
	double tSurf = 0, tRec = 0, tSeq = 0; size_t nSurf = 0, nLate = 0;
	auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	const size_t kBlock = (size_t)1 << 18;
	std::vector<uint8_t> surf;
	std::vector<LeafRec> recs;
	std::vector<uint32_t> recIndex;
	for (size_t b0 = 0; b0 < order.size(); b0 += kBlock) {
		const size_t bn = std::min(kBlock, order.size() - b0);
		// 2a. which leaves of the block contain surface (createTriangles' second early-out, :542-545)
		double tA = now();
		surf.assign(bn, 0);
		parallelFor(bn, 2048, nthreads, [&](size_t lo, size_t hi, int) {
			for (size_t i = lo; i < hi; i++) {
				const RtoGpuNode& n = nodes[order[b0 + i]];
				surf[i] = cellContainsSurface(g, n.x, n.y, n.z, n.size) ? 1 : 0;
			}
		});
		double tB = now();
		recIndex.clear();
		for (size_t i = 0; i < bn; i++) if (surf[i]) { recIndex.push_back((uint32_t)i); int32_t nd = order[b0 + i]; if (slotOf[nd] == -1) slotOf[nd] = -2; }
		recs.resize(recIndex.size());
		// 2b. per surface leaf: edge sign tests, joinable neighbours, and the dual vertices nobody will have cached
		parallelFor(recIndex.size(), 256, nthreads, [&](size_t lo, size_t hi, int tid) {
			std::vector<Hermite>& sc = scratch[(size_t)tid];
			for (size_t r = lo; r < hi; r++) {
				LeafRec& R = recs[r];
				R.node = order[b0 + recIndex[r]];
				const RtoGpuNode& n = nodes[R.node];
				const int x0 = n.x, y0 = n.y, z0 = n.z, size = n.size;
				R.edgeMask = 0; R.memoMask = 0;
				for (int dir = 0; dir < 3; dir++)
					for (int edge = 0; edge < 4; edge++) {
						int ex1 = x0, ey1 = y0, ez1 = z0;
						if (dir == 0) { ey1 += (edge & 1) ? size : 0; ez1 += (edge & 2) ? size : 0; }
						else if (dir == 1) { ex1 += (edge & 1) ? size : 0; ez1 += (edge & 2) ? size : 0; }
						else { ex1 += (edge & 1) ? size : 0; ey1 += (edge & 2) ? size : 0; }
						const int ex2 = ex1 + (dir == 0 ? size : 0), ey2 = ey1 + (dir == 1 ? size : 0), ez2 = ez1 + (dir == 2 ? size : 0);
						if (!g.inb(ex1, ey1, ez1) || !g.inb(ex2, ey2, ez2)) continue;
						if (g.filled(ex1, ey1, ez1) == g.filled(ex2, ey2, ez2)) continue;
						R.edgeMask |= (uint16_t)(1u << (dir * 4 + edge));
					}
				R.tgt[0] = R.node;
				for (int o = 1; o < 8; o++) {
					R.tgt[o] = -1;
					if (o == 7 || !R.edgeMask) continue;                         // (-s, -s, -s) is never asked for
					const int ax = x0 - ((o & 1) ? size : 0), ay = y0 - ((o & 2) ? size : 0), az = z0 - ((o & 4) ? size : 0);
					if (!g.inb(ax, ay, az)) continue;
					const int32_t k = leafAtOrigin(nodes, ax, ay, az);
					if (k < 0) continue;
					const int adjSize = nodes[k].size;
					if (std::max(size, adjSize) > std::min(size, adjSize) * 2) continue;   // MAX_SIZE_DIFFERENCE, :681-685
					R.tgt[o] = k;
				}
				if (slotOf[R.node] < 0) { R.memo[0] = dualVertex(g, x0, y0, z0, size, sc); R.memoMask |= 1; }
				// only the offsets some flagged edge really asks for
				uint8_t asked = 0;
				for (int dir = 0; dir < 3; dir++) for (int edge = 0; edge < 4; edge++) if (R.edgeMask & (1u << (dir * 4 + edge)))
					for (int adjIdx = 1; adjIdx < 4; adjIdx++) asked |= (uint8_t)(1u << adjOffsetBits(dir, edge, adjIdx));
				for (int o = 1; o < 7; o++) {
					if (!(asked & (1u << o)) || R.tgt[o] < 0 || slotOf[R.tgt[o]] != -1) continue;
					const int ax = x0 - ((o & 1) ? size : 0), ay = y0 - ((o & 2) ? size : 0), az = z0 - ((o & 4) ? size : 0);
					R.memo[o] = dualVertex(g, ax, ay, az, size, sc); R.memoMask |= (uint8_t)(1u << o);
				}
			}
		});
		double tC = now();
		// 3. the cache protocol, in visit order
		auto cached = [&](int32_t node, V3& v) { int32_t s = slotOf[node]; if (s < 0) return false; v = cacheVal[(size_t)s]; return true; };
		auto store = [&](int32_t node, V3 v) { slotOf[node] = (int32_t)cacheVal.size(); cacheVal.push_back(v); };
		auto emitIfArea = [&](V3 a, V3 b, V3 c) {
			V3 e1 = b - a, e2 = c - a;
			V3 cr = cross3(e1, e2);
			float area = 0.5f * sqrtf(dot3(cr, cr));
			if (area > 1e-6f) {
				RtoTriangle t; t.v0[0] = a.x; t.v0[1] = a.y; t.v0[2] = a.z; t.v1[0] = b.x; t.v1[1] = b.y; t.v1[2] = b.z; t.v2[0] = c.x; t.v2[1] = c.y; t.v2[2] = c.z;
				out.push_back(t);
			}
		};
		for (size_t r = 0; r < recs.size(); r++) {
			const LeafRec& R = recs[r];
			const RtoGpuNode& n = nodes[R.node];
			const int x0 = n.x, y0 = n.y, z0 = n.z, size = n.size;
			const size_t before = out.size();
			V3 cellVertex;
			if (!cached(R.node, cellVertex)) {
				if (R.memoMask & 1) cellVertex = R.memo[0]; else { cellVertex = dualVertex(g, x0, y0, z0, size, seqScratch); nLate++; }
				store(R.node, cellVertex);
			}
			for (int dir = 0; dir < 3; dir++)
				for (int edge = 0; edge < 4; edge++) {
					if (!(R.edgeMask & (1u << (dir * 4 + edge)))) continue;
					V3 adj[4]; int cnt = 0;
					adj[cnt++] = cellVertex;
					for (int adjIdx = 1; adjIdx < 4; adjIdx++) {
						const int o = adjOffsetBits(dir, edge, adjIdx);
						const int32_t k = R.tgt[o];
						if (k < 0) continue;
						V3 v;
						if (!cached(k, v)) {
							if (R.memoMask & (1u << o)) v = R.memo[o];
							else { v = dualVertex(g, x0 - ((o & 1) ? size : 0), y0 - ((o & 2) ? size : 0), z0 - ((o & 4) ? size : 0), size, seqScratch); nLate++; }
							store(k, v);
						}
						adj[cnt++] = v;
					}
					if (cnt == 3) emitIfArea(adj[0], adj[1], adj[2]);
					else if (cnt >= 4) { emitIfArea(adj[0], adj[1], adj[2]); emitIfArea(adj[0], adj[2], adj[3]); }
				}
			// fallback for boundary cells that produced nothing (:789-799 -> createFaceTriangles)
			if (out.size() == before && (x0 == 0 || y0 == 0 || z0 == 0 || (x0 + size) >= dimX || (y0 + size) >= dimY || (z0 + size) >= dimZ)) {
				const int faceDirections[6][3] = { {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1} };
				for (int face = 0; face < 6; face++) {
					const int nx = x0 + faceDirections[face][0] * size, ny = y0 + faceDirections[face][1] * size, nz = z0 + faceDirections[face][2] * size;
					if (!g.inb(nx, ny, nz)) continue;
					const bool currentSolid = n.isSolid != 0;
					bool neighborSolid = false;
					const int32_t k = leafAtOrigin(nodes, nx, ny, nz);
					if (k >= 0) {
						const int adjSize = nodes[k].size;
						if (std::max(size, adjSize) > std::min(size, adjSize) * 2) continue;
						neighborSolid = nodes[k].isSolid != 0;
					}
					else {
						int cx = nx + size / 2, cy = ny + size / 2, cz = nz + size / 2;
						cx = std::min(std::max(cx, 0), dimX - 1); cy = std::min(std::max(cy, 0), dimY - 1); cz = std::min(std::max(cz, 0), dimZ - 1);
						neighborSolid = g.filled(cx, cy, cz);
					}
					if (currentSolid == neighborSolid) continue;
					V3 neighborVertex;
					if (!(k >= 0 && cached(k, neighborVertex))) {
						float h = float(size) * 0.5f * voxelSize;
						neighborVertex = g.toWorld(nx, ny, nz) + mk3(h, h, h);
						if (k >= 0) store(k, neighborVertex);          // a key no leaf starts at is never read back
					}
					emitFaceFan(cellVertex, neighborVertex, faceDirections[face], size, voxelSize, out);
				}
			}
		}
		tSurf += tB - tA; tRec += tC - tB; tSeq += now() - tC; nSurf += recs.size();
	}
	if (verbose) std::fprintf(stderr, "rto_host_dc_mesh: %zu leaves visited, %zu with surface, %zu triangles; surface test %.2f s, records %.2f s (%d threads), replay %.2f s, %zu vertices computed in the replay\n",
		order.size(), nSurf, out.size(), tSurf, tRec, nthreads, tSeq, nLate);
	if (out.empty()) return RTO_OK;
	RtoTriangle* buf = (RtoTriangle*)std::malloc(out.size() * sizeof(RtoTriangle));
	if (!buf) return rto_fail(RTO_ERR_ALLOC, "rto_host_dc_mesh: out of memory");
	std::memcpy(buf, out.data(), out.size() * sizeof(RtoTriangle));
	*trisOut = buf; *numTris = out.size();
	return RTO_OK;
}
