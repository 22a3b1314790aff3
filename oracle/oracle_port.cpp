// oracle/oracle_port.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Stand-alone CPU restatement ("port") of the reference algorithms on the ray-casting hot path, written
// for clarity, not speed, with no dependency on /root/reference, glm, or the product library.  Every
// function cites the reference lines it follows.  Parity status: PINNED -- tests/test_oracle_pinning.py
// compares this port with oracle/_ref/libref.so (the reference's own translation units compiled in
// place) wherever the reference has real code (octree build, BFS flatten, octreeRaySkip, localMC mesh,
// the Dual-Contouring mesh, BVH build/query, Camera) and against the golden vectors under tests/golden/ generated from it.
// For the pieces the reference does not have (Moller-Trumbore closest hit, shadow rays, GLSL traversal
// on a CPU) the written-down rules of SURVEY.md section 8c are the specification; both this port and
// oracle/ref_harness.cpp (which uses real glm types) implement them and must agree bit for bit.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this.
// Build: oracle/Makefile `port` (-O2 -ffp-contract=off: x86-64 SSE2 scalar, no FMA).

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>
#include <string>
#include <queue>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../ray_tracing_octrees_b200/csrc/mc_tables.h"   // packed public-domain MC table (data only)

namespace {

// ---- glm-order fp32 helpers (thirdparty/glm-0.9.9.7/glm/detail/func_geometric.inl:48-90) -------------
struct V3 { float x, y, z; float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); } };
inline V3 v3(float a, float b, float c) { return V3{ a, b, c }; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
inline V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
inline float dot(V3 a, V3 b) { V3 t = a * b; return t.x + t.y + t.z; }                       // :48-55
inline V3 cross(V3 x, V3 y) { return v3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }  // :68-79
inline V3 normalize(V3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }                      // :82-90 + func_exponential.inl:136-139
inline float fmin2(float a, float b) { return (b < a) ? b : a; }   // glm::min / std::min / GLSL min
inline float fmax2(float a, float b) { return (a < b) ? b : a; }   // glm::max / std::max / GLSL max

// ---- reference-compatible PODs -----------------------------------------------------------------------
struct Grid {                        // VoxelGrid, OctreeVoxel.h:28-42 (x-fastest index)
	int dx = 0, dy = 0, dz = 0; float minX = 0, minY = 0, minZ = 0, voxel = 1; std::vector<uint8_t> data;
	uint8_t safe(int x, int y, int z) const {      // getVoxelSafe, OctreeVoxel.cpp:692-701 (OOB => EMPTY)
		if (x < 0 || y < 0 || z < 0 || x >= dx || y >= dy || z >= dz) return 0;
		return data[(size_t)x + (size_t)y * dx + (size_t)z * ((size_t)dx * dy)];
	}
};
struct ONode { int x, y, z, size; bool leaf, solid, uniform; ONode* child[8]; };   // OctreeNode, OctreeVoxel.h:45-62
struct GNode { int32_t x, y, z, size, isLeaf, isSolid, isUniform, child[8]; };     // GPUNodes, RayTracerBVH.h:21-26
static_assert(sizeof(GNode) == 60, "GPUNodes is 15 x int32");
struct Tri { V3 v0, v1, v2; };                                                       // Triangle, BVH.h:7-11
struct Box { V3 mn, mx; };                                                           // AABB, BVH.h:14-34
struct BNode { Box b; BNode* l = nullptr; BNode* r = nullptr; std::vector<const Tri*> tris; };   // BVHNode, BVH.h:37-42

struct Cam {                         // same layout as RtoCamera (include/rto_c.h)
	float camPos[3]; float invView[16]; float tanHalfFov; float aspect; int width, height;
};

// ---- octree build: buildOctreeRec, OctreeVoxel.cpp:704-762 -----------------------------------------
ONode* buildRec(const Grid& g, int x0, int y0, int z0, int size) {
	ONode* n = new ONode{ x0, y0, z0, size, false, false, false, {} };
	if (size == 1) { n->leaf = true; n->solid = g.safe(x0, y0, z0) == 1; n->uniform = true; return n; }
	uint8_t first = g.safe(x0, y0, z0);
	bool same = true;
	for (int z = z0; z < z0 + size && same; z++)
		for (int y = y0; y < y0 + size && same; y++)
			for (int x = x0; x < x0 + size; x++)
				if (g.safe(x, y, z) != first) { same = false; break; }
	if (same) { n->leaf = true; n->uniform = true; n->solid = first == 1; return n; }
	int half = size / 2;
	for (int i = 0; i < 8; i++)    // bit0 = x, bit1 = y, bit2 = z (:749-757)
		n->child[i] = buildRec(g, x0 + ((i & 1) ? half : 0), y0 + ((i & 2) ? half : 0), z0 + ((i & 4) ? half : 0), half);
	return n;
}
void freeRec(ONode* n) { if (!n) return; for (auto* c : n->child) freeRec(c); delete n; }

struct Octree {
	Grid g; ONode* root = nullptr; std::vector<GNode> flat; std::unordered_map<const ONode*, int> index;
	std::vector<GNode> culled;       // result of the last orc_cull_nodes (render mode 2 traverses it)
	~Octree() { freeRec(root); }
	void build() {                   // createOctreeFromVoxelGrid, OctreeVoxel.cpp:765-778
		if (g.dx == 0 || g.dy == 0 || g.dz == 0) return;
		int maxDim = std::max({ g.dx, g.dy, g.dz });
		int p2 = 1; while (p2 < maxDim) p2 <<= 1;
		root = buildRec(g, 0, 0, 0, p2);
		flatten();
	}
	void flatten() {                 // RayTracerBVH::setOctree BFS, RayTracerBVH.cpp:443-490
		flat.clear(); index.clear();
		std::queue<const ONode*> q; q.push(root); index[root] = 0;
		flat.push_back(GNode{ 0, 0, 0, 0, 0, 0, 0, { -1, -1, -1, -1, -1, -1, -1, -1 } });
		while (!q.empty()) {
			const ONode* nd = q.front(); q.pop();
			int idx = index[nd];
			flat[idx].x = nd->x; flat[idx].y = nd->y; flat[idx].z = nd->z; flat[idx].size = nd->size;
			flat[idx].isLeaf = nd->leaf; flat[idx].isSolid = nd->solid; flat[idx].isUniform = nd->uniform;
			for (int i = 0; i < 8; i++) flat[idx].child[i] = -1;
			if (nd->leaf) continue;
			for (int i = 0; i < 8; i++) if (nd->child[i]) {
				if (!index.count(nd->child[i])) {
					index[nd->child[i]] = (int)flat.size();
					flat.push_back(GNode{ 0, 0, 0, 0, 0, 0, 0, { -1, -1, -1, -1, -1, -1, -1, -1 } });
				}
				flat[idx].child[i] = index[nd->child[i]];
				q.push(nd->child[i]);
			}
		}
	}
};

// ---- marching cubes: localMC, OctreeVoxel.cpp:780-879; vertexInterp :633-640; render order Renderer.cpp:14-36
inline V3 vertexInterp(V3 p1, V3 p2, float a, float b) {
	if (std::fabs(0.0f - a) < 0.00001f) return p1;
	if (std::fabs(0.0f - b) < 0.00001f) return p2;
	if (std::fabs(a - b) < 0.00001f) return p1;
	float mu = (0.0f - a) / (b - a);
	return p1 + mu * (p2 - p1);
}
void localMC(const Grid& g, int x0, int y0, int z0, int size, std::vector<Tri>& out) {
	float vx = g.voxel;
	auto scalar = [&](int x, int y, int z) -> float {       // FILLED => -1, else (incl. OOB) +1 (:787-792)
		if (x < 0 || y < 0 || z < 0 || x >= g.dx || y >= g.dy || z >= g.dz) return 1.0f;
		return g.data[(size_t)x + (size_t)y * g.dx + (size_t)z * ((size_t)g.dx * g.dy)] == 1 ? -1.0f : 1.0f;
	};
	static const int off[8][3] = { {0,0,0},{1,0,0},{1,1,0},{0,1,0},{0,0,1},{1,0,1},{1,1,1},{0,1,1} };   // corner order :802-817
	for (int z = z0; z < z0 + size && z < g.dz - 1; z++)
		for (int y = y0; y < y0 + size && y < g.dy - 1; y++)
			for (int x = x0; x < x0 + size && x < g.dx - 1; x++) {
				V3 pos[8]; float val[8]; int cube = 0;
				for (int c = 0; c < 8; c++) {
					pos[c] = v3(g.minX + (x + off[c][0]) * vx, g.minY + (y + off[c][1]) * vx, g.minZ + (z + off[c][2]) * vx);
					val[c] = scalar(x + off[c][0], y + off[c][1], z + off[c][2]);
					if (val[c] < 0) cube |= 1 << c;
				}
				int flags = mc_edge_flags(cube);
				if (flags == 0) continue;
				V3 vert[12];
				for (int e = 0; e < 12; e++) if (flags & (1 << e)) {
					int a = kMcEdgeCorner[e][0], b = kMcEdgeCorner[e][1];
					vert[e] = vertexInterp(pos[a], pos[b], val[a], val[b]);
				}
				for (int i = 0; mc_tri_edge(cube, i) != -1; i += 3)
					out.push_back(Tri{ vert[mc_tri_edge(cube, i)], vert[mc_tri_edge(cube, i + 1)], vert[mc_tri_edge(cube, i + 2)] });
			}
}
void mcRender(const Grid& g, const ONode* n, int x0, int y0, int z0, int size, std::vector<Tri>& out) {
	if (!n) return;
	if (n->leaf) { localMC(g, x0, y0, z0, size, out); return; }
	int half = size / 2;
	for (int i = 0; i < 8; i++)
		mcRender(g, n->child[i], x0 + ((i & 1) ? half : 0), y0 + ((i & 2) ? half : 0), z0 + ((i & 4) ? half : 0), half, out);
}

// ---- BVH: BVH.cpp:19-113 ----------------------------------------------------------------------------
inline V3 vmin(V3 a, V3 b) { return v3(fmin2(a.x, b.x), fmin2(a.y, b.y), fmin2(a.z, b.z)); }
inline V3 vmax(V3 a, V3 b) { return v3(fmax2(a.x, b.x), fmax2(a.y, b.y), fmax2(a.z, b.z)); }
inline void expand(Box& b, V3 p) { b.mn = vmin(b.mn, p); b.mx = vmax(b.mx, p); }           // AABB::expand, BVH.h:24-27
inline Box emptyBox() { float m = std::numeric_limits<float>::max(); return Box{ v3(m, m, m), v3(-m, -m, -m) }; }
inline V3 centroid(const Tri* t) { return (t->v0 + t->v1 + t->v2) / 3.0f; }                // BVH.cpp:15-17

BNode* bvhBuild(std::vector<const Tri*>& tris) {      // BVH::build, BVH.cpp:33-71
	BNode* node = new BNode();
	Box bounds = emptyBox();
	for (const Tri* t : tris) {
		Box tb = emptyBox(); expand(tb, t->v0); expand(tb, t->v1); expand(tb, t->v2);
		expand(bounds, tb.mn); expand(bounds, tb.mx);
	}
	node->b = bounds;
	if (tris.size() <= 2) { node->tris = tris; return node; }
	V3 ext = bounds.mx - bounds.mn;
	int axis = 0;
	if (ext.y > ext.x) axis = 1;
	if (ext.z > ext[axis]) axis = 2;
	std::sort(tris.begin(), tris.end(), [axis](const Tri* a, const Tri* b) { return centroid(a)[axis] < centroid(b)[axis]; });
	size_t mid = tris.size() / 2;
	std::vector<const Tri*> L(tris.begin(), tris.begin() + mid), R(tris.begin() + mid, tris.end());
	node->l = bvhBuild(L);
	node->r = bvhBuild(R);
	return node;
}
void bvhFree(BNode* n) { if (!n) return; bvhFree(n->l); bvhFree(n->r); delete n; }

inline bool slab(const Box& box, V3 o, V3 inv, const int neg[3], float tmin, float tmax) {   // intersectAABB, BVH.cpp:76-87
	for (int i = 0; i < 3; i++) {
		float t0 = ((neg[i] ? box.mx[i] : box.mn[i]) - o[i]) * inv[i];
		float t1 = ((neg[i] ? box.mn[i] : box.mx[i]) - o[i]) * inv[i];
		tmin = t0 > tmin ? t0 : tmin;
		tmax = t1 < tmax ? t1 : tmax;
		if (tmax < tmin) return false;
	}
	return true;
}
void queryNode(const BNode* n, V3 o, V3 inv, const int neg[3], std::vector<const Tri*>& out, uint64_t* boxes) {   // BVH.cpp:89-105
	if (!n) return;
	if (boxes) (*boxes)++;
	if (!slab(n->b, o, inv, neg, 0.0f, std::numeric_limits<float>::max())) return;
	if (!n->l && !n->r) { for (auto* t : n->tris) out.push_back(t); return; }
	queryNode(n->l, o, inv, neg, out, boxes);
	queryNode(n->r, o, inv, neg, out, boxes);
}
struct Mesh {
	std::vector<Tri> tris; BNode* root = nullptr;
	~Mesh() { bvhFree(root); }
	void build() { std::vector<const Tri*> p; p.reserve(tris.size()); for (auto& t : tris) p.push_back(&t); root = bvhBuild(p); }   // BVH.cpp:19-27
	void query(V3 o, V3 d, std::vector<const Tri*>& out, uint64_t* boxes = nullptr) const {     // BVH::query, BVH.cpp:107-113
		V3 inv = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
		int neg[3] = { inv.x < 0, inv.y < 0, inv.z < 0 };
		queryNode(root, o, inv, neg, out, boxes);
	}
};

// ---- ray generation: GLSL generateRay, RayTracerBVH.cpp:338-355 (inverse(view), tan hoisted to host) -----
inline void genRay(const Cam& c, int px, int py, V3& o, V3& d) {
	float nx = (float(px) + 0.5f) / float(c.width) * 2.0f - 1.0f;
	float ny = 1.0f - (float(py) + 0.5f) / float(c.height) * 2.0f;
	nx *= c.aspect; nx *= c.tanHalfFov; ny *= c.tanHalfFov;
	// normalize(vec4(nx,ny,-1,0)): dot4 = (x*x + y*y) + (z*z + w*w)   (func_geometric.inl:58-66)
	float len2 = (nx * nx + ny * ny) + ((-1.0f) * (-1.0f) + 0.0f * 0.0f);
	float il = 1.0f / std::sqrt(len2);
	float vx = nx * il, vy = ny * il, vz = -1.0f * il, vw = 0.0f * il;
	// mat4 * vec4 = (m0*v0 + m1*v1) + (m2*v2 + m3*v3)   (type_mat4x4.inl:561-572), column-major m
	const float* m = c.invView;
	float w[3];
	for (int r = 0; r < 3; r++) w[r] = (m[0 + r] * vx + m[4 + r] * vy) + (m[8 + r] * vz + m[12 + r] * vw);
	o = v3(c.camPos[0], c.camPos[1], c.camPos[2]);
	d = normalize(v3(w[0], w[1], w[2]));
}

inline V3 shadeLambert(V3 n) {       // GLSL shade, RayTracerBVH.cpp:331-336
	V3 lightDir = normalize(v3(-1.0f, -1.0f, -1.0f));
	float ndotl = fmax2(0.0f, dot(n, -lightDir));
	return v3(1.0f, 0.8f, 0.6f) * ndotl + v3(0.1f, 0.1f, 0.1f);
}

// Moller-Trumbore, SURVEY.md 8c rule; all comparisons reject NaN.
inline bool mollerTrumbore(const Tri& tri, V3 o, V3 d, float& tOut) {
	V3 e1 = tri.v1 - tri.v0, e2 = tri.v2 - tri.v0;
	V3 p = cross(d, e2);
	float det = dot(e1, p);
	if (!(std::fabs(det) >= 1e-8f)) return false;
	float inv = 1.0f / det;
	V3 s = o - tri.v0;
	float u = dot(s, p) * inv;
	if (!(u >= 0.0f && u <= 1.0f)) return false;
	V3 q = cross(s, e1);
	float v = dot(d, q) * inv;
	if (!(v >= 0.0f && u + v <= 1.0f)) return false;
	float t = dot(e2, q) * inv;
	if (!(t > 1e-4f)) return false;
	tOut = t; return true;
}

// ---- octree traversal mode A: octreeRaySkip, VolumeRaycastRenderer.cpp:50-155 ------------------------
float raySkip(const ONode* node, V3 ro, V3 rd, float tMin, float tMax, const Grid& g, const ONode*& leafOut, uint64_t& visits) {
	if (!node) return 1e30f;
	visits++;
	float vx = g.voxel;
	float wx0 = g.minX + node->x * vx, wy0 = g.minY + node->y * vx, wz0 = g.minZ + node->z * vx;   // :70-74
	float wSize = node->size * vx;
	V3 bmin = v3(wx0, wy0, wz0), bmax = v3(wx0 + wSize, wy0 + wSize, wz0 + wSize);
	V3 inv = v3(1.0f / rd.x, 1.0f / rd.y, 1.0f / rd.z);
	const float small = 1e-10f;                                                                      // :84-87
	if (std::fabs(rd.x) < small) inv.x = rd.x >= 0 ? 1e10f : -1e10f;
	if (std::fabs(rd.y) < small) inv.y = rd.y >= 0 ? 1e10f : -1e10f;
	if (std::fabs(rd.z) < small) inv.z = rd.z >= 0 ? 1e10f : -1e10f;
	V3 t1 = (bmin - ro) * inv, t2 = (bmax - ro) * inv;
	V3 tN = vmin(t1, t2), tF = vmax(t1, t2);
	float enterT = fmax2(fmax2(tN.x, tN.y), fmax2(tN.z, tMin));                                     // :96
	float exitT = fmin2(fmin2(tF.x, tF.y), fmin2(tF.z, tMax));                                      // :97
	if (enterT > exitT) return 1e30f;
	if (node->leaf) { if (!node->solid) return 1e30f; leafOut = node; return enterT; }               // :105-110
	int dirMask = ((rd.x > 0) ? 1 : 0) | ((rd.y > 0) ? 2 : 0) | ((rd.z > 0) ? 4 : 0);               // :114-116
	float bestT = 1e30f;
	for (int dist = 0; dist <= 3; dist++)
		for (int oct = 0; oct < 8; oct++) {
			if (__builtin_popcount(oct ^ dirMask) != dist) continue;
			const ONode* c = node->child[oct];
			if (!c) continue;
			float ct = raySkip(c, ro, rd, enterT, exitT, g, leafOut, visits);
			if (ct < bestT) { bestT = ct; if (ct < 1e30f) return ct; }                               // :143-149
		}
	return bestT;
}

// ---- octree traversal mode B: GLSL intersectAABB + intersectOctreeIterative, RayTracerBVH.cpp:226-327 ------
struct HitB { bool hit; float t; int id; V3 normal; int steps; };
HitB traverseGLSL(const std::vector<GNode>& nodes, V3 gridMin, float voxel, V3 o, V3 d) {
	HitB r{ false, 1e30f, -1, v3(0, 0, 0), 0 };
	float closestT = 1e30f;
	int stack[128]; int sp = 0; stack[sp++] = 0;
	while (sp > 0 && r.steps < 512) {
		int idx = stack[--sp];
		if (idx < 0) continue;
		r.steps++;
		const GNode& n = nodes[idx];
		V3 nmin = gridMin + v3((float)n.x, (float)n.y, (float)n.z) * voxel;
		V3 nmax = nmin + v3((float)n.size, (float)n.size, (float)n.size) * voxel;
		V3 inv = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
		V3 t1 = (nmin - o) * inv, t2 = (nmax - o) * inv;
		V3 tMn = vmin(t1, t2), tMx = vmax(t1, t2);
		float tNear = fmax2(fmax2(tMn.x, tMn.y), tMn.z);
		float tFar = fmin2(fmin2(tMx.x, tMx.y), tMx.z);
		if (!(tNear <= tFar && tFar > 0.0f)) continue;
		if (tNear >= closestT) continue;
		if (n.isUniform == 1 || n.isLeaf == 1) {       // the two GLSL branches (:271-306) are identical
			if (n.isSolid == 1) {
				float tHit = fmax2(0.0f, tNear);
				if (tHit < closestT && tHit <= tFar) {
					closestT = tHit; r.hit = true; r.id = idx;
					V3 center = 0.5f * (nmin + nmax);
					V3 p = o + d * tHit;
					r.normal = normalize(p - center);
					break;
				}
			}
			continue;
		}
		for (int i = 0; i < 8; i++) if (n.child[i] >= 0) stack[sp++] = n.child[i];
	}
	r.t = closestT;
	return r;
}

// ---- camera: Camera.cpp:11-29, glm::lookAtRH (gtc/matrix_transform.inl), glm::inverse (func_matrix.inl) ----
void lookAtRH(V3 eye, V3 center, V3 up, float* m /*col-major*/) {
	V3 f = normalize(center - eye);
	V3 s = normalize(cross(f, up));
	V3 u = cross(s, f);
	float r[16] = { s.x, u.x, -f.x, 0,  s.y, u.y, -f.y, 0,  s.z, u.z, -f.z, 0,  -dot(s, eye), -dot(u, eye), dot(f, eye), 1 };
	std::memcpy(m, r, 64);
}
void inverse4(const float* mm, float* out) {      // glm compute_inverse<4,4>, func_matrix.inl
	auto m = [&](int c, int r) { return mm[c * 4 + r]; };
	float c00 = m(2, 2) * m(3, 3) - m(3, 2) * m(2, 3), c02 = m(1, 2) * m(3, 3) - m(3, 2) * m(1, 3), c03 = m(1, 2) * m(2, 3) - m(2, 2) * m(1, 3);
	float c04 = m(2, 1) * m(3, 3) - m(3, 1) * m(2, 3), c06 = m(1, 1) * m(3, 3) - m(3, 1) * m(1, 3), c07 = m(1, 1) * m(2, 3) - m(2, 1) * m(1, 3);
	float c08 = m(2, 1) * m(3, 2) - m(3, 1) * m(2, 2), c10 = m(1, 1) * m(3, 2) - m(3, 1) * m(1, 2), c11 = m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2);
	float c12 = m(2, 0) * m(3, 3) - m(3, 0) * m(2, 3), c14 = m(1, 0) * m(3, 3) - m(3, 0) * m(1, 3), c15 = m(1, 0) * m(2, 3) - m(2, 0) * m(1, 3);
	float c16 = m(2, 0) * m(3, 2) - m(3, 0) * m(2, 2), c18 = m(1, 0) * m(3, 2) - m(3, 0) * m(1, 2), c19 = m(1, 0) * m(2, 2) - m(2, 0) * m(1, 2);
	float c20 = m(2, 0) * m(3, 1) - m(3, 0) * m(2, 1), c22 = m(1, 0) * m(3, 1) - m(3, 0) * m(1, 1), c23 = m(1, 0) * m(2, 1) - m(2, 0) * m(1, 1);
	float F0[4] = { c00, c00, c02, c03 }, F1[4] = { c04, c04, c06, c07 }, F2[4] = { c08, c08, c10, c11 };
	float F3[4] = { c12, c12, c14, c15 }, F4[4] = { c16, c16, c18, c19 }, F5[4] = { c20, c20, c22, c23 };
	float V0[4] = { m(1, 0), m(0, 0), m(0, 0), m(0, 0) }, V1[4] = { m(1, 1), m(0, 1), m(0, 1), m(0, 1) };
	float V2[4] = { m(1, 2), m(0, 2), m(0, 2), m(0, 2) }, V3_[4] = { m(1, 3), m(0, 3), m(0, 3), m(0, 3) };
	const float sA[4] = { +1, -1, +1, -1 }, sB[4] = { -1, +1, -1, +1 };
	float inv[16];
	for (int i = 0; i < 4; i++) {
		float i0 = (V1[i] * F0[i] - V2[i] * F1[i]) + V3_[i] * F2[i];
		float i1 = (V0[i] * F0[i] - V2[i] * F3[i]) + V3_[i] * F4[i];
		float i2 = (V0[i] * F1[i] - V1[i] * F3[i]) + V3_[i] * F5[i];
		float i3 = (V0[i] * F2[i] - V1[i] * F4[i]) + V2[i] * F5[i];
		inv[0 * 4 + i] = i0 * sA[i]; inv[1 * 4 + i] = i1 * sB[i]; inv[2 * 4 + i] = i2 * sA[i]; inv[3 * 4 + i] = i3 * sB[i];
	}
	// Dot0 = m[0] * row0 ; Dot1 = (x + y) + (z + w)
	float d0 = m(0, 0) * inv[0 * 4 + 0], d1 = m(0, 1) * inv[1 * 4 + 0], d2 = m(0, 2) * inv[2 * 4 + 0], d3 = m(0, 3) * inv[3 * 4 + 0];
	float det = (d0 + d1) + (d2 + d3);
	float ood = 1.0f / det;
	for (int i = 0; i < 16; i++) out[i] = inv[i] * ood;
}

} // namespace

#define RTO_FLAG_SHADOWS 1u

extern "C" {

// ---- camera -----------------------------------------------------------------------------------------
void orc_camera_consts(float theta, float phi, float radius, const float* target, float fovDeg, float aspect,
	int w, int h, Cam* out, float* view16) {
	V3 tgt = v3(target[0], target[1], target[2]);
	V3 eye = radius * v3(std::cos(theta) * std::sin(phi), std::sin(theta), std::cos(theta) * std::cos(phi)) + tgt;   // Camera.cpp:13-17
	float view[16];
	lookAtRH(eye, tgt, v3(0, 1, 0), view);
	inverse4(view, out->invView);
	out->camPos[0] = eye.x; out->camPos[1] = eye.y; out->camPos[2] = eye.z;
	float fovRad = fovDeg * 0.01745329251994329576923690768489f;    // glm::radians
	out->tanHalfFov = std::tan(fovRad * 0.5f);
	out->aspect = aspect; out->width = w; out->height = h;
	if (view16) std::memcpy(view16, view, 64);
}

// ---- skip-distance estimate: restatement of VolumeRaycastRenderer.cpp:1598-1664 (49 probe rays through octreeRaySkip, 15th percentile
// x 0.75, temporal blend 0.4 : 0.6); glm::perspective, glm::inverse and mat4 * vec4 in glm's operation order.
float orc_skip_distance(void* hv, float theta, float phi, float radius, const float* target, float aspect, float last, float* tOut, float* oOut, float* dOut) {
	Octree* oc = (Octree*)hv;
	V3 tgt = v3(target[0], target[1], target[2]);
	V3 ro = radius * v3(std::cos(theta) * std::sin(phi), std::sin(theta), std::cos(theta) * std::cos(phi)) + tgt;
	float V[16], P[16] = { 0 }, invV[16], invP[16];
	lookAtRH(ro, tgt, v3(0, 1, 0), V);
	const float zn = 0.1f, zf = 5000.0f, th = std::tan(45.0f * 0.01745329251994329576923690768489f / 2.0f);       // :1608
	P[0] = 1.0f / (aspect * th); P[5] = 1.0f / th; P[10] = -(zf + zn) / (zf - zn); P[11] = -1.0f; P[14] = -(2.0f * zf * zn) / (zf - zn);
	inverse4(V, invV); inverse4(P, invP);
	auto mul = [](const float* m, const float* v, float* o) { for (int r = 0; r < 4; r++) o[r] = (m[r] * v[0] + m[4 + r] * v[1]) + (m[8 + r] * v[2] + m[12 + r] * v[3]); };
	std::vector<float> valid;
	int k = 0;
	for (int y = 0; y < 7; y++) for (int x = 0; x < 7; x++, k++) {
		float ndcX = ((float)x / 6 - 0.5f) * 2.0f * 0.2f, ndcY = ((float)y / 6 - 0.5f) * 2.0f * 0.2f;            // :1619-1620
		float clip[4] = { ndcX, ndcY, 1.0f, 1.0f }, vp[4], wp[4];
		mul(invP, clip, vp);
		float w = vp[3]; for (int c = 0; c < 4; c++) vp[c] /= w;                                                   // :1625
		mul(invV, vp, wp);
		V3 rd = normalize(v3(wp[0], wp[1], wp[2]) - ro);
		const ONode* leaf = nullptr; uint64_t visits = 0;
		float t = raySkip(oc->root, ro, rd, 0.0f, 1e30f, oc->g, leaf, visits);
		if (tOut) tOut[k] = t;
		if (oOut) { oOut[3 * k] = ro.x; oOut[3 * k + 1] = ro.y; oOut[3 * k + 2] = ro.z; }
		if (dOut) { dOut[3 * k] = rd.x; dOut[3 * k + 1] = rd.y; dOut[3 * k + 2] = rd.z; }
		if (t < 1e30f && t > 0.0f) valid.push_back(t);
	}
	float skip = 0.0f;
	if (!valid.empty()) { std::sort(valid.begin(), valid.end()); int idx = std::max(0, (int)(valid.size() * 0.15f)); skip = valid[idx]; skip *= 0.75f; }   // :1646-1654
	const float blend = 0.4f;
	return last * blend + skip * (1.0f - blend);                                                                    // :1657-1661
}

// ---- frustum culling: restatement of the CPU part of RayTracerBVH::renderSceneComputeWithCulling (RayTracerBVH.cpp:724-813) over
// Frustum (Frustum.cpp:5-93), glm::perspective (matrix_clip_space.inl:249-262) and glm's mat4 * mat4 (type_mat4x4.inl:630-648).
// Returns the number of nodes kept; out (may be null) receives the compacted, remapped array (15 ints per node).
size_t orc_cull_nodes(void* hv, float theta, float phi, float radius, const float* target, float fovDeg, float aspect, int w, int h, int32_t* out) {
	Octree* o = (Octree*)hv;
	(void)w; (void)h;
	V3 tgt = v3(target[0], target[1], target[2]);
	V3 eye = radius * v3(std::cos(theta) * std::sin(phi), std::sin(theta), std::cos(theta) * std::cos(phi)) + tgt;   // Camera.cpp:13-17
	float view[16];
	lookAtRH(eye, tgt, v3(0, 1, 0), view);
	float P[16] = { 0 };                                                             // glm::perspective(radians(fovDeg), aspect, 0.01f, 5000.f), :733
	const float zn = 0.01f, zf = 5000.f, th = std::tan(fovDeg * 0.01745329251994329576923690768489f / 2.0f);
	P[0] = 1.0f / (aspect * th); P[5] = 1.0f / th; P[10] = -(zf + zn) / (zf - zn); P[11] = -1.0f; P[14] = -(2.0f * zf * zn) / (zf - zn);
	float M[16];                                                                     // proj * view
	for (int j = 0; j < 4; j++) for (int r = 0; r < 4; r++)
		M[j * 4 + r] = ((P[r] * view[j * 4] + P[4 + r] * view[j * 4 + 1]) + P[8 + r] * view[j * 4 + 2]) + P[12 + r] * view[j * 4 + 3];
	float pl[6][4];                                                                  // Frustum.cpp:5-48: left, right, bottom, top, near, far
	for (int i = 0; i < 6; i++) {
		int row = i / 2; bool plus = (i % 2) == 0;
		for (int c = 0; c < 4; c++) pl[i][c] = plus ? M[c * 4 + 3] + M[c * 4 + row] : M[c * 4 + 3] - M[c * 4 + row];
		float len = std::sqrt(dot(v3(pl[i][0], pl[i][1], pl[i][2]), v3(pl[i][0], pl[i][1], pl[i][2])));
		for (int c = 0; c < 4; c++) pl[i][c] /= len;
	}
	const std::vector<GNode>& flat = o->flat;
	std::vector<int> newIdx(flat.size(), -1);
	int kept = 0;
	const float margin = 150.0f;                                                     // :746
	for (size_t i = 0; i < flat.size(); i++) {
		const GNode& n = flat[i];
		V3 mn = v3(o->g.minX + n.x * o->g.voxel, o->g.minY + n.y * o->g.voxel, o->g.minZ + n.z * o->g.voxel);      // :737-742
		float s = n.size * o->g.voxel;
		V3 mx = v3(mn.x + s, mn.y + s, mn.z + s);
		V3 emn = v3(mn.x - margin, mn.y - margin, mn.z - margin), emx = v3(mx.x + margin, mx.y + margin, mx.z + margin);   // Frustum.cpp:58-59
		bool outside = false;
		for (int k = 0; k < 6 && !outside; k++) {
			V3 p = v3(pl[k][0] > 0 ? emx.x : emn.x, pl[k][1] > 0 ? emx.y : emn.y, pl[k][2] > 0 ? emx.z : emn.z);
			if (dot(v3(pl[k][0], pl[k][1], pl[k][2]), p) + pl[k][3] < 0) outside = true;
		}
		if (!outside) newIdx[i] = kept++;
	}
	o->culled.assign((size_t)kept, GNode());
	{
		for (size_t i = 0; i < flat.size(); i++) {
			if (newIdx[i] < 0) continue;
			GNode n = flat[i];
			if (!n.isLeaf) for (int c = 0; c < 8; c++) { int oc = n.child[c]; n.child[c] = (oc >= 0 && oc < (int)flat.size() && newIdx[oc] >= 0) ? newIdx[oc] : -1; }   // :783-799
			o->culled[(size_t)newIdx[i]] = n;
		}
	}
	if (out && kept) std::memcpy(out, o->culled.data(), (size_t)kept * sizeof(GNode));
	return (size_t)kept;
}

// ---- octree -----------------------------------------------------------------------------------------
void* orc_grid_create(int dx, int dy, int dz, float minX, float minY, float minZ, float voxel, const uint8_t* data) {
	Octree* o = new Octree();
	o->g.dx = dx; o->g.dy = dy; o->g.dz = dz; o->g.minX = minX; o->g.minY = minY; o->g.minZ = minZ; o->g.voxel = voxel;
	o->g.data.assign(data, data + (size_t)dx * dy * dz);
	return o;
}
// ---- CSV voxeliser: restatement of loadCSVDataIntoVoxelGrid, BuildingLoader.cpp:153-290 (+ the loaders :35-129 and
// isPointInTriangle :131-150).  Vertex line: mesh, vertex, easting, northing, elevation, latitude, longitude, elevMin; face line: mesh, v1, v2, v3.
static void csvSplit(const std::string& line, std::vector<std::string>& tok) {          // :51-56 + trim :28-33
	tok.clear();
	std::istringstream ss(line); std::string t;
	while (std::getline(ss, t, ',')) {
		size_t a = t.find_first_not_of(" \t\n\r"), b = t.find_last_not_of(" \t\n\r");
		tok.push_back(a == std::string::npos ? "" : t.substr(a, b - a + 1));
	}
}
static bool pointInTri(V3 p, V3 a, V3 b, V3 c) {                                        // :131-150
	V3 v0 = c - a, v1 = b - a, v2 = p - a;
	float d00 = dot(v0, v0), d01 = dot(v0, v1), d02 = dot(v0, v2), d11 = dot(v1, v1), d12 = dot(v1, v2);
	float inv = d00 * d11 - d01 * d01;
	if (std::fabs(inv) < 1e-7f) return false;
	inv = 1.0f / inv;
	float u = (d11 * d02 - d01 * d12) * inv, v = (d00 * d12 - d01 * d02) * inv;
	return (u >= 0) && (v >= 0) && (u + v <= 1);
}
void* orc_grid_from_csv(const char* vertsPath, const char* facesPath, float voxelSize) {
	Octree* o = new Octree();
	struct VRow { int mesh, num; double e, n, h; };
	struct FRow { int mesh, a, b, c; };
	std::vector<VRow> vr; std::vector<FRow> fr; std::vector<std::string> tok; std::string line;
	{ std::ifstream f(vertsPath); if (f) { std::getline(f, line);                    // :44 header
		while (std::getline(f, line)) { if (line.empty()) continue; csvSplit(line, tok); if (tok.size() < 8) continue;
			try { VRow v; v.mesh = std::stoi(tok[0]); v.num = std::stoi(tok[1]); v.e = std::stod(tok[2]); v.n = std::stod(tok[3]); v.h = std::stod(tok[4]);
				std::stod(tok[5]); std::stod(tok[6]); std::stod(tok[7]); vr.push_back(v); } catch (const std::exception&) {} } } }
	{ std::ifstream f(facesPath); if (f) { std::getline(f, line);
		while (std::getline(f, line)) { if (line.empty()) continue; csvSplit(line, tok); if (tok.size() < 4) continue;
			try { FRow q; q.mesh = std::stoi(tok[0]); q.a = std::stoi(tok[1]); q.b = std::stoi(tok[2]); q.c = std::stoi(tok[3]); fr.push_back(q); } catch (const std::exception&) {} } } }
	if (vr.empty() || fr.empty()) return o;                                             // :158 (empty grid)
	std::unordered_map<int, std::unordered_map<int, VRow>> vmap;                        // :161-164
	for (const VRow& v : vr) vmap[v.mesh][v.num] = v;
	const double M = std::numeric_limits<double>::max();
	double mn[3] = { M, M, M }, mx[3] = { -M, -M, -M };
	for (const VRow& v : vr) if (std::isfinite(v.e) && std::isfinite(v.n) && std::isfinite(v.h)) {      // :174-183
		mn[0] = std::min(mn[0], v.e); mn[1] = std::min(mn[1], v.n); mn[2] = std::min(mn[2], v.h);
		mx[0] = std::max(mx[0], v.e); mx[1] = std::max(mx[1], v.n); mx[2] = std::max(mx[2], v.h);
	}
	const double pad = voxelSize;                                                       // :185-191
	for (int a = 0; a < 3; a++) { mn[a] -= pad; mx[a] += pad; }
	size_t d[3];
	for (int a = 0; a < 3; a++) d[a] = (size_t)std::ceil((mx[a] - mn[a]) / voxelSize);     // :194-196
	const size_t MAXD = 1000;                                                           // :199-208
	if (d[0] > MAXD || d[1] > MAXD || d[2] > MAXD) {
		float scale = std::max({ d[0] / MAXD, d[1] / MAXD, d[2] / MAXD });
		voxelSize *= scale;
		for (int a = 0; a < 3; a++) d[a] = (size_t)std::ceil((mx[a] - mn[a]) / voxelSize);
	}
	Grid& g = o->g;
	g.dx = (int)d[0]; g.dy = (int)d[1]; g.dz = (int)d[2]; g.minX = (float)mn[0]; g.minY = (float)mn[1]; g.minZ = (float)mn[2]; g.voxel = voxelSize;
	g.data.assign(d[0] * d[1] * d[2], 0);
	for (const FRow& q : fr) {                                                          // :230-287
		auto m = vmap.find(q.mesh); if (m == vmap.end()) continue;
		auto ia = m->second.find(q.a), ib = m->second.find(q.b), ic = m->second.find(q.c);
		if (ia == m->second.end() || ib == m->second.end() || ic == m->second.end()) continue;
		V3 v1 = { (float)ia->second.e, (float)ia->second.n, (float)ia->second.h }, v2 = { (float)ib->second.e, (float)ib->second.n, (float)ib->second.h },
			v3 = { (float)ic->second.e, (float)ic->second.n, (float)ic->second.h };
		float tmn[3] = { std::min({ v1.x, v2.x, v3.x }), std::min({ v1.y, v2.y, v3.y }), std::min({ v1.z, v2.z, v3.z }) };
		float tmx[3] = { std::max({ v1.x, v2.x, v3.x }), std::max({ v1.y, v2.y, v3.y }), std::max({ v1.z, v2.z, v3.z }) };
		const float gm[3] = { g.minX, g.minY, g.minZ }; const int dd[3] = { g.dx, g.dy, g.dz };
		int s[3], e[3];
		for (int a = 0; a < 3; a++) { s[a] = std::max(0, (int)((tmn[a] - gm[a]) / voxelSize)); e[a] = std::min(dd[a] - 1, (int)((tmx[a] - gm[a]) / voxelSize) + 1); }
		if (e[0] < s[0] || e[1] < s[1] || e[2] < s[2]) continue;
		for (int z = s[2]; z <= e[2]; z++) for (int y = s[1]; y <= e[1]; y++) for (int x = s[0]; x <= e[0]; x++) {
			V3 c = { g.minX + (x + 0.5f) * voxelSize, g.minY + (y + 0.5f) * voxelSize, g.minZ + (z + 0.5f) * voxelSize };
			if (pointInTri(c, v1, v2, v3)) {
				size_t idx = (size_t)x + (size_t)y * g.dx + (size_t)z * g.dx * g.dy;
				if (idx < g.data.size()) g.data[idx] = 1;
			}
		}
	}
	return o;
}

void* orc_grid_load(const char* path) {     // sceneCache.bin format, CacheUtils.cpp:33-59
	FILE* f = std::fopen(path, "rb");
	if (!f) return nullptr;
	Octree* o = new Octree();
	size_t n = 0; bool ok = true;
	ok &= std::fread(&o->g.dx, 4, 1, f) == 1; ok &= std::fread(&o->g.dy, 4, 1, f) == 1; ok &= std::fread(&o->g.dz, 4, 1, f) == 1;
	ok &= std::fread(&o->g.minX, 4, 1, f) == 1; ok &= std::fread(&o->g.minY, 4, 1, f) == 1; ok &= std::fread(&o->g.minZ, 4, 1, f) == 1;
	ok &= std::fread(&o->g.voxel, 4, 1, f) == 1; ok &= std::fread(&n, sizeof(size_t), 1, f) == 1;
	if (ok) { o->g.data.resize(n); ok = std::fread(o->g.data.data(), 1, n, f) == n; }
	std::fclose(f);
	if (!ok) { delete o; return nullptr; }
	return o;
}
void orc_grid_info(void* h, int* dims, float* minVoxel) {
	Octree* o = (Octree*)h;
	dims[0] = o->g.dx; dims[1] = o->g.dy; dims[2] = o->g.dz;
	minVoxel[0] = o->g.minX; minVoxel[1] = o->g.minY; minVoxel[2] = o->g.minZ; minVoxel[3] = o->g.voxel;
}
void orc_grid_data(void* h, uint8_t* out) { Octree* o = (Octree*)h; std::memcpy(out, o->g.data.data(), o->g.data.size()); }
int orc_octree_build(void* h) { Octree* o = (Octree*)h; o->build(); return (int)o->flat.size(); }
void orc_octree_flat(void* h, int32_t* out15) { Octree* o = (Octree*)h; std::memcpy(out15, o->flat.data(), o->flat.size() * 60); }
void orc_octree_free(void* h) { delete (Octree*)h; }

double orc_render_octree(void* h, const Cam* cam, int mode, int y0, int y1, float* rgba, int32_t* leafId, float* tOut,
	uint64_t* stats, int nthreads) {
	Octree* oc = (Octree*)h;
	const int W = cam->width;
	V3 gridMin = v3(oc->g.minX, oc->g.minY, oc->g.minZ);
	uint64_t sVisits = 0;
#ifdef _OPENMP
	if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
	auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 4) reduction(+:sVisits)
	for (int py = y0; py < y1; py++)
		for (int px = 0; px < W; px++) {
			size_t pix = (size_t)(py - y0) * W + px;
			V3 o, d; genRay(*cam, px, py, o, d);
			float tRes = 1e30f; int id = -1; V3 color = v3(0, 0, 0);
			if (mode == 0) {
				const ONode* leaf = nullptr; uint64_t visits = 0;
				tRes = raySkip(oc->root, o, d, 0.0f, 1e30f, oc->g, leaf, visits);
				sVisits += visits;
				if (leaf) {     // extension: box-centre normal + Lambert, same formulas as the GLSL hit (:279-283)
					id = oc->index[leaf];
					V3 nmin = gridMin + v3((float)leaf->x, (float)leaf->y, (float)leaf->z) * oc->g.voxel;
					V3 nmax = nmin + v3((float)leaf->size, (float)leaf->size, (float)leaf->size) * oc->g.voxel;
					V3 center = 0.5f * (nmin + nmax);
					V3 p = o + d * tRes;
					color = shadeLambert(normalize(p - center));
				}
			}
			else {
				HitB r = traverseGLSL(mode == 2 ? oc->culled : oc->flat, gridMin, oc->g.voxel, o, d);
				sVisits += (uint64_t)r.steps;
				if (r.hit) { tRes = r.t; id = r.id; color = shadeLambert(r.normal); }
			}
			if (rgba) { rgba[4 * pix] = color.x; rgba[4 * pix + 1] = color.y; rgba[4 * pix + 2] = color.z; rgba[4 * pix + 3] = 1.0f; }
			if (leafId) leafId[pix] = id;
			if (tOut) tOut[pix] = tRes;
		}
	double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	if (stats) { stats[0] = sVisits; stats[1] = 0; }
	return sec;
}

void orc_octree_rayskip(void* h, const float* o3, const float* d3, size_t n, float tMin, float tMax, float* tOut, int32_t* idOut) {
	Octree* oc = (Octree*)h;
	for (size_t i = 0; i < n; i++) {
		const ONode* leaf = nullptr; uint64_t visits = 0;
		tOut[i] = raySkip(oc->root, v3(o3[3 * i], o3[3 * i + 1], o3[3 * i + 2]), v3(d3[3 * i], d3[3 * i + 1], d3[3 * i + 2]), tMin, tMax, oc->g, leaf, visits);
		if (idOut) idOut[i] = leaf ? oc->index[leaf] : -1;
	}
}

// ---- Adaptive Dual Contouring: renderOctree (main.cpp:95-208) over AdaptiveDualContouringRenderer::createTriangles -----------------
// Literal, sequential restatement of AdaptiveDualContouringRenderer.cpp with the reference's two maps: g_octreeMap (every node under
// x << 20 | y << 10 | z in buildOctreeRec's insertion order, OctreeVoxel.cpp:551-553, 712-713: a child 0 overwrites its parent) and
// dualVertexCache (first writer wins).  The edge-intersection cache holds pure values and is left out.
struct DcPort {
	const Grid& g;
	std::unordered_map<long long, const ONode*> octreeMap;
	std::unordered_map<long long, V3> dualVertexCache;
	struct Hermite { V3 position, normal; };
	explicit DcPort(const Grid& grid) : g(grid) {}
	static long long key(int x, int y, int z) { return ((long long)x << 20) | ((long long)y << 10) | (long long)z; }
	void fillMap(const ONode* n) { if (!n) return; octreeMap[key(n->x, n->y, n->z)] = n; for (auto* c : n->child) fillMap(c); }
	bool filled(int x, int y, int z) const { return g.data[(size_t)x + (size_t)y * g.dx + (size_t)z * ((size_t)g.dx * g.dy)] == 1; }
	bool inb(int x, int y, int z) const { return !(x < 0 || y < 0 || z < 0 || x >= g.dx || y >= g.dy || z >= g.dz); }
	V3 gridToWorld(int x, int y, int z) const { return v3(g.minX + x * g.voxel, g.minY + y * g.voxel, g.minZ + z * g.voxel); }     // :1359-1365
	static V3 clampv(V3 x, V3 lo, V3 hi) { return v3(fmin2(fmax2(x.x, lo.x), hi.x), fmin2(fmax2(x.y, lo.y), hi.y), fmin2(fmax2(x.z, lo.z), hi.z)); }
	static V3 mixv(V3 x, V3 y, float a) { return x * (1.0f - a) + y * a; }                                                          // func_common.inl:104-112

	Hermite calculateIntersection(int x1, int y1, int z1, int x2, int y2, int z2) const {                                          // :1236-1357
		bool isFilled1 = filled(x1, y1, z1), isFilled2 = filled(x2, y2, z2);
		float v1 = isFilled1 ? -1.0f : 1.0f, v2 = isFilled2 ? -1.0f : 1.0f;
		V3 p1 = gridToWorld(x1, y1, z1), p2 = gridToWorld(x2, y2, z2);
		float t = v1 / (v1 - v2);
		t = fmin2(fmax2(t, 0.0f), 1.0f);
		V3 position = p1 + t * (p2 - p1);
		int dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;                                     // always one unit step along one axis here
		auto getScalar = [&](int x, int y, int z) -> float { if (!inb(x, y, z)) return 1.0f; return filled(x, y, z) ? -1.0f : 1.0f; };
		V3 normal = v3((float)dx, (float)dy, (float)dz);
		if (dx != 0) normal = v3(0.0f, getScalar(x1, y1 + 1, z1) - getScalar(x1, y1 - 1, z1), getScalar(x1, y1, z1 + 1) - getScalar(x1, y1, z1 - 1));
		else if (dy != 0) normal = v3(getScalar(x1 + 1, y1, z1) - getScalar(x1 - 1, y1, z1), 0.0f, getScalar(x1, y1, z1 + 1) - getScalar(x1, y1, z1 - 1));
		else normal = v3(getScalar(x1 + 1, y1, z1) - getScalar(x1 - 1, y1, z1), getScalar(x1, y1 + 1, z1) - getScalar(x1, y1 - 1, z1), 0.0f);
		if (dot(normal, normal) < 1e-10) normal = v3((float)dx, (float)dy, (float)dz);
		else normal = normalize(normal);
		float dotProduct = normal.x * dx + normal.y * dy + normal.z * dz;
		if ((dotProduct > 0) == isFilled2) normal = -normal;
		return Hermite{ position, normal };
	}
	std::vector<Hermite> gatherHermiteData(int x0, int y0, int z0, int size) const {                                              // :1090-1144
		int maxX = std::min(x0 + size, g.dx - 1), maxY = std::min(y0 + size, g.dy - 1), maxZ = std::min(z0 + size, g.dz - 1);
		int minX = std::max(x0, 0), minY = std::max(y0, 0), minZ = std::max(z0, 0);
		int stride = (size > 8) ? 2 : 1;
		if (size <= 4) stride = 1;
		std::vector<Hermite> points;
		for (int z = minZ; z <= maxZ; z += stride) for (int y = minY; y <= maxY; y += stride) for (int x = minX; x <= maxX; x += stride) {
			bool currentFilled = filled(x, y, z);
			const int dirs[3][3] = { {1, 0, 0}, {0, 1, 0}, {0, 0, 1} };
			for (int d = 0; d < 3; d++) {
				int nx = x + dirs[d][0], ny = y + dirs[d][1], nz = z + dirs[d][2];
				if (!inb(nx, ny, nz)) continue;
				if (currentFilled != filled(nx, ny, nz)) points.push_back(calculateIntersection(x, y, z, nx, ny, nz));
			}
		}
		return points;
	}
	struct Qef {                                                                                                                   // :46-160
		float ata[3][3] = { {0, 0, 0}, {0, 0, 0}, {0, 0, 0} }; V3 atb = v3(0, 0, 0), pointSum = v3(0, 0, 0); int numPoints = 0;
		void addPoint(V3 point, V3 normalIn) {
			V3 n = normalize(normalIn);
			const float c[3] = { n.x, n.y, n.z };
			for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) ata[i][j] += c[i] * c[j];
			float d = -dot(n, point);
			atb.x += n.x * d; atb.y += n.y * d; atb.z += n.z * d;
			pointSum = pointSum + point; numPoints++;
		}
		V3 solve(V3 cellCenter, float cellSize) const {
			V3 masspoint = (numPoints > 0) ? pointSum / (float)numPoints : cellCenter;
			if (numPoints <= 2) return masspoint;
			float m[3][3];
			for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = ata[i][j];
			m[0][0] += 0.3f; m[1][1] += 0.3f; m[2][2] += 0.3f;
			auto det3 = [&]() { return m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2]) + m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]); };   // func_matrix.inl:211-220
			bool invertible = true;
			float inv[3][3];
			if (std::abs(det3()) < 1e-10) invertible = false;
			else {
				float ood = 1.0f / det3();                                                                                         // func_matrix.inl:269-291
				inv[0][0] = +(m[1][1] * m[2][2] - m[2][1] * m[1][2]) * ood; inv[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]) * ood; inv[2][0] = +(m[1][0] * m[2][1] - m[2][0] * m[1][1]) * ood;
				inv[0][1] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]) * ood; inv[1][1] = +(m[0][0] * m[2][2] - m[2][0] * m[0][2]) * ood; inv[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]) * ood;
				inv[0][2] = +(m[0][1] * m[1][2] - m[1][1] * m[0][2]) * ood; inv[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]) * ood; inv[2][2] = +(m[0][0] * m[1][1] - m[1][0] * m[0][1]) * ood;
				for (int i = 0; i < 3 && invertible; i++) for (int j = 0; j < 3 && invertible; j++)
					if (std::isnan(inv[i][j]) || std::isinf(inv[i][j]) || std::abs(inv[i][j]) > 1e6) invertible = false;
			}
			if (invertible) {
				V3 solution = v3(inv[0][0] * atb.x + inv[1][0] * atb.y + inv[2][0] * atb.z, inv[0][1] * atb.x + inv[1][1] * atb.y + inv[2][1] * atb.z, inv[0][2] * atb.x + inv[1][2] * atb.y + inv[2][2] * atb.z);
				solution = masspoint + 0.7f * (solution - masspoint);
				if (!std::isnan(solution.x) && !std::isnan(solution.y) && !std::isnan(solution.z)) {
					V3 dl = masspoint - solution;
					if (dot(dl, dl) < cellSize * cellSize) return mixv(solution, masspoint, 0.2f);
				}
			}
			return masspoint;
		}
	};
	V3 generateDualVertex(const std::vector<Hermite>& hermiteData, V3 cellCenter, float cellSize) const {                          // :1146-1234
		if (hermiteData.empty()) return cellCenter;
		V3 halfSize = v3(cellSize * 0.5f, cellSize * 0.5f, cellSize * 0.5f);
		V3 minBound = cellCenter - halfSize, maxBound = cellCenter + halfSize;
		float inset = cellSize * 0.001f;
		minBound = minBound + v3(inset, inset, inset); maxBound = maxBound - v3(inset, inset, inset);
		V3 massPoint = v3(0, 0, 0);
		for (const auto& hp : hermiteData) massPoint = massPoint + hp.position;
		massPoint = massPoint / (float)hermiteData.size();
		V3 avgNormal = v3(0, 0, 0);
		for (const auto& hp : hermiteData) avgNormal = avgNormal + hp.normal;
		if (std::sqrt(dot(avgNormal, avgNormal)) > 0.0001f) {
			avgNormal = normalize(avgNormal);
			V3 a = v3(std::fabs(avgNormal.x), std::fabs(avgNormal.y), std::fabs(avgNormal.z));
			float maxComp = fmax2(fmax2(a.x, a.y), a.z);
			if (maxComp > 0.85f) {
				if (a.x == maxComp) avgNormal = v3(avgNormal.x > 0 ? 1.0f : -1.0f, 0, 0);
				else if (a.y == maxComp) avgNormal = v3(0, avgNormal.y > 0 ? 1.0f : -1.0f, 0);
				else avgNormal = v3(0, 0, avgNormal.z > 0 ? 1.0f : -1.0f);
				V3 planePoint = v3(0, 0, 0); int planePointCount = 0;
				for (const auto& hp : hermiteData) if (dot(normalize(hp.normal), avgNormal) > 0.7f) { planePoint = planePoint + hp.position; planePointCount++; }
				if (planePointCount > 0) {
					planePoint = planePoint / (float)planePointCount;
					float d = -dot(avgNormal, planePoint);
					float t = -(dot(avgNormal, cellCenter) + d);
					return clampv(cellCenter + t * avgNormal, minBound, maxBound);
				}
			}
		}
		Qef qef;
		for (const auto& hp : hermiteData) qef.addPoint(hp.position, hp.normal);
		V3 c = (minBound + maxBound) * 0.5f;                                                                                       // solveConstrained, :151-161
		V3 qefSolution = clampv(qef.solve(c, maxBound.x - minBound.x), minBound, maxBound);
		return mixv(qefSolution, massPoint, 0.1f);
	}
	bool cellContainsSurface(int x0, int y0, int z0, int size) const {                                                            // :1367-1530
		int maxX = std::min(x0 + size, g.dx), maxY = std::min(y0 + size, g.dy), maxZ = std::min(z0 + size, g.dz);
		int minX = std::max(x0, 0), minY = std::max(y0, 0), minZ = std::max(z0, 0);
		if (minX >= maxX || minY >= maxY || minZ >= maxZ) return false;
		bool anyFilled = false, anyEmpty = false;
		const int corners[8][3] = { {minX, minY, minZ}, {maxX - 1, minY, minZ}, {maxX - 1, maxY - 1, minZ}, {minX, maxY - 1, minZ},
			{minX, minY, maxZ - 1}, {maxX - 1, minY, maxZ - 1}, {maxX - 1, maxY - 1, maxZ - 1}, {minX, maxY - 1, maxZ - 1} };
		for (int i = 0; i < 8; i++) {
			if (!inb(corners[i][0], corners[i][1], corners[i][2])) continue;
			if (filled(corners[i][0], corners[i][1], corners[i][2])) anyFilled = true; else anyEmpty = true;
			if (anyFilled && anyEmpty) return true;
		}
		for (int dir = 0; dir < 3; dir++) {
			int stride = std::max(1, size / 4);
			for (int offset = 0; offset < size; offset += stride) {
				int lo[3] = { minX, minY, minZ }, hi[3] = { maxX, maxY, maxZ }, dim[3] = { g.dx, g.dy, g.dz };
				int p[3];                                                // the two axes other than dir step along the diagonal
				bool skip = false;
				for (int a = 0; a < 3; a++) if (a != dir) { p[a] = lo[a] + offset; if (p[a] >= hi[a]) skip = true; }
				if (skip) continue;
				for (int side = 0; side < 2; side++) {                   // the face at min, then the face at max
					int c1 = side == 0 ? lo[dir] - 1 : hi[dir] - 1, c2 = side == 0 ? lo[dir] : hi[dir];
					if (c1 >= 0 && c2 < dim[dir]) {
						int q1[3] = { p[0], p[1], p[2] }, q2[3] = { p[0], p[1], p[2] };
						q1[dir] = c1; q2[dir] = c2;
						if (filled(q1[0], q1[1], q1[2]) != filled(q2[0], q2[1], q2[2])) return true;
					}
				}
			}
		}
		if (size <= 4)
			for (int z = minZ; z < maxZ - 1; z++) for (int y = minY; y < maxY - 1; y++) for (int x = minX; x < maxX - 1; x++) {
				bool s = filled(x, y, z);
				if (s != filled(x + 1, y, z) || s != filled(x, y + 1, z) || s != filled(x, y, z + 1)) return true;
			}
		return false;
	}
	static void pushIfArea(std::vector<Tri>& out, V3 a, V3 b, V3 c) {                                                             // :727-739
		V3 cr = cross(b - a, c - a);
		if (0.5f * std::sqrt(dot(cr, cr)) > 1e-6f) out.push_back(Tri{ a, b, c });
	}
	void createFaceTriangles(const ONode* node, int x0, int y0, int z0, int size, std::vector<Tri>& triangles) {                  // :805-1088
		long long cellKey = key(x0, y0, z0);
		V3 cellVertex;
		auto it = dualVertexCache.find(cellKey);
		if (it != dualVertexCache.end()) cellVertex = it->second;
		else { cellVertex = gridToWorld(x0, y0, z0) + v3(size * 0.5f * g.voxel, size * 0.5f * g.voxel, size * 0.5f * g.voxel); dualVertexCache[cellKey] = cellVertex; }
		const int faceDirections[6][3] = { {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1} };
		for (int face = 0; face < 6; face++) {
			int nx = x0 + faceDirections[face][0] * size, ny = y0 + faceDirections[face][1] * size, nz = z0 + faceDirections[face][2] * size;
			if (!inb(nx, ny, nz)) continue;
			bool currentSolid = node->solid, neighborSolid = false;
			long long neighborKey = key(nx, ny, nz);
			auto nodeIt = octreeMap.find(neighborKey);
			bool leafThere = nodeIt != octreeMap.end() && nodeIt->second->leaf;
			if (leafThere) {
				int adjSize = nodeIt->second->size;
				if (std::max(size, adjSize) > std::min(size, adjSize) * 2) continue;
				neighborSolid = nodeIt->second->solid;
			}
			else {
				int cx = std::min(std::max(nx + size / 2, 0), g.dx - 1), cy = std::min(std::max(ny + size / 2, 0), g.dy - 1), cz = std::min(std::max(nz + size / 2, 0), g.dz - 1);
				neighborSolid = filled(cx, cy, cz);
			}
			if (currentSolid == neighborSolid) continue;
			V3 neighborVertex; bool hasNeighborVertex = false;
			if (leafThere) { auto f = dualVertexCache.find(neighborKey); if (f != dualVertexCache.end()) { neighborVertex = f->second; hasNeighborVertex = true; } }
			if (!hasNeighborVertex) { neighborVertex = gridToWorld(nx, ny, nz) + v3(size * 0.5f * g.voxel, size * 0.5f * g.voxel, size * 0.5f * g.voxel); dualVertexCache[neighborKey] = neighborVertex; }
			float halfSize = size * g.voxel * 0.5f;
			V3 faceNormal = v3((float)faceDirections[face][0], (float)faceDirections[face][1], (float)faceDirections[face][2]);
			V3 faceCenter = (cellVertex + neighborVertex) * 0.5f;
			V3 tangent1, tangent2;
			if (std::fabs(faceNormal.x) > 0.5f) { tangent1 = v3(0, 1, 0); tangent2 = v3(0, 0, 1); }
			else if (std::fabs(faceNormal.y) > 0.5f) { tangent1 = v3(1, 0, 0); tangent2 = v3(0, 0, 1); }
			else { tangent1 = v3(1, 0, 0); tangent2 = v3(0, 1, 0); }
			const int divisions = 2;
			std::vector<V3> gridPoints;
			for (int i = 0; i <= divisions; i++) {
				float u = i / float(divisions);
				for (int j = 0; j <= divisions; j++) {
					float v = j / float(divisions);
					float mappedU = 2.0f * u - 1.0f, mappedV = 2.0f * v - 1.0f;
					V3 point = faceCenter + tangent1 * (mappedU * halfSize) + tangent2 * (mappedV * halfSize);
					float distFromCenter = std::sqrt(mappedU * mappedU + mappedV * mappedV);
					float bulge = 0.05f * halfSize * (1.0f - distFromCenter * distFromCenter);
					point = point + faceNormal * bulge;
					gridPoints.push_back(point);
				}
			}
			for (int pass = 0; pass < 2; pass++)
				for (int i = 0; i < divisions; i++) for (int j = 0; j < divisions; j++) {
					int idx00 = i * (divisions + 1) + j, idx10 = (i + 1) * (divisions + 1) + j, idx01 = i * (divisions + 1) + (j + 1), idx11 = (i + 1) * (divisions + 1) + (j + 1);
					if (pass == 0) {
						triangles.push_back(Tri{ cellVertex, gridPoints[idx00], gridPoints[idx10] }); triangles.push_back(Tri{ cellVertex, gridPoints[idx10], gridPoints[idx11] });
						triangles.push_back(Tri{ cellVertex, gridPoints[idx11], gridPoints[idx01] }); triangles.push_back(Tri{ cellVertex, gridPoints[idx01], gridPoints[idx00] });
					}
					else {
						triangles.push_back(Tri{ neighborVertex, gridPoints[idx10], gridPoints[idx00] }); triangles.push_back(Tri{ neighborVertex, gridPoints[idx11], gridPoints[idx10] });
						triangles.push_back(Tri{ neighborVertex, gridPoints[idx01], gridPoints[idx11] }); triangles.push_back(Tri{ neighborVertex, gridPoints[idx00], gridPoints[idx01] });
					}
				}
		}
	}
	void createTriangles(const ONode* node, int x0, int y0, int z0, int size, std::vector<Tri>& result) {                         // :528-803
		std::vector<Tri> triangles;
		if (!node || !node->leaf) return;
		if (!cellContainsSurface(x0, y0, z0, size)) return;
		float cellSizeWorld = size * g.voxel;
		V3 cellCenter = gridToWorld(x0, y0, z0) + v3(size * 0.5f * g.voxel, size * 0.5f * g.voxel, size * 0.5f * g.voxel);
		long long cellKey = key(x0, y0, z0);
		V3 cellVertex = cellCenter;
		auto it = dualVertexCache.find(cellKey);
		if (it != dualVertexCache.end()) cellVertex = it->second;
		else {
			std::vector<Hermite> hermiteData = gatherHermiteData(x0, y0, z0, size);
			if (!hermiteData.empty()) cellVertex = generateDualVertex(hermiteData, cellCenter, cellSizeWorld);
			dualVertexCache[cellKey] = cellVertex;
		}
		static const int edgeDirections[3][3] = { {1, 0, 0}, {0, 1, 0}, {0, 0, 1} };
		for (int dir = 0; dir < 3; dir++) for (int edge = 0; edge < 4; edge++) {
			int ex1 = x0, ey1 = y0, ez1 = z0;
			if (dir == 0) { ey1 += (edge & 1) ? size : 0; ez1 += (edge & 2) ? size : 0; }
			else if (dir == 1) { ex1 += (edge & 1) ? size : 0; ez1 += (edge & 2) ? size : 0; }
			else { ex1 += (edge & 1) ? size : 0; ey1 += (edge & 2) ? size : 0; }
			int ex2 = ex1 + edgeDirections[dir][0] * size, ey2 = ey1 + edgeDirections[dir][1] * size, ez2 = ez1 + edgeDirections[dir][2] * size;
			if (!inb(ex1, ey1, ez1) || !inb(ex2, ey2, ez2)) continue;
			if (filled(ex1, ey1, ez1) == filled(ex2, ey2, ez2)) continue;
			std::vector<V3> adjacent;
			adjacent.push_back(cellVertex);
			for (int adjIdx = 1; adjIdx < 4; adjIdx++) {
				int adjX = x0, adjY = y0, adjZ = z0;
				if (dir == 0) { if (adjIdx == 1) adjY = ey1 - size; else if (adjIdx == 2) adjZ = ez1 - size; else { adjY = ey1 - size; adjZ = ez1 - size; } }
				else if (dir == 1) { if (adjIdx == 1) adjX = ex1 - size; else if (adjIdx == 2) adjZ = ez1 - size; else { adjX = ex1 - size; adjZ = ez1 - size; } }
				else { if (adjIdx == 1) adjX = ex1 - size; else if (adjIdx == 2) adjY = ey1 - size; else { adjX = ex1 - size; adjY = ey1 - size; } }
				if (!inb(adjX, adjY, adjZ)) continue;
				long long k = key(adjX, adjY, adjZ);
				auto nodeIt = octreeMap.find(k);
				if (nodeIt == octreeMap.end() || !nodeIt->second->leaf) continue;
				int adjSize = nodeIt->second->size, mySize = node->size;
				if (std::max(mySize, adjSize) > std::min(mySize, adjSize) * 2) continue;
				V3 adjVertex;
				auto found = dualVertexCache.find(k);
				if (found != dualVertexCache.end()) adjVertex = found->second;
				else {
					V3 adjCenter = gridToWorld(adjX, adjY, adjZ) + v3(size * 0.5f * g.voxel, size * 0.5f * g.voxel, size * 0.5f * g.voxel);
					std::vector<Hermite> adjHermite = gatherHermiteData(adjX, adjY, adjZ, size);
					adjVertex = !adjHermite.empty() ? generateDualVertex(adjHermite, adjCenter, cellSizeWorld) : adjCenter;
					dualVertexCache[k] = adjVertex;
				}
				adjacent.push_back(adjVertex);
			}
			if (adjacent.size() == 3) pushIfArea(triangles, adjacent[0], adjacent[1], adjacent[2]);
			else if (adjacent.size() >= 4) { pushIfArea(triangles, adjacent[0], adjacent[1], adjacent[2]); pushIfArea(triangles, adjacent[0], adjacent[2], adjacent[3]); }
		}
		if (triangles.empty() && (x0 == 0 || y0 == 0 || z0 == 0 || (x0 + size) >= g.dx || (y0 + size) >= g.dy || (z0 + size) >= g.dz))
			createFaceTriangles(node, x0, y0, z0, size, triangles);
		result.insert(result.end(), triangles.begin(), triangles.end());
	}
};

// ---- mesh + BVH -------------------------------------------------------------------------------------
void* orc_mesh_from_octree(void* h) {
	Octree* oc = (Octree*)h; Mesh* m = new Mesh();
	mcRender(oc->g, oc->root, 0, 0, 0, oc->root->size, m->tris);
	return m;
}
// renderOctree's traverse lambda (main.cpp:152-187): depth first, children 0..7, nodes whose box + extraMargin is outside the frustum
// of viewProj16 (column-major proj * view; NULL: no culling) are dropped with their subtrees; every leaf goes through createTriangles.
void* orc_dc_mesh_from_octree(void* h, const float* viewProj16, float extraMargin) {
	Octree* oc = (Octree*)h; Mesh* m = new Mesh();
	if (!oc->root) return m;
	DcPort dc(oc->g);
	dc.fillMap(oc->root);
	float pl[6][4];
	if (viewProj16) for (int i = 0; i < 6; i++) {                                   // Frustum.cpp:5-48
		int row = i / 2; bool plus = (i % 2) == 0;
		for (int c = 0; c < 4; c++) pl[i][c] = plus ? viewProj16[c * 4 + 3] + viewProj16[c * 4 + row] : viewProj16[c * 4 + 3] - viewProj16[c * 4 + row];
		float len = std::sqrt(dot(v3(pl[i][0], pl[i][1], pl[i][2]), v3(pl[i][0], pl[i][1], pl[i][2])));
		for (int c = 0; c < 4; c++) pl[i][c] /= len;
	}
	struct Walk {
		Octree* oc; DcPort& dc; Mesh* m; const float* vp; float (*pl)[4]; float margin;
		void go(const ONode* n) {
			if (!n) return;
			if (vp) {                                                                   // Frustum::testAABB == -1, Frustum.cpp:52-75
				const Grid& g = oc->g;
				V3 mn = v3(g.minX + n->x * g.voxel, g.minY + n->y * g.voxel, g.minZ + n->z * g.voxel);
				float s = n->size * g.voxel;
				V3 mx = v3(mn.x + s, mn.y + s, mn.z + s);
				V3 emn = v3(mn.x - margin, mn.y - margin, mn.z - margin), emx = v3(mx.x + margin, mx.y + margin, mx.z + margin);
				for (int k = 0; k < 6; k++) {
					V3 p = v3(pl[k][0] > 0 ? emx.x : emn.x, pl[k][1] > 0 ? emx.y : emn.y, pl[k][2] > 0 ? emx.z : emn.z);
					if (dot(v3(pl[k][0], pl[k][1], pl[k][2]), p) + pl[k][3] < 0) return;
				}
			}
			if (n->leaf) dc.createTriangles(n, n->x, n->y, n->z, n->size, m->tris);
			else for (auto* c : n->child) go(c);
		}
	} walk{ oc, dc, m, viewProj16, pl, extraMargin };
	walk.go(oc->root);
	return m;
}
void* orc_mesh_from_tris(const float* xyz9, size_t n) { Mesh* m = new Mesh(); m->tris.resize(n); std::memcpy((void*)m->tris.data(), xyz9, n * 36); return m; }
size_t orc_mesh_count(void* mv) { return ((Mesh*)mv)->tris.size(); }
void orc_mesh_tris(void* mv, float* out9) { Mesh* m = (Mesh*)mv; std::memcpy(out9, m->tris.data(), m->tris.size() * 36); }
double orc_bvh_build(void* mv) {
	auto t0 = std::chrono::steady_clock::now();
	((Mesh*)mv)->build();
	return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
void orc_mesh_free(void* mv) { delete (Mesh*)mv; }

static void exportNode(const Mesh* m, const BNode* n, std::vector<float>& boxes, std::vector<int32_t>& meta) {
	for (int i = 0; i < 3; i++) boxes.push_back(n->b.mn[i]);
	for (int i = 0; i < 3; i++) boxes.push_back(n->b.mx[i]);
	bool leaf = !n->l && !n->r;
	meta.push_back(leaf); meta.push_back((int32_t)n->tris.size());
	for (int i = 0; i < 2; i++) meta.push_back(i < (int)n->tris.size() ? (int32_t)(n->tris[i] - &m->tris[0]) : -1);
	if (!leaf) { exportNode(m, n->l, boxes, meta); exportNode(m, n->r, boxes, meta); }
}
size_t orc_bvh_export(void* mv, float* boxes6, int32_t* meta4, size_t cap) {
	Mesh* m = (Mesh*)mv; std::vector<float> b; std::vector<int32_t> me;
	exportNode(m, m->root, b, me);
	size_t n = me.size() / 4;
	if (boxes6 && meta4 && n <= cap) { std::memcpy(boxes6, b.data(), b.size() * 4); std::memcpy(meta4, me.data(), me.size() * 4); }
	return n;
}
size_t orc_bvh_query(void* mv, const float* o3, const float* d3, size_t nrays, int64_t* offsets, int32_t* ids, size_t cap) {
	Mesh* m = (Mesh*)mv; std::vector<const Tri*> cand; size_t total = 0;
	for (size_t r = 0; r < nrays; r++) {
		cand.clear();
		m->query(v3(o3[3 * r], o3[3 * r + 1], o3[3 * r + 2]), v3(d3[3 * r], d3[3 * r + 1], d3[3 * r + 2]), cand);
		offsets[r] = (int64_t)total;
		for (auto* t : cand) { if (ids && total < cap) ids[total] = (int32_t)(t - &m->tris[0]); total++; }
	}
	offsets[nrays] = (int64_t)total;
	return total;
}

double orc_render_bvh(void* mv, const Cam* cam, unsigned flags, float shadowBias, int y0, int y1,
	float* rgba, int32_t* hitId, float* tOut, uint64_t* stats, int nthreads) {
	Mesh* m = (Mesh*)mv;
	const int W = cam->width;
	uint64_t sB = 0, sC = 0, sBs = 0, sCs = 0, nShadow = 0;
#ifdef _OPENMP
	if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
	auto tstart = std::chrono::steady_clock::now();
#pragma omp parallel reduction(+:sB,sC,sBs,sCs,nShadow)
	{
		std::vector<const Tri*> cand;
#pragma omp for schedule(dynamic, 4)
		for (int py = y0; py < y1; py++)
			for (int px = 0; px < W; px++) {
				size_t pix = (size_t)(py - y0) * W + px;
				V3 o, d; genRay(*cam, px, py, o, d);
				cand.clear();
				m->query(o, d, cand, stats ? &sB : nullptr);
				sC += cand.size();
				float best = 1e30f; const Tri* bestTri = nullptr;
				for (const Tri* tri : cand) { float t; if (mollerTrumbore(*tri, o, d, t) && t < best) { best = t; bestTri = tri; } }
				V3 color = v3(0, 0, 0);
				if (bestTri) {
					V3 e1 = bestTri->v1 - bestTri->v0, e2 = bestTri->v2 - bestTri->v0;
					V3 n = normalize(cross(e1, e2));
					if (dot(n, d) > 0.0f) n = -n;
					V3 hit = o + d * best;
					bool shadowed = false;
					if (flags & RTO_FLAG_SHADOWS) {
						V3 so = hit + n * shadowBias;
						V3 sd = normalize(v3(1.0f, 1.0f, 1.0f));
						cand.clear();
						m->query(so, sd, cand, stats ? &sBs : nullptr);
						sCs += cand.size(); nShadow++;
						for (const Tri* tri : cand) { float t; if (mollerTrumbore(*tri, so, sd, t)) { shadowed = true; break; } }
					}
					color = shadowed ? v3(0.1f, 0.1f, 0.1f) : shadeLambert(n);
				}
				if (rgba) { rgba[4 * pix] = color.x; rgba[4 * pix + 1] = color.y; rgba[4 * pix + 2] = color.z; rgba[4 * pix + 3] = 1.0f; }
				if (hitId) hitId[pix] = bestTri ? (int32_t)(bestTri - &m->tris[0]) : -1;
				if (tOut) tOut[pix] = best;
			}
	}
	double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - tstart).count();
	if (stats) { stats[0] = sB; stats[1] = sC; stats[2] = sBs; stats[3] = sCs; stats[4] = nShadow; }
	return sec;
}

int orc_num_threads() {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

} // extern "C"
