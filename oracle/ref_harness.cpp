// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin extern "C" harness around the UNMODIFIED reference translation units, compiled
// in place from /root/reference by oracle/Makefile into oracle/_ref/libref.so.
// Nothing from the reference is copied: this file only #includes / links it.
//
// What is real reference code here (called, never restated):
//   createOctreeFromVoxelGrid / freeOctree      453-skeleton/OctreeVoxel.cpp:704-778,881-888
//   RayTracerBVH::setOctree (BFS flatten)       453-skeleton/RayTracerBVH.cpp:430-505 (GL entry points stubbed)
//   static octreeRaySkip                        453-skeleton/VolumeRaycastRenderer.cpp:50-155 (via #include of the .cpp)
//   MarchingCubesRenderer::render / localMC     453-skeleton/Renderer.cpp:14-36, OctreeVoxel.cpp:780-879
//   BVH::BVH / BVH::query                       453-skeleton/BVH.cpp:19-113
//   Camera::getView / getPos                    453-skeleton/Camera.cpp:11-29
//   loadVoxelGrid                               453-skeleton/CacheUtils.cpp:33-59
//   AdaptiveDualContouringRenderer::render      453-skeleton/AdaptiveDualContouringRenderer.cpp:489-1530 (createTriangles per leaf)
//   loadCSVDataIntoVoxelGrid (CSV voxeliser)    453-skeleton/BuildingLoader.cpp:153-290
//   glm 0.9.9.7 arithmetic                      thirdparty/glm-0.9.9.7
//
// What is an EXTENSION RULE written here with glm types (the reference
// has no such code, so these rules *define* the expected answer):
//   pixel ray generation on the host           restates GLSL RayTracerBVH.cpp:338-355
//   Moller-Trumbore closest hit over BVH::query candidates, shadow any-hit, Lambert shade
//                                               shade restates GLSL RayTracerBVH.cpp:331-336
//   GLSL intersectOctreeIterative in C++        restates RayTracerBVH.cpp:226-236, 239-327
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.

#include "VolumeRaycastRenderer.cpp"   // brings in static octreeRaySkip + file-local intersectAABB

#define private public
#include "BVH.h"
#include "RayTracerBVH.h"
#undef private

#include "Renderer.h"
#include "AdaptiveDualContouringRenderer.h"
#include "Frustum.h"
#include "CacheUtils.h"
#include "Camera.h"

#include <glm/gtc/matrix_inverse.hpp>
#include <cstring>
#include <cstdint>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

// ---------------------------------------------------------------------------------------------
// GL stubs so that RayTracerBVH::setOctree can run headless (it only uploads an SSBO).
// ---------------------------------------------------------------------------------------------
static void APIENTRY stubGenBuffers(GLsizei n, GLuint* b) { for (GLsizei i = 0; i < n; i++) b[i] = 1u + (GLuint)i; }
static void APIENTRY stubBindBuffer(GLenum, GLuint) {}
static void APIENTRY stubBufferData(GLenum, GLsizeiptr, const void*, GLenum) {}
static void APIENTRY stubBindBufferBase(GLenum, GLuint, GLuint) {}
static void APIENTRY stubDeleteBuffers(GLsizei, const GLuint*) {}

static void installGLStubs() {
	glad_glGenBuffers = stubGenBuffers;
	glad_glBindBuffer = stubBindBuffer;
	glad_glBufferData = stubBufferData;
	glad_glBindBufferBase = stubBindBufferBase;
	glad_glDeleteBuffers = stubDeleteBuffers;
}

// Stubs for everything RayTracerBVH::renderSceneComputeWithCulling touches (RayTracerBVH.cpp:706-892), so that its CPU part -- the
// frustum test of every node, the compaction and the child-index remap (:724-813) -- runs headless.  The SSBO re-upload (:808-812)
// is captured: that byte array is exactly what the compute shader would traverse.
static std::vector<uint8_t> g_lastBufferData;
static void APIENTRY captureBufferData(GLenum, GLsizeiptr size, const void* data, GLenum) {
	g_lastBufferData.assign((const uint8_t*)data, (const uint8_t*)data + (data ? size : 0));
}
static void APIENTRY stubGenTextures(GLsizei n, GLuint* t) { for (GLsizei i = 0; i < n; i++) t[i] = 7u + (GLuint)i; }
static void APIENTRY stubActiveTexture(GLenum) {}
static void APIENTRY stubBindTexture(GLenum, GLuint) {}
static void APIENTRY stubTexImage2D(GLenum, GLint, GLint, GLsizei, GLsizei, GLint, GLenum, GLenum, const void*) {}
static void APIENTRY stubTexParameteri(GLenum, GLenum, GLint) {}
static void APIENTRY stubUseProgram(GLuint) {}
static void APIENTRY stubBindImageTexture(GLuint, GLuint, GLint, GLboolean, GLint, GLenum, GLenum) {}
static GLint APIENTRY stubGetUniformLocation(GLuint, const GLchar*) { return 0; }
static void APIENTRY stubUniform1i(GLint, GLint) {}
static void APIENTRY stubUniform1f(GLint, GLfloat) {}
static void APIENTRY stubUniform3f(GLint, GLfloat, GLfloat, GLfloat) {}
static void APIENTRY stubUniformMatrix4fv(GLint, GLsizei, GLboolean, const GLfloat*) {}
static void APIENTRY stubDispatchCompute(GLuint, GLuint, GLuint) {}
static void APIENTRY stubMemoryBarrier(GLbitfield) {}
static void APIENTRY stubBindVertexArray(GLuint) {}
static void APIENTRY stubDrawArrays(GLenum, GLint, GLsizei) {}
static void installRenderStubs() {
	installGLStubs();
	glad_glBufferData = captureBufferData;
	glad_glGenTextures = stubGenTextures; glad_glActiveTexture = stubActiveTexture; glad_glBindTexture = stubBindTexture;
	glad_glTexImage2D = stubTexImage2D; glad_glTexParameteri = stubTexParameteri; glad_glUseProgram = stubUseProgram;
	glad_glBindImageTexture = stubBindImageTexture; glad_glGetUniformLocation = stubGetUniformLocation;
	glad_glUniform1i = stubUniform1i; glad_glUniform1f = stubUniform1f; glad_glUniform3f = stubUniform3f;
	glad_glUniformMatrix4fv = stubUniformMatrix4fv; glad_glDispatchCompute = stubDispatchCompute; glad_glMemoryBarrier = stubMemoryBarrier;
	glad_glBindVertexArray = stubBindVertexArray; glad_glDrawArrays = stubDrawArrays;
}

// ---------------------------------------------------------------------------------------------
// Handles
// ---------------------------------------------------------------------------------------------
struct RefOctree {
	VoxelGrid grid;
	OctreeNode* root = nullptr;
	RayTracerBVH* tracer = nullptr;                       // owns m_flatNodes (real setOctree output)
	std::unordered_map<const OctreeNode*, int> bfsIndex;  // node -> index in m_flatNodes
};

struct RefMesh {
	std::vector<Triangle> tris;
	BVH* bvh = nullptr;
};

struct RefCamConsts {        // layout shared with tests (ctypes)
	float camPos[3];
	float invView[16];       // column-major, glm::inverse(Camera::getView())
	float tanHalfFov;        // std::tan(glm::radians(fovDeg) * 0.5f)
	float aspect;
	int   width, height;
};

extern "C" {

// ---- voxel grids --------------------------------------------------------------------------
void* ref_grid_create(int dx, int dy, int dz, float minX, float minY, float minZ, float voxelSize,
	const uint8_t* data) {
	RefOctree* h = new RefOctree();
	h->grid.dimX = dx; h->grid.dimY = dy; h->grid.dimZ = dz;
	h->grid.minX = minX; h->grid.minY = minY; h->grid.minZ = minZ;
	h->grid.voxelSize = voxelSize;
	h->grid.data.resize((size_t)dx * dy * dz);
	std::memcpy(h->grid.data.data(), data, h->grid.data.size());
	return h;
}

// real loadCSVDataIntoVoxelGrid (BuildingLoader.cpp:153-290); its std::cout chatter is silenced for the duration of the call
void* ref_grid_from_csv(const char* verts, const char* faces, float voxelSize) {
	RefOctree* h = new RefOctree();
	std::streambuf* keepOut = std::cout.rdbuf(nullptr);
	std::streambuf* keepErr = std::cerr.rdbuf(nullptr);
	h->grid = loadCSVDataIntoVoxelGrid(verts, faces, voxelSize);
	std::cout.rdbuf(keepOut); std::cerr.rdbuf(keepErr);
	return h;
}

void* ref_grid_load(const char* path) {   // real loadVoxelGrid
	RefOctree* h = new RefOctree();
	if (!loadVoxelGrid(path, h->grid)) { delete h; return nullptr; }
	return h;
}

void ref_grid_info(void* hv, int* dims, float* minVoxel) {
	RefOctree* h = (RefOctree*)hv;
	dims[0] = h->grid.dimX; dims[1] = h->grid.dimY; dims[2] = h->grid.dimZ;
	minVoxel[0] = h->grid.minX; minVoxel[1] = h->grid.minY; minVoxel[2] = h->grid.minZ;
	minVoxel[3] = h->grid.voxelSize;
}

void ref_grid_data(void* hv, uint8_t* out) {
	RefOctree* h = (RefOctree*)hv;
	std::memcpy(out, h->grid.data.data(), h->grid.data.size());
}

// ---- octree: real build + real BFS flatten ---------------------------------------------------
int ref_octree_build(void* hv) {
	RefOctree* h = (RefOctree*)hv;
	installGLStubs();
	h->root = createOctreeFromVoxelGrid(h->grid);
	h->tracer = new RayTracerBVH();
	h->tracer->setOctree(h->root, h->grid);
	// Recover node -> flat index with the same BFS discovery order setOctree uses
	// (RayTracerBVH.cpp:443-490); verified against m_flatNodes below.
	h->bfsIndex.clear();
	std::vector<const OctreeNode*> order;
	if (h->root) { order.push_back(h->root); h->bfsIndex[h->root] = 0; }
	for (size_t i = 0; i < order.size(); i++) {
		const OctreeNode* nd = order[i];
		if (nd->isLeaf) continue;
		for (int c = 0; c < 8; c++) if (nd->children[c]) {
			h->bfsIndex[nd->children[c]] = (int)order.size();
			order.push_back(nd->children[c]);
		}
	}
	const std::vector<GPUNodes>& fn = h->tracer->m_flatNodes;
	if (order.size() != fn.size()) return -1;
	for (size_t i = 0; i < order.size(); i++) {
		if (fn[i].x != order[i]->x || fn[i].y != order[i]->y || fn[i].z != order[i]->z || fn[i].size != order[i]->size) return -2;
		for (int c = 0; c < 8; c++) {
			int expect = (!order[i]->isLeaf && order[i]->children[c]) ? h->bfsIndex[order[i]->children[c]] : -1;
			if (fn[i].child[c] != expect) return -3;
		}
	}
	return (int)fn.size();
}

void ref_octree_flat(void* hv, int32_t* out15) {   // GPUNodes[] as 15 x int32 per node
	RefOctree* h = (RefOctree*)hv;
	static_assert(sizeof(GPUNodes) == 60, "GPUNodes must be 15 x int32");
	std::memcpy(out15, h->tracer->m_flatNodes.data(), h->tracer->m_flatNodes.size() * sizeof(GPUNodes));
}

void ref_octree_free(void* hv) {
	RefOctree* h = (RefOctree*)hv;
	if (!h) return;
	delete h->tracer;
	freeOctree(h->root);
	delete h;
}

// ---- camera: real Camera + glm::inverse ------------------------------------------------------
// The REAL RayTracerBVH::renderSceneComputeWithCulling(camera, w, h, aspect, fovDeg, true) (RayTracerBVH.cpp:706-892) up to its GL
// dispatch: returns the number of nodes of the culled array it uploaded; out (may be null) receives them, 15 ints each.
// Call with out == nullptr first to learn the count.  Needs ref_octree_build before.
size_t ref_cull_nodes(void* hv, float theta, float phi, float radius, const float* target, float fovDeg, float aspect, int w, int h, int32_t* out) {
	RefOctree* o = (RefOctree*)hv;
	if (!o->tracer) return 0;
	installRenderStubs();
	o->tracer->m_computeInited = true; o->tracer->m_computeProg = 1; o->tracer->m_fsqProg = 1;
	Camera cam(theta, phi, radius);
	cam.setTarget(glm::vec3(target[0], target[1], target[2]));
	g_lastBufferData.clear();
	std::streambuf* keepOut = std::cout.rdbuf(nullptr);
	o->tracer->renderSceneComputeWithCulling(cam, w, h, aspect, fovDeg, true);
	std::cout.rdbuf(keepOut);
	installGLStubs();
	size_t n = g_lastBufferData.size() / sizeof(GPUNodes);
	if (out) std::memcpy(out, g_lastBufferData.data(), n * sizeof(GPUNodes));
	return n;
}

// The skip-distance estimate of VolumeRaycastRenderer::drawRaycast (VolumeRaycastRenderer.cpp:1598-1664).  That function is GL code
// from top to bottom and cannot be called headless, so its 60 lines around the octreeRaySkip calls are restated here with the same
// glm expressions; the 49 octreeRaySkip calls themselves are the reference's own function.  `last` plays the role of the
// function-local static lastSkipDistance.  tOut (49), o, d (147 each) may be null.
float ref_skip_distance(void* hv, float theta, float phi, float radius, const float* target, float aspect, float last, float* tOut, float* oOut, float* dOut) {
	RefOctree* h = (RefOctree*)hv;
	Camera cam(theta, phi, radius);
	cam.setTarget(glm::vec3(target[0], target[1], target[2]));
	float skipDistance = 0.0f;
	const int gridSize = 7; const float sampleOffset = 0.2f;
	std::vector<float> validSkipDistances;
	glm::mat4 V = cam.getView();
	glm::mat4 P = glm::perspective(glm::radians(45.0f), aspect, 0.1f, 5000.0f);
	glm::mat4 invV = glm::inverse(V);
	glm::mat4 invP = glm::inverse(P);
	glm::vec3 ro = cam.getPos();
	int k = 0;
	for (int y = 0; y < gridSize; y++) {
		for (int x = 0; x < gridSize; x++, k++) {
			float ndcX = ((float)x / (gridSize - 1) - 0.5f) * 2.0f * sampleOffset;
			float ndcY = ((float)y / (gridSize - 1) - 0.5f) * 2.0f * sampleOffset;
			glm::vec4 clipPos(ndcX, ndcY, 1.f, 1.f);
			glm::vec4 viewPos = invP * clipPos;
			viewPos /= viewPos.w;
			glm::vec4 worldPos4 = invV * viewPos;
			glm::vec3 rd = glm::normalize(glm::vec3(worldPos4) - ro);
			float raySkip = octreeRaySkip(h->root, ro, rd, 0.0f, 1e30f, h->grid);
			if (tOut) tOut[k] = raySkip;
			if (oOut) { oOut[3 * k] = ro.x; oOut[3 * k + 1] = ro.y; oOut[3 * k + 2] = ro.z; }
			if (dOut) { dOut[3 * k] = rd.x; dOut[3 * k + 1] = rd.y; dOut[3 * k + 2] = rd.z; }
			if (raySkip < 1e30f && raySkip > 0.0f) validSkipDistances.push_back(raySkip);
		}
	}
	if (!validSkipDistances.empty()) {
		std::sort(validSkipDistances.begin(), validSkipDistances.end());
		int safeIndex = std::max(0, (int)(validSkipDistances.size() * 0.15f));
		skipDistance = validSkipDistances[safeIndex];
		skipDistance *= 0.75f;
	}
	float blendFactor = 0.4f;
	skipDistance = last * blendFactor + skipDistance * (1.0f - blendFactor);
	return skipDistance;
}

void ref_camera_consts(float theta, float phi, float radius, const float* target, float fovDeg, float aspect,
	int w, int h, RefCamConsts* out, float* view16) {
	Camera cam(theta, phi, radius);
	cam.setTarget(glm::vec3(target[0], target[1], target[2]));
	glm::mat4 view = cam.getView();
	glm::mat4 inv = glm::inverse(view);          // the GLSL recomputes inverse(view) per thread; rule: host glm
	glm::vec3 pos = cam.getPos();
	out->camPos[0] = pos.x; out->camPos[1] = pos.y; out->camPos[2] = pos.z;
	std::memcpy(out->invView, &inv[0][0], 64);
	if (view16) std::memcpy(view16, &view[0][0], 64);
	float fovRad = glm::radians(fovDeg);
	out->tanHalfFov = std::tan(fovRad * 0.5f);
	out->aspect = aspect;
	out->width = w; out->height = h;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------
// Extension rules (glm op order, fp32, no FMA: build with -ffp-contract=off)
// ---------------------------------------------------------------------------------------------
static inline void genRay(const RefCamConsts& c, int px, int py, glm::vec3& o, glm::vec3& d) {
	// GLSL generateRay, RayTracerBVH.cpp:338-355, with inverse(view) and tan() hoisted to the host.
	float nx = (float(px) + 0.5f) / float(c.width) * 2.0f - 1.0f;
	float ny = 1.0f - (float(py) + 0.5f) / float(c.height) * 2.0f;
	nx *= c.aspect;
	nx *= c.tanHalfFov;
	ny *= c.tanHalfFov;
	glm::mat4 invView;
	std::memcpy(&invView[0][0], c.invView, 64);
	glm::vec4 rayDirView = glm::normalize(glm::vec4(nx, ny, -1.0f, 0.0f));
	glm::vec4 rayDirWorld = invView * rayDirView;
	o = glm::vec3(c.camPos[0], c.camPos[1], c.camPos[2]);
	d = glm::normalize(glm::vec3(rayDirWorld));
}

static inline glm::vec3 shadeLambert(const glm::vec3& normal) {   // GLSL shade, RayTracerBVH.cpp:331-336
	glm::vec3 lightDir = glm::normalize(glm::vec3(-1.0f, -1.0f, -1.0f));
	float ndotl = glm::max(0.0f, glm::dot(normal, -lightDir));
	return glm::vec3(1.0f, 0.8f, 0.6f) * ndotl + glm::vec3(0.1f, 0.1f, 0.1f);
}

// Moller-Trumbore per SURVEY.md 8c; every comparison is written so that NaN rejects.
static inline bool mollerTrumbore(const Triangle& tri, const glm::vec3& o, const glm::vec3& d, float& tOut) {
	glm::vec3 e1 = tri.v1 - tri.v0;
	glm::vec3 e2 = tri.v2 - tri.v0;
	glm::vec3 p = glm::cross(d, e2);
	float det = glm::dot(e1, p);
	if (!(std::fabs(det) >= 1e-8f)) return false;
	float inv = 1.0f / det;
	glm::vec3 s = o - tri.v0;
	float u = glm::dot(s, p) * inv;
	if (!(u >= 0.0f && u <= 1.0f)) return false;
	glm::vec3 q = glm::cross(s, e1);
	float v = glm::dot(d, q) * inv;
	if (!(v >= 0.0f && u + v <= 1.0f)) return false;
	float t = glm::dot(e2, q) * inv;
	if (!(t > 1e-4f)) return false;
	tOut = t;
	return true;
}

#define RTO_FLAG_SHADOWS 1u

extern "C" {

// ---- mesh + BVH: real MarchingCubesRenderer + real BVH --------------------------------------
void* ref_mesh_from_octree(void* hv) {
	RefOctree* h = (RefOctree*)hv;
	MarchingCubesRenderer mc;
	std::vector<MCTriangle> mct = mc.render(h->root, h->grid, 0, 0, 0, h->root->size);
	RefMesh* m = new RefMesh();
	m->tris.resize(mct.size());
	for (size_t i = 0; i < mct.size(); i++) { m->tris[i].v0 = mct[i].v[0]; m->tris[i].v1 = mct[i].v[1]; m->tris[i].v2 = mct[i].v[2]; }
	return m;
}

// The Dual-Contouring mesh as the application produces it: renderOctree (main.cpp:95-208) walks the octree depth first (children 0..7),
// skips every node whose box (+ extraMargin) is outside the frustum and calls AdaptiveDualContouringRenderer::render on each leaf, one
// after the other on the calling thread (the renderer's thread pool is never fed by this path), so the renderer's dual-vertex cache
// fills in that order.  main.cpp is the GL application and cannot be compiled here; its 30-line traversal lambda (main.cpp:152-187)
// is restated below with the same glm expressions, everything it calls is the reference's own code.  The compute-shader attempt of
// render() (taken only when the root itself is a leaf) is disabled: there is no GL context.
// viewProj16 == NULL: no culling (every leaf is visited).
static std::vector<MCTriangle> refDcTriangles(void* hv, const float* viewProj16, float extraMargin) {
	RefOctree* h = (RefOctree*)hv;
	AdaptiveDualContouringRenderer dc;
	dc.m_useComputeShader = false;
	dc.m_useAdaptiveLOD = true;
	dc.m_detailThreshold = 0.1f;
	std::vector<MCTriangle> result;
	Frustum* frustum = nullptr;
	if (viewProj16) { glm::mat4 VP; std::memcpy(&VP[0][0], viewProj16, 64); frustum = new Frustum(VP); }
	const VoxelGrid& grid = h->grid;
	Renderer& renderer = dc;
	std::function<void(const OctreeNode*)> traverse = [&](const OctreeNode* node) {
		if (!node) return;
		float voxelSize = grid.voxelSize;
		glm::vec3 minPoint(grid.minX + node->x * voxelSize, grid.minY + node->y * voxelSize, grid.minZ + node->z * voxelSize);
		glm::vec3 maxPoint = minPoint + glm::vec3(node->size * voxelSize);
		if (frustum) {
			int frustumTest = frustum->testAABB(minPoint, maxPoint, extraMargin);
			if (frustumTest == -1) return;
		}
		if (node->isLeaf) {
			auto tris = renderer.render(node, grid, node->x, node->y, node->z, node->size);
			if (!tris.empty()) result.insert(result.end(), tris.begin(), tris.end());
		}
		else {
			for (auto child : node->children) traverse(child);
		}
	};
	traverse(h->root);
	delete frustum;
	return result;
}
void* ref_dc_mesh_from_octree(void* hv, const float* viewProj16, float extraMargin) {
	std::vector<MCTriangle> result = refDcTriangles(hv, viewProj16, extraMargin);
	RefMesh* m = new RefMesh();
	m->tris.resize(result.size());
	for (size_t i = 0; i < result.size(); i++) { m->tris[i].v0 = result[i].v[0]; m->tris[i].v1 = result[i].v[1]; m->tris[i].v2 = result[i].v[2]; }
	return m;
}
// The same run with the MCTriangles whole (3 vertices + 3 normals = 18 floats each, the record of the triangle cache, main.cpp:27-67).
// Returns the count; fills out18 when cap is large enough.
size_t ref_dc_mesh_full(void* hv, const float* viewProj16, float extraMargin, float* out18, size_t cap) {
	std::vector<MCTriangle> result = refDcTriangles(hv, viewProj16, extraMargin);
	static_assert(sizeof(MCTriangle) == 72, "MCTriangle must be 18 floats");
	if (out18 && cap >= result.size() && !result.empty()) std::memcpy(out18, result.data(), result.size() * 72);
	return result.size();
}

void* ref_mesh_from_tris(const float* xyz9, size_t n) {
	RefMesh* m = new RefMesh();
	m->tris.resize(n);
	static_assert(sizeof(Triangle) == 36, "Triangle must be 9 floats");
	std::memcpy(m->tris.data(), xyz9, n * 36);
	return m;
}

size_t ref_mesh_count(void* mv) { return ((RefMesh*)mv)->tris.size(); }
void ref_mesh_tris(void* mv, float* out9) { RefMesh* m = (RefMesh*)mv; std::memcpy(out9, m->tris.data(), m->tris.size() * 36); }

double ref_bvh_build(void* mv) {   // returns build seconds
	RefMesh* m = (RefMesh*)mv;
	auto t0 = std::chrono::steady_clock::now();
	m->bvh = new BVH(m->tris);
	return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

void ref_mesh_free(void* mv) { RefMesh* m = (RefMesh*)mv; if (!m) return; delete m->bvh; delete m; }

// Tree export in DFS pre-order (left before right), for structural comparison with the product builder:
// per node: bounds(6 floats), isLeaf, triCount, tri ids (<=2). Returns node count.
static void exportNode(const RefMesh* m, const BVHNode* n, std::vector<float>& boxes, std::vector<int32_t>& meta) {
	for (int i = 0; i < 3; i++) boxes.push_back(n->bounds.min[i]);
	for (int i = 0; i < 3; i++) boxes.push_back(n->bounds.max[i]);
	bool leaf = !n->left && !n->right;
	meta.push_back(leaf ? 1 : 0);
	meta.push_back((int32_t)n->triangles.size());
	for (int i = 0; i < 2; i++) meta.push_back(i < (int)n->triangles.size() ? (int32_t)(n->triangles[i] - &m->tris[0]) : -1);
	if (!leaf) { exportNode(m, n->left, boxes, meta); exportNode(m, n->right, boxes, meta); }
}

size_t ref_bvh_export(void* mv, float* boxes6, int32_t* meta4, size_t cap) {
	RefMesh* m = (RefMesh*)mv;
	std::vector<float> b; std::vector<int32_t> me;
	exportNode(m, m->bvh->root, b, me);
	size_t n = me.size() / 4;
	if (boxes6 && meta4 && n <= cap) { std::memcpy(boxes6, b.data(), b.size() * 4); std::memcpy(meta4, me.data(), me.size() * 4); }
	return n;
}

// real BVH::query for a list of rays; out ids are triangle indices in query order. offsets has n+1 entries.
size_t ref_bvh_query(void* mv, const float* o3, const float* d3, size_t nrays, int64_t* offsets, int32_t* ids, size_t cap) {
	RefMesh* m = (RefMesh*)mv;
	std::vector<const Triangle*> cand;
	size_t total = 0;
	for (size_t r = 0; r < nrays; r++) {
		cand.clear();
		m->bvh->query(glm::vec3(o3[3 * r], o3[3 * r + 1], o3[3 * r + 2]), glm::vec3(d3[3 * r], d3[3 * r + 1], d3[3 * r + 2]), cand);
		offsets[r] = (int64_t)total;
		for (auto* t : cand) { if (ids && total < cap) ids[total] = (int32_t)(t - &m->tris[0]); total++; }
	}
	offsets[nrays] = (int64_t)total;
	return total;
}

// Box-test counter: replays queryNode's visit pattern (BVH.cpp:89-105) with the same slab test to count
// intersectAABB calls; the candidate count it produces is asserted equal to the real query's by the caller.
static void countNode(const BVHNode* n, const glm::vec3& o, const glm::vec3& inv, const int neg[3], uint64_t& boxes, uint64_t& cands) {
	if (!n) return;
	boxes++;
	float tmin = 0.0f, tmax = std::numeric_limits<float>::max();
	for (int i = 0; i < 3; i++) {
		float t0 = ((neg[i] ? n->bounds.max[i] : n->bounds.min[i]) - o[i]) * inv[i];
		float t1 = ((neg[i] ? n->bounds.min[i] : n->bounds.max[i]) - o[i]) * inv[i];
		tmin = t0 > tmin ? t0 : tmin;
		tmax = t1 < tmax ? t1 : tmax;
		if (tmax < tmin) return;
	}
	if (!n->left && !n->right) { cands += n->triangles.size(); return; }
	countNode(n->left, o, inv, neg, boxes, cands);
	countNode(n->right, o, inv, neg, boxes, cands);
}

// Render through BVH::query + MT closest hit (+ optional shadow). rgba: 4 floats/pixel, id: int32, t: float.
// rows [y0, y1). stats (optional, 4 x uint64): box tests (primary), candidates (primary), box tests (shadow), candidates (shadow)
// -- only filled when stats != nullptr (slower).
double ref_render_bvh(void* mv, const RefCamConsts* cam, unsigned flags, float shadowBias, int y0, int y1,
	float* rgba, int32_t* hitId, float* tOut, uint64_t* stats, int nthreads) {
	RefMesh* m = (RefMesh*)mv;
	const int W = cam->width;
	uint64_t sB = 0, sC = 0, sBs = 0, sCs = 0, nShadow = 0;
#ifdef _OPENMP
	if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
	auto tstart = std::chrono::steady_clock::now();
#pragma omp parallel reduction(+:sB,sC,sBs,sCs,nShadow)
	{
		std::vector<const Triangle*> cand;
#pragma omp for schedule(dynamic, 4)
		for (int py = y0; py < y1; py++) {
			for (int px = 0; px < W; px++) {
				size_t pix = (size_t)(py - y0) * W + px;
				glm::vec3 o, d;
				genRay(*cam, px, py, o, d);
				cand.clear();
				m->bvh->query(o, d, cand);
				if (stats) {
					glm::vec3 inv(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
					int neg[3] = { inv.x < 0, inv.y < 0, inv.z < 0 };
					uint64_t c2 = 0;
					countNode(m->bvh->root, o, inv, neg, sB, c2);
					if (c2 != cand.size()) sB = (uint64_t)-1;
					sC += c2;
				}
				float best = 1e30f; const Triangle* bestTri = nullptr;
				for (const Triangle* tri : cand) {
					float t;
					if (mollerTrumbore(*tri, o, d, t) && t < best) { best = t; bestTri = tri; }
				}
				glm::vec3 color(0.0f);
				if (bestTri) {
					glm::vec3 e1 = bestTri->v1 - bestTri->v0;
					glm::vec3 e2 = bestTri->v2 - bestTri->v0;
					glm::vec3 n = glm::normalize(glm::cross(e1, e2));
					if (glm::dot(n, d) > 0.0f) n = -n;
					glm::vec3 hit = o + d * best;
					bool shadowed = false;
					if (flags & RTO_FLAG_SHADOWS) {
						glm::vec3 so = hit + n * shadowBias;
						glm::vec3 sd = glm::normalize(glm::vec3(1.0f, 1.0f, 1.0f));
						cand.clear();
						m->bvh->query(so, sd, cand);
						nShadow++;
						if (stats) {
							glm::vec3 inv(1.0f / sd.x, 1.0f / sd.y, 1.0f / sd.z);
							int neg[3] = { inv.x < 0, inv.y < 0, inv.z < 0 };
							countNode(m->bvh->root, so, inv, neg, sBs, sCs);
						}
						for (const Triangle* tri : cand) { float t; if (mollerTrumbore(*tri, so, sd, t)) { shadowed = true; break; } }
					}
					color = shadowed ? glm::vec3(0.1f, 0.1f, 0.1f) : shadeLambert(n);
				}
				if (rgba) { rgba[4 * pix] = color.x; rgba[4 * pix + 1] = color.y; rgba[4 * pix + 2] = color.z; rgba[4 * pix + 3] = 1.0f; }
				if (hitId) hitId[pix] = bestTri ? (int32_t)(bestTri - &m->tris[0]) : -1;
				if (tOut) tOut[pix] = best;
			}
		}
	}
	double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - tstart).count();
	if (stats) { stats[0] = sB; stats[1] = sC; stats[2] = sBs; stats[3] = sCs; stats[4] = nShadow; }
	return sec;
}

// ---- octree mode A: REAL octreeRaySkip for t; instrumented replay only to name the leaf -------
struct SkipResult { float t; const OctreeNode* leaf; uint64_t visits; };

static float replaySkip(const OctreeNode* node, const glm::vec3& ro, const glm::vec3& rd, float tMin, float tMax,
	const VoxelGrid& grid, const OctreeNode*& leafOut, uint64_t& visits) {
	if (!node) return 1e30f;
	visits++;
	float vx = grid.voxelSize;
	float wx0 = grid.minX + node->x * vx;
	float wy0 = grid.minY + node->y * vx;
	float wz0 = grid.minZ + node->z * vx;
	float wSize = node->size * vx;
	glm::vec3 bmin(wx0, wy0, wz0);
	glm::vec3 bmax(wx0 + wSize, wy0 + wSize, wz0 + wSize);
	glm::vec3 invRd = 1.0f / rd;
	const float smallValue = 1e-10f;
	if (std::abs(rd.x) < smallValue) invRd.x = rd.x >= 0 ? 1e10f : -1e10f;
	if (std::abs(rd.y) < smallValue) invRd.y = rd.y >= 0 ? 1e10f : -1e10f;
	if (std::abs(rd.z) < smallValue) invRd.z = rd.z >= 0 ? 1e10f : -1e10f;
	glm::vec3 t1 = (bmin - ro) * invRd;
	glm::vec3 t2 = (bmax - ro) * invRd;
	glm::vec3 tNear = glm::min(t1, t2);
	glm::vec3 tFar = glm::max(t1, t2);
	float enterT = std::max(std::max(tNear.x, tNear.y), std::max(tNear.z, tMin));
	float exitT = std::min(std::min(tFar.x, tFar.y), std::min(tFar.z, tMax));
	if (enterT > exitT) return 1e30f;
	if (node->isLeaf) {
		if (!node->isSolid) return 1e30f;
		leafOut = node;
		return enterT;
	}
	int dirMask = ((rd.x > 0) ? 1 : 0) | ((rd.y > 0) ? 2 : 0) | ((rd.z > 0) ? 4 : 0);
	float bestT = 1e30f;
	for (int dist = 0; dist <= 3; dist++) {
		for (int octant = 0; octant < 8; octant++) {
			if (__builtin_popcount(octant ^ dirMask) != dist) continue;
			const OctreeNode* child = node->children[octant];
			if (!child) continue;
			float childT = replaySkip(child, ro, rd, enterT, exitT, grid, leafOut, visits);
			if (childT < bestT) { bestT = childT; if (childT < 1e30f) return childT; }
		}
	}
	return bestT;
}

// mode 0 = octreeRaySkip semantics (A), mode 1 = GLSL intersectOctreeIterative semantics (B)
double ref_render_octree(void* hv, const RefCamConsts* cam, int mode, int y0, int y1,
	float* rgba, int32_t* leafId, float* tOut, uint64_t* stats, int nthreads) {
	RefOctree* h = (RefOctree*)hv;
	const int W = cam->width;
	// mode 2: the GLSL traversal over the frustum-culled array the last ref_cull_nodes call made renderSceneComputeWithCulling build
	const std::vector<GPUNodes>& nodes = (mode == 2) ? h->tracer->m_visibleNodes : h->tracer->m_flatNodes;
	if (mode == 2) mode = 1;
	const glm::vec3 gridMin(h->grid.minX, h->grid.minY, h->grid.minZ);
	const float voxelSize = h->grid.voxelSize;
	uint64_t sVisits = 0, sMismatch = 0;
#ifdef _OPENMP
	if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
	auto tstart = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 4) reduction(+:sVisits,sMismatch)
	for (int py = y0; py < y1; py++) {
		for (int px = 0; px < W; px++) {
			size_t pix = (size_t)(py - y0) * W + px;
			glm::vec3 o, d;
			genRay(*cam, px, py, o, d);
			float tRes = 1e30f; int id = -1; glm::vec3 color(0.0f);
			if (mode == 0) {
				float tReal = octreeRaySkip(h->root, o, d, 0.0f, 1e30f, h->grid);      // the reference itself
				if (leafId || rgba || stats) {
					const OctreeNode* leaf = nullptr; uint64_t visits = 0;
					float tReplay = replaySkip(h->root, o, d, 0.0f, 1e30f, h->grid, leaf, visits);
					if (std::memcmp(&tReal, &tReplay, 4) != 0) sMismatch++;
					sVisits += visits;
					if (leaf) {
						id = h->bfsIndex[leaf];
						// shading of a mode-A hit is an extension: same box-centre normal + Lambert as the GLSL path
						glm::vec3 nodeMin = gridMin + glm::vec3(leaf->x, leaf->y, leaf->z) * voxelSize;
						glm::vec3 nodeMax = nodeMin + glm::vec3(leaf->size) * voxelSize;
						glm::vec3 center = 0.5f * (nodeMin + nodeMax);
						glm::vec3 p = o + d * tReal;
						color = shadeLambert(glm::normalize(p - center));
					}
				}
				tRes = tReal;
			}
			else {
				// GLSL intersectOctreeIterative restated (RayTracerBVH.cpp:239-327)
				float closestT = 1e30f; bool hitFound = false; glm::vec3 bestNormal(0.0f);
				int stack[128]; int sp = 0; stack[sp++] = 0; int steps = 0;
				while (sp > 0 && steps < 512) {
					sp--;
					int nodeIdx = stack[sp];
					if (nodeIdx < 0) continue;
					steps++;
					const GPUNodes& node = nodes[nodeIdx];
					glm::vec3 nodeMin = gridMin + glm::vec3(node.x, node.y, node.z) * voxelSize;
					glm::vec3 nodeMax = nodeMin + glm::vec3(node.size) * voxelSize;
					glm::vec3 invDir = 1.0f / d;
					glm::vec3 t1 = (nodeMin - o) * invDir;
					glm::vec3 t2 = (nodeMax - o) * invDir;
					glm::vec3 tMin = glm::min(t1, t2);
					glm::vec3 tMax = glm::max(t1, t2);
					float tNear = glm::max(glm::max(tMin.x, tMin.y), tMin.z);
					float tFar = glm::min(glm::min(tMax.x, tMax.y), tMax.z);
					if (!(tNear <= tFar && tFar > 0.0f)) continue;
					if (tNear >= closestT) continue;
					if (node.isUniform == 1 || node.isLeaf == 1) {
						if (node.isSolid == 1) {
							float tHit = glm::max(0.0f, tNear);
							if (tHit < closestT && tHit <= tFar) {
								closestT = tHit; hitFound = true; id = nodeIdx;
								glm::vec3 center = 0.5f * (nodeMin + nodeMax);
								glm::vec3 p = o + d * tHit;
								bestNormal = glm::normalize(p - center);
								break;
							}
						}
						continue;
					}
					for (int i = 0; i < 8; i++) { int c = node.child[i]; if (c >= 0) stack[sp++] = c; }
				}
				sVisits += (uint64_t)steps;
				if (hitFound) { tRes = closestT; color = shadeLambert(bestNormal); }
			}
			if (rgba) { rgba[4 * pix] = color.x; rgba[4 * pix + 1] = color.y; rgba[4 * pix + 2] = color.z; rgba[4 * pix + 3] = 1.0f; }
			if (leafId) leafId[pix] = id;
			if (tOut) tOut[pix] = tRes;
		}
	}
	double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - tstart).count();
	if (stats) { stats[0] = sVisits; stats[1] = sMismatch; }
	return sec;
}

// real octreeRaySkip on an explicit ray list (used for edge cases: axis-parallel / zero components)
void ref_octree_rayskip(void* hv, const float* o3, const float* d3, size_t n, float tMin, float tMax, float* tOut, int32_t* idOut) {
	RefOctree* h = (RefOctree*)hv;
	for (size_t i = 0; i < n; i++) {
		glm::vec3 o(o3[3 * i], o3[3 * i + 1], o3[3 * i + 2]), d(d3[3 * i], d3[3 * i + 1], d3[3 * i + 2]);
		tOut[i] = octreeRaySkip(h->root, o, d, tMin, tMax, h->grid);      // the reference itself
		if (idOut) {                                                        // leaf name from the replay; t must agree bitwise
			const OctreeNode* leaf = nullptr; uint64_t visits = 0;
			float tr = replaySkip(h->root, o, d, tMin, tMax, h->grid, leaf, visits);
			idOut[i] = (std::memcmp(&tr, &tOut[i], 4) != 0) ? -2 : (leaf ? h->bfsIndex[leaf] : -1);
		}
	}
}

int ref_num_threads() {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

} // extern "C"
