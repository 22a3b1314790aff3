// RayTracerBVH.h (shim) -- the reference's ray caster class (453-skeleton/RayTracerBVH.h:15-80) without OpenGL.
// setOctree / ensureComputeInitialized / renderSceneCompute keep their signatures; the frame lands in a host Framebuffer
// (frame()) instead of a GL texture.  render(..., Framebuffer&) and setMesh() are the north-star extensions.
#pragma once
#include "BVH.h"
#include "Camera.h"
#include "OctreeVoxel.h"
#include <cstdio>
#include <cstdlib>
#include <new>
#include <queue>
#include <unordered_map>

struct Ray { rto_shim::vec3 origin, direction; };                          // RayTracerBVH.h:15-18
struct GPUNodes { int x, y, z, size; int isLeaf, isSolid; int isUniform; int child[8]; };   // RayTracerBVH.h:21-26
static_assert(sizeof(GPUNodes) == sizeof(RtoGpuNode), "GPUNodes must be 15 x int32");

// Frame planes live in page-locked host memory (rto_host_alloc_pinned): the device -> host copy of a frame then runs at the speed of the
// link instead of through the driver's staging of pageable memory.  Without a CUDA device (host-only use of the builders) the allocator
// falls back to ordinary memory; a 64-byte header in front of every block remembers which kind it is.
template <class T> struct RtoFrameAllocator {
	using value_type = T;
	RtoFrameAllocator() = default;
	template <class U> RtoFrameAllocator(const RtoFrameAllocator<U>&) {}
	T* allocate(size_t n) {
		const size_t bytes = n * sizeof(T) + 64;
		void* p = nullptr; unsigned kind = 1;
		if (rto_host_alloc_pinned(bytes, &p) != RTO_OK) { p = std::malloc(bytes); kind = 0; }
		if (!p) throw std::bad_alloc();
		*static_cast<unsigned*>(p) = kind;
		return reinterpret_cast<T*>(static_cast<char*>(p) + 64);
	}
	void deallocate(T* q, size_t) noexcept {
		void* p = reinterpret_cast<char*>(q) - 64;
		if (*static_cast<unsigned*>(p)) rto_host_free_pinned(p); else std::free(p);
	}
	template <class U> bool operator==(const RtoFrameAllocator<U>&) const { return true; }
	template <class U> bool operator!=(const RtoFrameAllocator<U>&) const { return false; }
};
template <class T> using RtoFrameVector = std::vector<T, RtoFrameAllocator<T>>;

struct Framebuffer {                       // host copy of one frame: row 0 = top scanline
	int width = 0, height = 0;
	RtoFrameVector<float> rgba;            // 4 floats per pixel (RGBA32F like the reference's output texture)
	RtoFrameVector<int32_t> hitId;         // octree: node index == leaf id of setOctree's BFS numbering; mesh: triangle index; -1 miss
	RtoFrameVector<float> t;               // hit distance, 1e30f on miss
};

class RayTracerBVH {
public:
	RayTracerBVH() = default;
	~RayTracerBVH() { rto_scene_destroy(m_scene); rto_scene_destroy(m_culled); rto_group_destroy(m_group); }
	RayTracerBVH(const RayTracerBVH&) = delete;
	RayTracerBVH& operator=(const RayTracerBVH&) = delete;

	// RayTracerBVH.cpp:430-505: BFS flatten (root = 0, discovery order) and upload.  setOctree(nullptr, ...) clears.
	void setOctree(OctreeNode* root, const VoxelGrid& grid) {
		m_flatNodes.clear(); m_visibleToFlat.clear();
		rto_scene_destroy(m_scene); m_scene = nullptr; m_mode = RTO_MODE_OCTREE_GLSL;
		rto_scene_destroy(m_culled); m_culled = nullptr;
		m_gridMin[0] = grid.minX; m_gridMin[1] = grid.minY; m_gridMin[2] = grid.minZ; m_voxelSize = grid.voxelSize;
		if (!root) return;
		std::queue<OctreeNode*> q; std::unordered_map<OctreeNode*, int> index;
		q.push(root); index[root] = 0;
		m_flatNodes.push_back(GPUNodes{ 0, 0, 0, 0, 0, 0, 0, { -1, -1, -1, -1, -1, -1, -1, -1 } });
		while (!q.empty()) {
			OctreeNode* nd = q.front(); q.pop();
			int idx = index[nd];
			GPUNodes g{ nd->x, nd->y, nd->z, nd->size, nd->isLeaf ? 1 : 0, nd->isSolid ? 1 : 0, nd->isUniform ? 1 : 0, { -1, -1, -1, -1, -1, -1, -1, -1 } };
			if (!nd->isLeaf)
				for (int i = 0; i < 8; i++) if (OctreeNode* c = nd->children[i]) {
					auto it = index.find(c);
					if (it == index.end()) { it = index.emplace(c, (int)m_flatNodes.size()).first; m_flatNodes.push_back(GPUNodes{ 0, 0, 0, 0, 0, 0, 0, { -1, -1, -1, -1, -1, -1, -1, -1 } }); }
					g.child[i] = it->second; q.push(c);
				}
			m_flatNodes[idx] = g;
		}
		float gmin[3] = { grid.minX, grid.minY, grid.minZ };
		if (rto_scene_create_octree(reinterpret_cast<const RtoGpuNode*>(m_flatNodes.data()), m_flatNodes.size(), gmin, grid.voxelSize, &m_scene) != RTO_OK)
			std::fprintf(stderr, "[RayTracerBVH] %s\n", rto_last_error());
	}
	// extension: ray cast a triangle soup through the reference-shaped BVH
	void setMesh(const BVH& bvh, float sceneScale) {
		rto_scene_destroy(m_scene); m_scene = nullptr; rto_scene_destroy(m_culled); m_culled = nullptr; m_culledEmpty = false;
		m_external = bvh.scene(); m_mode = RTO_MODE_BVH; m_shadowBias = 1e-3f * sceneScale; m_bvh = &bvh; m_groupHasMesh = false;
	}
	// extension: use several GPUs of the box for renderBatch (rto_group_*: scene replicated, frames dealt to the devices, hit codes
	// gathered on devices[0] over NVLink and expanded there).  An empty list goes back to one device.
	bool setDevices(const std::vector<int>& devices) {
		rto_group_destroy(m_group); m_group = nullptr; m_groupHasMesh = false;
		if (devices.empty()) return true;
		if (rto_group_create(devices.data(), (int)devices.size(), &m_group) != RTO_OK) { std::fprintf(stderr, "[RayTracerBVH] %s\n", rto_last_error()); return false; }
		return true;
	}
	// extension: many cameras in one submission (a camera orbit, main.cpp's per-frame loop unrolled): mesh scenes only
	bool renderBatch(const std::vector<Camera>& cameras, int width, int height, float aspect, float fovDeg, std::vector<Framebuffer>& out) {
		if (!m_inited || m_mode != RTO_MODE_BVH || !m_bvh || cameras.empty()) return false;
		std::vector<RtoCamera> cams(cameras.size());
		for (size_t i = 0; i < cameras.size(); i++) if (cameras[i].consts(fovDeg, aspect, width, height, cams[i], nullptr) != RTO_OK) return false;
		const size_t npix = (size_t)width * height, n = cameras.size();
		// the batch lands in page-locked planes kept from call to call (pinning memory costs more than the copy it speeds up)
		RtoFrameVector<float>& rgba = m_batchRgba; RtoFrameVector<float>& t = m_batchT; RtoFrameVector<int32_t>& ids = m_batchIds;
		if (rgba.size() < n * npix * 4) { rgba.resize(n * npix * 4); t.resize(n * npix); ids.resize(n * npix); }
		RtoFrame fr{ rgba.data(), ids.data(), t.data(), RTO_MEM_HOST };
		int rc;
		if (m_group) {
			if (!m_groupHasMesh) {
				const std::vector<Triangle>& tris = m_bvh->triangles();
				if (rto_group_scene_bvh(m_group, reinterpret_cast<const RtoTriangle*>(tris.data()), tris.size(), m_bvh->host()) != RTO_OK) { std::fprintf(stderr, "[RayTracerBVH] %s\n", rto_last_error()); return false; }
				m_groupHasMesh = true;
			}
			rc = rto_group_render_batch(m_group, cams.data(), (int)n, m_flags, m_shadowBias, &fr);
		}
		else rc = rto_render_batch(m_external, cams.data(), (int)n, RTO_MODE_BVH, m_flags, m_shadowBias, 0, height, &fr);
		if (rc != RTO_OK) { std::fprintf(stderr, "[RayTracerBVH] %s\n", rto_last_error()); return false; }
		out.resize(n);
		for (size_t i = 0; i < n; i++) {
			out[i].width = width; out[i].height = height;
			out[i].rgba.assign(rgba.begin() + i * npix * 4, rgba.begin() + (i + 1) * npix * 4);
			out[i].hitId.assign(ids.begin() + i * npix, ids.begin() + (i + 1) * npix);
			out[i].t.assign(t.begin() + i * npix, t.begin() + (i + 1) * npix);
		}
		return true;
	}

	void ensureComputeInitialized() { if (!m_inited && rto_init(0) == RTO_OK) m_inited = true; else if (!m_inited) std::fprintf(stderr, "[RayTracerBVH] %s\n", rto_last_error()); }
	void setFrustumCullingEnabled(bool on) { m_cullingEnabled = on; }       // RayTracerBVH.h:44 (stored, like the reference; the culling call below does not consult it either)
	void setTraversalMode(int mode) { m_mode = mode; }                      // RTO_MODE_OCTREE_GLSL (reference shader) or RTO_MODE_OCTREE_SKIP
	void setShadows(bool on) { m_flags = on ? RTO_FLAG_SHADOWS : 0u; }

	// RayTracerBVH.cpp:614-704.  Like the reference, errors are printed and the frame is left untouched.
	void renderSceneCompute(const Camera& camera, int width, int height, float aspect, float fovDeg) { render(camera, width, height, aspect, fovDeg, m_frame); }
	// RayTracerBVH.cpp:706-892.  With updateFrustum the node array is culled against the camera's frustum (margin 150), compacted and
	// re-uploaded (:724-813), here on the GPU; like the reference's SSBO, the culled array then stays the one every later render
	// traverses until the next update or setOctree.  frame().hitId indexes that array; visibleToFlat() maps it back to flatNodes().
	void renderSceneComputeWithCulling(const Camera& camera, int width, int height, float aspect, float fovDeg, bool updateFrustum) {
		if (updateFrustum && m_scene && !m_flatNodes.empty()) {
			RtoCamera cam; rto_shim::mat4 viewM; float vp[16];
			const float* view = reinterpret_cast<const float*>(&viewM);
			RtoGpuNode* culled = nullptr; size_t n = 0; int32_t* back = nullptr;
			if (camera.consts(fovDeg, aspect, width, height, cam, &viewM) == RTO_OK && rto_host_view_proj(view, fovDeg, aspect, 0.01f, 5000.f, vp) == RTO_OK &&
				rto_device_frustum_cull(reinterpret_cast<const RtoGpuNode*>(m_flatNodes.data()), m_flatNodes.size(), m_gridMin, m_voxelSize, vp, 150.0f, &culled, &n, &back) == RTO_OK) {
				rto_scene_destroy(m_culled); m_culled = nullptr;
				m_visibleToFlat.assign(back, back + n);
				if (n && rto_scene_create_octree(culled, n, m_gridMin, m_voxelSize, &m_culled) != RTO_OK) std::fprintf(stderr, "[RayTracerBVH] %s\n", rto_last_error());
				m_culledEmpty = (n == 0);
			}
			else std::fprintf(stderr, "[RayTracerBVH] %s\n", rto_last_error());
			rto_host_free(culled); rto_host_free(back);
		}
		renderSceneCompute(camera, width, height, aspect, fovDeg);
	}
	const std::vector<int32_t>& visibleToFlat() const { return m_visibleToFlat; }
	bool render(const Camera& camera, int width, int height, float aspect, float fovDeg, Framebuffer& fb) {
		RtoScene* sc = m_culled ? m_culled : (m_scene ? m_scene : m_external);
		if (!m_inited) { std::fprintf(stderr, "[RayTracerBVH] Compute pipeline not initialized or failed.\n"); return false; }
		if (!sc || (m_culledEmpty && m_scene)) return false;               // no data (or everything culled): skip (RayTracerBVH.cpp:624-627)
		RtoCamera cam;
		if (camera.consts(fovDeg, aspect, width, height, cam, nullptr) != RTO_OK) { std::fprintf(stderr, "[RayTracerBVH] %s\n", rto_last_error()); return false; }
		fb.width = width; fb.height = height;
		fb.rgba.resize((size_t)width * height * 4); fb.hitId.resize((size_t)width * height); fb.t.resize((size_t)width * height);
		RtoFrame fr{ fb.rgba.data(), fb.hitId.data(), fb.t.data(), RTO_MEM_HOST };
		if (rto_render(sc, &cam, m_mode, m_flags, m_shadowBias, 0, height, &fr) != RTO_OK) { std::fprintf(stderr, "[RayTracerBVH] %s\n", rto_last_error()); return false; }
		return true;
	}
	// extension: the uploaded scene as a file and back (rto_scene_save / rto_scene_load) -- the cache of what the reference rebuilds at
	// every start (octree flatten, main.cpp:1127-1131), next to its own sceneCache.bin and triangle cache.  A loaded octree scene renders
	// at once; flatNodes() stays empty, so renderSceneComputeWithCulling(.., updateFrustum = true) needs a setOctree() first.
	bool saveSceneCache(const char* path) const {
		RtoScene* sc = m_scene ? m_scene : m_external;
		if (!sc || rto_scene_save(sc, path) != RTO_OK) { std::fprintf(stderr, "[RayTracerBVH] %s\n", sc ? rto_last_error() : "no scene to save"); return false; }
		return true;
	}
	bool loadSceneCache(const char* path, const VoxelGrid& grid) {
		RtoScene* sc = nullptr;
		if (rto_scene_load(path, &sc) != RTO_OK) { std::fprintf(stderr, "[RayTracerBVH] %s\n", rto_last_error()); return false; }
		int kind = 0;
		rto_scene_info(sc, &kind, nullptr, nullptr, nullptr, nullptr);
		if (kind == RTO_MODE_BVH) { rto_scene_destroy(sc); std::fprintf(stderr, "[RayTracerBVH] %s holds a mesh scene; setOctree's cache is an octree scene\n", path); return false; }
		m_flatNodes.clear(); m_visibleToFlat.clear();
		rto_scene_destroy(m_scene); m_scene = sc; m_mode = RTO_MODE_OCTREE_GLSL;
		rto_scene_destroy(m_culled); m_culled = nullptr; m_culledEmpty = false;
		m_gridMin[0] = grid.minX; m_gridMin[1] = grid.minY; m_gridMin[2] = grid.minZ; m_voxelSize = grid.voxelSize;
		return true;
	}
	const Framebuffer& frame() const { return m_frame; }
	const std::vector<GPUNodes>& flatNodes() const { return m_flatNodes; }

private:
	std::vector<GPUNodes> m_flatNodes;
	RtoScene* m_scene = nullptr;
	RtoFrameVector<float> m_batchRgba, m_batchT; RtoFrameVector<int32_t> m_batchIds;      // renderBatch's landing planes
	RtoScene* m_culled = nullptr;         // the frustum-culled array of the last renderSceneComputeWithCulling(..., true)
	std::vector<int32_t> m_visibleToFlat;
	float m_gridMin[3] = { 0, 0, 0 }, m_voxelSize = 1.0f;
	bool m_cullingEnabled = false, m_culledEmpty = false;
	RtoScene* m_external = nullptr;       // owned by a BVH (setMesh)
	const BVH* m_bvh = nullptr;
	RtoGroup* m_group = nullptr;          // setDevices
	bool m_groupHasMesh = false;
	int m_mode = RTO_MODE_OCTREE_GLSL;
	unsigned m_flags = 0;
	float m_shadowBias = 0.0f;
	bool m_inited = false;
	Framebuffer m_frame;
};

// flatten a pointer tree for librto (any order works for the meshers as long as child links are consistent: BFS like setOctree)
inline std::vector<RtoGpuNode> rto_shim_flatten(const OctreeNode* root) {
	std::vector<RtoGpuNode> flat; std::vector<const OctreeNode*> order;
	if (!root) return flat;
	order.push_back(root);
	for (size_t i = 0; i < order.size(); i++) {
		const OctreeNode* nd = order[i];
		RtoGpuNode g{ nd->x, nd->y, nd->z, nd->size, nd->isLeaf ? 1 : 0, nd->isSolid ? 1 : 0, nd->isUniform ? 1 : 0, { -1, -1, -1, -1, -1, -1, -1, -1 } };
		if (!nd->isLeaf) for (int c = 0; c < 8; c++) if (nd->children[c]) { g.child[c] = (int)order.size(); order.push_back(nd->children[c]); }
		flat.push_back(g);
	}
	return flat;
}

// renderOctree(root, grid, dcRenderer, camera, aspect, extraMargin) (main.cpp:95-208) for the Dual-Contouring renderer: the triangle
// soup AdaptiveDualContouringRenderer::createTriangles emits leaf by leaf, in the same order, as the Triangles BVH takes
// (main.cpp:1275 -> BVH).  viewProj16 (column-major proj * view, e.g. rto_host_view_proj(view, 45, aspect, 0.01, 5000)) culls like
// renderOctree; nullptr visits every leaf.  The reference passes extraMargin 50.
inline std::vector<Triangle> rto_shim_dual_contouring(const OctreeNode* root, const VoxelGrid& grid, const float* viewProj16 = nullptr, float extraMargin = 50.0f) {
	std::vector<Triangle> out;
	if (!root) return out;
	std::vector<RtoGpuNode> flat = rto_shim_flatten(root);
	RtoTriangle* tris = nullptr; size_t n = 0;
	float gmin[3] = { grid.minX, grid.minY, grid.minZ };
	if (rto_host_dc_mesh(reinterpret_cast<const uint8_t*>(grid.data.data()), grid.dimX, grid.dimY, grid.dimZ, gmin, grid.voxelSize, flat.data(), flat.size(),
		viewProj16, extraMargin, &tris, &n) != RTO_OK) { std::fprintf(stderr, "[AdaptiveDualContouringRenderer] %s\n", rto_last_error()); return out; }
	out.resize(n);
	for (size_t i = 0; i < n; i++) {
		out[i].v0 = rto_shim::vec3(tris[i].v0[0], tris[i].v0[1], tris[i].v0[2]);
		out[i].v1 = rto_shim::vec3(tris[i].v1[0], tris[i].v1[1], tris[i].v1[2]);
		out[i].v2 = rto_shim::vec3(tris[i].v2[0], tris[i].v2[1], tris[i].v2[2]);
	}
	rto_host_free(tris);
	return out;
}

inline std::vector<MCTriangle> rto_shim_marching_cubes(const OctreeNode* root, const VoxelGrid& grid) {
	std::vector<MCTriangle> out;
	if (!root) return out;
	std::vector<RtoGpuNode> flat = rto_shim_flatten(root);
	RtoTriangle* tris = nullptr; size_t n = 0;
	float gmin[3] = { grid.minX, grid.minY, grid.minZ };
	if (rto_host_mc_mesh(reinterpret_cast<const uint8_t*>(grid.data.data()), grid.dimX, grid.dimY, grid.dimZ, gmin, grid.voxelSize, flat.data(), flat.size(), &tris, &n) != RTO_OK) return out;
	out.resize(n);
	for (size_t i = 0; i < n; i++) {
		const float* v = tris[i].v0;
		for (int k = 0; k < 3; k++) out[i].v[k] = rto_shim::vec3(v[3 * k], v[3 * k + 1], v[3 * k + 2]);
		// flat normal as localMC computes it: normalize(cross(e1, e2)) in glm operation order
		float e1[3] = { v[3] - v[0], v[4] - v[1], v[5] - v[2] }, e2[3] = { v[6] - v[0], v[7] - v[1], v[8] - v[2] };
		float cx = e1[1] * e2[2] - e2[1] * e1[2], cy = e1[2] * e2[0] - e2[2] * e1[0], cz = e1[0] * e2[1] - e2[0] * e1[1];
		float il = 1.0f / __builtin_sqrtf((cx * cx + cy * cy) + cz * cz);
		for (int k = 0; k < 3; k++) out[i].normal[k] = rto_shim::vec3(cx * il, cy * il, cz * il);
	}
	rto_host_free(tris);
	return out;
}
