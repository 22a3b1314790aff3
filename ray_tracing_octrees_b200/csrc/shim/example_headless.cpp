// example_headless.cpp -- what the reference's main() does around the hot path (main.cpp:1029-1131, 1357-1363), headless,
// written against the shim headers.  Build:  g++ -std=c++17 example_headless.cpp -L../.. -lrto -Wl,-rpath,'$ORIGIN/../..'
// Without a GPU the host-side steps still run and the render step reports RTO_ERR_NO_DEVICE (there is no CPU fallback).
#include "RayTracerBVH.h"
#include "Frustum.h"
#include "VolumeRaycastRenderer.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

// FNV-1a over the three planes of a frame (tests/test_shim_cpp.py computes the same over the Python binding's render)
static unsigned long long frameHash(const Framebuffer& fb) {
	unsigned long long h = 1469598103934665603ull;
	auto eat = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } };
	eat(fb.hitId.data(), fb.hitId.size() * 4); eat(fb.t.data(), fb.t.size() * 4); eat(fb.rgba.data(), fb.rgba.size() * 4);
	return h;
}

int main() {
	const int dim = 32;                                         // generateTestVolume(dim) shell sphere (main.cpp:337-372, 1050-1070)
	VoxelGrid grid;
	grid.dimX = grid.dimY = grid.dimZ = dim; grid.minX = grid.minY = grid.minZ = -0.5f; grid.voxelSize = 1.f / dim;
	grid.data.resize(dim * dim * dim, VoxelState::EMPTY);
	float c = 0.5f * (dim - 1), rO = 0.4f * dim, rI = 0.2f * dim;
	for (int z = 0; z < dim; z++) for (int y = 0; y < dim; y++) for (int x = 0; x < dim; x++) {
		float dx = x - c, dy = y - c, dz = z - c, dist = std::sqrt(dx * dx + dy * dy + dz * dz);
		if (!(dist < rI || dist > rO)) grid.data[grid.index(x, y, z)] = VoxelState::FILLED;
	}
	OctreeNode* root = createOctreeFromVoxelGrid(grid);
	std::vector<MCTriangle> mc = rto_shim_marching_cubes(root, grid);
	std::vector<Triangle> tris(mc.size());
	for (size_t i = 0; i < mc.size(); i++) { tris[i].v0 = mc[i].v[0]; tris[i].v1 = mc[i].v[1]; tris[i].v2 = mc[i].v[2]; }
	BVH bvh(tris);
	RayTracerBVH tracer;
	tracer.ensureComputeInitialized();
	tracer.setOctree(root, grid);
	std::printf("octree nodes %zu, triangles %zu, bvh nodes %zu\n", tracer.flatNodes().size(), tris.size(), rto_host_bvh_num_nodes(bvh.host()));
	std::vector<Triangle> dc = rto_shim_dual_contouring(root, grid);      // the other mesher of the application (main.cpp:1275)
	std::printf("dual contouring triangles %zu\n", dc.size());
	Camera cam(0.5235988f, 0.6981317f, 1.2f);
	Framebuffer fb;
	// host-side types of the boundary: the BVH's pointer tree (BVH.h:37-42), the frustum (Frustum.h:6-24), OctreeNode's side table
	const BVHNode* bn = bvh.root();
	size_t leaves = 0, leafTris = 0; std::vector<const BVHNode*> todo{ bn };
	while (!todo.empty()) { const BVHNode* q = todo.back(); todo.pop_back(); if (!q) continue; if (!q->left && !q->right) { leaves++; leafTris += q->triangles.size(); } else { todo.push_back(q->left); todo.push_back(q->right); } }
	std::printf("bvh pointer tree: %zu leaves holding %zu triangles, root box [%g %g %g]-[%g %g %g]\n", leaves, leafTris, bn->bounds.min.x, bn->bounds.min.y, bn->bounds.min.z, bn->bounds.max.x, bn->bounds.max.y, bn->bounds.max.z);
	{
		Camera fc(0.5235988f, 0.6981317f, 1.2f);
		RtoCamera rc; rto_shim::mat4 view; float vp[16];
		fc.consts(45.f, 160.f / 120.f, 160, 120, rc, &view);
		rto_host_view_proj(reinterpret_cast<const float*>(&view), 45.f, 160.f / 120.f, 0.01f, 5000.f, vp);
		rto_shim::mat4 vpm; for (int i = 0; i < 16; i++) vpm.m[i] = vp[i];
		Frustum fr(vpm);
		std::printf("frustum: unit box %d, box behind the camera %d, flat index of the root %d, of its child 3 %d\n", fr.testAABB(rto_shim::vec3(-0.5f), rto_shim::vec3(0.5f), 0.f),
			fr.testAABB(rto_shim::vec3(50.f), rto_shim::vec3(51.f), 0.f), rto_shim_flat_index(root), rto_shim_flat_index(root->children[3]));
	}
	if (tracer.render(cam, 160, 120, 160.f / 120.f, 45.f, fb)) {
		size_t hits = 0; for (int id : fb.hitId) hits += id >= 0;
		std::printf("octree frame: %zu of %zu pixels hit\n", hits, fb.hitId.size());
		std::printf("octree frame hash %016llx\n", frameHash(fb));
		{	// the uploaded scene as a file and back: a second tracer starts from the cache instead of createOctreeFromVoxelGrid + setOctree
			const char* tmpdir = std::getenv("TMPDIR");
			const std::string cache = std::string(tmpdir ? tmpdir : "/tmp") + "/rto_example_octree.rtoscene";
			RayTracerBVH cached; Framebuffer fb2;
			cached.ensureComputeInitialized();
			if (tracer.saveSceneCache(cache.c_str()) && cached.loadSceneCache(cache.c_str(), grid) && cached.render(cam, 160, 120, 160.f / 120.f, 45.f, fb2))
				std::printf("scene cache: frame hash %016llx\n", frameHash(fb2));
			std::remove(cache.c_str());
		}
		{	// the single-ray call VolumeRaycastRenderer makes 49 times per frame (VolumeRaycastRenderer.cpp:1630), through the centre pixel's ray
			rto_shim::vec3 ro = cam.getPos(), rd(-ro.x, -ro.y, -ro.z);
			float len = std::sqrt(rd.x * rd.x + rd.y * rd.y + rd.z * rd.z); rd = rto_shim::vec3(rd.x / len, rd.y / len, rd.z / len);
			std::printf("octreeRaySkip towards the centre: %.9g\n", octreeRaySkip(root, ro, rd, 0.0f, 1e30f, grid));
			rto_shim_forget_octree(root);
		}
		// the call the reference's main loop makes (main.cpp:1357): cull against the frustum, then render the culled array
		tracer.renderSceneComputeWithCulling(cam, 160, 120, 160.f / 120.f, 45.f, true);
		size_t chits = 0; for (int id : tracer.frame().hitId) chits += id >= 0;
		std::printf("culled frame: %zu of %zu nodes visible, %zu pixels hit\n", tracer.visibleToFlat().size(), tracer.flatNodes().size(), chits);
		try {
			tracer.setMesh(bvh, grid.voxelSize); tracer.setShadows(true);
			if (tracer.render(cam, 160, 120, 160.f / 120.f, 45.f, fb)) { hits = 0; for (int id : fb.hitId) hits += id >= 0; std::printf("mesh frame: %zu of %zu pixels hit\n", hits, fb.hitId.size()); std::printf("mesh frame hash %016llx\n", frameHash(fb)); }
			// a batch of cameras, on one device and through a device group (every GPU of the box the caller lists; here the first one or two)
			std::vector<Camera> orbit; for (int k = 0; k < 3; k++) orbit.emplace_back(0.5235988f, 0.6981317f + 0.5f * k, 1.2f);
			std::vector<Framebuffer> one, many;
			int ndev = 0; { size_t dummy; int sm; if (rto_device_info(&sm, nullptr, nullptr, &dummy, &dummy) == RTO_OK) ndev = 1; }
			if (ndev && tracer.renderBatch(orbit, 160, 120, 160.f / 120.f, 45.f, one)) {
				std::vector<int> devs{ 0 }; { RtoGroup* probe = nullptr; int two[2] = { 0, 1 }; if (rto_group_create(two, 2, &probe) == RTO_OK) { devs.push_back(1); rto_group_destroy(probe); } }
				tracer.setDevices(devs);
				bool same = tracer.renderBatch(orbit, 160, 120, 160.f / 120.f, 45.f, many) && many.size() == one.size();
				for (size_t i = 0; same && i < one.size(); i++) same = frameHash(one[i]) == frameHash(many[i]);
				std::printf("batch of %zu frames: first frame hash %016llx, device group of %zu: %s\n", one.size(), frameHash(one[0]), devs.size(), same ? "identical" : "DIFFERENT");
				tracer.setDevices({});
			}
			std::vector<const Triangle*> cand;
			bvh.query(cam.getPos(), rto_shim::vec3(-0.55f, -0.5f, -0.66f), cand);
			std::printf("BVH::query candidates: %zu\n", cand.size());
		} catch (const std::exception& e) { std::printf("mesh path: %s\n", e.what()); }
	}
	else std::printf("render skipped: %s\n", rto_last_error());
	freeOctree(root);
	return 0;
}
