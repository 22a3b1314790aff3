// example_headless.cpp -- what the reference's main() does around the hot path (main.cpp:1029-1131, 1357-1363), headless,
// written against the shim headers.  Build:  g++ -std=c++17 example_headless.cpp -L../.. -lrto -Wl,-rpath,'$ORIGIN/../..'
// Without a GPU the host-side steps still run and the render step reports RTO_ERR_NO_DEVICE (there is no CPU fallback).
#include "RayTracerBVH.h"
#include <cmath>
#include <cstdio>

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
	if (tracer.render(cam, 160, 120, 160.f / 120.f, 45.f, fb)) {
		size_t hits = 0; for (int id : fb.hitId) hits += id >= 0;
		std::printf("octree frame: %zu of %zu pixels hit\n", hits, fb.hitId.size());
		// the call the reference's main loop makes (main.cpp:1357): cull against the frustum, then render the culled array
		tracer.renderSceneComputeWithCulling(cam, 160, 120, 160.f / 120.f, 45.f, true);
		size_t chits = 0; for (int id : tracer.frame().hitId) chits += id >= 0;
		std::printf("culled frame: %zu of %zu nodes visible, %zu pixels hit\n", tracer.visibleToFlat().size(), tracer.flatNodes().size(), chits);
		try {
			tracer.setMesh(bvh, grid.voxelSize); tracer.setShadows(true);
			if (tracer.render(cam, 160, 120, 160.f / 120.f, 45.f, fb)) { hits = 0; for (int id : fb.hitId) hits += id >= 0; std::printf("mesh frame: %zu of %zu pixels hit\n", hits, fb.hitId.size()); }
			std::vector<const Triangle*> cand;
			bvh.query(cam.getPos(), rto_shim::vec3(-0.55f, -0.5f, -0.66f), cand);
			std::printf("BVH::query candidates: %zu\n", cand.size());
		} catch (const std::exception& e) { std::printf("mesh path: %s\n", e.what()); }
	}
	else std::printf("render skipped: %s\n", rto_last_error());
	freeOctree(root);
	return 0;
}
