// tests/dc_asan/driver.cpp -- TEST INFRASTRUCTURE.  The Dual-Contouring builders (host_dc.cpp over rto_dc.h, the per-cell code the CUDA
// kernels of rto_dc.cu share) compiled with AddressSanitizer + UBSan and run on exactly-sized heap buffers: every voxel / node /
// record index the mesher forms is bounds-checked here, which stands in for compute-sanitizer on the device side.
#include "../../include/rto_c.h"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

int rto_fail(int code, const char* fmt, ...) { va_list ap; va_start(ap, fmt); std::vfprintf(stderr, fmt, ap); va_end(ap); std::fputc('\n', stderr); return code; }

static int run(const char* name, int dx, int dy, int dz, const std::vector<uint8_t>& vox, const float* vp) {
	const float gmin[3] = { -1.5f, 0.25f, 3.0f };
	const float voxel = 0.37f;
	std::vector<uint8_t> grid(vox);                       // exactly sized: ASan guards both ends
	RtoGpuNode* nodes = nullptr; size_t numNodes = 0;
	if (rto_host_octree_build(grid.data(), dx, dy, dz, &nodes, &numNodes) != RTO_OK) return 1;
	std::vector<RtoGpuNode> tight(nodes, nodes + numNodes);
	rto_host_free(nodes);
	RtoTriangle *a = nullptr, *b = nullptr, *c = nullptr; float* nrm = nullptr; size_t na = 0, nb = 0, nc = 0;
	if (rto_host_dc_mesh(grid.data(), dx, dy, dz, gmin, voxel, tight.data(), numNodes, vp, 0.5f, &a, &na) != RTO_OK) return 2;
	if (rto_host_dc_mesh_replay(grid.data(), dx, dy, dz, gmin, voxel, tight.data(), numNodes, vp, 0.5f, &b, &nb) != RTO_OK) return 3;
	if (rto_host_dc_mesh_normals(grid.data(), dx, dy, dz, gmin, voxel, tight.data(), numNodes, vp, 0.5f, &c, &nrm, &nc) != RTO_OK) return 4;
	int rc = 0;
	if (na != nb || na != nc || (na && (std::memcmp(a, b, na * sizeof(RtoTriangle)) || std::memcmp(a, c, na * sizeof(RtoTriangle))))) rc = 5;
	std::printf("%-14s %3d x %3d x %3d  %7zu nodes  %7zu triangles  %s\n", name, dx, dy, dz, numNodes, na, rc ? "MISMATCH" : "ok");
	rto_host_free(a); rto_host_free(b); rto_host_free(c); rto_host_free(nrm);
	return rc;
}

int main() {
	std::mt19937 rng(12345);
	auto noise = [&](int dx, int dy, int dz, double p) { std::vector<uint8_t> v((size_t)dx * dy * dz); std::bernoulli_distribution d(p); for (auto& x : v) x = d(rng) ? 1 : 0; return v; };
	int bad = 0;
	// a view-projection looking down -z from (0, 0, 20): perspective(45 deg, 1.5, 0.01, 5000) * translate(0, 0, -20), column-major
	const float vp[16] = { 1.6094757f, 0, 0, 0,  0, 2.4142136f, 0, 0,  0, 0, -1.000004f, -1.0f,  0, 0, 19.98008f, 20.0f };
	bad += run("noise 30%", 13, 7, 20, noise(13, 7, 20, 0.3), nullptr);
	bad += run("noise 3%", 33, 18, 9, noise(33, 18, 9, 0.03), nullptr);
	bad += run("noise 95%", 17, 17, 17, noise(17, 17, 17, 0.95), nullptr);
	bad += run("noise culled", 21, 30, 24, noise(21, 30, 24, 0.4), vp);
	bad += run("full ragged", 5, 9, 2, std::vector<uint8_t>(90, 1), nullptr);
	bad += run("one voxel", 1, 1, 1, std::vector<uint8_t>(1, 1), nullptr);
	bad += run("thin", 1, 40, 3, noise(1, 40, 3, 0.5), nullptr);
	{	// blocks: large uniform leaves next to surfaces (strided gathers, uniform-box stepping)
		const int d = 48; std::vector<uint8_t> v((size_t)d * d * d, 0);
		for (int k = 0; k < 10; k++) { int x0 = rng() % d, y0 = rng() % d, z0 = rng() % d, sx = 1 + rng() % 30, sy = 1 + rng() % 30, sz = 1 + rng() % 30;
			for (int z = z0; z < std::min(d, z0 + sz); z++) for (int y = y0; y < std::min(d, y0 + sy); y++) for (int x = x0; x < std::min(d, x0 + sx); x++) v[(size_t)x + (size_t)y * d + (size_t)z * d * d] = 1; }
		bad += run("blocks", d, d, d, v, nullptr);
		bad += run("blocks culled", d, d, d, v, vp);
	}
	{	// sphere shell (main.cpp:337-372)
		const int d = 24; std::vector<uint8_t> v((size_t)d * d * d, 0); const float c = 0.5f * (d - 1);
		for (int z = 0; z < d; z++) for (int y = 0; y < d; y++) for (int x = 0; x < d; x++) { float r = std::sqrt((x - c) * (x - c) + (y - c) * (y - c) + (z - c) * (z - c)); if (!(r < 0.2f * d || r > 0.4f * d)) v[(size_t)x + (size_t)y * d + (size_t)z * d * d] = 1; }
		bad += run("sphere", d, d, d, v, nullptr);
	}
	std::printf(bad ? "FAILED\n" : "all ok\n");
	return bad ? 1 : 0;
}
