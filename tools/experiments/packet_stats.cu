// experiment: per-ray vs packet traversal work on the DT scene (host emulation)
#include "/root/repo/ray_tracing_octrees_b200/csrc/rto_kernels.cuh"
#include <cstdarg>
#include <cstdio>
#include <vector>
#include <set>
#include <algorithm>
#include <zlib.h>
using namespace rto;
int rto_fail(int code, const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); return code; }

struct Counters { double inner = 0, leaves = 0, tris = 0; };

// per-ray pruned traversal recording visited inner nodes and leaves
static void trace_ray(const BvhDev& S, V3 o, V3 d, std::vector<int>& inner, std::vector<int>& leaves, float& bestT, int& bestPos) {
	bestT = kMissT; bestPos = -1;
	RayBox rb = make_raybox(o, d);
	float te;
	if (!slab_ref(rb, S.rootLo[0], S.rootLo[1], S.rootLo[2], S.rootHi[0], S.rootHi[1], S.rootHi[2], te)) return;
	StackEnt stack[64]; int sp = 0; int cur = S.rootRef; float tcut = kMissT * kPruneSlack;
	for (;;) {
		bool pop = true;
		if (cur >= 0) {
			inner.push_back(cur);
			const float4* n = S.nodes + 4 * (size_t)cur;
			float4 a = n[0], b = n[1], c = n[2]; float2 r = *reinterpret_cast<const float2*>(n + 3);
			float e0, e1; bool h0, h1;
			node_boxes<kOctGeneric>(rb, make_raybox2(rb), S.paired, a, b, c, tcut, h0, h1, e0, e1);
			int r0 = f2i(r.x), r1 = f2i(r.y);
			if (h0 && h1) { bool sw = e1 < e0; StackEnt e; e.ref = sw ? r0 : r1; e.t = sw ? e0 : e1; stack[sp++] = e; cur = sw ? r1 : r0; pop = false; }
			else if (h0) { cur = r0; pop = false; } else if (h1) { cur = r1; pop = false; }
		} else {
			leaves.push_back(cur);
			int ref = ~cur; int pos = ref >> 1, cnt = (ref & 1) + 1;
			for (int k = 0; k < cnt; k++) { TriV tri = load_tri(S.tris, pos + k); float t; if ((!S.leafBox || ref_leaf_box_passes<kOctGeneric>(S, rb, pos + k, tcut)) && moller_trumbore(tri, o, d, t)) if (t < bestT || (t == bestT && pos + k < bestPos)) { bestT = t; bestPos = pos + k; tcut = t * kPruneSlack; } }
		}
		if (pop) { bool got = false; while (sp > 0) { StackEnt e = stack[--sp]; if (e.t <= tcut) { cur = e.ref; got = true; break; } } if (!got) break; }
	}
}

int main(int argc, char** argv) {
	// load DT grid (gz)
	gzFile f = gzopen("/root/repo/tests/golden/dt_sceneCache.bin.gz", "rb");
	int dims[3]; float mv[4]; size_t n;
	gzread(f, dims, 12); gzread(f, mv, 16); gzread(f, &n, 8);
	std::vector<uint8_t> vox(n); gzread(f, vox.data(), (unsigned)n); gzclose(f);
	RtoGpuNode* nodes; size_t nn; rto_host_octree_build(vox.data(), dims[0], dims[1], dims[2], &nodes, &nn);
	RtoTriangle* tris; size_t nt; rto_host_mc_mesh(vox.data(), dims[0], dims[1], dims[2], mv, mv[3], nodes, nn, &tris, &nt);
	RtoHostBvh* hb; rto_host_bvh_build(tris, nt, &hb);
	BvhLayout L; rto_build_bvh_layout(*hb, L);
	BvhDev S; S.numTris = (int)nt; S.rootRef = L.fastRoot; S.leafBox = 1; S.grow = L.fastGrow; S.paired = 1; S.exactPaired = 0; S.exactNodes = (const float4*)L.refNodes.data(); S.exactRoot = L.refRoot; S.exactLeafBox = 0; S.nodes = (const float4*)L.fastNodes.data(); S.tris = (const float4*)L.tris.data();
	for (int k = 0; k < 3; k++) { S.rootLo[k] = L.rootLo[k]; S.rootHi[k] = L.rootHi[k]; }
	printf("tris %zu\n", nt);
	const int W = 1920, H = 1080;
	RtoCamera cam; float tgt[3] = { 0, 0, 0 };
	float th = 35.0f * 3.14159265f / 180.0f, ph = (argc > 1 ? atof(argv[1]) : 40.0f) * 3.14159265f / 180.0f;
	rto_host_camera_orbit(th, ph, 0.6f * 4250.0f, tgt, 45.0f, float(W) / float(H), W, H, &cam, nullptr);
	double sumInner = 0, sumLeaves = 0, rays = 0, unionInner = 0, unionLeaves = 0, tiles = 0, pkInner = 0, pkLeaves = 0, pkLeafLanes = 0, hitTiles = 0;
	for (int ty = 0; ty < H / 4; ty += 5) for (int tx = 0; tx < W / 8; tx += 5) {
		std::set<int> uI, uL; Ray rr[32]; bool anyHit = false;
		for (int l = 0; l < 32; l++) {
			int px = tx * 8 + (l & 7), py = ty * 4 + (l >> 3);
			Ray ray = gen_ray(cam, px, py); rr[l] = ray;
			std::vector<int> in, lv; float bt; int bp;
			trace_ray(S, ray.o, ray.d, in, lv, bt, bp);
			sumInner += in.size(); sumLeaves += lv.size(); rays += 1;
			uI.insert(in.begin(), in.end()); uL.insert(lv.begin(), lv.end());
			anyHit |= bp >= 0;
		}
		unionInner += uI.size(); unionLeaves += uL.size(); tiles += 1; hitTiles += anyHit;
		// packet traversal: shared stack, node entered if any lane passes (own tcut); near child = the one more lanes prefer
		{
			RayBox rb[32]; float tcut[32], bestT[32]; int bestPos[32]; bool alive[32];
			for (int l = 0; l < 32; l++) { rb[l] = make_raybox(rr[l].o, rr[l].d); tcut[l] = kMissT * kPruneSlack; bestT[l] = kMissT; bestPos[l] = -1; float te; alive[l] = slab_ref(rb[l], S.rootLo[0], S.rootLo[1], S.rootLo[2], S.rootHi[0], S.rootHi[1], S.rootHi[2], te); }
			struct PE { int ref; float tminAll; }; PE st[128]; int sp = 0; int cur = S.rootRef; bool have = false; for (int l = 0; l < 32; l++) have |= alive[l];
			while (have) {
				bool pop = true;
				if (cur >= 0) {
					pkInner += 1;
					const float4* nd = S.nodes + 4 * (size_t)cur; float4 a = nd[0], b = nd[1], c = nd[2]; float2 r = *reinterpret_cast<const float2*>(nd + 3);
					int r0 = f2i(r.x), r1 = f2i(r.y);
					int any0 = 0, any1 = 0, pref1 = 0; float min0 = 1e38f, min1 = 1e38f;
					for (int l = 0; l < 32; l++) if (alive[l]) { float e0, e1; bool h0, h1; node_boxes<kOctGeneric>(rb[l], make_raybox2(rb[l]), S.paired, a, b, c, tcut[l], h0, h1, e0, e1); if (h0) { any0++; min0 = std::min(min0, e0); } if (h1) { any1++; min1 = std::min(min1, e1); } if (h0 && h1 && e1 < e0) pref1++; else if (h1 && !h0) pref1++; }
					if (any0 && any1) { bool sw = min1 < min0; st[sp++] = { sw ? r0 : r1, sw ? min0 : min1 }; cur = sw ? r1 : r0; pop = false; }
					else if (any0) { cur = r0; pop = false; } else if (any1) { cur = r1; pop = false; }
				} else {
					pkLeaves += 1;
					int ref = ~cur; int pos = ref >> 1, cnt = (ref & 1) + 1;
					// lanes test their own ray against the leaf box?  (the parent's test already told which lanes pass; count them via MT directly)
					for (int l = 0; l < 32; l++) if (alive[l]) {
						bool touched = false;
						for (int k = 0; k < cnt; k++) { TriV tri = load_tri(S.tris, pos + k); float t; touched = true; if (moller_trumbore(tri, rr[l].o, rr[l].d, t)) if (t < bestT[l] || (t == bestT[l] && pos + k < bestPos[l])) { bestT[l] = t; bestPos[l] = pos + k; tcut[l] = t * kPruneSlack; } }
						pkLeafLanes += touched;
					}
				}
				if (pop) { bool got = false; while (sp > 0) { PE e = st[--sp]; bool need = false; for (int l = 0; l < 32; l++) if (alive[l] && e.tminAll <= tcut[l]) need = true; if (need) { cur = e.ref; got = true; break; } } if (!got) break; }
			}
		}
	}
	printf("per ray: inner %.1f leaves %.1f | per tile(32 rays): union inner %.1f union leaves %.1f | packet: inner %.1f leaves %.1f (hit tiles %.0f of %.0f)\n",
		sumInner / rays, sumLeaves / rays, unionInner / tiles, unionLeaves / tiles, pkInner / tiles, pkLeaves / tiles, hitTiles, tiles);
	return 0;
}
