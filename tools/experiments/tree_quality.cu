// experiment: quality of the production BVH topology on the DT scene, measured on the CPU -- surface-area cost of the tree and the
// work the ordered traversal does on it (inner nodes entered, triangles tested per primary and per shadow ray) over sampled 4x8 tiles
// of bench cameras.  Host builders' tuning knobs are read from the environment by the builders themselves (RTO_BVH_*).
#include "/root/repo/ray_tracing_octrees_b200/csrc/rto_kernels.cuh"
#include <cstdarg>
#include <cstdio>
#include <chrono>
#include <vector>
#include <algorithm>
#include <zlib.h>
#include "treeopt.h"
using namespace rto;
int rto_fail(int code, const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); return code; }

struct Work { double inner = 0, tris = 0, rays = 0, maxInner = 0; };

// the production closest-hit walk (bvh_closest_loop<true, generic>) with counters
static void closest(const BvhDev& S, V3 o, V3 d, float& bestT, int& bestPos, Work& w) {
	bestT = kMissT; bestPos = -1;
	w.rays += 1;
	RayBox rb = make_raybox(o, d);
	RayBox2 rb2 = make_raybox2(rb);
	float te;
	if (!slab_ref(rb, S.rootLo[0], S.rootLo[1], S.rootLo[2], S.rootHi[0], S.rootHi[1], S.rootHi[2], te)) return;
	StackEnt stack[kBvhStack]; int sp = 0; int cur = S.rootRef; float tcut = kMissT * kPruneSlack;
	double mine = 0;
	for (;;) {
		bool pop = true;
		if (cur >= 0) {
			w.inner += 1; mine += 1;
			const float4* n = S.nodes + 4 * (size_t)cur;
			float4 a = n[0], b = n[1], c = n[2]; float2 r = *reinterpret_cast<const float2*>(n + 3);
			float e0, e1; bool h0, h1;
			node_boxes<kOctGeneric>(rb, rb2, S.paired, a, b, c, tcut, h0, h1, e0, e1);
			int r0 = f2i(r.x), r1 = f2i(r.y);
			if (h0 && h1) { bool sw = e1 < e0; StackEnt e; e.ref = sw ? r0 : r1; e.t = sw ? e0 : e1; stack[sp++] = e; cur = sw ? r1 : r0; pop = false; }
			else if (h0) { cur = r0; pop = false; } else if (h1) { cur = r1; pop = false; }
		} else {
			int pos = (~cur) >> 1;
			w.tris += 1;
			TriV tri = load_tri(S.tris, pos); float t;
			if (moller_trumbore(tri, o, d, t) && (t < bestT || (t == bestT && pos < bestPos)) && ref_leaf_box_passes<kOctGeneric>(S, rb, pos, FLT_MAX)) { bestT = t; bestPos = pos; tcut = t * kPruneSlack; }
		}
		if (pop) { bool got = false; while (sp > 0) { StackEnt e = stack[--sp]; if (e.t <= tcut) { cur = e.ref; got = true; break; } } if (!got) break; }
	}
	w.maxInner = std::max(w.maxInner, mine);
}
static int g_anyOrder = 0;
static bool anyhit(const BvhDev& S, V3 o, V3 d, Work& w) {
	w.rays += 1;
	RayBox rb = make_raybox(o, d);
	RayBox2 rb2 = make_raybox2(rb);
	float te;
	if (!slab_ref(rb, S.rootLo[0], S.rootLo[1], S.rootLo[2], S.rootHi[0], S.rootHi[1], S.rootHi[2], te)) return false;
	int stack[kBvhStack]; int sp = 0; int cur = S.rootRef;
	for (;;) {
		if (cur >= 0) {
			w.inner += 1;
			const float4* n = S.nodes + 4 * (size_t)cur;
			float4 a = n[0], b = n[1], c = n[2]; float2 r = *reinterpret_cast<const float2*>(n + 3);
			float e0, e1; bool h0, h1;
			node_boxes<kOctGeneric>(rb, rb2, S.paired, a, b, c, FLT_MAX, h0, h1, e0, e1);
			int r0 = f2i(r.x), r1 = f2i(r.y);
			if (h0 && h1) {
				bool sw = e1 < e0;
				if (g_anyOrder == 1) sw = !sw;
				else if (g_anyOrder == 2) sw = false;
				else if (g_anyOrder == 3) {
					float x0, x1, t;
					slab_ref(rb, a.x, a.z, b.x, b.z, c.x, c.z, t); // exit distances recomputed below
					auto exitT = [&](float lox, float loy, float loz, float hix, float hiy, float hiz) {
						float tx = ((rb.nx ? lox : hix) - rb.o.x) * rb.inv.x, ty = ((rb.ny ? loy : hiy) - rb.o.y) * rb.inv.y, tz = ((rb.nz ? loz : hiz) - rb.o.z) * rb.inv.z;
						return fminf(fminf(tx, ty), tz); };
					x0 = exitT(a.x, a.z, b.x, b.z, c.x, c.z); x1 = exitT(a.y, a.w, b.y, b.w, c.y, c.w);
					sw = (x1 - e1) > (x0 - e0);
				}
				else if (g_anyOrder == 4) sw = (r1 < 0) && (r0 >= 0);   // leaves first
				stack[sp++] = sw ? r0 : r1; cur = sw ? r1 : r0; continue; }
			if (h0) { cur = r0; continue; }
			if (h1) { cur = r1; continue; }
		} else {
			int pos = (~cur) >> 1;
			w.tris += 1;
			TriV tri = load_tri(S.tris, pos); float t;
			if (moller_trumbore(tri, o, d, t) && ref_leaf_box_passes<kOctGeneric>(S, rb, pos, FLT_MAX)) return true;
		}
		if (sp == 0) return false;
		cur = stack[--sp];
	}
}

// surface-area cost: sum over inner nodes of (area of child 0's box + area of child 1's box) / area of the root box -- the expected number
// of child boxes a random line through the root box enters; depth statistics
static void tree_cost(const BvhLayout& L, double& cost, int& maxDepth, double& meanLeafDepth) {
	auto area = [](const float* lo, const float* hi) { double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2]; return dx * dy + dy * dz + dz * dx; };
	const double rootA = area(L.rootLo, L.rootHi);
	cost = 0; maxDepth = 0; meanLeafDepth = 0; double leaves = 0;
	std::vector<std::pair<int, int>> st; if (L.fastRoot >= 0) st.push_back({ L.fastRoot, 1 });
	while (!st.empty()) {
		auto [n, dep] = st.back(); st.pop_back();
		const float* d = &L.fastNodes[(size_t)n * 16];
		for (int c = 0; c < 2; c++) {
			float lo[3] = { d[0 + c], d[2 + c], d[4 + c] }, hi[3] = { d[6 + c], d[8 + c], d[10 + c] };
			cost += area(lo, hi) / rootA;
			int32_t ref; std::memcpy(&ref, &d[12 + c], 4);
			if (ref >= 0) st.push_back({ ref, dep + 1 }); else { maxDepth = std::max(maxDepth, dep); meanLeafDepth += dep; leaves += 1; }
		}
	}
	meanLeafDepth /= std::max(1.0, leaves);
}

int main(int argc, char** argv) {
	gzFile f = gzopen("/root/repo/tests/golden/dt_sceneCache.bin.gz", "rb");
	int dims[3]; float mv[4]; size_t n;
	gzread(f, dims, 12); gzread(f, mv, 16); gzread(f, &n, 8);
	std::vector<uint8_t> vox(n); gzread(f, vox.data(), (unsigned)n); gzclose(f);
	RtoGpuNode* nodes; size_t nn; rto_host_octree_build(vox.data(), dims[0], dims[1], dims[2], &nodes, &nn);
	RtoTriangle* tris; size_t nt; rto_host_mc_mesh(vox.data(), dims[0], dims[1], dims[2], mv, mv[3], nodes, nn, &tris, &nt);
	RtoHostBvh* hb; rto_host_bvh_build(tris, nt, &hb);
	auto t0 = std::chrono::steady_clock::now();
	BvhLayout L; rto_build_bvh_layout(*hb, L);
	const double layoutS = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	BvhDev S{}; S.numTris = (int)nt; S.rootRef = L.fastRoot; S.leafBox = 1; S.grow = L.fastGrow; S.paired = 1; S.exactPaired = 0; S.exactNodes = (const float4*)L.refNodes.data(); S.exactRoot = L.refRoot; S.exactLeafBox = 0; S.nodes = (const float4*)L.fastNodes.data(); S.tris = (const float4*)L.tris.data();
	for (int k = 0; k < 3; k++) { S.rootLo[k] = L.rootLo[k]; S.rootHi[k] = L.rootHi[k]; }
	double cost; int maxDepth; double meanDepth;
	if (getenv("TREE_OPT_PASSES")) {
		tree_cost(L, cost, maxDepth, meanDepth);
		printf("before: SA cost %.3f  max depth %d  mean leaf depth %.2f\n", cost, maxDepth, meanDepth);
		auto t1 = std::chrono::steady_clock::now();
		rto_treeopt::Tree T; T.load(L.fastNodes, L.fastRoot);
		printf("loaded: cost %.3f\n", T.cost());
		rto_treeopt::optimize(T, atoi(getenv("TREE_OPT_PASSES")), getenv("TREE_OPT_FRACTION") ? atof(getenv("TREE_OPT_FRACTION")) : 1.0);
		int dep = T.store(L.fastNodes, L.fastRoot);
		S.rootRef = L.fastRoot;
		printf("optimised in %.2f s: cost %.3f, depth %d\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count(), T.cost(), dep);
	}
	tree_cost(L, cost, maxDepth, meanDepth);
	printf("tris %zu  layout %.2f s  SA cost %.3f  max depth %d  mean leaf depth %.2f\n", nt, layoutS, cost, maxDepth, meanDepth);
	if (getenv("ANY_ORDER")) g_anyOrder = atoi(getenv("ANY_ORDER"));
	const int W = 1920, H = 1080, step = argc > 1 ? atoi(argv[1]) : 6;
	const float bias = 1e-3f * mv[3];
	Work P, Sh; double hits = 0, shadowed = 0; unsigned long long idSum = 0; double tSum = 0;
	double sumP = 0, slotP = 0, sumS = 0, slotS = 0, slotMerged = 0;       // lane-steps used / lane-steps a warp in lockstep pays (32 x the longest walk of the tile)
	for (int k = 0; k < 64; k += 8) {                      // 8 of the bench's 64 orbit cameras
		RtoCamera cam; float tgt[3] = { 0, 0, 0 };
		float th = 35.0f * 3.14159265f / 180.0f, ph = (40.0f + 360.0f / 64 * k) * 3.14159265f / 180.0f;
		rto_host_camera_orbit(th, ph, 0.6f * 4250.0f, tgt, 45.0f, float(W) / float(H), W, H, &cam, nullptr);
		for (int ty = 0; ty < H / 8; ty += step) for (int tx = 0; tx < W / 4; tx += step) {
			double pl[32], sl[32];
			for (int l = 0; l < 32; l++) {
				pl[l] = sl[l] = 0;
				const double p0 = P.inner + 2 * P.tris, s0 = Sh.inner + 2 * Sh.tris;
				struct Fin { double& a; double& b; const Work& P; const Work& Sh; double p0, s0; ~Fin() { a = P.inner + 2 * P.tris - p0; b = Sh.inner + 2 * Sh.tris - s0; } } fin{ pl[l], sl[l], P, Sh, p0, s0 };
				int px = tx * 4 + (l & 3), py = ty * 8 + (l >> 2);
				Ray ray = gen_ray(cam, px, py);
				float bt; int bp;
				closest(S, ray.o, ray.d, bt, bp, P);
				if (bp < 0) continue;
				hits += 1; idSum += (unsigned long long)bp; tSum += bt;
				TriV tri = load_tri(S.tris, bp);
				V3 nrm = normalize3(cross3(tri.e1, tri.e2));
				if (dot3(nrm, ray.d) > 0.0f) nrm = -nrm;
				V3 so = (ray.o + ray.d * bt) + nrm * bias;
				if (anyhit(S, so, normalize3(mk3(1.0f, 1.0f, 1.0f)), Sh)) shadowed += 1;
			}
			double mp = 0, ms = 0, mm = 0;
			for (int l = 0; l < 32; l++) { sumP += pl[l]; sumS += sl[l]; mp = std::max(mp, pl[l]); ms = std::max(ms, sl[l]); mm = std::max(mm, pl[l] + sl[l]); }
			slotP += 32 * mp; slotS += 32 * ms; slotMerged += 32 * mm;
		}
	}
	printf("primary: %.0f rays, hit %.3f, inner %.2f, tris %.2f per ray (deepest walk %.0f inner)\n", P.rays, hits / P.rays, P.inner / P.rays, P.tris / P.rays, P.maxInner);
	printf("shadow : %.0f rays, shadowed %.3f, inner %.2f, tris %.2f per ray\n", Sh.rays, shadowed / std::max(1.0, Sh.rays), Sh.inner / std::max(1.0, Sh.rays), Sh.tris / std::max(1.0, Sh.rays));
	printf("lockstep: primary %.3f of the lane-steps used, shadow %.3f, both phases %.3f; if a lane went on to its shadow ray at once: %.3f\n",
		sumP / slotP, sumS / slotS, (sumP + sumS) / (slotP + slotS), (sumP + sumS) / slotMerged);
	printf("total  : inner %.2f, tris %.2f per primary ray  (checksum id %llu t %.6e)\n", (P.inner + Sh.inner) / P.rays, (P.tris + Sh.tris) / P.rays, idSum, tSum);
	return 0;
}
