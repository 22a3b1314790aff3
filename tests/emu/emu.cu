// tests/emu/emu.cu -- TEST INFRASTRUCTURE: runs the product's traversal functions (rto_kernels.cuh, RTO_DEV = host + device)
// on the CPU over the product's own device layouts (host_layouts.cpp), pixel by pixel.  Purpose: find logic errors
// (wrong links, endless walks, mismatches against the oracle) in this GPU-less container before spending GPU minutes.
// This is NOT a fallback: it is not part of librto.so and nothing in the package loads it.
#include <functional>
#include "../../ray_tracing_octrees_b200/csrc/rto_kernels.cuh"
#include "../../ray_tracing_octrees_b200/csrc/rto_sahchunk.h"
#include <cstdarg>
#include <cstdio>
#include <vector>

using namespace rto;

int rto_fail(int code, const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); return code; }

struct EmuOct { OctLayout L; OctDev D; };
struct EmuBvh { BvhLayout L; BvhDev ref, fast; std::vector<RtoTriangle> tris; };

extern "C" {

void* emu_octree_create(const RtoGpuNode* nodes, size_t n, const float* gmin, float voxel) {
	EmuOct* e = new EmuOct();
	if (rto_build_octree_layout(nodes, n, e->L) != RTO_OK) { delete e; return nullptr; }
	OctDev& D = e->D;
	D.numNodes = (int)n; D.rootSize = nodes[0].size; D.compact = e->L.compact;
	D.gmin[0] = gmin[0]; D.gmin[1] = gmin[1]; D.gmin[2] = gmin[2]; D.voxel = voxel;
	D.desc = e->L.compact ? e->L.desc.data() + 7 : nullptr; D.up = e->L.up.data();
	D.inner = (const int4*)e->L.inner.data(); D.nodes16 = (const int4*)e->L.padded.data();
	return e;
}
void emu_octree_free(void* h) { delete (EmuOct*)h; }

// count = 1: per-node paths with visit counters (what rto_render_stats runs); count = 0: production (fast) paths
void emu_render_octree(void* h, const RtoCamera* cam, int mode, int count, int y0, int y1, float* rgba, int32_t* ids, float* t, uint64_t* visits, uint32_t* perPixelVisits) {
	EmuOct* e = (EmuOct*)h;
	uint64_t v = 0;
	for (int py = y0; py < y1; py++)
		for (int px = 0; px < cam->width; px++) {
			size_t pix = (size_t)(py - y0) * cam->width + px;
			Ray ray = gen_ray(*cam, px, py);
			OctHit hit = count ? oct_trace<true>(e->D, mode, ray.o, ray.d, 0.0f, 1e30f) : oct_trace<false>(e->D, mode, ray.o, ray.d, 0.0f, 1e30f);
			v += hit.visits;
			if (perPixelVisits) perPixelVisits[pix] = hit.visits;
			V3 color = mk3(0.0f, 0.0f, 0.0f);
			if (hit.id >= 0) color = shade_lambert(hit.normal);
			if (rgba) { rgba[4 * pix] = color.x; rgba[4 * pix + 1] = color.y; rgba[4 * pix + 2] = color.z; rgba[4 * pix + 3] = 1.0f; }
			if (ids) ids[pix] = hit.id;
			if (t) t[pix] = hit.t;
		}
	if (visits) *visits = v;
}
void emu_trace_octree(void* h, int mode, int count, const float* o3, const float* d3, size_t n, float tMin, float tMax, float* t, int32_t* ids) {
	EmuOct* e = (EmuOct*)h;
	for (size_t i = 0; i < n; i++) {
		V3 o = mk3(o3[3 * i], o3[3 * i + 1], o3[3 * i + 2]), d = mk3(d3[3 * i], d3[3 * i + 1], d3[3 * i + 2]);
		OctHit hit = count ? oct_trace<true>(e->D, mode, o, d, tMin, tMax) : oct_trace<false>(e->D, mode, o, d, tMin, tMax);
		t[i] = hit.t; ids[i] = hit.id;
	}
}

void* emu_bvh_create(const RtoTriangle* tris, size_t n) {
	EmuBvh* e = new EmuBvh();
	e->tris.assign(tris, tris + n);
	RtoHostBvh* hb = nullptr;
	if (rto_host_bvh_build(e->tris.data(), n, &hb) != RTO_OK) { delete e; return nullptr; }
	rto_build_bvh_layout(*hb, e->L);
	rto_host_bvh_free(hb);
	BvhDev D{};
	D.numTris = (int)n; D.rootRef = e->L.refRoot;
	for (int k = 0; k < 3; k++) { D.rootLo[k] = e->L.rootLo[k]; D.rootHi[k] = e->L.rootHi[k]; }
	D.nodes = (const float4*)e->L.refNodes.data(); D.tris = (const float4*)e->L.tris.data();
	D.leafBox = 0; D.grow = 0.0f; D.paired = 0; D.exactPaired = 0;
	D.exactNodes = D.nodes; D.exactRoot = D.rootRef; D.exactLeafBox = 0;
	e->ref = D; e->fast = D;
	if (!e->L.fastNodes.empty()) { e->fast.nodes = (const float4*)e->L.fastNodes.data(); e->fast.rootRef = e->L.fastRoot; e->fast.leafBox = 1; e->fast.grow = e->L.fastGrow; e->fast.paired = 1; }
	if (!e->L.wideNodes.empty() && e->L.wideRoot >= 0) {      // (built when RTO_BVH_WIDE=1 is set at creation)
		e->fast.wide = (const uint4*)e->L.wideNodes.data(); e->fast.wideRoot = e->L.wideRoot; e->fast.wideStep = e->L.wideStep;
		for (int k = 0; k < 3; k++) e->fast.wideLo[k] = e->L.wideLo[k];
	}
	return e;
}
void emu_bvh_free(void* h) { delete (EmuBvh*)h; }
int emu_bvh_wide_nodes(void* h) { return (int)(((EmuBvh*)h)->L.wideNodes.size() / 16); }

void emu_render_bvh(void* h, const RtoCamera* cam, unsigned flags, float bias, int y0, int y1, float* rgba, int32_t* ids, float* t) {
	EmuBvh* e = (EmuBvh*)h;
	const bool prune = !(flags & RTO_FLAG_NO_PRUNE);
	const BvhDev& S = prune ? e->fast : e->ref;
	for (int py = y0; py < y1; py++)
		for (int px = 0; px < cam->width; px++) {
			size_t pix = (size_t)(py - y0) * cam->width + px;
			Ray ray = gen_ray(*cam, px, py);
			float bestT; int bestPos;
			const bool wide = prune && (flags & 0x100u) && S.wide;      // the 4-wide quantised form of the same tree
			if (wide) bvh_closest<true, true>(S, ray.o, ray.d, bestT, bestPos);
			else if (prune) bvh_closest<true>(S, ray.o, ray.d, bestT, bestPos); else bvh_closest<false>(S, ray.o, ray.d, bestT, bestPos);
			V3 color = mk3(0.0f, 0.0f, 0.0f); int id = -1;
			if (bestPos >= 0) {
				TriV tri = load_tri(S.tris, bestPos);
				id = tri.id;
				V3 e1 = tri.e1, e2 = tri.e2;
				V3 n = normalize3(cross3(e1, e2));
				if (dot3(n, ray.d) > 0.0f) n = -n;
				V3 hp = ray.o + ray.d * bestT;
				bool shadowed = false;
				if (flags & RTO_FLAG_SHADOWS) shadowed = wide ? bvh_any<true>(S, hp + n * bias, normalize3(mk3(1.0f, 1.0f, 1.0f))) : bvh_any(S, hp + n * bias, normalize3(mk3(1.0f, 1.0f, 1.0f)));
				color = shadowed ? mk3(0.1f, 0.1f, 0.1f) : shade_lambert(n);
			}
			if (rgba) { rgba[4 * pix] = color.x; rgba[4 * pix + 1] = color.y; rgba[4 * pix + 2] = color.z; rgba[4 * pix + 3] = 1.0f; }
			if (ids) ids[pix] = id;
			if (t) t[pix] = bestT;
		}
}

// rto_sahchunk.h on the CPU: rebuild one subtree over m leaves placed at leaf offset `first`, with its root on the first or on the
// last index of its range, and check the result as a tree: every leaf exactly once, every owned slot exactly once, parent links
// consistent, nothing written outside the owned slots.  Returns the sum of the half areas of all internal boxes (the quantity the
// splits minimise), or a negative error code.
double emu_sah_chunk(const float* leafBox6, int m, int first, int rootAtEnd) {
	const int last = first + m - 1, root = rootAtEnd ? last : first;
	const int total = last + 8;
	std::vector<float> boxes((size_t)total * 6, 0.0f);
	std::memcpy(&boxes[(size_t)first * 6], leafBox6, sizeof(float) * 6 * m);
	std::vector<float4> nodes((size_t)total * 4, make_float4(-7.0f, -7.0f, -7.0f, -7.0f));
	std::vector<int> pInner(total, -7), pLeaf(total, -7);
	sah_rebuild_chunk(boxes.data(), first, last, root, nodes.data(), pInner.data(), pLeaf.data());
	const int slotLo = rootAtEnd ? first + 1 : first, slotHi = rootAtEnd ? last : last - 1;
	for (int i = 0; i < total; i++) {
		const bool owned = i >= slotLo && i <= slotHi;
		if (!owned && (nodes[4 * (size_t)i + 3].x != -7.0f || (pInner[i] != -7))) return -1;             // wrote outside its slots
		if (owned && nodes[4 * (size_t)i + 3].x == -7.0f && nodes[4 * (size_t)i + 3].y == -7.0f) return -2;  // a slot was left unused
		if ((i < first || i > last) && pLeaf[i] != -7) return -3;
	}
	if (pInner[root] != -7) return -4;                                                                      // the root's own link belongs to the radix tree above
	std::vector<int> seenLeaf(total, 0), seenNode(total, 0);
	double cost = 0;
	struct Fr { int node; };
	std::vector<int> st{ root };
	std::vector<SahBox> box(total);
	// post-order via explicit recursion
	std::function<SahBox(int, int, int)> walk = [&](int ref, int parent, int side) -> SahBox {
		SahBox b; sah_box_reset(b);
		if (ref < 0) {
			const int leaf = (~ref) >> 1;
			if (leaf < first || leaf > last || ((~ref) & 1)) { cost = -5; return b; }
			if (seenLeaf[leaf]++) { cost = -6; return b; }
			if (pLeaf[leaf] != 2 * parent + side) { cost = -7; return b; }
			sah_box_add(b, &boxes[(size_t)leaf * 6]);
			return b;
		}
		if (ref < slotLo || ref > slotHi) { cost = -8; return b; }
		if (seenNode[ref]++) { cost = -9; return b; }
		if (ref != root && pInner[ref] != 2 * parent + side) { cost = -10; return b; }
		int r0, r1; std::memcpy(&r0, &nodes[4 * (size_t)ref + 3].x, 4); std::memcpy(&r1, &nodes[4 * (size_t)ref + 3].y, 4);
		SahBox l = walk(r0, ref, 0); if (cost < 0) return b;
		SahBox r = walk(r1, ref, 1); if (cost < 0) return b;
		b = l; sah_box_join(b, r);
		cost += sah_half_area(b);
		return b;
	};
	walk(root, -1, 0);
	if (cost < 0) return cost;
	for (int i = first; i <= last; i++) if (seenLeaf[i] != 1) return -11;
	for (int i = slotLo; i <= slotHi; i++) if (seenNode[i] != 1) return -12;
	return cost;
}

} // extern "C"
