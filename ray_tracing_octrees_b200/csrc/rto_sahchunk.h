// rto_sahchunk.h -- surface-area rebuild of the bottom of the device-built BVH (rto_build.cu), one thread per subtree.
//
// The device route sorts the triangles along a Morton curve and takes Karras' binary radix tree over them: built in a few passes at
// memory speed, but every split is a Morton split, and a tree of Morton splits traces 15-25 % slower than the host route's
// surface-area tree (profiles/README.md).  Measurements in round 1 showed that neither the node order nor the top of the tree makes
// that difference -- what is left is the choice of every split below.  So the radix tree keeps its top, and every maximal subtree
// with at most kSahChunk leaves is rebuilt from scratch by binned surface-area splits over its leaves, into the node slots that
// subtree owns: the internal nodes of a radix subtree over the sorted leaves [first, last] are the indices of that interval except the
// end the subtree's root does not sit on (every internal node of Karras' tree sits on an end of its own range), so the subtree's root
// stays where its parent points and the other m - 2 slots are handed out in pre-order from that end.  Only topology is written
// (child references and parent links); the bottom-up fit that follows computes every box.  Any binary tree over the same leaves is a
// valid result -- the choice of splits decides speed, never hits.
//
// Written host+device: tests/emu runs it on the CPU (tests/test_emu_traversal.py) before the GPU ever sees it.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

namespace rto {

// Leaves per rebuilt subtree.  Measured (B200, 16 x 1080p primary + shadow per launch, ms; host route's binned-SAH tree for comparison):
//                    radix tree    64     256    1024    4096    host SAH
//   DT mesh             3.733    3.315   3.114   2.953   2.893    2.828
//   512^3 city mesh     5.608    5.219   5.045   4.911   4.832    4.822
//   1024^3 MC mesh (4K) 3.567    3.302   3.179   3.099   3.031      -
// 4096 leaves = 12 levels of surface-area splits under the radix tree's top; the whole build of the 98.7 M-triangle scene stays under a second.
#ifndef RTO_SAH_CHUNK
#define RTO_SAH_CHUNK 4096
#endif
constexpr int kSahChunk = RTO_SAH_CHUNK;
constexpr int kSahBins = 16;

#if defined(__CUDACC__)
#define RTO_SAH_HD __host__ __device__ inline
#else
#define RTO_SAH_HD inline
#endif

struct SahBox { float lo[3], hi[3]; };
RTO_SAH_HD void sah_box_reset(SahBox& b) { for (int a = 0; a < 3; a++) { b.lo[a] = FLT_MAX; b.hi[a] = -FLT_MAX; } }
RTO_SAH_HD void sah_box_add(SahBox& b, const float* box6) { for (int a = 0; a < 3; a++) { b.lo[a] = fminf(b.lo[a], box6[a]); b.hi[a] = fmaxf(b.hi[a], box6[3 + a]); } }
RTO_SAH_HD void sah_box_join(SahBox& b, const SahBox& o) { for (int a = 0; a < 3; a++) { b.lo[a] = fminf(b.lo[a], o.lo[a]); b.hi[a] = fmaxf(b.hi[a], o.hi[a]); } }
RTO_SAH_HD float sah_half_area(const SahBox& b) {
	const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
	return dx * dy + dy * dz + dz * dx;
}

// leafBox: 6 floats per leaf (lo xyz, hi xyz), leaves in sorted order; the subtree spans leaves [first, last] (2 <= m <= kSahChunk) and its
// root is internal node `root` (== first or == last).  nodes: 4 x float4 per internal node, the references live in [3].xy;
// parentOfInner / parentOfLeaf: 2 * parent + side.  Leaves are single triangles: reference ~(position << 1).
RTO_SAH_HD void sah_rebuild_chunk(const float* __restrict__ leafBox, int first, int last, int root, float4* __restrict__ nodes,
	int* __restrict__ parentOfInner, int* __restrict__ parentOfLeaf) {
	const int m = last - first + 1;
	uint16_t idx[kSahChunk];
	for (int i = 0; i < m; i++) idx[i] = (uint16_t)i;
	const int dir = (root == first) ? 1 : -1;      // slots are handed out upwards from `first` or downwards from `last`
	struct Job { int a, b, slot; };
	Job stack[24];                                  // the smaller half is split first: at most log2(kSahChunk) + 1 halves wait
	int sp = 0;
	stack[sp++] = Job{ 0, m, root };
	while (sp > 0) {
		const Job J = stack[--sp];
		const int n = J.b - J.a;
		// bounds of the centroids (twice the centroid: lo + hi, the factor changes nothing)
		float cmin[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, cmax[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
		for (int i = J.a; i < J.b; i++) {
			const float* bx = leafBox + 6 * (size_t)(first + idx[i]);
			for (int a = 0; a < 3; a++) { const float c = bx[a] + bx[3 + a]; cmin[a] = fminf(cmin[a], c); cmax[a] = fmaxf(cmax[a], c); }
		}
		int bestAxis = -1, bestSplit = 0; float bestCost = FLT_MAX;
		for (int ax = 0; ax < 3; ax++) {
			const float ext = cmax[ax] - cmin[ax];
			if (!(ext > 0.0f)) continue;
			const float scale = (float)kSahBins / ext;
			int cnt[kSahBins]; SahBox bb[kSahBins];
			for (int k = 0; k < kSahBins; k++) { cnt[k] = 0; sah_box_reset(bb[k]); }
			for (int i = J.a; i < J.b; i++) {
				const float* bx = leafBox + 6 * (size_t)(first + idx[i]);
				int k = (int)((bx[ax] + bx[3 + ax] - cmin[ax]) * scale);
				k = k < 0 ? 0 : (k >= kSahBins ? kSahBins - 1 : k);
				cnt[k]++; sah_box_add(bb[k], bx);
			}
			float rightArea[kSahBins]; int rightCnt[kSahBins];
			SahBox acc; sah_box_reset(acc); int c = 0;
			for (int k = kSahBins - 1; k > 0; k--) { sah_box_join(acc, bb[k]); c += cnt[k]; rightCnt[k] = c; rightArea[k] = c ? sah_half_area(acc) : 0.0f; }
			sah_box_reset(acc); c = 0;
			for (int k = 0; k < kSahBins - 1; k++) {
				sah_box_join(acc, bb[k]); c += cnt[k];
				if (c == 0 || rightCnt[k + 1] == 0) continue;
				const float cost = sah_half_area(acc) * (float)c + rightArea[k + 1] * (float)rightCnt[k + 1];
				if (cost < bestCost) { bestCost = cost; bestAxis = ax; bestSplit = k + 1; }
			}
		}
		int mid;
		if (bestAxis < 0) mid = J.a + n / 2;           // all centroids coincide: halve the list
		else {
			const float scale = (float)kSahBins / (cmax[bestAxis] - cmin[bestAxis]), base = cmin[bestAxis];
			int i = J.a, j = J.b - 1;
			while (i <= j) {
				const float* bx = leafBox + 6 * (size_t)(first + idx[i]);
				int k = (int)((bx[bestAxis] + bx[3 + bestAxis] - base) * scale);
				k = k < 0 ? 0 : (k >= kSahBins ? kSahBins - 1 : k);
				if (k < bestSplit) i++;
				else { const uint16_t t = idx[i]; idx[i] = idx[j]; idx[j] = t; j--; }
			}
			mid = i;
			if (mid == J.a || mid == J.b) mid = J.a + n / 2;
		}
		const int nl = mid - J.a, nr = J.b - mid;
		// slots of the two subtrees: a subtree over k leaves owns k - 1 slots, its root at the end nearest to its parent
		const int leftSlot = J.slot + dir, rightSlot = J.slot + dir * nl;      // (dir = +1: [slot + 1, slot + nl - 1] and [slot + nl, ...]; dir = -1 mirrored)
		int r0, r1;
		if (nl == 1) { const int leaf = first + idx[J.a]; r0 = ~(leaf << 1); parentOfLeaf[leaf] = 2 * J.slot; }
		else { r0 = leftSlot; parentOfInner[leftSlot] = 2 * J.slot; }
		if (nr == 1) { const int leaf = first + idx[mid]; r1 = ~(leaf << 1); parentOfLeaf[leaf] = 2 * J.slot + 1; }
		else { r1 = rightSlot; parentOfInner[rightSlot] = 2 * J.slot + 1; }
		float4 refs;
#if defined(__CUDA_ARCH__)
		refs = make_float4(__int_as_float(r0), __int_as_float(r1), 0.0f, 0.0f);
#else
		refs.z = refs.w = 0.0f; memcpy(&refs.x, &r0, 4); memcpy(&refs.y, &r1, 4);
#endif
		nodes[4 * (size_t)J.slot + 3] = refs;
		// the larger half waits, the smaller one is split next
		const Job L{ J.a, mid, leftSlot }, R{ mid, J.b, rightSlot };
		if (nl >= nr) { if (nl > 1) stack[sp++] = L; if (nr > 1) stack[sp++] = R; }
		else { if (nr > 1) stack[sp++] = R; if (nl > 1) stack[sp++] = L; }
	}
}

} // namespace rto
