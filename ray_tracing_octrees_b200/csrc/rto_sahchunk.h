// rto_sahchunk.h -- surface-area rebuild of the bottom of the device-built BVH (rto_build.cu), one thread block per subtree
// (sah_rebuild_chunk_block; sah_rebuild_chunk is the one-thread form of the same algorithm).
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
// The one-thread form is written host+device: tests/emu runs it on the CPU (tests/test_emu_traversal.py) before the GPU ever sees it, and the
// block form is compared with it on the GPU (tests/test_gpu_builders.py).
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
//
// sah_rebuild_range splits the leaves idx[a .. b) (indices relative to `first`) below node slot `slot`; dir = +1 / -1: slots are handed out
// upwards from `first` or downwards from `last`.  One thread; idx may live in local, shared or global memory.
template <typename IdxPtr>
RTO_SAH_HD void sah_rebuild_range(const float* __restrict__ leafBox, int first, IdxPtr idx, int a0, int b0, int slot0, int dir, float4* __restrict__ nodes,
	int* __restrict__ parentOfInner, int* __restrict__ parentOfLeaf) {
	struct Job { int a, b, slot; };
	Job stack[24];                                  // the smaller half is split first: at most log2(kSahChunk) + 1 halves wait
	int sp = 0;
	stack[sp++] = Job{ a0, b0, slot0 };
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

// one thread, the whole subtree (the form tests/emu checks on the CPU; the device runs sah_rebuild_chunk_block below)
RTO_SAH_HD void sah_rebuild_chunk(const float* __restrict__ leafBox, int first, int last, int root, float4* __restrict__ nodes,
	int* __restrict__ parentOfInner, int* __restrict__ parentOfLeaf) {
	const int m = last - first + 1;
	uint16_t idx[kSahChunk];
	for (int i = 0; i < m; i++) idx[i] = (uint16_t)i;
	sah_rebuild_range(leafBox, first, idx, 0, m, root, (root == first) ? 1 : -1, nodes, parentOfInner, parentOfLeaf);
}

#if defined(__CUDACC__)
// ---- the same rebuild by one thread block per subtree --------------------------------------------------------------------------
// One thread per subtree made the rebuild a constant quarter of a second whatever the scene (a 4096-leaf subtree is twelve levels of five
// passes over its leaves, all on one lane).  Here a block of kSahBlock threads owns the subtree: ranges of more than kSahSerial leaves are
// split by the whole block -- centroid bounds and the 3 x 16 bins through shared-memory atomics on order-preserving integer images of the
// floats (min / max of floats are exact whatever the order, so the bins, and with them every split, are the ones the serial code finds),
// a stable partition by warp ballots -- and the ranges of at most kSahSerial leaves that fall out are queued in shared memory and
// finished afterwards by one thread each with sah_rebuild_range itself.  The tree is the serial one node for node, except below ranges
// whose centroids all coincide, where "halve the list" depends on the order inside the list (any halving is a valid tree).
constexpr int kSahBlock = 128;
constexpr int kSahSerial = 64;
static_assert(kSahChunk <= 4096 && kSahSerial <= 64, "the queue entries of sah_rebuild_chunk_block pack 12-bit offsets and a 6-bit length");

__device__ __forceinline__ unsigned sah_enc(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float sah_dec(unsigned e) { return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e); }

__device__ inline void sah_rebuild_chunk_block(const float* __restrict__ leafBox, int first, int last, int root, float4* __restrict__ nodes,
	int* __restrict__ parentOfInner, int* __restrict__ parentOfLeaf) {
	__shared__ uint16_t sIdx[kSahChunk], sTmp[kSahChunk];
	__shared__ unsigned sSmall[kSahChunk / 2];                  // a (12 bits) | n - 1 (6 bits) << 12 | slot - first (12 bits) << 18
	__shared__ int sNumSmall, sSp, sCur[3];
	__shared__ int sStack[16][3];                               // the smaller half is split first: at most log2(kSahChunk) + 1 halves wait
	__shared__ unsigned sCmin[3], sCmax[3], sLo[3][kSahBins][3], sHi[3][kSahBins][3];
	__shared__ int sCnt[3][kSahBins], sAxisSplit[3], sWarpL[kSahBlock / 32];
	__shared__ float sAxisCost[3];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int m = last - first + 1, dir = (root == first) ? 1 : -1;
	for (int i = tid; i < m; i += kSahBlock) sIdx[i] = (uint16_t)i;
	if (tid == 0) {
		sNumSmall = 0; sSp = 0;
		if (m > kSahSerial) { sStack[0][0] = 0; sStack[0][1] = m; sStack[0][2] = root; sSp = 1; }
		else { sSmall[0] = 0u | ((unsigned)(m - 1) << 12) | ((unsigned)(root - first) << 18); sNumSmall = 1; }
	}
	for (;;) {
		__syncthreads();
		if (tid == 0) {
			if (sSp == 0) sCur[0] = -1;
			else { sSp--; sCur[0] = sStack[sSp][0]; sCur[1] = sStack[sSp][1]; sCur[2] = sStack[sSp][2]; }
		}
		if (tid < 3) { sCmin[tid] = sah_enc(FLT_MAX); sCmax[tid] = sah_enc(-FLT_MAX); sAxisCost[tid] = FLT_MAX; sAxisSplit[tid] = 0; }
		for (int i = tid; i < 3 * kSahBins; i += kSahBlock) {
			(&sCnt[0][0])[i] = 0;
			for (int c = 0; c < 3; c++) { (&sLo[0][0][0])[3 * i + c] = sah_enc(FLT_MAX); (&sHi[0][0][0])[3 * i + c] = sah_enc(-FLT_MAX); }
		}
		__syncthreads();
		const int a = sCur[0], b = sCur[1], slot = sCur[2];
		if (a < 0) break;
		const int n = b - a;
		// bounds of the centroids
		{
			float cmin[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, cmax[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
			for (int i = a + tid; i < b; i += kSahBlock) {
				const float* bx = leafBox + 6 * (size_t)(first + sIdx[i]);
				for (int c = 0; c < 3; c++) { const float v = bx[c] + bx[3 + c]; cmin[c] = fminf(cmin[c], v); cmax[c] = fmaxf(cmax[c], v); }
			}
			for (int c = 0; c < 3; c++) {
				for (int o = 16; o > 0; o >>= 1) { cmin[c] = fminf(cmin[c], __shfl_xor_sync(0xffffffffu, cmin[c], o)); cmax[c] = fmaxf(cmax[c], __shfl_xor_sync(0xffffffffu, cmax[c], o)); }
				if (lane == 0) { atomicMin(&sCmin[c], sah_enc(cmin[c])); atomicMax(&sCmax[c], sah_enc(cmax[c])); }
			}
		}
		__syncthreads();
		float base[3], scale[3]; bool live[3];
		for (int c = 0; c < 3; c++) {
			const float lo = sah_dec(sCmin[c]), hi = sah_dec(sCmax[c]);
			const float ext = hi - lo;
			live[c] = ext > 0.0f; base[c] = lo; scale[c] = live[c] ? (float)kSahBins / ext : 0.0f;
		}
		// the bins of the three axes
		for (int i = a + tid; i < b; i += kSahBlock) {
			const float* bx = leafBox + 6 * (size_t)(first + sIdx[i]);
			float v[6];
			for (int c = 0; c < 6; c++) v[c] = bx[c];
			for (int ax = 0; ax < 3; ax++) {
				if (!live[ax]) continue;
				int k = (int)((v[ax] + v[3 + ax] - base[ax]) * scale[ax]);
				k = k < 0 ? 0 : (k >= kSahBins ? kSahBins - 1 : k);
				atomicAdd(&sCnt[ax][k], 1);
				for (int c = 0; c < 3; c++) { atomicMin(&sLo[ax][k][c], sah_enc(v[c])); atomicMax(&sHi[ax][k][c], sah_enc(v[3 + c])); }
			}
		}
		__syncthreads();
		// one thread per axis sweeps its 16 bins (the serial code's sweep on the same numbers)
		if (tid < 3 && live[tid]) {
			const int ax = tid;
			SahBox bb[kSahBins]; int cnt[kSahBins];
			for (int k = 0; k < kSahBins; k++) {
				cnt[k] = sCnt[ax][k];
				for (int c = 0; c < 3; c++) { bb[k].lo[c] = sah_dec(sLo[ax][k][c]); bb[k].hi[c] = sah_dec(sHi[ax][k][c]); }
			}
			float rightArea[kSahBins]; int rightCnt[kSahBins];
			SahBox acc; sah_box_reset(acc); int c = 0;
			for (int k = kSahBins - 1; k > 0; k--) { sah_box_join(acc, bb[k]); c += cnt[k]; rightCnt[k] = c; rightArea[k] = c ? sah_half_area(acc) : 0.0f; }
			sah_box_reset(acc); c = 0;
			float bestCost = FLT_MAX; int bestSplit = 0;
			for (int k = 0; k < kSahBins - 1; k++) {
				sah_box_join(acc, bb[k]); c += cnt[k];
				if (c == 0 || rightCnt[k + 1] == 0) continue;
				const float cost = sah_half_area(acc) * (float)c + rightArea[k + 1] * (float)rightCnt[k + 1];
				if (cost < bestCost) { bestCost = cost; bestSplit = k + 1; }
			}
			sAxisCost[ax] = bestCost; sAxisSplit[ax] = bestSplit;
		}
		__syncthreads();
		int bestAxis = -1, bestSplit = 0; float bestCost = FLT_MAX;
		for (int ax = 0; ax < 3; ax++) if (sAxisCost[ax] < bestCost) { bestCost = sAxisCost[ax]; bestAxis = ax; bestSplit = sAxisSplit[ax]; }
		int nl = 0;
		if (bestAxis >= 0) for (int k = 0; k < bestSplit; k++) nl += sCnt[bestAxis][k];
		int mid;
		if (bestAxis < 0 || nl == 0 || nl == n) mid = a + n / 2;      // all centroids coincide: halve the list as it stands
		else {
			// stable partition: the leaves of bins below the split keep their order in front, the others behind them
			mid = a + nl;
			int doneL = 0, doneR = 0;
			for (int tile = a; tile < b; tile += kSahBlock) {
				const int i = tile + tid;
				bool valid = i < b, left = false; uint16_t id = 0;
				if (valid) {
					id = sIdx[i];
					const float* bx = leafBox + 6 * (size_t)(first + id);
					int k = (int)((bx[bestAxis] + bx[3 + bestAxis] - base[bestAxis]) * scale[bestAxis]);
					k = k < 0 ? 0 : (k >= kSahBins ? kSahBins - 1 : k);
					left = k < bestSplit;
				}
				const unsigned bl = __ballot_sync(0xffffffffu, valid && left);
				if (lane == 0) sWarpL[warp] = __popc(bl);
				__syncthreads();
				int before = 0, total = 0;
				for (int w = 0; w < kSahBlock / 32; w++) { const int c = sWarpL[w]; if (w < warp) before += c; total += c; }
				const int rankL = before + __popc(bl & ((1u << lane) - 1u));
				const int posInTile = warp * 32 + lane;
				if (valid) {
					if (left) sTmp[a + doneL + rankL] = id;
					else sTmp[mid + doneR + (posInTile - rankL)] = id;
				}
				const int tileN = min(kSahBlock, b - tile);
				doneL += total; doneR += tileN - total;
				__syncthreads();
			}
			for (int i = a + tid; i < b; i += kSahBlock) sIdx[i] = sTmp[i];
			__syncthreads();
		}
		if (tid == 0) {
			const int nr = b - mid; nl = mid - a;
			const int leftSlot = slot + dir, rightSlot = slot + dir * nl;
			int r0, r1;
			if (nl == 1) { const int leaf = first + sIdx[a]; r0 = ~(leaf << 1); parentOfLeaf[leaf] = 2 * slot; }
			else { r0 = leftSlot; parentOfInner[leftSlot] = 2 * slot; }
			if (nr == 1) { const int leaf = first + sIdx[mid]; r1 = ~(leaf << 1); parentOfLeaf[leaf] = 2 * slot + 1; }
			else { r1 = rightSlot; parentOfInner[rightSlot] = 2 * slot + 1; }
			nodes[4 * (size_t)slot + 3] = make_float4(__int_as_float(r0), __int_as_float(r1), 0.0f, 0.0f);
			// ranges the block splits go on the stack (larger half first, so that the smaller one is split next), small ones into the queue
			const int ja[2] = { a, mid }, jn[2] = { nl, nr }, js[2] = { leftSlot, rightSlot };
			const int o0 = nl >= nr ? 0 : 1;
			for (int q = 0; q < 2; q++) {
				const int h = q == 0 ? o0 : 1 - o0;
				if (jn[h] > kSahSerial) { sStack[sSp][0] = ja[h]; sStack[sSp][1] = ja[h] + jn[h]; sStack[sSp][2] = js[h]; sSp++; }
				else if (jn[h] > 1) sSmall[sNumSmall++] = (unsigned)ja[h] | ((unsigned)(jn[h] - 1) << 12) | ((unsigned)(js[h] - first) << 18);
			}
		}
	}
	// (the loop was left behind a barrier: queue and index list are complete)
	for (int j = tid; j < sNumSmall; j += kSahBlock) {
		const unsigned e = sSmall[j];
		const int a = (int)(e & 4095u), n = (int)((e >> 12) & 63u) + 1, slot = first + (int)(e >> 18);
		sah_rebuild_range(leafBox, first, sIdx, a, a + n, slot, dir, nodes, parentOfInner, parentOfLeaf);
	}
}
#endif

} // namespace rto
