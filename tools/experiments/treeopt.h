// treeopt.h -- EXPERIMENT (tree_quality.cu): insertion-based optimisation of the production BVH topology.  Measured: not worth it, see profiles/README.md.
//
// The binned surface-area builder decides every split greedily, top-down; what it cannot see is that a subtree would sit better somewhere
// else entirely.  Here subtrees are taken out and put back where they add the least surface area (Bittner, Hapala, Havran, "Fast
// insertion-based optimization of bounding volume hierarchies", 2013): remove node X and its parent P (X's sibling takes P's place), re-fit
// the boxes above, search the tree best-first for the node Y next to which X costs the least -- area(Y u X) plus what the boxes above Y
// grow by -- and re-use P as the new parent of Y and X.  Any tree over the same leaves with union boxes is a legal production tree (the
// kernels decide candidates at the leaves, rto_internal.h), so results do not change; only the number of boxes a ray enters does.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <queue>
#include <vector>

namespace rto_treeopt {

struct Tree {
	// nodes 0 .. numInner - 1 are inner nodes, numInner .. 2 * numInner are leaves
	int numInner = 0, root = -1;
	std::vector<int> parent, left, right;
	std::vector<int32_t> leafRef;                 // per node; meaningful for leaves
	std::vector<float> lo, hi;                    // 3 floats per node

	static inline float area(const float* l, const float* h) { float dx = h[0] - l[0], dy = h[1] - l[1], dz = h[2] - l[2]; return dx * dy + dy * dz + dz * dx; }
	inline float areaOf(int n) const { return area(&lo[3 * (size_t)n], &hi[3 * (size_t)n]); }
	inline float unionArea(int a, const float* l, const float* h) const {
		const float* al = &lo[3 * (size_t)a]; const float* ah = &hi[3 * (size_t)a];
		float ul[3] = { std::min(al[0], l[0]), std::min(al[1], l[1]), std::min(al[2], l[2]) }, uh[3] = { std::max(ah[0], h[0]), std::max(ah[1], h[1]), std::max(ah[2], h[2]) };
		return area(ul, uh);
	}
	inline bool isLeaf(int n) const { return n >= numInner; }
	// box of an inner node from its children; true if it changed
	inline bool refit(int n) {
		const int a = left[n], b = right[n];
		bool changed = false;
		for (int k = 0; k < 3; k++) {
			float l = std::min(lo[3 * (size_t)a + k], lo[3 * (size_t)b + k]), h = std::max(hi[3 * (size_t)a + k], hi[3 * (size_t)b + k]);
			if (l != lo[3 * (size_t)n + k] || h != hi[3 * (size_t)n + k]) { lo[3 * (size_t)n + k] = l; hi[3 * (size_t)n + k] = h; changed = true; }
		}
		return changed;
	}
	inline void refitUp(int n) { while (n >= 0 && refit(n)) n = parent[n]; }

	// paired node array (host_builders.cpp SahBuilder::build) -> pointer tree
	void load(const std::vector<float>& buf, int32_t rootRef) {
		numInner = (int)(buf.size() / 16);
		const int T = 2 * numInner + 1;
		parent.assign(T, -1); left.assign(T, -1); right.assign(T, -1); leafRef.assign(T, 0); lo.assign(3 * (size_t)T, 0.0f); hi.assign(3 * (size_t)T, 0.0f);
		root = rootRef;
		int nextLeaf = numInner;
		std::vector<int> st{ rootRef };
		while (!st.empty()) {
			const int n = st.back(); st.pop_back();
			const float* d = &buf[(size_t)n * 16];
			for (int c = 0; c < 2; c++) {
				int32_t ref; std::memcpy(&ref, &d[12 + c], 4);
				int id;
				if (ref >= 0) { id = ref; st.push_back(ref); }
				else { id = nextLeaf++; leafRef[id] = ref; }
				(c ? right[n] : left[n]) = id; parent[id] = n;
				for (int k = 0; k < 3; k++) { lo[3 * (size_t)id + k] = d[2 * k + c]; hi[3 * (size_t)id + k] = d[6 + 2 * k + c]; }
			}
		}
		// the root's own box is not stored anywhere: union of its children (bottom-up order does not matter for one node)
		refit(root);
	}

	// pointer tree -> paired node array in pre-order (left subtree directly behind its parent); returns the deepest level
	int store(std::vector<float>& buf, int32_t& rootRef) const {
		std::vector<int> newIndex(numInner, -1);
		// pre-order numbering
		{
			int next = 0;
			std::vector<int> st{ root };
			while (!st.empty()) {
				const int n = st.back(); st.pop_back();
				newIndex[n] = next++;
				if (!isLeaf(right[n])) st.push_back(right[n]);
				if (!isLeaf(left[n])) st.push_back(left[n]);
			}
		}
		int maxDepth = 0;
		std::vector<std::pair<int, int>> st{ { root, 1 } };
		while (!st.empty()) {
			const auto [n, dep] = st.back(); st.pop_back();
			maxDepth = std::max(maxDepth, dep);
			float* d = &buf[(size_t)newIndex[n] * 16];
			const int ch[2] = { left[n], right[n] };
			for (int c = 0; c < 2; c++) {
				const int id = ch[c];
				for (int k = 0; k < 3; k++) { d[2 * k + c] = lo[3 * (size_t)id + k]; d[6 + 2 * k + c] = hi[3 * (size_t)id + k]; }
				int32_t ref = isLeaf(id) ? leafRef[id] : (int32_t)newIndex[id];
				std::memcpy(&d[12 + c], &ref, 4);
				if (!isLeaf(id)) st.push_back({ id, dep + 1 });
			}
			d[14] = 0.0f; d[15] = 0.0f;
		}
		rootRef = (int32_t)newIndex[root];
		return maxDepth;
	}

	double cost() const {          // sum of the areas of all nodes below the root / area of the root
		double s = 0; const int T = 2 * numInner + 1;
		for (int n = 0; n < T; n++) if (n != root) s += areaOf(n);
		return s / areaOf(root);
	}

	// take X out (its parent P goes with it, X's sibling moves up), find the best place, put it back under P.  Returns the area saved.
	float reinsert(int X, std::vector<std::pair<float, int>>& heap) {
		const int P = parent[X];
		if (P < 0) return 0.0f;
		const int S = left[P] == X ? right[P] : left[P];
		const int G = parent[P];
		const float xl[3] = { lo[3 * (size_t)X], lo[3 * (size_t)X + 1], lo[3 * (size_t)X + 2] }, xh[3] = { hi[3 * (size_t)X], hi[3 * (size_t)X + 1], hi[3 * (size_t)X + 2] };
		const float aX = area(xl, xh);
		// --- remove ---
		parent[S] = G;
		if (G >= 0) { (left[G] == P ? left[G] : right[G]) = S; refitUp(G); }
		else root = S;
		// --- search, best first on the induced cost (the growth of the boxes above the candidate) ---
		float best = 3.4e38f; int bestY = S;
		heap.clear();
		auto cmp = [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first > b.first; };
		heap.push_back({ 0.0f, root });
		while (!heap.empty()) {
			std::pop_heap(heap.begin(), heap.end(), cmp);
			const auto [ci, Y] = heap.back(); heap.pop_back();
			if (ci + aX >= best) break;
			const float cd = unionArea(Y, xl, xh);
			const float c = ci + cd;
			if (c < best) { best = c; bestY = Y; }
			const float cc = c - areaOf(Y);
			if (!isLeaf(Y) && cc + aX < best) {
				heap.push_back({ cc, left[Y] }); std::push_heap(heap.begin(), heap.end(), cmp);
				heap.push_back({ cc, right[Y] }); std::push_heap(heap.begin(), heap.end(), cmp);
			}
		}
		// --- insert: P takes bestY's place and holds bestY and X ---
		const int Y = bestY, Q = parent[Y];
		left[P] = Y; right[P] = X; parent[Y] = P; parent[X] = P; parent[P] = Q;
		if (Q >= 0) (left[Q] == Y ? left[Q] : right[Q]) = P; else root = P;
		refit(P);
		refitUp(Q);
		return 0.0f;
	}
};

// `passes` sweeps over the nodes in order of decreasing area; `fraction` of them (the largest) are re-inserted in every sweep
inline void optimize(Tree& t, int passes, double fraction) {
	const int T = 2 * t.numInner + 1;
	std::vector<std::pair<float, int>> order, heap;
	for (int p = 0; p < passes; p++) {
		order.clear();
		for (int n = 0; n < T; n++) if (n != t.root) order.push_back({ t.areaOf(n), n });
		std::sort(order.begin(), order.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first > b.first || (a.first == b.first && a.second < b.second); });
		const size_t count = (size_t)(order.size() * fraction);
		for (size_t i = 0; i < count; i++) if (order[i].second != t.root) t.reinsert(order[i].second, heap);
	}
}

} // namespace rto_treeopt
