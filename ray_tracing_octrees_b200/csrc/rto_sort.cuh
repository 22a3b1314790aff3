// rto_sort.cuh -- stable radix sort of 32-bit key / 32-bit value pairs on the device: the sort behind RTO_FLAG_SORT_RAYS (ray coherence
// sorting of explicit ray lists, rto_device.cu) and rto_device_sort_pairs.
//
// Least significant digit first, 8 bits per pass, four passes, three kernels per pass:
//   k_sort_count    one block per tile of 4096 keys: how many keys of the tile carry each digit            -> hist[digit][tile]
//   k_sort_scan     one block per digit: exclusive prefix over the tiles (in place) and the digit's total  -> hist, totals[digit]
//   k_sort_scatter  one block per tile: every key to  base(digit) + prefix(digit, tile) + its rank among the tile's keys of that digit
// A pass has to be stable for the next one to build on it, so the rank of a key counts the keys of its digit that come BEFORE it in
// the input: a tile is cut into eight runs of 512 consecutive keys, one per warp, which a warp reads 32 at a time; inside a warp the
// lanes holding the same digit find each other with __match_any_sync (rank = lanes below me in the group), the group's lowest lane
// advances the warp's counter for that digit in shared memory and hands the old value to the others by a shuffle; afterwards one
// thread per digit turns the eight per-warp counts into offsets.  (warp, round, lane) order is index order, so the sort is stable and
// the permutation it yields is the one any stable sort yields.
// No atomics on global memory, no inter-block waiting: the three kernels of a pass are ordered by the stream.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace rto {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortRounds = 16;                               // keys per thread
constexpr int kSortTile = kSortThreads * kSortRounds;         // 4096 keys per block
constexpr int kSortRun = 32 * kSortRounds;                    // 512 consecutive keys per warp

inline size_t sort_num_tiles(size_t n) { return (n + kSortTile - 1) / kSortTile; }
// scratch: hist[256][tiles] + totals[256]
inline size_t sort_scratch_bytes(size_t n) { return (256 * sort_num_tiles(n) + 256) * sizeof(uint32_t); }

__global__ void __launch_bounds__(kSortThreads) k_sort_count(const uint32_t* __restrict__ keys, size_t n, int shift, uint32_t* __restrict__ hist, size_t numTiles) {
	__shared__ uint32_t cnt[256];
	cnt[threadIdx.x] = 0;
	__syncthreads();
	const size_t base = (size_t)blockIdx.x * kSortTile;
	const int lane = threadIdx.x & 31;
	for (int j = 0; j < kSortRounds; j++) {
		const size_t i = base + (size_t)j * kSortThreads + threadIdx.x;
		const bool valid = i < n;
		const unsigned d = valid ? ((keys[i] >> shift) & 255u) : (256u + (unsigned)lane);
		const unsigned peers = __match_any_sync(0xffffffffu, d);
		if (valid && (peers & ((1u << lane) - 1u)) == 0) atomicAdd(&cnt[d], (uint32_t)__popc(peers));
	}
	__syncthreads();
	hist[(size_t)threadIdx.x * numTiles + blockIdx.x] = cnt[threadIdx.x];
}

// exclusive scan of 256 values held one per thread (kSortThreads == 256); returns the value for this thread, the block's total in *total
__device__ __forceinline__ uint32_t sort_block_exclusive(uint32_t v, uint32_t* warpSums /* [kSortWarps] shared */, uint32_t* total) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t inc = v;
	for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
	if (lane == 31) warpSums[warp] = inc;
	__syncthreads();
	uint32_t before = 0, all = 0;
	for (int w = 0; w < kSortWarps; w++) { const uint32_t s = warpSums[w]; if (w < warp) before += s; all += s; }
	__syncthreads();                                         // warpSums may be reused by the caller's next call
	*total = all;
	return before + inc - v;
}

__global__ void __launch_bounds__(kSortThreads) k_sort_scan(uint32_t* __restrict__ hist, size_t numTiles, uint32_t* __restrict__ totals) {
	__shared__ uint32_t warpSums[kSortWarps];
	uint32_t* row = hist + (size_t)blockIdx.x * numTiles;
	uint32_t carry = 0;
	for (size_t t0 = 0; t0 < numTiles; t0 += kSortThreads) {
		const size_t t = t0 + threadIdx.x;
		const uint32_t v = t < numTiles ? row[t] : 0u;
		uint32_t total;
		const uint32_t ex = sort_block_exclusive(v, warpSums, &total);
		if (t < numTiles) row[t] = carry + ex;
		carry += total;
	}
	if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(kSortThreads) k_sort_scatter(const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
	uint32_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut, size_t n, int shift,
	const uint32_t* __restrict__ hist, size_t numTiles, const uint32_t* __restrict__ totals) {
	__shared__ uint32_t wcount[kSortWarps][256];             // per warp: keys of each digit seen so far; later: keys of that digit in the warps before
	__shared__ uint32_t tileBase[256];                       // where the tile's keys of a digit start in the output
	__shared__ uint32_t warpSums[kSortWarps];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int w = 0; w < kSortWarps; w++) wcount[w][threadIdx.x] = 0;
	__syncthreads();
	const size_t runStart = (size_t)blockIdx.x * kSortTile + (size_t)warp * kSortRun;
	uint32_t key[kSortRounds];
	uint16_t rank[kSortRounds];
#pragma unroll
	for (int j = 0; j < kSortRounds; j++) {
		const size_t i = runStart + (size_t)j * 32 + lane;
		const bool valid = i < n;
		key[j] = valid ? keysIn[i] : 0u;
		const unsigned d = valid ? ((key[j] >> shift) & 255u) : (256u + (unsigned)lane);
		const unsigned peers = __match_any_sync(0xffffffffu, d);
		const unsigned below = peers & ((1u << lane) - 1u);
		uint32_t seen = 0;
		if (valid && below == 0) { seen = wcount[warp][d]; wcount[warp][d] = seen + (uint32_t)__popc(peers); }
		seen = __shfl_sync(0xffffffffu, seen, __ffs(peers) - 1);
		rank[j] = (uint16_t)(seen + (uint32_t)__popc(below));
		__syncwarp();
	}
	__syncthreads();
	// one thread per digit: counts of the eight warps -> offsets of the warps inside the tile's share of that digit
	{
		const int d = threadIdx.x;
		uint32_t running = 0;
		for (int w = 0; w < kSortWarps; w++) { const uint32_t c = wcount[w][d]; wcount[w][d] = running; running += c; }
		uint32_t all;
		const uint32_t digitBase = sort_block_exclusive(totals[d], warpSums, &all);       // keys with a smaller digit, whole input
		tileBase[d] = digitBase + hist[(size_t)d * numTiles + blockIdx.x];
	}
	__syncthreads();
#pragma unroll
	for (int j = 0; j < kSortRounds; j++) {
		const size_t i = runStart + (size_t)j * 32 + lane;
		if (i < n) {
			const unsigned d = (key[j] >> shift) & 255u;
			const size_t pos = (size_t)tileBase[d] + wcount[warp][d] + rank[j];
			keysOut[pos] = key[j];
			valsOut[pos] = valsIn[i];
		}
	}
}

// Sorts n pairs by key, stable.  keys0 / vals0 hold the input AND the result; keys1 / vals1 are buffers of the same size, scratch is
// sort_scratch_bytes(n).  Twelve launches on `st`, no host synchronisation.  n < 2^32.
inline cudaError_t sort_pairs_u32(uint32_t* keys0, uint32_t* vals0, uint32_t* keys1, uint32_t* vals1, size_t n, void* scratch, cudaStream_t st, uint64_t* launches = nullptr) {
	if (n < 2) return cudaSuccess;
	const size_t tiles = sort_num_tiles(n);
	uint32_t* hist = (uint32_t*)scratch;
	uint32_t* totals = hist + 256 * tiles;
	uint32_t *ki = keys0, *vi = vals0, *ko = keys1, *vo = vals1;
	for (int pass = 0; pass < 4; pass++) {
		const int shift = 8 * pass;
		k_sort_count<<<(unsigned)tiles, kSortThreads, 0, st>>>(ki, n, shift, hist, tiles);
		k_sort_scan<<<256, kSortThreads, 0, st>>>(hist, tiles, totals);
		k_sort_scatter<<<(unsigned)tiles, kSortThreads, 0, st>>>(ki, vi, ko, vo, n, shift, hist, tiles, totals);
		uint32_t* t = ki; ki = ko; ko = t; t = vi; vi = vo; vo = t;
	}
	if (launches) *launches += 12;
	return cudaGetLastError();                               // an even number of passes: the result is back in keys0 / vals0
}

} // namespace rto
