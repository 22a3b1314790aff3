// rto_group.cu -- several GPUs of one box behind one handle (include/rto_c.h "device groups").
//
// The path shards without communication until the frames have to be in ONE place: the scene is replicated on every device, the rows
// of a batch of frames are dealt to the devices as contiguous ranges, and every device but the first writes 4-byte hit codes -- not
// 24-byte pixels -- straight into the first device's memory from the trace kernel's epilogue (NVLink peer stores, one 128-byte line
// per warp).  The first device rebuilds the planes from the codes (k_resolve_bvh: same ray, same Moller-Trumbore arithmetic, same
// shading => the same bits as a direct render) on a second stream while it traces its own, smaller share.  Shares follow the
// measured time of every device (the first one also pays for the expansion), so the group finishes together.
//
// One host thread drives all devices; nothing here blocks except rto_group_sync and host-memory frames.
#include "rto_scene.cuh"
#include "rto_nvtx.h"

#include <cmath>
#include <cstring>
#include <new>
#include <vector>

namespace {

struct Member {
	int device = -1;
	RtoScene* scene = nullptr;
	cudaEvent_t evBegin = nullptr, evEnd = nullptr;       // timing of the last batch on this device (lead: incl. the expansion)
	std::vector<cudaEvent_t> evCodes;                      // one per chunk: "the codes of this chunk are in the lead's memory"
	double weight = 1.0;                                   // share of the rows of a batch
	bool timedOnce = false;
};

struct Range { int frame0, frames, y0, y1; };             // frames [frame0, frame0 + frames), rows [y0, y1) of each

// rows [g0, g1) of the batch's row space (frame f, row y  <->  f * H + y) as at most three launches: a partial first frame, whole
// frames, a partial last frame
void split_rows(long long g0, long long g1, int H, std::vector<Range>& out) {
	out.clear();
	if (g1 <= g0) return;
	long long f0 = g0 / H, f1 = (g1 - 1) / H;
	int ya = (int)(g0 - f0 * H), yb = (int)(g1 - f1 * H);
	if (f0 == f1) { out.push_back({ (int)f0, 1, ya, yb }); return; }
	if (ya != 0) { out.push_back({ (int)f0, 1, ya, H }); f0++; }
	long long fullEnd = (yb == H) ? f1 + 1 : f1;
	if (fullEnd > f0) out.push_back({ (int)f0, (int)(fullEnd - f0), 0, H });
	if (yb != H) out.push_back({ (int)f1, 1, 0, yb });
}

} // namespace

struct RtoGroup {
	std::vector<Member> m;
	cudaStream_t resolveStream = nullptr;                  // on the lead device
	cudaEvent_t evResolved[2] = { nullptr, nullptr };      // the lead has expanded everything out of code buffer k
	uint32_t* codes[2] = { nullptr, nullptr }; size_t codeWords = 0;   // on the lead device; batches alternate between the two
	int parity = 0;
	int chunks = 4;                                        // each device's share is traced in this many launches so that the lead can expand
	                                                       // chunk c while chunk c + 1 is traced
	void* planes[3] = { nullptr, nullptr, nullptr }; size_t planeBytes[3] = { 0, 0, 0 };   // lead-device staging for host-memory frames
	bool balance = true;
};

static int group_fail_cleanup(RtoGroup* g, int rc) { rto_group_destroy(g); return rc; }

extern "C" int rto_group_create(const int* devices, int numDevices, RtoGroup** out) try {
	RTO_RANGE("rto_group_create");
	if (!out) return rto_fail(RTO_ERR_INVALID, "rto_group_create: null output");
	*out = nullptr;
	if (!devices || numDevices <= 0 || numDevices > 64) return rto_fail(RTO_ERR_INVALID, "rto_group_create: 1 to 64 devices");
	int rc = rto_require_device(); if (rc) return rc;
	int have = 0; CUDA_TRY(cudaGetDeviceCount(&have));
	for (int i = 0; i < numDevices; i++) {
		if (devices[i] < 0 || devices[i] >= have) return rto_fail(RTO_ERR_INVALID, "rto_group_create: device %d does not exist (%d visible)", devices[i], have);
		for (int j = 0; j < i; j++) if (devices[j] == devices[i]) return rto_fail(RTO_ERR_INVALID, "rto_group_create: device %d listed twice", devices[i]);
	}
	RtoGroup* g = new (std::nothrow) RtoGroup();
	if (!g) return rto_fail(RTO_ERR_ALLOC, "out of host memory");
	g->m.resize(numDevices);
	for (int i = 0; i < numDevices; i++) {
		Member& M = g->m[i];
		M.device = devices[i];
		if ((rc = rto_init(M.device))) return group_fail_cleanup(g, rc);
		cudaError_t e = cudaEventCreate(&M.evBegin);
		if (e == cudaSuccess) e = cudaEventCreate(&M.evEnd);
		if (e != cudaSuccess) return group_fail_cleanup(g, rto_fail(RTO_ERR_CUDA, "rto_group_create: %s", cudaGetErrorString(e)));
		M.weight = (i == 0 && numDevices > 1) ? 0.55 : 1.0;       // the lead also expands everyone else's codes; refined from measured times
		if (i > 0) {
			int can = 0;
			CUDA_TRY(cudaDeviceCanAccessPeer(&can, M.device, devices[0]));
			if (!can) return group_fail_cleanup(g, rto_fail(RTO_ERR_UNSUPPORTED, "rto_group_create: device %d cannot access device %d's memory (no NVLink / PCIe peer path)", M.device, devices[0]));
			e = cudaDeviceEnablePeerAccess(devices[0], 0);
			if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
			if (e != cudaSuccess) return group_fail_cleanup(g, rto_fail(RTO_ERR_CUDA, "rto_group_create: cudaDeviceEnablePeerAccess(%d -> %d): %s", M.device, devices[0], cudaGetErrorString(e)));
		}
	}
	CUDA_TRY(cudaSetDevice(devices[0]));
	cudaError_t e = cudaStreamCreateWithFlags(&g->resolveStream, cudaStreamNonBlocking);
	for (int k = 0; k < 2 && e == cudaSuccess; k++) e = cudaEventCreateWithFlags(&g->evResolved[k], cudaEventDisableTiming);
	if (e != cudaSuccess) return group_fail_cleanup(g, rto_fail(RTO_ERR_CUDA, "rto_group_create: %s", cudaGetErrorString(e)));
	*out = g;
	return RTO_OK;
} RTO_CATCH_ALL("rto_group_create")

extern "C" void rto_group_destroy(RtoGroup* g) {
	if (!g) return;
	for (Member& M : g->m) {
		if (M.device < 0) continue;
		cudaSetDevice(M.device);
		if (M.scene) rto_scene_destroy(M.scene);
		if (M.evBegin) cudaEventDestroy(M.evBegin);
		if (M.evEnd) cudaEventDestroy(M.evEnd);
		for (cudaEvent_t e : M.evCodes) cudaEventDestroy(e);
	}
	if (!g->m.empty() && g->m[0].device >= 0) {
		cudaSetDevice(g->m[0].device);
		if (g->resolveStream) { cudaStreamSynchronize(g->resolveStream); cudaStreamDestroy(g->resolveStream); }
		for (int k = 0; k < 2; k++) { if (g->evResolved[k]) cudaEventDestroy(g->evResolved[k]); if (g->codes[k]) cudaFree(g->codes[k]); }
		for (void* p : g->planes) if (p) cudaFree(p);
	}
	delete g;
}

extern "C" int rto_group_size(const RtoGroup* g) { return g ? (int)g->m.size() : 0; }

extern "C" int rto_group_scene_bvh(RtoGroup* g, const RtoTriangle* tris, size_t numTris, const RtoHostBvh* prebuilt) try {
	RTO_RANGE("rto_group_scene_bvh");
	if (!g) return rto_fail(RTO_ERR_INVALID, "rto_group_scene_bvh: null group");
	if (numTris && !tris) return rto_fail(RTO_ERR_INVALID, "rto_group_scene_bvh: null triangles");
	BvhLayout L; size_t numRefNodes = 0;
	int rc = rto_bvh_layout_from_tris(tris, numTris, prebuilt, L, &numRefNodes); if (rc) return rc;      // trees built once, on the host
	for (Member& M : g->m) {
		CUDA_TRY(cudaSetDevice(M.device));
		if (M.scene) { rto_scene_destroy(M.scene); M.scene = nullptr; }
		if ((rc = rto_scene_from_bvh_layout(L, numTris, numRefNodes, &M.scene))) return rc;                 // ... uploaded to every device
	}
	return RTO_OK;
} RTO_CATCH_ALL("rto_group_scene_bvh")

extern "C" RtoScene* rto_group_scene(RtoGroup* g, int member) {
	return (g && member >= 0 && member < (int)g->m.size()) ? g->m[member].scene : nullptr;
}

extern "C" int rto_group_set_balancing(RtoGroup* g, int enabled, const float* weights, int chunks) try {
	if (!g) return rto_fail(RTO_ERR_INVALID, "rto_group_set_balancing: null group");
	if (chunks < 0 || chunks > 64) return rto_fail(RTO_ERR_INVALID, "rto_group_set_balancing: 1 to 64 chunks (0 keeps the setting)");
	if (chunks) g->chunks = chunks;
	g->balance = enabled != 0;
	if (weights) for (size_t i = 0; i < g->m.size(); i++) {
		if (!(weights[i] > 0.0f)) return rto_fail(RTO_ERR_INVALID, "rto_group_set_balancing: weights must be positive");
		g->m[i].weight = weights[i];
	}
	return RTO_OK;
} RTO_CATCH_ALL("rto_group_set_balancing")

extern "C" int rto_group_last_ms(RtoGroup* g, float* msPerDevice) try {
	if (!g || !msPerDevice) return rto_fail(RTO_ERR_INVALID, "rto_group_last_ms: null argument");
	for (size_t i = 0; i < g->m.size(); i++) {
		Member& M = g->m[i];
		msPerDevice[i] = 0.0f;
		if (!M.timedOnce) continue;
		CUDA_TRY(cudaSetDevice(M.device));
		CUDA_TRY(cudaEventSynchronize(M.evEnd));
		CUDA_TRY(cudaEventElapsedTime(&msPerDevice[i], M.evBegin, M.evEnd));
	}
	return RTO_OK;
} RTO_CATCH_ALL("rto_group_last_ms")

extern "C" int rto_group_sync(RtoGroup* g) try {
	if (!g) return rto_fail(RTO_ERR_INVALID, "rto_group_sync: null group");
	for (Member& M : g->m) if (M.scene) {
		CUDA_TRY(cudaSetDevice(M.device));
		CUDA_TRY(cudaStreamSynchronize(M.scene->stream));
	}
	CUDA_TRY(cudaSetDevice(g->m[0].device));
	CUDA_TRY(cudaStreamSynchronize(g->resolveStream));
	return RTO_OK;
} RTO_CATCH_ALL("rto_group_sync")

// shares from the measured times of the previous batch, if that batch has finished (never waits)
static void rebalance(RtoGroup* g) {
	const size_t n = g->m.size();
	if (!g->balance || n < 2) return;
	std::vector<float> ms(n, 0.0f);
	for (size_t i = 0; i < n; i++) {
		Member& M = g->m[i];
		if (!M.timedOnce) return;
		cudaSetDevice(M.device);
		if (cudaEventQuery(M.evEnd) != cudaSuccess) { cudaGetLastError(); return; }
		if (cudaEventElapsedTime(&ms[i], M.evBegin, M.evEnd) != cudaSuccess) { cudaGetLastError(); return; }
		if (!(ms[i] > 0.0f)) return;
	}
	double mean = 0; for (float v : ms) mean += v; mean /= n;
	double sum = 0;
	for (size_t i = 0; i < n; i++) { g->m[i].weight *= std::pow(mean / ms[i], 0.7); sum += g->m[i].weight; }
	for (size_t i = 0; i < n; i++) { g->m[i].weight = g->m[i].weight * n / sum; if (g->m[i].weight < 0.02) g->m[i].weight = 0.02; }
}

extern "C" int rto_group_render_batch(RtoGroup* g, const RtoCamera* cams, int numCams, uint32_t flags, float shadowBias, const RtoFrame* frame) try {
	RTO_RANGE("rto_group_render_batch");
	if (!g || !cams || !frame || numCams <= 0) return rto_fail(RTO_ERR_INVALID, "rto_group_render_batch: null argument");
	const size_t n = g->m.size();
	for (Member& M : g->m) if (!M.scene) return rto_fail(RTO_ERR_INVALID, "rto_group_render_batch: no scene (rto_group_scene_bvh first)");
	if (frame->memory != RTO_MEM_HOST && frame->memory != RTO_MEM_DEVICE) return rto_fail(RTO_ERR_INVALID, "rto_group_render_batch: bad RtoFrame.memory");
	const int W = cams[0].width, H = cams[0].height;
	if (W <= 0 || H <= 0) return rto_fail(RTO_ERR_INVALID, "rto_group_render_batch: bad image size");
	for (int c = 1; c < numCams; c++) if (cams[c].width != W || cams[c].height != H) return rto_fail(RTO_ERR_INVALID, "rto_group_render_batch: all cameras must share one image size");
	if (numCams > 65535 || (unsigned long long)numCams * H * W >= 0xffffffffull) return rto_fail(RTO_ERR_UNSUPPORTED, "rto_group_render_batch: batch too large");
	Member& lead = g->m[0];
	const bool host = frame->memory == RTO_MEM_HOST;
	const size_t npix = (size_t)numCams * H * W;
	int rc;

	rebalance(g);
	CUDA_TRY(cudaSetDevice(lead.device));
	// planes on the lead device
	float4* rgba = (float4*)frame->rgba; int32_t* hid = frame->hitId; float* t = frame->t;
	if (host) {
		void** want[3] = { (void**)&rgba, (void**)&hid, (void**)&t };
		const void* asked[3] = { frame->rgba, frame->hitId, frame->t };
		const size_t bytes[3] = { npix * 16, npix * 4, npix * 4 };
		for (int k = 0; k < 3; k++) {
			if (!asked[k]) continue;
			if (g->planeBytes[k] < bytes[k]) {
				CUDA_TRY(cudaStreamSynchronize(lead.scene->stream));
				if (g->planes[k]) cudaFree(g->planes[k]);
				g->planes[k] = nullptr; g->planeBytes[k] = 0;
				cudaError_t e = cudaMalloc(&g->planes[k], bytes[k]);
				if (e != cudaSuccess) return rto_fail(RTO_ERR_ALLOC, "rto_group_render_batch: cudaMalloc(%zu) failed: %s", bytes[k], cudaGetErrorString(e));
				g->planeBytes[k] = bytes[k];
			}
			*want[k] = g->planes[k];
		}
	}
	const size_t frameWords = rto_codes_frame_words(W, H);
	if (n > 1 && g->codeWords < frameWords * numCams) {
		(void)rto_group_sync(g);
		CUDA_TRY(cudaSetDevice(lead.device));
		for (int k = 0; k < 2; k++) {
			if (g->codes[k]) cudaFree(g->codes[k]);
			g->codes[k] = nullptr;
		}
		g->codeWords = 0;
		for (int k = 0; k < 2; k++) {
			cudaError_t e = cudaMalloc((void**)&g->codes[k], frameWords * numCams * 4);
			if (e != cudaSuccess) return rto_fail(RTO_ERR_ALLOC, "rto_group_render_batch: cudaMalloc(%zu) failed: %s", frameWords * numCams * 4, cudaGetErrorString(e));
		}
		g->codeWords = frameWords * numCams;
	}
	const int par = g->parity;
	g->parity ^= 1;
	uint32_t* codes = g->codes[par];
	const int C = g->chunks;

	// contiguous row ranges (in units of the 8-row tiles the kernels work in), proportional to the weights
	const long long tilesPerFrame = (H + 7) / 8, tiles = tilesPerFrame * numCams;
	std::vector<long long> cut(n + 1, 0);
	double wsum = 0; for (Member& M : g->m) wsum += M.weight;
	double acc = 0;
	for (size_t i = 0; i < n; i++) {
		acc += g->m[i].weight;
		long long c = (i + 1 == n) ? tiles : (long long)std::llround(acc / wsum * (double)tiles);
		if (c < cut[i]) c = cut[i];
		if (c > tiles) c = tiles;
		cut[i + 1] = c;
	}
	auto rowOf = [&](long long tile) { long long f = tile / tilesPerFrame, k = tile % tilesPerFrame; return f * H + k * 8; };      // (tile == tiles gives numCams * H)
	auto chunkCut = [&](size_t i, int c) { return cut[i] + (cut[i + 1] - cut[i]) * c / C; };

	std::vector<Range> ranges;
	// 1. every other device: its rows, chunk by chunk, codes into the lead's memory over NVLink
	for (size_t i = 1; i < n; i++) {
		Member& M = g->m[i];
		CUDA_TRY(cudaSetDevice(M.device));
		cudaStream_t st = M.scene->stream;
		while ((int)M.evCodes.size() < 2 * C) { cudaEvent_t e; CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); M.evCodes.push_back(e); }
		// the lead may still be expanding the batch before last out of this code buffer
		CUDA_TRY(cudaStreamWaitEvent(st, g->evResolved[par], 0));
		CUDA_TRY(cudaEventRecord(M.evBegin, st));
		for (int c = 0; c < C; c++) {
			split_rows(rowOf(chunkCut(i, c)), rowOf(chunkCut(i, c + 1)), H, ranges);
			for (const Range& R : ranges)
				if ((rc = rto_enqueue_render(M.scene, cams + R.frame0, R.frames, RTO_MODE_BVH, flags, shadowBias, R.y0, R.y1, nullptr, nullptr, nullptr, codes, (size_t)R.frame0, st))) return rc;
			CUDA_TRY(cudaEventRecord(M.evCodes[par * C + c], st));
		}
		CUDA_TRY(cudaEventRecord(M.evEnd, st));
		M.timedOnce = true;
	}
	// 2. the lead: its own rows straight into the planes ...
	CUDA_TRY(cudaSetDevice(lead.device));
	cudaStream_t ls = lead.scene->stream;
	CUDA_TRY(cudaEventRecord(lead.evBegin, ls));
	split_rows(rowOf(cut[0]), rowOf(cut[1]), H, ranges);
	for (const Range& R : ranges) {
		const size_t off = ((size_t)R.frame0 * H + R.y0) * W;
		if ((rc = rto_enqueue_render(lead.scene, cams + R.frame0, R.frames, RTO_MODE_BVH, flags, shadowBias, R.y0, R.y1,
			rgba ? rgba + off : nullptr, hid ? hid + off : nullptr, t ? t + off : nullptr, nullptr, 0, ls))) return rc;
	}
	CUDA_TRY(cudaEventRecord(lead.evEnd, ls));                    // (the lead's share is judged by when its own trace ends, with the expansion running beside it)
	lead.timedOnce = true;
	// 3. ... while its second stream expands the others' codes chunk by chunk as they arrive
	for (int c = 0; c < C && n > 1; c++)
		for (size_t i = 1; i < n; i++) {
			CUDA_TRY(cudaStreamWaitEvent(g->resolveStream, g->m[i].evCodes[par * C + c], 0));
			split_rows(rowOf(chunkCut(i, c)), rowOf(chunkCut(i, c + 1)), H, ranges);
			for (const Range& R : ranges) {
				const size_t off = ((size_t)R.frame0 * H + R.y0) * W;
				if ((rc = rto_enqueue_resolve(lead.scene, cams + R.frame0, R.frames, R.y0, R.y1, codes, (size_t)R.frame0,
					rgba ? rgba + off : nullptr, hid ? hid + off : nullptr, t ? t + off : nullptr, g->resolveStream))) return rc;
			}
		}
	CUDA_TRY(cudaEventRecord(g->evResolved[par], g->resolveStream));
	CUDA_TRY(cudaStreamWaitEvent(ls, g->evResolved[par], 0));     // the lead's stream is the one callers order their own work after
	if (host) {
		if (frame->rgba) CUDA_TRY(cudaMemcpyAsync(frame->rgba, rgba, npix * 16, cudaMemcpyDeviceToHost, ls));
		if (frame->hitId) CUDA_TRY(cudaMemcpyAsync(frame->hitId, hid, npix * 4, cudaMemcpyDeviceToHost, ls));
		if (frame->t) CUDA_TRY(cudaMemcpyAsync(frame->t, t, npix * 4, cudaMemcpyDeviceToHost, ls));
		CUDA_TRY(cudaStreamSynchronize(ls));
	}
	return RTO_OK;
} RTO_CATCH_ALL("rto_group_render_batch")

extern "C" void* rto_group_stream(const RtoGroup* g) { return (g && !g->m.empty() && g->m[0].scene) ? (void*)g->m[0].scene->stream : nullptr; }
