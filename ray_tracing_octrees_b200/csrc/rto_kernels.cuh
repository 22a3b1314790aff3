// rto_kernels.cuh -- hand-written sm_100a ray-casting kernels (device code only).
//
// Three traversal semantics, each bit-compatible with the CPU oracle (see oracle/):
//   BVH      BVH::query candidate set (BVH.cpp:76-113) + Moller-Trumbore closest hit / shadow any-hit
//            (SURVEY.md 8c rule), ordered traversal with t-pruning; RTO_FLAG_NO_PRUNE replays the
//            reference's full visit pattern.
//   mode A   octreeRaySkip (VolumeRaycastRenderer.cpp:50-155): back-to-front Hamming order, first finite
//            leaf result wins, (enterT, exitT) clamps handed from parent to child.
//   mode B   GLSL intersectOctreeIterative (RayTracerBVH.cpp:239-327): LIFO order 7..0, first solid box hit
//            wins, 512-step budget.
// Exactness rules: this file is compiled with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false; every
// expression keeps the reference's operation order (rto_math.h); min/max are spelled as selects with the
// reference's operand order so NaN/inf cases agree.
#pragma once
#include "rto_devtypes.h"
#include <cfloat>
#include <cstring>
#include <cuda_runtime.h>

// Traversal code is written once and is callable from device code (the product) and from host code (tests/emu, which runs
// the very same functions on the CPU so that logic errors are found without a GPU).  librto.so never calls them on the host.
#define RTO_DEV __host__ __device__ __forceinline__
#if defined(__CUDA_ARCH__)
#define RTO_LDG(p) __ldg(p)
#else
#define RTO_LDG(p) (*(p))
#endif

namespace rto {

RTO_DEV int f2i(float f) {
#if defined(__CUDA_ARCH__)
	return __float_as_int(f);
#else
	int i; std::memcpy(&i, &f, 4); return i;
#endif
}
RTO_DEV int clz32(unsigned v) {
#if defined(__CUDA_ARCH__)
	return __clz((int)v);
#else
	return v ? __builtin_clz(v) : 32;
#endif
}
RTO_DEV int popc32(unsigned v) {
#if defined(__CUDA_ARCH__)
	return __popc(v);
#else
	return __builtin_popcount(v);
#endif
}
RTO_DEV int ffs32(unsigned v) {
#if defined(__CUDA_ARCH__)
	return __ffs((int)v);
#else
	return __builtin_ffs((int)v);
#endif
}

// Loop form of the BVH walks (measured on B200, 16 x 1080p per launch, single loop -> while-while: closest hit C2 2.909 -> 2.829 ms,
// primary only 2.021 -> 1.940, 512^3 city mesh 5.025 -> 4.826, sphere 0.901 -> 0.875; any-hit in the same form: C2 +1 % slower, city
// 2 % faster -- left in the single-loop form; parking the leaves of the any-hit walk and testing them behind the loop, where the lanes
// have re-joined: 8-36 % slower, a shadowed ray that walks on costs more than the fuller Moller-Trumbore block saves).  profiles/README.md, round 2.
#ifndef RTO_BVH_WHILE_WHILE
#define RTO_BVH_WHILE_WHILE 1
#endif
#ifndef RTO_RESOLVE_HOIST_RAY
#define RTO_RESOLVE_HOIST_RAY 0
#endif
#ifndef RTO_BVH_ANY_WHILE_WHILE
#define RTO_BVH_ANY_WHILE_WHILE 0
#endif

__device__ __host__ __forceinline__ void load_node(const float4* __restrict__ n, float4& a, float4& b, float4& c, float2& r) {
	// (two 256-bit loads -- LDG.E.256, new on sm_100 -- instead of these four were measured 1-4 % slower: profiles/README.md, round 2)
	a = RTO_LDG(n); b = RTO_LDG(n + 1); c = RTO_LDG(n + 2);
	r = RTO_LDG(reinterpret_cast<const float2*>(n + 3));
}

struct Ray { V3 o, d; };
struct alignas(8) StackEnt { int ref; float t; };     // postponed far child of the ordered BVH traversal and its box entry distance

// 3-input min / max: one FMNMX3 on sm_100 (the compiler fuses nested fmaxf only sometimes; when it sees common sub-expressions
// across the 8 children of an octree node it prefers 24 two-input operations to 16 three-input ones, and the ALU pipe that
// executes them is what bounds the octree kernels).  NaN-free inputs only: callers guarantee it.
RTO_DEV float fmax3f(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
	float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
#else
	return fmaxf(fmaxf(a, b), c);
#endif
}
RTO_DEV float fmin3f(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
	float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
#else
	return fminf(fminf(a, b), c);
#endif
}

// ------------------------------------------------------------------------------------------------
// Pixel ray: GLSL generateRay (RayTracerBVH.cpp:338-355) with inverse(view), tan(fov/2) from the host
// ------------------------------------------------------------------------------------------------
RTO_DEV Ray gen_ray(const RtoCamera& c, int px, int py) {
	float nx = (float(px) + 0.5f) / float(c.width) * 2.0f - 1.0f;
	float ny = 1.0f - (float(py) + 0.5f) / float(c.height) * 2.0f;
	nx *= c.aspect;
	nx *= c.tanHalfFov;
	ny *= c.tanHalfFov;
	// normalize(vec4(nx, ny, -1, 0)): dot4 = (x*x + y*y) + (z*z + w*w)
	float len2 = (nx * nx + ny * ny) + ((-1.0f) * (-1.0f) + 0.0f * 0.0f);
	float il = 1.0f / sqrtf(len2);
	float vx = nx * il, vy = ny * il, vz = -1.0f * il, vw = 0.0f * il;
	// mat4 * vec4 = (m0*v0 + m1*v1) + (m2*v2 + m3*v3), column-major
	const float* m = c.invView;
	float wx = (m[0] * vx + m[4] * vy) + (m[8] * vz + m[12] * vw);
	float wy = (m[1] * vx + m[5] * vy) + (m[9] * vz + m[13] * vw);
	float wz = (m[2] * vx + m[6] * vy) + (m[10] * vz + m[14] * vw);
	Ray r;
	r.o = mk3(c.camPos[0], c.camPos[1], c.camPos[2]);
	r.d = normalize3(mk3(wx, wy, wz));
	return r;
}

// ------------------------------------------------------------------------------------------------
// BVH
// ------------------------------------------------------------------------------------------------
struct RayBox {            // per-ray constants of BVH::query (BVH.cpp:107-113)
	V3 o, inv; bool nx, ny, nz;
	V3 noi;                // -(o * inv), for the fused conservative node test
};
RTO_DEV RayBox make_raybox(V3 o, V3 d) {
	RayBox r; r.o = o;
	r.inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
	r.nx = r.inv.x < 0; r.ny = r.inv.y < 0; r.nz = r.inv.z < 0;
	r.noi = -(o * r.inv);
	return r;
}

// intersectAABB (BVH.cpp:76-87) with tmin = 0, tmax = FLT_MAX.  The per-axis early-outs of the reference are
// equivalent to one final test because tmin only grows and tmax only shrinks (NaN candidates are never taken).
RTO_DEV bool slab_ref(const RayBox& r, float lox, float loy, float loz, float hix, float hiy, float hiz, float& tEntry) {
	float tmin = 0.0f, tmax = FLT_MAX;
	float t0 = ((r.nx ? hix : lox) - r.o.x) * r.inv.x;
	float t1 = ((r.nx ? lox : hix) - r.o.x) * r.inv.x;
	tmin = t0 > tmin ? t0 : tmin; tmax = t1 < tmax ? t1 : tmax;
	t0 = ((r.ny ? hiy : loy) - r.o.y) * r.inv.y;
	t1 = ((r.ny ? loy : hiy) - r.o.y) * r.inv.y;
	tmin = t0 > tmin ? t0 : tmin; tmax = t1 < tmax ? t1 : tmax;
	t0 = ((r.nz ? hiz : loz) - r.o.z) * r.inv.z;
	t1 = ((r.nz ? loz : hiz) - r.o.z) * r.inv.z;
	tmin = t0 > tmin ? t0 : tmin; tmax = t1 < tmax ? t1 : tmax;
	tEntry = tmin;
	return !(tmax < tmin);
}

// Rays with a zero or infinite reciprocal component keep the select form above (NaN cases); all others may use the
// octant-specialised form below.
RTO_DEV bool ray_needs_exact_box(const RayBox& rb) {
	float ax = fabsf(rb.inv.x), ay = fabsf(rb.inv.y), az = fabsf(rb.inv.z);
	return !(ax > 0.0f && ax <= FLT_MAX && ay > 0.0f && ay <= FLT_MAX && az > 0.0f && az <= FLT_MAX);
}

// The device record holds v0 and the two edges e1 = v1 - v0, e2 = v2 - v0 the rule starts from (one IEEE subtraction each,
// done once on the host: same bits as computing them per test), so a test costs 6 subtractions less.
struct TriV { V3 v0, e1, e2; int id; };
RTO_DEV TriV load_tri(const float4* __restrict__ tris, int pos) {
	float4 a = RTO_LDG(tris + 4 * (size_t)pos), b = RTO_LDG(tris + 4 * (size_t)pos + 1);
	float2 c = RTO_LDG(reinterpret_cast<const float2*>(tris + 4 * (size_t)pos + 2));
	TriV t;
	t.v0 = mk3(a.x, a.y, a.z); t.e1 = mk3(a.w, b.x, b.y); t.e2 = mk3(b.z, b.w, c.x); t.id = f2i(c.y);
	return t;
}

// Moller-Trumbore, SURVEY.md 8c rule (every comparison rejects NaN)
RTO_DEV bool moller_trumbore(const TriV& tri, V3 o, V3 d, float& tOut) {
	const V3 e1 = tri.e1, e2 = tri.e2;
	V3 p = cross3(d, e2);
	float det = dot3(e1, p);
	if (!(fabsf(det) >= 1e-8f)) return false;
	float inv = 1.0f / det;
	V3 s = o - tri.v0;
	float u = dot3(s, p) * inv;
	if (!(u >= 0.0f && u <= 1.0f)) return false;
	V3 q = cross3(s, e1);
	float v = dot3(d, q) * inv;
	if (!(v >= 0.0f && u + v <= 1.0f)) return false;
	float t = dot3(e2, q) * inv;
	if (!(t > 1e-4f)) return false;
	tOut = t;
	return true;
}

// Octant-specialised CONSERVATIVE form for trees whose boxes are grown (BvhDev::grow > 0).  With the signs of 1/d known at compile
// time (OCT bit a set <=> 1/d negative on axis a) the near and far plane of every axis are known, so the selects of BVH.cpp:78-86
// disappear and the running max/min collapse into 3-input min/max (FMNMX3 on sm_100); and because no decision the reference
// makes hangs on these boxes any more -- candidates are decided at the leaf, by the exact box of the reference leaf and the exact
// Moller-Trumbore test -- each plane distance is ONE fused multiply-add, plane * (1/d) - o * (1/d), instead of a subtraction and
// a product.  Compared with (plane - o) * (1/d) the fused form is off by at most the rounding of o * (1/d), which equals moving
// the plane by |o| * 2^-24: bvh_fused_ok() admits a ray only when that is below grow / 4 (an origin within 16 scene extents).  A warp of primary rays almost
// always shares one octant, and every shadow ray does.
template <int OCT>
RTO_DEV bool slab_oct(V3 noi, V3 inv, float lox, float loy, float loz, float hix, float hiy, float hiz, float tcap, float& tEntry) {
	float t0x = fmaf((OCT & 1) ? hix : lox, inv.x, noi.x), t1x = fmaf((OCT & 1) ? lox : hix, inv.x, noi.x);
	float t0y = fmaf((OCT & 2) ? hiy : loy, inv.y, noi.y), t1y = fmaf((OCT & 2) ? loy : hiy, inv.y, noi.y);
	float t0z = fmaf((OCT & 4) ? hiz : loz, inv.z, noi.z), t1z = fmaf((OCT & 4) ? loz : hiz, inv.z, noi.z);
	float tmin = fmaxf(fmaxf(fmaxf(t0x, t0y), t0z), 0.0f);
	// tcap = FLT_MAX gives intersectAABB's tmax; tcap = the pruning distance (< FLT_MAX) folds "entry <= tcap" into the same
	// comparison: tmin <= min(tmax, FLT_MAX) && tmin <= tcap  <=>  tmin <= min(tmax, tcap)
	float tmax = fminf(fminf(fminf(t1x, t1y), t1z), tcap);
	tEntry = tmin;
	return !(tmax < tmin);
}
constexpr int kOctGeneric = 8;
RTO_DEV bool bvh_fused_ok(const BvhDev& S, const RayBox& rb) {
	const float far = fmaxf(fmaxf(fabsf(rb.o.x), fabsf(rb.o.y)), fabsf(rb.o.z));
	// o / d must not overflow either (a direction component below ~1e-35): inf - inf would make the fused form miss boxes
	const bool finite = fabsf(rb.noi.x) <= FLT_MAX && fabsf(rb.noi.y) <= FLT_MAX && fabsf(rb.noi.z) <= FLT_MAX;
	return S.grow > 0.0f && finite && far * (1.0f / 4194304.0f) <= S.grow;      // |o| * 2^-24 <= grow / 4 (NaN origins fail the test)
}
// 0..7: fused octant-specialised node tests; kOctGeneric: the reference's exact select form
RTO_DEV int ray_octant(const BvhDev& S, const RayBox& rb) {
	if (ray_needs_exact_box(rb) || !bvh_fused_ok(S, rb)) return kOctGeneric;
	return (rb.nx ? 1 : 0) | (rb.ny ? 2 : 0) | (rb.nz ? 4 : 0);
}

// packed fp32 x 2 fused multiply-add: one FFMA2 on sm_100 (fma.rn.f32x2), two fmaf elsewhere (tests/emu).  Each half is an IEEE fma.
RTO_DEV float2 fma2(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__)
	unsigned long long ra, rb, rc, rd;
	asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
	asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
	asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
	float2 d;
	asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
	return d;
#else
	return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
// per-ray constants of the paired node test: 1/d and -o/d of every axis twice
struct RayBox2 { float2 ix, iy, iz, nx, ny, nz; };
RTO_DEV RayBox2 make_raybox2(const RayBox& rb) {
	RayBox2 r;
	r.ix = make_float2(rb.inv.x, rb.inv.x); r.iy = make_float2(rb.inv.y, rb.inv.y); r.iz = make_float2(rb.inv.z, rb.inv.z);
	r.nx = make_float2(rb.noi.x, rb.noi.x); r.ny = make_float2(rb.noi.y, rb.noi.y); r.nz = make_float2(rb.noi.z, rb.noi.z);
	return r;
}
// slab_oct for both children of a node in the paired layout (BvhDev::paired): six FFMA2 instead of twelve FFMA
template <int OCT>
RTO_DEV void slab_oct_pair(const RayBox2& r, float4 a, float4 b, float4 c, float tcap, bool& h0, bool& h1, float& e0, float& e1) {
	const float2 loX = make_float2(a.x, a.y), loY = make_float2(a.z, a.w), loZ = make_float2(b.x, b.y);
	const float2 hiX = make_float2(b.z, b.w), hiY = make_float2(c.x, c.y), hiZ = make_float2(c.z, c.w);
	const float2 t0x = fma2((OCT & 1) ? hiX : loX, r.ix, r.nx), t1x = fma2((OCT & 1) ? loX : hiX, r.ix, r.nx);
	const float2 t0y = fma2((OCT & 2) ? hiY : loY, r.iy, r.ny), t1y = fma2((OCT & 2) ? loY : hiY, r.iy, r.ny);
	const float2 t0z = fma2((OCT & 4) ? hiZ : loZ, r.iz, r.nz), t1z = fma2((OCT & 4) ? loZ : hiZ, r.iz, r.nz);
	e0 = fmaxf(fmaxf(fmaxf(t0x.x, t0y.x), t0z.x), 0.0f);
	e1 = fmaxf(fmaxf(fmaxf(t0x.y, t0y.y), t0z.y), 0.0f);
	const float x0 = fminf(fminf(fminf(t1x.x, t1y.x), t1z.x), tcap), x1 = fminf(fminf(fminf(t1x.y, t1y.y), t1z.y), tcap);
	h0 = !(x0 < e0); h1 = !(x1 < e1);
}

// both child boxes of an inner node (layout in BvhDev)
// h0/h1: the child's box passes intersectAABB and is entered no later than tcap (tcap = FLT_MAX: no pruning)
// Trees with grown boxes (the only ones the fused tests run on) are stored in the paired layout; the reference-shaped tree is not.
template <int OCT>
RTO_DEV void node_boxes(const RayBox& rb, const RayBox2& rb2, int paired, float4 a, float4 b, float4 c, float tcap, bool& h0, bool& h1, float& e0, float& e1) {
	if (OCT < kOctGeneric) slab_oct_pair<OCT>(rb2, a, b, c, tcap, h0, h1, e0, e1);
	else if (paired) {
		h0 = slab_ref(rb, a.x, a.z, b.x, b.z, c.x, c.z, e0) && (e0 <= tcap);
		h1 = slab_ref(rb, a.y, a.w, b.y, b.w, c.y, c.w, e1) && (e1 <= tcap);
	}
	else {
		h0 = slab_ref(rb, a.x, a.y, a.z, a.w, b.x, b.y, e0) && (e0 <= tcap);
		h1 = slab_ref(rb, b.z, b.w, c.x, c.y, c.z, c.w, e1) && (e1 <= tcap);
	}
}

// the box of the reference leaf a triangle belongs to (triangle record, BvhDev): the test BVH::queryNode makes before it emits the
// leaf's triangles, for trees whose own leaves are finer than the reference's (BvhDev::leafBox)
template <int OCT>
RTO_DEV bool ref_leaf_box_passes(const BvhDev& S, const RayBox& rb, int pos, float tcap) {
	const float4* rec = S.tris + 4 * (size_t)pos;
	float2 c = RTO_LDG(reinterpret_cast<const float2*>(rec + 2) + 1);
	float4 d = RTO_LDG(rec + 3);
	float e;
	return slab_ref(rb, c.x, c.y, d.x, d.y, d.z, d.w, e) && (e <= tcap);       // the reference's own arithmetic: this one decides
}

// ------------------------------------------------------------------------------------------------
// 4-wide quantised form of the production tree (BvhDev::wide)
// ------------------------------------------------------------------------------------------------
// A plane of a child box is a 16-bit grid position q: plane = wideLo + q * step.  Its distance along the ray,
// (plane - o) / d = q * (step / d) + (wideLo - o) / d, is ONE fused multiply-add per plane once q is a float, and q becomes a float
// without a conversion instruction: the bit pattern 0x4B000000 | q is the float 2^23 + q (one PRMT / LOP3), and the 2^23 is taken out
// again inside the per-ray constant, b' = (wideLo - o)/d - 2^23 * (step/d).  The quantised boxes were rounded outwards and widened
// by what this arithmetic and the leaf-level test can be off by (host_builders.cpp rto_build_wide_topology), so a ray that passes
// the test of a triangle's own box passes the test of every wide node above it: candidates are decided at the leaves, exactly as
// with the binary form, and every result is the same bit for bit.  Per node: 64 bytes instead of 64 bytes per pair of children,
// and half as many dependent fetches on the way down.
struct WideRay { float2 sx, sy, sz, bx, by, bz; };
RTO_DEV bool make_wideray(const BvhDev& S, const RayBox& rb, WideRay& w) {
	const float sx = S.wideStep * rb.inv.x, sy = S.wideStep * rb.inv.y, sz = S.wideStep * rb.inv.z;
	const float bx = fmaf(-8388608.0f, sx, (S.wideLo[0] - rb.o.x) * rb.inv.x);
	const float by = fmaf(-8388608.0f, sy, (S.wideLo[1] - rb.o.y) * rb.inv.y);
	const float bz = fmaf(-8388608.0f, sz, (S.wideLo[2] - rb.o.z) * rb.inv.z);
	w.sx = make_float2(sx, sx); w.sy = make_float2(sy, sy); w.sz = make_float2(sz, sz);
	w.bx = make_float2(bx, bx); w.by = make_float2(by, by); w.bz = make_float2(bz, bz);
	// 2^23 * step / d must not overflow (a direction component below ~1e-30): such rays stay on the binary form
	return fabsf(bx) <= FLT_MAX && fabsf(by) <= FLT_MAX && fabsf(bz) <= FLT_MAX &&
		fabsf(sx) <= 1e30f && fabsf(sy) <= 1e30f && fabsf(sz) <= 1e30f;
}
RTO_DEV float i2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
	return __uint_as_float(u);
#else
	float f; std::memcpy(&f, &u, 4); return f;
#endif
}
// (inline PTX with the selector as an immediate and 0x4B000000 in a register: left to itself the compiler builds the low half from two
// LOP3 and re-materialises the selector of the high half over and over -- 44 instructions per node for 24 conversions)
RTO_DEV float wide_lo(uint32_t w) {                                                  // 2^23 + low half
#if defined(__CUDA_ARCH__)
	uint32_t r; asm("prmt.b32 %0, %1, %2, 0x7610;" : "=r"(r) : "r"(w), "r"(0x4B000000u)); return __uint_as_float(r);
#else
	return i2f(0x4B000000u | (w & 0xffffu));
#endif
}
RTO_DEV float wide_hi(uint32_t w) {                                                  // 2^23 + high half
#if defined(__CUDA_ARCH__)
	uint32_t r; asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(w), "r"(0x4B000000u)); return __uint_as_float(r);
#else
	return i2f(0x4B000000u | (w >> 16));
#endif
}
// entry distances of the four children (clamped at 0) and the mask of those whose box is hit no later than tcap
template <int OCT>
RTO_DEV unsigned wide_test(const WideRay& R, uint4 X, uint4 Y, uint4 Z, float tcap, float tn[4]) {
	const uint32_t xs[4] = { X.x, X.y, X.z, X.w }, ys[4] = { Y.x, Y.y, Y.z, Y.w }, zs[4] = { Z.x, Z.y, Z.z, Z.w };
	unsigned mask = 0;
#pragma unroll
	for (int p = 0; p < 4; p += 2) {          // two children per packed fused multiply-add
		const float2 nx = fma2(make_float2((OCT & 1) ? wide_hi(xs[p]) : wide_lo(xs[p]), (OCT & 1) ? wide_hi(xs[p + 1]) : wide_lo(xs[p + 1])), R.sx, R.bx);
		const float2 fx = fma2(make_float2((OCT & 1) ? wide_lo(xs[p]) : wide_hi(xs[p]), (OCT & 1) ? wide_lo(xs[p + 1]) : wide_hi(xs[p + 1])), R.sx, R.bx);
		const float2 ny = fma2(make_float2((OCT & 2) ? wide_hi(ys[p]) : wide_lo(ys[p]), (OCT & 2) ? wide_hi(ys[p + 1]) : wide_lo(ys[p + 1])), R.sy, R.by);
		const float2 fy = fma2(make_float2((OCT & 2) ? wide_lo(ys[p]) : wide_hi(ys[p]), (OCT & 2) ? wide_lo(ys[p + 1]) : wide_hi(ys[p + 1])), R.sy, R.by);
		const float2 nz = fma2(make_float2((OCT & 4) ? wide_hi(zs[p]) : wide_lo(zs[p]), (OCT & 4) ? wide_hi(zs[p + 1]) : wide_lo(zs[p + 1])), R.sz, R.bz);
		const float2 fz = fma2(make_float2((OCT & 4) ? wide_lo(zs[p]) : wide_hi(zs[p]), (OCT & 4) ? wide_lo(zs[p + 1]) : wide_hi(zs[p + 1])), R.sz, R.bz);
		const float e0 = fmaxf(fmax3f(nx.x, ny.x, nz.x), 0.0f), e1 = fmaxf(fmax3f(nx.y, ny.y, nz.y), 0.0f);
		const float x0 = fminf(fmin3f(fx.x, fy.x, fz.x), tcap), x1 = fminf(fmin3f(fx.y, fy.y, fz.y), tcap);
		tn[p] = e0; tn[p + 1] = e1;
		if (!(x0 < e0)) mask |= 1u << p;
		if (!(x1 < e1)) mask |= 2u << p;
	}
	return mask;
}

constexpr int kWideDone = (int)0x80000000;        // == kWideEmpty: never pushed (an empty slot is never hit), never a leaf reference

// closest hit on the wide form: same rule, same leaves, same results as bvh_closest_loop<true, OCT>
template <int OCT>
RTO_DEV void bvh_closest_wide_loop(const BvhDev& S, const RayBox& rb, const WideRay& wr, V3 o, V3 d, float& bestT, int& bestPos) {
	StackEnt stack[kWideStackDev];
	int sp = 0;
	int cur = S.wideRoot;
	float tcut = kMissT * kPruneSlack;
	for (;;) {
		while (cur >= 0) {
			const uint4* n = S.wide + 4 * (size_t)cur;
			const uint4 X = RTO_LDG(n), Y = RTO_LDG(n + 1), Z = RTO_LDG(n + 2), R = RTO_LDG(n + 3);
			float tn[4];
			const unsigned mask = wide_test<OCT>(wr, X, Y, Z, tcut, tn);
			const int refs[4] = { (int)R.x, (int)R.y, (int)R.z, (int)R.w };
			// on to the nearest child that is hit, the others wait with their entry distances (selects only: no array is indexed at run time)
			int best = -1, bestRef = kWideDone; float bt = FLT_MAX;
#pragma unroll
			for (int k = 0; k < 4; k++) { const bool nearer = ((mask >> k) & 1u) && tn[k] < bt; bt = nearer ? tn[k] : bt; best = nearer ? k : best; bestRef = nearer ? refs[k] : bestRef; }
#pragma unroll
			for (int k = 0; k < 4; k++) if (((mask >> k) & 1u) && k != best && sp < kWideStackDev) { StackEnt e; e.ref = refs[k]; e.t = tn[k]; stack[sp++] = e; }
			if (best >= 0) cur = bestRef;
			else {
				cur = kWideDone;
				while (sp > 0) {
					StackEnt e = stack[--sp];
					if (e.t <= tcut) { cur = e.ref; break; }
				}
			}
		}
		if (cur == kWideDone) break;
		{
			const int pos = (~cur) >> 1;                  // single-triangle leaves
			TriV tri = load_tri(S.tris, pos);
			float t;
			if (moller_trumbore(tri, o, d, t)) {
				if ((t < bestT || (t == bestT && pos < bestPos)) && ref_leaf_box_passes<OCT>(S, rb, pos, FLT_MAX)) {
					bestT = t; bestPos = pos; tcut = t * kPruneSlack;
				}
			}
		}
		cur = kWideDone;
		while (sp > 0) {
			StackEnt e = stack[--sp];
			if (e.t <= tcut) { cur = e.ref; break; }
		}
		if (cur == kWideDone) break;
	}
}

template <int OCT>
RTO_DEV bool bvh_any_wide_loop(const BvhDev& S, const RayBox& rb, const WideRay& wr, V3 o, V3 d) {
	int stackRef[kWideStackDev];
	int sp = 0;
	int cur = S.wideRoot;
	while (true) {
		if (cur >= 0) {
			const uint4* n = S.wide + 4 * (size_t)cur;
			const uint4 X = RTO_LDG(n), Y = RTO_LDG(n + 1), Z = RTO_LDG(n + 2), R = RTO_LDG(n + 3);
			float tn[4];
			const unsigned mask = wide_test<OCT>(wr, X, Y, Z, FLT_MAX, tn);
			const int refs[4] = { (int)R.x, (int)R.y, (int)R.z, (int)R.w };
			int best = -1, bestRef = kWideDone; float bt = FLT_MAX;
#pragma unroll
			for (int k = 0; k < 4; k++) { const bool nearer = ((mask >> k) & 1u) && tn[k] < bt; bt = nearer ? tn[k] : bt; best = nearer ? k : best; bestRef = nearer ? refs[k] : bestRef; }
#pragma unroll
			for (int k = 0; k < 4; k++) if (((mask >> k) & 1u) && k != best && sp < kWideStackDev) stackRef[sp++] = refs[k];
			if (best >= 0) { cur = bestRef; continue; }
		}
		else {
			const int pos = (~cur) >> 1;
			TriV tri = load_tri(S.tris, pos);
			float t;
			if (moller_trumbore(tri, o, d, t) && ref_leaf_box_passes<OCT>(S, rb, pos, FLT_MAX)) return true;
		}
		if (sp == 0) break;
		cur = stackRef[--sp];
	}
	return false;
}

// rays admitted to the fused tests (oct < 8) whose wide constants are finite walk the wide form; false: the caller goes on with the binary form
template <bool WIDE>
RTO_DEV bool bvh_closest_wide(const BvhDev& S, const RayBox& rb, int oct, V3 o, V3 d, float& bestT, int& bestPos) {
	if (!WIDE || oct >= kOctGeneric || !S.wide) return false;
	WideRay wr;
	if (!make_wideray(S, rb, wr)) return false;
	switch (oct) {
	case 0: bvh_closest_wide_loop<0>(S, rb, wr, o, d, bestT, bestPos); break;
	case 1: bvh_closest_wide_loop<1>(S, rb, wr, o, d, bestT, bestPos); break;
	case 2: bvh_closest_wide_loop<2>(S, rb, wr, o, d, bestT, bestPos); break;
	case 3: bvh_closest_wide_loop<3>(S, rb, wr, o, d, bestT, bestPos); break;
	case 4: bvh_closest_wide_loop<4>(S, rb, wr, o, d, bestT, bestPos); break;
	case 5: bvh_closest_wide_loop<5>(S, rb, wr, o, d, bestT, bestPos); break;
	case 6: bvh_closest_wide_loop<6>(S, rb, wr, o, d, bestT, bestPos); break;
	default: bvh_closest_wide_loop<7>(S, rb, wr, o, d, bestT, bestPos); break;
	}
	return true;
}
template <bool WIDE>
RTO_DEV bool bvh_any_wide(const BvhDev& S, const RayBox& rb, int oct, V3 o, V3 d, bool& hit) {
	if (!WIDE || oct >= kOctGeneric || !S.wide) return false;
	WideRay wr;
	if (!make_wideray(S, rb, wr)) return false;
	switch (oct) {
	case 0: hit = bvh_any_wide_loop<0>(S, rb, wr, o, d); break;
	case 1: hit = bvh_any_wide_loop<1>(S, rb, wr, o, d); break;
	case 2: hit = bvh_any_wide_loop<2>(S, rb, wr, o, d); break;
	case 3: hit = bvh_any_wide_loop<3>(S, rb, wr, o, d); break;
	case 4: hit = bvh_any_wide_loop<4>(S, rb, wr, o, d); break;
	case 5: hit = bvh_any_wide_loop<5>(S, rb, wr, o, d); break;
	case 6: hit = bvh_any_wide_loop<6>(S, rb, wr, o, d); break;
	default: hit = bvh_any_wide_loop<7>(S, rb, wr, o, d); break;
	}
	return true;
}

// Closest hit.  Result = min over the reference's candidate set of (t, position in candidate order), i.e. the
// oracle's "strict <, first candidate wins".  PRUNE: near-child-first order and subtrees entered only while their
// box entry <= best * kPruneSlack.  !PRUNE: every box the reference's queryNode would test is tested.
template <bool PRUNE, int OCT>
RTO_DEV void bvh_closest_loop(const BvhDev& S, const RayBox& rb, V3 o, V3 d, float& bestT, int& bestPos) {
	StackEnt stack[kBvhStack];                 // one 8-byte local store / load per push / pop
	int sp = 0;
	int cur = S.rootRef;
	const RayBox2 rb2 = make_raybox2(rb);
	float tcut = kMissT * kPruneSlack;
#if RTO_BVH_WHILE_WHILE
	// while-while form: the lanes of a warp walk inner nodes until each of them stands at a leaf (or is done) and re-join at the end of
	// the inner loop, so that the Moller-Trumbore block below runs once for all of them instead of once per lane that happens to reach
	// a leaf (in the single-loop form that block ran with 4-7 of 32 lanes and took a third of all issue slots).  No votes: the
	// re-joining is the hardware's own convergence barrier at the loop exit.  A lane that waits has nothing to catch up with -- its
	// pruning distance only changes at its own leaves -- so the visit pattern, and with it every result, is the same.
	constexpr int kDone = (int)0x80000000;        // never a leaf reference: ~kDone >> 1 is beyond the 2^30 triangles a scene can have
	for (;;) {
		while (cur >= 0) {
			const float4* n = S.nodes + 4 * (size_t)cur;
			float4 a, b, c; float2 r;
			load_node(n, a, b, c, r);
			float e0, e1;
			bool h0, h1;
			node_boxes<OCT>(rb, rb2, S.paired, a, b, c, PRUNE ? tcut : FLT_MAX, h0, h1, e0, e1);
			int r0 = f2i(r.x), r1 = f2i(r.y);
			if (h0 && h1) {
				bool swap = PRUNE && (e1 < e0);
				int farRef = swap ? r0 : r1;
				float farT = swap ? e0 : e1;
				if (sp < kBvhStack) { StackEnt e; e.ref = farRef; e.t = farT; stack[sp++] = e; }
				cur = swap ? r1 : r0;
			}
			else if (h0) cur = r0;
			else if (h1) cur = r1;
			else {
				cur = kDone;
				while (sp > 0) {
					StackEnt e = stack[--sp];
					if (!PRUNE || e.t <= tcut) { cur = e.ref; break; }
				}
			}
		}
		if (cur == kDone) break;
		{
			int ref = ~cur;
			int pos = ref >> 1, cnt = (ref & 1) + 1;
			for (int k = 0; k < cnt; k++) {
				TriV tri = load_tri(S.tris, pos + k);
				float t;
				if (moller_trumbore(tri, o, d, t)) {
					if ((t < bestT || (t == bestT && pos + k < bestPos)) && (!S.leafBox || ref_leaf_box_passes<OCT>(S, rb, pos + k, FLT_MAX))) {
						bestT = t; bestPos = pos + k; tcut = t * kPruneSlack;
					}
				}
			}
		}
		cur = kDone;
		while (sp > 0) {
			StackEnt e = stack[--sp];
			if (!PRUNE || e.t <= tcut) { cur = e.ref; break; }
		}
		if (cur == kDone) break;
	}
#else
	// (a per-step warp vote that re-joins the lanes costs 5-9 % here)
	{
		for (;;) {
			bool pop = true;
			if (cur >= 0) {
				const float4* n = S.nodes + 4 * (size_t)cur;
				float4 a, b, c; float2 r;
				load_node(n, a, b, c, r);
				float e0, e1;
				bool h0, h1;
				node_boxes<OCT>(rb, rb2, S.paired, a, b, c, PRUNE ? tcut : FLT_MAX, h0, h1, e0, e1);
				int r0 = f2i(r.x), r1 = f2i(r.y);
				if (h0 && h1) {
					bool swap = PRUNE && (e1 < e0);
					int nearRef = swap ? r1 : r0, farRef = swap ? r0 : r1;
					float farT = swap ? e0 : e1;
					if (sp < kBvhStack) { StackEnt e; e.ref = farRef; e.t = farT; stack[sp++] = e; }
					cur = nearRef; pop = false;
				}
				else if (h0) { cur = r0; pop = false; }
				else if (h1) { cur = r1; pop = false; }
			}
			else {
				int ref = ~cur;
				int pos = ref >> 1, cnt = (ref & 1) + 1;
				for (int k = 0; k < cnt; k++) {
					TriV tri = load_tri(S.tris, pos + k);
					float t;
					if (moller_trumbore(tri, o, d, t)) {
						// (the reference-leaf box is checked only for a hit that would be taken: two out of three tests miss anyway, and
						// the box, which contains the triangle, almost never fails)
						if ((t < bestT || (t == bestT && pos + k < bestPos)) && (!S.leafBox || ref_leaf_box_passes<OCT>(S, rb, pos + k, FLT_MAX))) {
							bestT = t; bestPos = pos + k; tcut = t * kPruneSlack;
						}
					}
				}
			}
			if (pop) {
				bool got = false;
				while (sp > 0) {
					StackEnt e = stack[--sp];
					if (!PRUNE || e.t <= tcut) { cur = e.ref; got = true; break; }
				}
				if (!got) break;
			}
		}
	}
#endif
}

template <bool PRUNE, bool WIDE = false>
RTO_DEV void bvh_closest(const BvhDev& S, V3 o, V3 d, float& bestT, int& bestPos) {
	bestT = kMissT; bestPos = -1;
	if (S.numTris <= 0) return;
	RayBox rb = make_raybox(o, d);
	float te;
	if (!slab_ref(rb, S.rootLo[0], S.rootLo[1], S.rootLo[2], S.rootHi[0], S.rootHi[1], S.rootHi[2], te)) return;
	if (!PRUNE) { bvh_closest_loop<false, kOctGeneric>(S, rb, o, d, bestT, bestPos); return; }      // verification path: one generic loop
	const int oct = ray_octant(S, rb);
	if (WIDE && PRUNE && bvh_closest_wide<WIDE>(S, rb, oct, o, d, bestT, bestPos)) return;
	switch (oct) {
	case 0: bvh_closest_loop<PRUNE, 0>(S, rb, o, d, bestT, bestPos); break;
	case 1: bvh_closest_loop<PRUNE, 1>(S, rb, o, d, bestT, bestPos); break;
	case 2: bvh_closest_loop<PRUNE, 2>(S, rb, o, d, bestT, bestPos); break;
	case 3: bvh_closest_loop<PRUNE, 3>(S, rb, o, d, bestT, bestPos); break;
	case 4: bvh_closest_loop<PRUNE, 4>(S, rb, o, d, bestT, bestPos); break;
	case 5: bvh_closest_loop<PRUNE, 5>(S, rb, o, d, bestT, bestPos); break;
	case 6: bvh_closest_loop<PRUNE, 6>(S, rb, o, d, bestT, bestPos); break;
	case 7: bvh_closest_loop<PRUNE, 7>(S, rb, o, d, bestT, bestPos); break;
	default: {
		// not admitted to the fused tests (zero / infinite 1/d, far origin): the reference's own arithmetic on the exact tree
		BvhDev E = S; E.nodes = S.exactNodes; E.rootRef = S.exactRoot; E.leafBox = S.exactLeafBox; E.grow = 0.0f; E.paired = S.exactPaired;
		bvh_closest_loop<PRUNE, kOctGeneric>(E, rb, o, d, bestT, bestPos);
		break; }
	}
}

// Shadow / any-hit: true iff some candidate of the reference's query passes Moller-Trumbore.
template <int OCT>
RTO_DEV bool bvh_any_loop(const BvhDev& S, const RayBox& rb, V3 o, V3 d) {
	int stackRef[kBvhStack];
	int sp = 0;
	int cur = S.rootRef;
	const RayBox2 rb2 = make_raybox2(rb);
#if RTO_BVH_ANY_WHILE_WHILE
	constexpr int kDone = (int)0x80000000;
	for (;;) {
		while (cur >= 0) {
			const float4* n = S.nodes + 4 * (size_t)cur;
			float4 a, b, c; float2 r;
			load_node(n, a, b, c, r);
			float e0, e1;
			bool h0, h1;
			node_boxes<OCT>(rb, rb2, S.paired, a, b, c, FLT_MAX, h0, h1, e0, e1);
			int r0 = f2i(r.x), r1 = f2i(r.y);
			if (h0 && h1) {
				bool swap = e1 < e0;
				if (sp < kBvhStack) stackRef[sp++] = swap ? r0 : r1;
				cur = swap ? r1 : r0;
			}
			else if (h0) cur = r0;
			else if (h1) cur = r1;
			else cur = sp > 0 ? stackRef[--sp] : kDone;
		}
		if (cur == kDone) return false;
		{
			int ref = ~cur;
			int pos = ref >> 1, cnt = (ref & 1) + 1;
			for (int k = 0; k < cnt; k++) {
				TriV tri = load_tri(S.tris, pos + k);
				float t;
				if (moller_trumbore(tri, o, d, t) && (!S.leafBox || ref_leaf_box_passes<OCT>(S, rb, pos + k, FLT_MAX))) return true;
			}
		}
		if (sp == 0) return false;
		cur = stackRef[--sp];
	}
#else
	while (true) {
		if (cur >= 0) {
			const float4* n = S.nodes + 4 * (size_t)cur;
			float4 a, b, c; float2 r;
			load_node(n, a, b, c, r);
			float e0, e1;
			bool h0, h1;
			node_boxes<OCT>(rb, rb2, S.paired, a, b, c, FLT_MAX, h0, h1, e0, e1);
			int r0 = f2i(r.x), r1 = f2i(r.y);
			if (h0 && h1) {
				bool swap = e1 < e0;
				if (sp < kBvhStack) stackRef[sp++] = swap ? r0 : r1;
				cur = swap ? r1 : r0;
				continue;
			}
			if (h0) { cur = r0; continue; }
			if (h1) { cur = r1; continue; }
		}
		else {
			int ref = ~cur;
			int pos = ref >> 1, cnt = (ref & 1) + 1;
			for (int k = 0; k < cnt; k++) {
				TriV tri = load_tri(S.tris, pos + k);
				float t;
				if (moller_trumbore(tri, o, d, t) && (!S.leafBox || ref_leaf_box_passes<OCT>(S, rb, pos + k, FLT_MAX))) return true;
			}
		}
		if (sp == 0) break;
		cur = stackRef[--sp];
	}
	return false;
#endif
}

template <bool WIDE = false>
RTO_DEV bool bvh_any(const BvhDev& S, V3 o, V3 d) {
	if (S.numTris <= 0) return false;
	RayBox rb = make_raybox(o, d);
	float te;
	if (!slab_ref(rb, S.rootLo[0], S.rootLo[1], S.rootLo[2], S.rootHi[0], S.rootHi[1], S.rootHi[2], te)) return false;
	const int oct = ray_octant(S, rb);
	if (WIDE) { bool hit; if (bvh_any_wide<WIDE>(S, rb, oct, o, d, hit)) return hit; }
	switch (oct) {
	case 0: return bvh_any_loop<0>(S, rb, o, d);
	case 1: return bvh_any_loop<1>(S, rb, o, d);
	case 2: return bvh_any_loop<2>(S, rb, o, d);
	case 3: return bvh_any_loop<3>(S, rb, o, d);
	case 4: return bvh_any_loop<4>(S, rb, o, d);
	case 5: return bvh_any_loop<5>(S, rb, o, d);
	case 6: return bvh_any_loop<6>(S, rb, o, d);
	case 7: return bvh_any_loop<7>(S, rb, o, d);
	default: {
		BvhDev E = S; E.nodes = S.exactNodes; E.rootRef = S.exactRoot; E.leafBox = S.exactLeafBox; E.grow = 0.0f; E.paired = S.exactPaired;
		return bvh_any_loop<kOctGeneric>(E, rb, o, d); }
	}
}

// Reference-order replay (left before right, nothing pruned): emits candidate positions in BVH::query order and
// counts intersectAABB calls the reference would make (1 for the root + 2 per internal node entered).
template <typename Emit>
RTO_DEV void bvh_replay(const BvhDev& S, V3 o, V3 d, unsigned long long& boxTests, Emit emit) {
	if (S.numTris < 0) return;
	boxTests += 1;
	RayBox rb = make_raybox(o, d);
	float te;
	if (!slab_ref(rb, S.rootLo[0], S.rootLo[1], S.rootLo[2], S.rootHi[0], S.rootHi[1], S.rootHi[2], te)) return;
	if (S.numTris == 0) return;
	int stackRef[kBvhStack];
	int sp = 0;
	int cur = S.rootRef;
	while (true) {
		if (cur >= 0) {
			const float4* n = S.nodes + 4 * (size_t)cur;
			float4 a = RTO_LDG(n), b = RTO_LDG(n + 1), c = RTO_LDG(n + 2), r = RTO_LDG(n + 3);
			float e0, e1;
			boxTests += 2;
			bool h0 = slab_ref(rb, a.x, a.y, a.z, a.w, b.x, b.y, e0);
			bool h1 = slab_ref(rb, b.z, b.w, c.x, c.y, c.z, c.w, e1);
			int r0 = f2i(r.x), r1 = f2i(r.y);
			if (h0 && h1) { if (sp < kBvhStack) stackRef[sp++] = r1; cur = r0; continue; }
			if (h0) { cur = r0; continue; }
			if (h1) { cur = r1; continue; }
		}
		else {
			int ref = ~cur;
			int pos = ref >> 1, cnt = (ref & 1) + 1;
			for (int k = 0; k < cnt; k++) emit(pos + k);
		}
		if (sp == 0) break;
		cur = stackRef[--sp];
	}
}

// ------------------------------------------------------------------------------------------------
// Octree helpers
// ------------------------------------------------------------------------------------------------
// order of octants for octreeRaySkip: ascending popcount(octant ^ dirMask), ascending octant within a class
// (VolumeRaycastRenderer.cpp:122-134).  nibble j of kSkipOrder[m] = j-th octant; nibble k of kSkipRank[m] = position of octant k.
__host__ __device__ constexpr int popc3(int v) { return (v & 1) + ((v >> 1) & 1) + ((v >> 2) & 1); }
__host__ __device__ constexpr uint32_t skip_order(int m) {
	uint32_t r = 0; int j = 0;
	for (int dist = 0; dist <= 3; dist++)
		for (int o = 0; o < 8; o++)
			if (popc3(o ^ m) == dist) { r |= (uint32_t)o << (4 * j); j++; }
	return r;
}
__host__ __device__ constexpr uint32_t skip_rank(int m) {
	uint32_t ord = skip_order(m), r = 0;
	for (int j = 0; j < 8; j++) r |= (uint32_t)j << (4 * ((ord >> (4 * j)) & 7u));
	return r;
}
// The mask of the children a mode-A walk may skip (empty leaves), re-ordered from octant numbering into VISIT order: bit j of the result
// = bit skip_order(m)[j] of the mask.  Done with shifts and logic operations this costs 14 instructions per node on the ALU pipe, the
// pipe that bounds the octree kernels; here it is one byte load from a 2 KB table (the LSU pipe idles in these kernels).
#ifndef RTO_OCTA_SKIP_LUT
#define RTO_OCTA_SKIP_LUT 1
#endif
#ifndef RTO_OCT_FADD_COORDS
#define RTO_OCT_FADD_COORDS 1
#endif
struct SkipPermTable { unsigned char v[8][256]; };
constexpr SkipPermTable make_skip_perm() {
	SkipPermTable t{};
	for (int m = 0; m < 8; m++) {
		const uint32_t ord = skip_order(m);
		for (unsigned k = 0; k < 256; k++) {
			unsigned r = 0;
			for (int j = 0; j < 8; j++) r |= ((k >> ((ord >> (4 * j)) & 7u)) & 1u) << j;
			t.v[m][k] = (unsigned char)r;
		}
	}
	return t;
}
constexpr SkipPermTable kSkipPermInit = make_skip_perm();
#if defined(__CUDACC__)
static __device__ const SkipPermTable g_skipPermDev = kSkipPermInit;
#endif
static const SkipPermTable g_skipPermHost = kSkipPermInit;
RTO_DEV unsigned skip_perm(int m, unsigned mask) {
#if defined(__CUDA_ARCH__)
	return (unsigned)__ldg(&g_skipPermDev.v[m][mask]);
#else
	return g_skipPermHost.v[m][mask];
#endif
}
RTO_DEV void skip_tables(int dirMask, uint32_t& order, uint32_t& rank) {
	constexpr uint32_t ord[8] = { skip_order(0), skip_order(1), skip_order(2), skip_order(3), skip_order(4), skip_order(5), skip_order(6), skip_order(7) };
	constexpr uint32_t rnk[8] = { skip_rank(0), skip_rank(1), skip_rank(2), skip_rank(3), skip_rank(4), skip_rank(5), skip_rank(6), skip_rank(7) };
	order = ord[dirMask]; rank = rnk[dirMask];
}

struct OctBox { V3 mn, mx; };
// world box of a node: nodeMin = gridMin + vec3(x,y,z)*voxelSize; nodeMax = nodeMin + vec3(size)*voxelSize
// (RayTracerBVH.cpp:262-263; identical operations in VolumeRaycastRenderer.cpp:70-77)
RTO_DEV OctBox oct_box(const OctDev& S, int x, int y, int z, int size) {
	OctBox b;
	b.mn = mk3(S.gmin[0] + float(x) * S.voxel, S.gmin[1] + float(y) * S.voxel, S.gmin[2] + float(z) * S.voxel);
	float w = float(size) * S.voxel;
	b.mx = mk3(b.mn.x + w, b.mn.y + w, b.mn.z + w);
	return b;
}

struct OctHit { float t; int id; V3 normal; unsigned visits; };

// ---- mode B: GLSL intersectAABB (RayTracerBVH.cpp:226-236) -----------------------------------------------
RTO_DEV bool glsl_box(const OctBox& b, V3 o, V3 inv, float& tNear, float& tFar) {
	V3 t1 = (b.mn - o) * inv, t2 = (b.mx - o) * inv;
	V3 tmn = min3(t1, t2), tmx = max3(t1, t2);
	tNear = maxf(maxf(tmn.x, tmn.y), tmn.z);
	tFar = minf(minf(tmx.x, tmx.y), tmx.z);
	return (tNear <= tFar && tFar > 0.0f);
}

RTO_DEV V3 box_normal(const OctBox& b, V3 o, V3 d, float t) {     // RayTracerBVH.cpp:279-282
	V3 center = 0.5f * (b.mn + b.mx);
	V3 p = o + d * t;
	return normalize3(p - center);
}

// compact layout: stackless walk; sibling order 7..0 == the pop order of the GLSL stack
RTO_DEV OctHit octB_compact(const OctDev& S, V3 o, V3 d) {
	OctHit h; h.t = kMissT; h.id = -1; h.normal = mk3(0.0f, 0.0f, 0.0f); h.visits = 0;
	V3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
	const float closestT = kMissT;
	int node = 0, x = 0, y = 0, z = 0, size = S.rootSize;
	int steps = 0;
	while (steps < 512) {
		steps++;
		OctBox b = oct_box(S, x, y, z, size);
		float tNear, tFar;
		if (glsl_box(b, o, inv, tNear, tFar) && !(tNear >= closestT)) {
			uint32_t dsc = RTO_LDG(S.desc + node);
			if (dsc & kOctLeaf) {
				if (dsc & kOctSolid) {
					float tHit = maxf(0.0f, tNear);
					if (tHit < closestT && tHit <= tFar) { h.t = tHit; h.id = node; h.normal = box_normal(b, o, d, tHit); break; }
				}
			}
			else {       // children pushed 0..7, so 7 is visited first
				size >>= 1; x += size; y += size; z += size;
				node = (int)dsc + 7;
				continue;
			}
		}
		// next sibling (k-1) or climb
		bool done = false;
		while (true) {
			if (node == 0) { done = true; break; }
			int k = (node - 1) & 7;
			if (k > 0) {
				int nk = k - 1;
				x = (x & ~size) | ((nk & 1) ? size : 0); y = (y & ~size) | ((nk & 2) ? size : 0); z = (z & ~size) | ((nk & 4) ? size : 0);
				node -= 1;
				break;
			}
			node = RTO_LDG(S.up + ((node - 1) >> 3));
			x &= ~size; y &= ~size; z &= ~size; size <<= 1;
		}
		if (done) break;
	}
	h.visits = (unsigned)steps;
	return h;
}

// general layout: the GLSL loop verbatim over 16-int nodes
RTO_DEV OctHit octB_general(const OctDev& S, V3 o, V3 d) {
	OctHit h; h.t = kMissT; h.id = -1; h.normal = mk3(0.0f, 0.0f, 0.0f); h.visits = 0;
	V3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
	const float closestT = kMissT;
	int stack[128]; int sp = 0; stack[sp++] = 0;
	int steps = 0;
	while (sp > 0 && steps < 512) {
		int idx = stack[--sp];
		if (idx < 0 || idx >= S.numNodes) continue;       // (the reference does not guard idx >= numNodes; such arrays are rejected at upload)
		steps++;
		const int4* n = S.nodes16 + 4 * (size_t)idx;
		int4 a = RTO_LDG(n), f = RTO_LDG(n + 1);
		OctBox b = oct_box(S, a.x, a.y, a.z, a.w);
		float tNear, tFar;
		if (!glsl_box(b, o, inv, tNear, tFar)) continue;
		if (tNear >= closestT) continue;
		if (f.z == 1 || f.x == 1) {                       // isUniform == 1, or isLeaf == 1 (RayTracerBVH.cpp:271,291)
			if (f.y == 1) {
				float tHit = maxf(0.0f, tNear);
				if (tHit < closestT && tHit <= tFar) { h.t = tHit; h.id = idx; h.normal = box_normal(b, o, d, tHit); break; }
			}
			continue;
		}
		int4 c0 = RTO_LDG(n + 2), c1 = RTO_LDG(n + 3);
		int ch[8] = { f.w, c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z };
#pragma unroll
		for (int i = 0; i < 8; i++) if (ch[i] >= 0 && sp < 128) stack[sp++] = ch[i];
	}
	h.visits = (unsigned)steps;
	return h;
}

// ---- mode A: octreeRaySkip (VolumeRaycastRenderer.cpp:50-155) ---------------------------------------------
struct SkipRay { V3 o, inv; uint32_t order, rank; };
RTO_DEV SkipRay make_skipray(V3 o, V3 d) {
	SkipRay r; r.o = o;
	r.inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
	const float smallValue = 1e-10f;                                       // :84-87
	if (fabsf(d.x) < smallValue) r.inv.x = d.x >= 0 ? 1e10f : -1e10f;
	if (fabsf(d.y) < smallValue) r.inv.y = d.y >= 0 ? 1e10f : -1e10f;
	if (fabsf(d.z) < smallValue) r.inv.z = d.z >= 0 ? 1e10f : -1e10f;
	int dirMask = ((d.x > 0) ? 1 : 0) | ((d.y > 0) ? 2 : 0) | ((d.z > 0) ? 4 : 0);   // :114-116
	skip_tables(dirMask, r.order, r.rank);
	return r;
}
RTO_DEV bool skip_box(const OctBox& b, const SkipRay& r, float tMin, float tMax, float& enterT, float& exitT) {
	V3 t1 = (b.mn - r.o) * r.inv, t2 = (b.mx - r.o) * r.inv;
	V3 tN = min3(t1, t2), tF = max3(t1, t2);
	enterT = maxf(maxf(tN.x, tN.y), maxf(tN.z, tMin));                     // :96
	exitT = minf(minf(tF.x, tF.y), minf(tF.z, tMax));                      // :97
	return !(enterT > exitT);
}

// The same test with the octant of 1/d known (OCT < 8, voxel > 0, no NaN anywhere): the near/far plane of each axis is picked
// without a comparison and the clamps use FMNMX; values equal skip_box's except possibly the sign of a zero, which no
// comparison sees (a zero that would be RETURNED re-runs the ray through the per-node path, see octA_fast_loop).
template <int OCT>
RTO_DEV bool skip_box_oct(const OctBox& b, const SkipRay& r, float tMin, float tMax, float& enterT, float& exitT) {
	if (OCT >= 8) return skip_box(b, r, tMin, tMax, enterT, exitT);
	V3 t1 = (b.mn - r.o) * r.inv, t2 = (b.mx - r.o) * r.inv;
	V3 tN = mk3((OCT & 1) ? t2.x : t1.x, (OCT & 2) ? t2.y : t1.y, (OCT & 4) ? t2.z : t1.z);
	V3 tF = mk3((OCT & 1) ? t1.x : t2.x, (OCT & 2) ? t1.y : t2.y, (OCT & 4) ? t1.z : t2.z);
	enterT = fmaxf(fmaxf(tN.x, tN.y), fmaxf(tN.z, tMin));
	exitT = fminf(fminf(tF.x, tF.y), fminf(tF.z, tMax));
	return !(enterT > exitT);
}

RTO_DEV OctHit octA_compact(const OctDev& S, V3 o, V3 d, float tMin, float tMax) {
	OctHit h; h.t = kMissT; h.id = -1; h.normal = mk3(0.0f, 0.0f, 0.0f); h.visits = 0;
	SkipRay r = make_skipray(o, d);
	float minS[kMaxOctDepth], maxS[kMaxOctDepth];
	int depth = 0;
	float curMin = tMin, curMax = tMax;
	int node = 0, x = 0, y = 0, z = 0, size = S.rootSize;
	unsigned visits = 0;
	while (true) {
		visits++;
		OctBox b = oct_box(S, x, y, z, size);
		float enterT, exitT;
		if (skip_box(b, r, curMin, curMax, enterT, exitT)) {
			uint32_t dsc = RTO_LDG(S.desc + node);
			if (dsc & kOctLeaf) {
				// a solid leaf returns enterT; the callers keep it only if it is < 1e30f (:143-149)
				if ((dsc & kOctSolid) && enterT < 1e30f) { h.t = enterT; h.id = node; h.normal = box_normal(b, o, d, enterT); break; }
			}
			else {
				minS[depth] = curMin; maxS[depth] = curMax; depth++;
				curMin = enterT; curMax = exitT;
				int k = (int)(r.order & 7u);
				size >>= 1;
				x += (k & 1) ? size : 0; y += (k & 2) ? size : 0; z += (k & 4) ? size : 0;
				node = (int)dsc + k;
				continue;
			}
		}
		bool done = false;
		while (true) {
			if (node == 0) { done = true; break; }
			int k = (node - 1) & 7;
			int j = (int)((r.rank >> (4 * k)) & 7u);
			if (j < 7) {
				int nk = (int)((r.order >> (4 * (j + 1))) & 7u);
				x = (x & ~size) | ((nk & 1) ? size : 0); y = (y & ~size) | ((nk & 2) ? size : 0); z = (z & ~size) | ((nk & 4) ? size : 0);
				node = node - k + nk;
				break;
			}
			node = RTO_LDG(S.up + ((node - 1) >> 3));
			x &= ~size; y &= ~size; z &= ~size; size <<= 1;
			depth--; curMin = minS[depth]; curMax = maxS[depth];
		}
		if (done) break;
	}
	h.visits = visits;
	return h;
}

RTO_DEV OctHit octA_general(const OctDev& S, V3 o, V3 d, float tMin, float tMax) {
	OctHit h; h.t = kMissT; h.id = -1; h.normal = mk3(0.0f, 0.0f, 0.0f); h.visits = 0;
	SkipRay r = make_skipray(o, d);
	// explicit recursion frames: node being expanded, next order position, the (tMin, tMax) its children receive
	int   nodeS[kMaxOctDepth]; int posS[kMaxOctDepth]; float minS[kMaxOctDepth], maxS[kMaxOctDepth];
	int depth = 0;
	unsigned visits = 0;
	int cur = 0; float curMin = tMin, curMax = tMax;
	while (true) {
		if (cur >= 0 && cur < S.numNodes) {
			// a walk over a tree visits no node twice; arrays whose child graph is not a tree are refused for this mode before any launch
			// (rto_device.cu check_mode), the budget keeps even a corrupted device array from walking forever
			if (++visits > (unsigned)S.numNodes) break;
			const int4* n = S.nodes16 + 4 * (size_t)cur;
			int4 a = RTO_LDG(n), f = RTO_LDG(n + 1);
			OctBox b = oct_box(S, a.x, a.y, a.z, a.w);
			float enterT, exitT;
			if (skip_box(b, r, curMin, curMax, enterT, exitT)) {
				if (f.x != 0) {                              // node->isLeaf (:105)
					if (f.y != 0 && enterT < 1e30f) { h.t = enterT; h.id = cur; h.normal = box_normal(b, o, d, enterT); break; }
				}
				else if (depth < kMaxOctDepth) {
					nodeS[depth] = cur; posS[depth] = 0; minS[depth] = enterT; maxS[depth] = exitT; depth++;
				}
			}
		}
		// take the next child of the innermost open frame
		bool found = false;
		while (depth > 0) {
			int f = depth - 1;
			if (posS[f] < 8) {
				int k = (int)((r.order >> (4 * posS[f])) & 7u);
				posS[f]++;
				const int* ch = reinterpret_cast<const int*>(S.nodes16 + 4 * (size_t)nodeS[f]) + 7;
				int c = RTO_LDG(ch + k);
				if (c < 0) continue;                         // null child: skipped without a call (:136-137)
				cur = c; curMin = minS[f]; curMax = maxS[f];
				found = true;
				break;
			}
			depth--;
		}
		if (!found) break;
	}
	h.visits = visits;
	return h;
}


// ------------------------------------------------------------------------------------------------
// Octree fast paths (compact layout): all 8 children of an internal node are classified at once.
//
// The reference tests every child box separately; its arithmetic per child is (plane - o) * inv on planes computed as
// gridMin + float(coord) * voxel (+ float(size) * voxel).  Here the 12 plane distances of the 8 children (2 per axis for each
// half) are computed once per internal node with exactly those operations and combined per child with min/max, which yields
// the same tNear/tFar values as the per-child code as long as no NaN occurs -- guaranteed when 1/d is finite and non-zero.
// (FMNMX and the reference's select forms can differ only in the sign of a zero, which no comparison sees; the one place a
// zero could surface -- mode A returning enterT == 0 -- re-runs the ray through the exact per-node path.)  One 16-byte record
// per internal node carries first-child index, child leaf/solid masks and the links needed for a stackless walk; the masks
// of children still to visit live in registers, 8 bits per level.  Rays with an infinite/zero reciprocal component, trees
// deeper than 16 levels and the visit counters of rto_render_stats use the per-node paths above.
// ------------------------------------------------------------------------------------------------
struct ChildPlanes { float n0[3], f0[3], n1[3], f1[3]; };     // per axis: near/far distances of the low and the high half

// OCT < 8: bit a set <=> 1/d negative on axis a (and voxel > 0), so the smaller of the two plane distances is known without a
// min/max: the products are monotonic in the plane position.  OCT == 8: generic min/max.
template <int OCT>
RTO_DEV ChildPlanes oct_child_planes(const OctDev& S, V3 o, V3 inv, int x, int y, int z, int h) {
	ChildPlanes P;
	const float fh = float(h);
	const float w = fh * S.voxel;
	const int c[3] = { x, y, z };
	const float oo[3] = { o.x, o.y, o.z }, ii[3] = { inv.x, inv.y, inv.z };
#pragma unroll
	for (int a = 0; a < 3; a++) {
		// float(c + h) == float(c) + float(h): both are integers below 2^24 on the fast paths (rootSize <= 65536), so the sum is exact --
		// an addition on the FMA pipe instead of a second conversion on the ALU pipe, which is the one that bounds these kernels
#if RTO_OCT_FADD_COORDS
		const float fc = float(c[a]);
		float lo0 = S.gmin[a] + fc * S.voxel, lo1 = S.gmin[a] + (fc + fh) * S.voxel;
#else
		float lo0 = S.gmin[a] + float(c[a]) * S.voxel, lo1 = S.gmin[a] + float(c[a] + h) * S.voxel;
#endif
		float t1 = (lo0 - oo[a]) * ii[a], t2 = ((lo0 + w) - oo[a]) * ii[a];
		float u1 = (lo1 - oo[a]) * ii[a], u2 = ((lo1 + w) - oo[a]) * ii[a];
		if (OCT < 8) {
			const bool neg = (OCT >> a) & 1;
			P.n0[a] = neg ? t2 : t1; P.f0[a] = neg ? t1 : t2;
			P.n1[a] = neg ? u2 : u1; P.f1[a] = neg ? u1 : u2;
		}
		else {
			P.n0[a] = fminf(t1, t2); P.f0[a] = fmaxf(t1, t2);
			P.n1[a] = fminf(u1, u2); P.f1[a] = fmaxf(u1, u2);
		}
	}
	return P;
}

RTO_DEV bool inv_is_regular(V3 inv) {
	float ax = fabsf(inv.x), ay = fabsf(inv.y), az = fabsf(inv.z);
	return ax > 0.0f && ax <= FLT_MAX && ay > 0.0f && ay <= FLT_MAX && az > 0.0f && az <= FLT_MAX;
}

RTO_DEV void oct_child_coords(int k, int h, int& x, int& y, int& z) {
	x += (k & 1) ? h : 0; y += (k & 2) ? h : 0; z += (k & 4) ? h : 0;
}

// mode A: clamps of the ancestors (dynamically indexed: local memory)
struct OctClamps { float mn[16], mx[16]; unsigned char mask[16]; };     // + M of the ancestors

// ---- mode B ------------------------------------------------------------------------------------------------------
// One loop, state in plain locals (an init/step form with a warp vote between steps, tried in round 1, costs this walk 10-25 %).
template <int OCT>
RTO_DEV OctHit octB_fast_loop(const OctDev& S, V3 o, V3 d, V3 inv) {
	OctHit hit; hit.t = kMissT; hit.id = -1; hit.normal = mk3(0.0f, 0.0f, 0.0f); hit.visits = 0;
	int x = 0, y = 0, z = 0, size = S.rootSize;
	int steps = 1;                                             // the root is popped first
	{
		OctBox b = oct_box(S, 0, 0, 0, size);
		float tNear, tFar;
		hit.visits = 1;
		if (!glsl_box(b, o, inv, tNear, tFar) || tNear >= kMissT) return hit;
		uint32_t dsc = RTO_LDG(S.desc);
		if (dsc & kOctLeaf) {
			if (dsc & kOctSolid) { float tHit = maxf(0.0f, tNear); if (tHit < kMissT && tHit <= tFar) { hit.t = tHit; hit.id = 0; hit.normal = box_normal(b, o, d, tHit); } }
			return hit;
		}
	}
	unsigned char maskS[16];                                   // remaining-children masks of the ancestors (local memory: the LSU pipe idles,
	                                                           // the ALU pipe that would shift them in and out of registers is the bound)
	int level = 0, rank = 0, pos = 8;
	int4 e = RTO_LDG(S.inner);
	bool entering = true;
	unsigned M = 0;
	while (true) {
		int h = size >> 1;
		unsigned leafMask = (unsigned)e.w & 0xffu;
		if (entering) {
			const unsigned solidMask = ((unsigned)e.w >> 8) & 0xffu;
			ChildPlanes P = oct_child_planes<OCT>(S, o, inv, x, y, z, h);
			// "tNear < closestT" (closestT stays 1e30 until the hit that ends the walk) is folded into the far distance of one axis:
			// tn <= tf && tf > 0 && tn < 1e30  <=>  tn <= min(tf, kBelowMissT) && min(tf, kBelowMissT) > 0
			P.f0[2] = fminf(P.f0[2], kBelowMissT); P.f1[2] = fminf(P.f1[2], kBelowMissT);
			// and "tFar > 0" into the near distance of the same axis: tf > 0 <=> tf >= the smallest positive float (no flush-to-zero
			// in this build), so tn <= tf && tf > 0  <=>  max(tn, kMinPositive) <= tf
			P.n0[2] = fmaxf(P.n0[2], kMinPositive); P.n1[2] = fmaxf(P.n1[2], kMinPositive);
			unsigned hits = 0;
#pragma unroll
			for (int k = 0; k < 8; k++) {
				float tn = fmax3f((k & 1) ? P.n1[0] : P.n0[0], (k & 2) ? P.n1[1] : P.n0[1], (k & 4) ? P.n1[2] : P.n0[2]);
				float tf = fmin3f((k & 1) ? P.f1[0] : P.f0[0], (k & 2) ? P.f1[1] : P.f0[1], (k & 4) ? P.f1[2] : P.f0[2]);
				hits |= (tn <= tf) ? (1u << k) : 0u;
			}
			M = hits & ~(leafMask & ~solidMask);               // children that can do more than burn a step: solid leaves and internal nodes
			pos = 8;
		}
		unsigned below = M & ((1u << pos) - 1u);
		// climbs happen inside the iteration: the lanes of a warp run this loop in lockstep, and a lane that hands back through k levels
		// would otherwise pay k iterations, each as long as the classification some other lane runs in it (measured, 16 x 1080p: 512^3
		// city 6.44 -> 5.43 ms, DT 3.96 -> 3.46)
		bool out = false;
		while (below == 0u) {                                  // the remaining `pos` children are popped, tested and dropped
			steps += pos;
			if (steps >= 512 || level == 0) { out = true; break; }
			pos = ((unsigned)e.w >> 16) & 7u;
			rank = e.z;
			x &= ~size; y &= ~size; z &= ~size; size <<= 1;
			level--;
			M = maskS[level];
			e = RTO_LDG(S.inner + rank);
			below = M & ((1u << pos) - 1u);
		}
		if (out) break;
		h = size >> 1;
		leafMask = (unsigned)e.w & 0xffu;
		const int j = 31 - clz32(below);
		steps += pos - 1 - j;
		if (steps >= 512) break;
		steps++;
		pos = j;
		int cx = x, cy = y, cz = z;
		oct_child_coords(j, h, cx, cy, cz);
		if ((leafMask >> j) & 1u) {                            // solid leaf whose box the ray hits: the reference's exact per-node values
			OctBox b = oct_box(S, cx, cy, cz, h);
			float tNear, tFar;
			if (glsl_box(b, o, inv, tNear, tFar) && !(tNear >= kMissT)) {
				float tHit = maxf(0.0f, tNear);
				if (tHit < kMissT && tHit <= tFar) { hit.t = tHit; hit.id = e.x + j; hit.normal = box_normal(b, o, d, tHit); break; }
			}
			entering = false;
			continue;
		}
		maskS[level] = (unsigned char)M;
		level++;
		rank = e.y + popc32(~leafMask & ((1u << j) - 1u) & 0xffu);
		x = cx; y = cy; z = cz; size = h;
		e = RTO_LDG(S.inner + rank);
		entering = true;
	}
	hit.visits = (unsigned)(steps > 512 ? 512 : steps);
	return hit;
}

// dispatch on the octant of 1/d (one instantiation per sign pattern; a warp of primary rays almost always shares one)
#define RTO_OCT_DISPATCH(oct, CALL) \
	switch (oct) { \
	case 0: CALL(0); break; case 1: CALL(1); break; case 2: CALL(2); break; case 3: CALL(3); break; \
	case 4: CALL(4); break; case 5: CALL(5); break; case 6: CALL(6); break; case 7: CALL(7); break; \
	default: CALL(8); break; }

RTO_DEV int inv_octant(V3 inv, float voxel) {
	if (!(voxel > 0.0f)) return 8;
	return ((inv.x < 0) ? 1 : 0) | ((inv.y < 0) ? 2 : 0) | ((inv.z < 0) ? 4 : 0);
}
RTO_DEV OctHit octB_fast(const OctDev& S, V3 o, V3 d) {
	V3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
	if (!inv_is_regular(inv) || S.rootSize > 65536) return octB_compact(S, o, d);
	const int oct = inv_octant(inv, S.voxel);
	OctHit hit;
#define RTO_CALL_B(K) hit = octB_fast_loop<K>(S, o, d, inv)
	RTO_OCT_DISPATCH(oct, RTO_CALL_B)
#undef RTO_CALL_B
	return hit;
}

// ---- mode A ------------------------------------------------------------------------------------------------------
// One loop with its state in plain locals.  Every trip of the outer loop starts with the classification of the 8 children of the node
// just entered -- the expensive part, which all lanes of the warp that are still walking run together -- and the inner loop does the
// cheap rest per lane: hand back to the parent while a node's children are used up, test a solid leaf, until the lane has a node to
// descend into (or is done).  The lanes re-join at the exit of the inner loop, the hardware's own convergence point; no votes.
// History (B200, 16 x 1080p per launch, 512^3 city / DT): init + step form with one warp vote per step, climbs as steps of their own
// 8.38 / 5.35 ms (round 1: the vote had made it 1.5-2.4x faster than lanes drifting apart); climbs folded into the step 6.44 / 4.36;
// without the vote 6.14 / 4.16; this loop 6.06 / 4.12, at 10 resident blocks per SM 5.75 / 3.97.  Same operations per ray in the same
// order throughout: same bits.
template <int OCT>
RTO_DEV OctHit octA_fast_loop(const OctDev& S, V3 o, V3 d, float tMin, float tMax, const SkipRay& r) {
	OctHit hit; hit.t = kMissT; hit.id = -1; hit.normal = mk3(0.0f, 0.0f, 0.0f); hit.visits = 0;
	const uint32_t order = (OCT < 8) ? skip_order((~OCT) & 7) : r.order;
	const uint32_t rnk = (OCT < 8) ? skip_rank((~OCT) & 7) : r.rank;
	int x = 0, y = 0, z = 0, size = S.rootSize;
	float curMin, curMax;
	{
		OctBox b = oct_box(S, 0, 0, 0, size);
		if (!skip_box(b, r, tMin, tMax, curMin, curMax)) return hit;
		uint32_t dsc = RTO_LDG(S.desc);
		if (dsc & kOctLeaf) {
			if ((dsc & kOctSolid) && curMin < 1e30f) { hit.t = curMin; hit.id = 0; hit.normal = box_normal(b, o, d, curMin); }
			return hit;
		}
	}
	OctClamps cs;
	int level = 0, rank = 0, steps = 0;
	int4 e = RTO_LDG(S.inner);
	for (;;) {
		int h = size >> 1;
		unsigned leafMask = (unsigned)e.w & 0xffu;
		unsigned M;
		{
			const unsigned solidMask = ((unsigned)e.w >> 8) & 0xffu;
			ChildPlanes P = oct_child_planes<OCT>(S, r.o, r.inv, x, y, z, h);
			const unsigned skipMask = leafMask & ~solidMask;       // empty leaves return 1e30f whatever their box test says
			P.n0[2] = fmaxf(P.n0[2], curMin); P.n1[2] = fmaxf(P.n1[2], curMin);
			P.f0[2] = fminf(P.f0[2], curMax); P.f1[2] = fminf(P.f1[2], curMax);
			unsigned Mo = 0, skipO = 0;                            // bit j = j-th child in VISIT order
			constexpr bool kLut = RTO_OCTA_SKIP_LUT && OCT < 8;
			if (kLut) skipO = skip_perm((~OCT) & 7, skipMask);
#pragma unroll
			for (int j = 0; j < 8; j++) {
				const int k = (int)((order >> (4 * j)) & 7u);
				float tn = fmax3f((k & 1) ? P.n1[0] : P.n0[0], (k & 2) ? P.n1[1] : P.n0[1], (k & 4) ? P.n1[2] : P.n0[2]);
				float tf = fmin3f((k & 1) ? P.f1[0] : P.f0[0], (k & 2) ? P.f1[1] : P.f0[1], (k & 4) ? P.f1[2] : P.f0[2]);
				Mo |= !(tn > tf) ? (1u << j) : 0u;
				if (!kLut) skipO |= ((skipMask >> k) & 1u) << j;
			}
			M = Mo & ~skipO;
		}
		unsigned rem = M;                                          // order positions still to visit in the current node
		bool done = false;
		for (;;) {
			while (rem == 0u) {                                    // children used up: back to the parent, after the child we came from
				if (level == 0) { done = true; break; }
				const int pos = (int)((rnk >> (4 * (((unsigned)e.w >> 16) & 7u))) & 7u);
				rank = e.z;
				x &= ~size; y &= ~size; z &= ~size; size <<= 1;
				level--;
				M = cs.mask[level];
				curMin = cs.mn[level]; curMax = cs.mx[level];
				e = RTO_LDG(S.inner + rank);
				rem = M & ~((2u << pos) - 1u);
			}
			if (done) break;
			h = size >> 1;
			leafMask = (unsigned)e.w & 0xffu;
			const int jO = ffs32(rem) - 1;
			const int k = (int)((order >> (4 * jO)) & 7u);
			rem &= rem - 1u;                                       // consumed
			int cx = x, cy = y, cz = z;
			oct_child_coords(k, h, cx, cy, cz);
			OctBox b = oct_box(S, cx, cy, cz, h);
			float enterT, exitT;
			const bool ok = skip_box_oct<OCT>(b, r, curMin, curMax, enterT, exitT);
			if ((leafMask >> k) & 1u) {                            // solid leaf
				if (ok && enterT < 1e30f) {
					if (enterT == 0.0f) return octA_compact(S, o, d, tMin, tMax);      // zero of either sign: take the reference's select forms
					hit.t = enterT; hit.id = e.x + k; hit.normal = box_normal(b, o, d, enterT);
					done = true; break;
				}
				continue;
			}
			if (!ok) continue;                                     // (cannot happen: same values as the batch test)
			cs.mask[level] = (unsigned char)M;
			cs.mn[level] = curMin; cs.mx[level] = curMax;
			level++;
			curMin = enterT; curMax = exitT;
			rank = e.y + popc32(~leafMask & ((1u << k) - 1u) & 0xffu);
			x = cx; y = cy; z = cz; size = h;
			e = RTO_LDG(S.inner + rank);
			// (a finite walk enters < nodes internal nodes; the bound only keeps a corrupted array from hanging the GPU)
			if (++steps >= (1 << 24)) done = true;
			break;
		}
		if (done) break;
	}
	return hit;
}

RTO_DEV bool octA_is_fast(const OctDev& S, const SkipRay& r, float tMin, float tMax) {
	return inv_is_regular(r.inv) && S.rootSize <= 65536 && (tMin == tMin) && (tMax == tMax);
}
RTO_DEV int octA_octant(const OctDev& S, const SkipRay& r, V3 d) {
	if (d.x == 0.0f || d.y == 0.0f || d.z == 0.0f) return 8;   // dirMask (d > 0) and the sign of the clamped reciprocal disagree
	return inv_octant(r.inv, S.voxel);
}

RTO_DEV OctHit octA_fast(const OctDev& S, V3 o, V3 d, float tMin, float tMax) {
	SkipRay r = make_skipray(o, d);
	if (!octA_is_fast(S, r, tMin, tMax)) return octA_compact(S, o, d, tMin, tMax);
	const int oct = octA_octant(S, r, d);
	OctHit hit;
#define RTO_CALL_A(K) hit = octA_fast_loop<K>(S, o, d, tMin, tMax, r)
	RTO_OCT_DISPATCH(oct, RTO_CALL_A)
#undef RTO_CALL_A
	return hit;
}

// COUNT: the caller needs OctHit::visits (rto_render_stats) -> per-node paths, which count what the reference visits.
template <bool COUNT>
RTO_DEV OctHit oct_trace(const OctDev& S, int mode, V3 o, V3 d, float tMin, float tMax) {
	if (mode == RTO_MODE_OCTREE_SKIP) {
		if (!S.compact) return octA_general(S, o, d, tMin, tMax);
		return COUNT ? octA_compact(S, o, d, tMin, tMax) : octA_fast(S, o, d, tMin, tMax);
	}
	if (!S.compact) return octB_general(S, o, d);
	return COUNT ? octB_compact(S, o, d) : octB_fast(S, o, d);
}

// ------------------------------------------------------------------------------------------------
// Pixel mapping: one warp = one 4x8 pixel tile (4 wide, 8 tall), one 128-thread block = 16x8 pixels, blockIdx.z = camera
// ------------------------------------------------------------------------------------------------
struct RenderArgs {
	RtoCamera cam0;               // used when cams == nullptr
	const RtoCamera* cams;        // device array for batches
	int y0, y1;
	float4* rgba; int* hitId; float* t;
	float shadowBias;
	unsigned flags;
	// compact hit codes (BVH scenes; include/rto_c.h rto_render_codes): one 32-bit word per pixel in TILE order -- the 128 pixels of a
	// block are consecutive, so a warp writes (and rto_resolve_codes reads) one full 128-byte line; the destination may be the memory
	// of another GPU (NVLink peer mapping), which is why the order is the kernel's own and not the image's.
	uint32_t* codes;
	unsigned codeFrame0;          // index, in the code buffer, of the frame blockIdx.z == 0 renders
	unsigned codeTilesY;          // blocks per image column = (height + 7) / 8
};

// word of one pixel: 0 = miss, else (position of the hit triangle in the scene's leaf order + 1) | (normal flipped << 30) | (shadowed << 31)
constexpr uint32_t kCodeShadowBit = 0x80000000u, kCodeFlipBit = 0x40000000u;
__device__ __forceinline__ size_t code_index(const RenderArgs& A) {
	const size_t tile = ((size_t)(A.codeFrame0 + blockIdx.z) * A.codeTilesY + (size_t)(A.y0 >> 3) + blockIdx.y) * gridDim.x + blockIdx.x;
	return tile * kRenderThreads + threadIdx.x;
}

__device__ __forceinline__ bool pixel_of_thread(const RenderArgs& A, const RtoCamera& cam, int& px, int& py, size_t& pix) {
	int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	// 4 wide x 8 tall per warp: measured 1-4 % faster than 8 x 4 on every kernel (16 x 2: 5-7 % slower)
	px = blockIdx.x * kRenderBlockW + warp * 4 + (lane & 3);
	py = A.y0 + blockIdx.y * 8 + (lane >> 2);
	if (px >= cam.width || py >= A.y1) return false;
	pix = (size_t)blockIdx.z * (size_t)(A.y1 - A.y0) * cam.width + (size_t)(py - A.y0) * cam.width + px;
	return true;
}

// Frame planes are written once and never read by the kernels: streaming stores (evict-first) keep the 400 MB a launch writes
// from pushing the scene out of L2.
RTO_DEV void store_pixel(const RenderArgs& A, size_t pix, V3 color, int id, float t) {
#if defined(__CUDA_ARCH__) && !defined(RTO_NO_STREAMING_STORES)
	if (A.rgba) __stcs(A.rgba + pix, make_float4(color.x, color.y, color.z, 1.0f));
	if (A.hitId) __stcs(A.hitId + pix, id);
	if (A.t) __stcs(A.t + pix, t);
#else
	if (A.rgba) A.rgba[pix] = make_float4(color.x, color.y, color.z, 1.0f);
	if (A.hitId) A.hitId[pix] = id;
	if (A.t) A.t[pix] = t;
#endif
}

#ifndef RTO_BVH_MIN_BLOCKS
#define RTO_BVH_MIN_BLOCKS 12     // 40 registers, 48 warps per SM.  Re-measured on the final kernel of round 2 (16 x 1080p, ms; DT primary + shadow / DT primary
                                  // only / 512^3 city mesh primary + shadow / sphere): 12 blocks 2.762 / 1.897 / 4.648 / 2.225 (11 blocks compiles to the same
                                  // 40 registers), 10 blocks (48 registers, the setting of rounds 1-2 until then) 2.829 / 1.939 / 4.826 / 2.276, 9 blocks
                                  // (56 registers) 2.923 / 2.008 / - / 2.343, 14 blocks (32 registers, spills) 2.923 / 1.914 on DT.  On the kernel of round 1 the order was the other way round (10 blocks 1.716
                                  // against 1.732 per 8 frames): since the while-while loop the node step waits on its loads more than on issue slots.
#endif
template <bool SHADOWS, bool PRUNE, bool WIDE = false>
__global__ void __launch_bounds__(kRenderThreads, RTO_BVH_MIN_BLOCKS * 128 / kRenderThreads) k_render_bvh(BvhDev S, RenderArgs A) {
	RtoCamera cam = A.cam0;
	if (A.cams) cam = A.cams[blockIdx.z];
	int px, py; size_t pix;
	if (!pixel_of_thread(A, cam, px, py, pix)) return;
	Ray ray = gen_ray(cam, px, py);
	float bestT; int bestPos;
	bvh_closest<PRUNE, WIDE>(S, ray.o, ray.d, bestT, bestPos);
	V3 color = mk3(0.0f, 0.0f, 0.0f);
	int id = -1;
	if (bestPos >= 0) {
		TriV tri = load_tri(S.tris, bestPos);
		id = tri.id;
		V3 n = normalize3(cross3(tri.e1, tri.e2));
		const bool flip = dot3(n, ray.d) > 0.0f;
		if (flip) n = -n;
		V3 hit = ray.o + ray.d * bestT;
		bool shadowed = false;
		if (SHADOWS) {
			V3 so = hit + n * A.shadowBias;
			V3 sd = normalize3(mk3(1.0f, 1.0f, 1.0f));
			shadowed = bvh_any<WIDE>(S, so, sd);
		}
		color = shadowed ? mk3(0.1f, 0.1f, 0.1f) : shade_lambert(n);
		if (A.codes) __stcs(A.codes + code_index(A), (uint32_t)(bestPos + 1) | (flip ? kCodeFlipBit : 0u) | (shadowed ? kCodeShadowBit : 0u));
	}
	else if (A.codes) __stcs(A.codes + code_index(A), 0u);
	store_pixel(A, pix, color, id, bestT);
}

// The frame planes from the hit codes: colour, hit id and t of a pixel are pure functions of (camera, pixel, hit triangle, flip and
// shadow bits) -- the same ray generation, the Moller-Trumbore distance computed by the same operations on the same record (the u / v
// tests decided nothing about its value), and the Lambert term of the triangle's normal, which is the same number for every pixel
// that sees the triangle and is therefore tabulated once per scene (k_shade_table; the flipped normal's term is its exact negation,
// products and sums of negated operands being negated results).  The planes equal a direct render bit for bit.
// 4 bytes in, 24 bytes out per pixel; about 170 instructions per hit pixel, most of them the IEEE divisions and square roots of the ray.
__global__ void __launch_bounds__(256) k_shade_table(const float4* __restrict__ tris, int numTris, float* __restrict__ ndotl) {
	const int pos = blockIdx.x * blockDim.x + threadIdx.x;
	if (pos >= numTris) return;
	TriV tri = load_tri(tris, pos);
	V3 n = normalize3(cross3(tri.e1, tri.e2));
	V3 lightDir = normalize3(mk3(-1.0f, -1.0f, -1.0f));
	ndotl[pos] = dot3(n, -lightDir);                  // shade_lambert(n) = (1, .8, .6) * max(0, this) + .1
}

RTO_DEV float mt_distance(const TriV& tri, V3 o, V3 d) {      // moller_trumbore()'s t, operation for operation
	V3 p = cross3(d, tri.e2);
	float det = dot3(tri.e1, p);
	float inv = 1.0f / det;
	V3 s = o - tri.v0;
	V3 q = cross3(s, tri.e1);
	return dot3(tri.e2, q) * inv;
}

#ifndef RTO_RESOLVE_MIN_BLOCKS
#define RTO_RESOLVE_MIN_BLOCKS 12
#endif
__global__ void __launch_bounds__(kRenderThreads, RTO_RESOLVE_MIN_BLOCKS * 128 / kRenderThreads) k_resolve_bvh(BvhDev S, RenderArgs A, const float* __restrict__ ndotl) {
	RtoCamera cam = A.cam0;
	if (A.cams) cam = A.cams[blockIdx.z];
	int px, py; size_t pix;
	if (!pixel_of_thread(A, cam, px, py, pix)) return;
	const uint32_t code = __ldcs(A.codes + code_index(A));
#if RTO_RESOLVE_HOIST_RAY
	Ray ray = gen_ray(cam, px, py);              // independent of the code word: its divisions and square roots run while that load is in flight
#endif
	V3 color = mk3(0.0f, 0.0f, 0.0f);
	int id = -1;
	float t = kMissT;
	const int pos = (int)(code & ~(kCodeShadowBit | kCodeFlipBit)) - 1;
	if (pos >= 0 && pos < S.numTris) {
#if !RTO_RESOLVE_HOIST_RAY
		Ray ray = gen_ray(cam, px, py);
#endif
		TriV tri = load_tri(S.tris, pos);
		id = tri.id;
		t = mt_distance(tri, ray.o, ray.d);
		const float nl = RTO_LDG(ndotl + pos);
		const float k = maxf(0.0f, (code & kCodeFlipBit) ? -nl : nl);
		color = (code & kCodeShadowBit) ? mk3(0.1f, 0.1f, 0.1f) : mk3(1.0f, 0.8f, 0.6f) * k + mk3(0.1f, 0.1f, 0.1f);
	}
	store_pixel(A, pix, color, id, t);
}

// One instantiation per traversal mode so that each gets its own register budget (mode A: 48 registers, mode B: 40).
#ifndef RTO_OCT_B_MIN_BLOCKS
#define RTO_OCT_B_MIN_BLOCKS 12    // re-measured: 12 blocks 2.12 / 3.31 ms, 8 blocks 2.20 / 3.45, 10 blocks 2.15 / 3.38, 14-16 blocks 2.18 / 3.41
#endif
#ifndef RTO_OCT_A_MIN_BLOCKS
#define RTO_OCT_A_MIN_BLOCKS 10    // re-measured on the one-loop walk (DT / 512^3 city, 16 x 1080p): 10 blocks (48 registers) 3.97 / 5.75 ms, 8 blocks 4.12 / 6.06, 6 blocks 4.16 / 6.21, 12 blocks 4.10 / 5.91
#endif
template <int MODE>
__global__ void __launch_bounds__(kRenderThreads, (MODE == RTO_MODE_OCTREE_SKIP ? RTO_OCT_A_MIN_BLOCKS : RTO_OCT_B_MIN_BLOCKS) * 128 / kRenderThreads) k_render_octree(OctDev S, RenderArgs A) {
	const int mode = MODE;
	RtoCamera cam = A.cam0;
	if (A.cams) cam = A.cams[blockIdx.z];
	int px, py; size_t pix;
	if (!pixel_of_thread(A, cam, px, py, pix)) return;
	Ray ray = gen_ray(cam, px, py);
	OctHit h = oct_trace<false>(S, mode, ray.o, ray.d, 0.0f, 1e30f);
	V3 color = mk3(0.0f, 0.0f, 0.0f);
	if (h.id >= 0) color = shade_lambert(h.normal);
	store_pixel(A, pix, color, h.id, h.t);
}

// ------------------------------------------------------------------------------------------------
// Explicit ray lists
// ------------------------------------------------------------------------------------------------
// Ray coherence sorting for explicit ray lists (RTO_FLAG_SORT_RAYS): key = octant of the direction (3 bits), Morton code of the point
// where the ray meets the scene box (7 bits per axis) and a coarse direction (4 + 4 bits).  Rays that share a key start in the same
// 1/128 cell of the scene and point the same way, so neighbouring lanes walk the same nodes.  The arithmetic here only decides the
// ORDER in which rays are traced (perm); every ray is traced by the same code and its result lands at its own index.
__global__ void __launch_bounds__(256) k_ray_sort_keys(const float* __restrict__ o3, const float* __restrict__ d3, size_t n,
	float lox, float loy, float loz, float hix, float hiy, float hiz, uint32_t* __restrict__ keys, uint32_t* __restrict__ idx) {
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float ox = o3[3 * i], oy = o3[3 * i + 1], oz = o3[3 * i + 2], dx = d3[3 * i], dy = d3[3 * i + 1], dz = d3[3 * i + 2];
	float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;
	float t0 = fmaxf(fmaxf(fminf((lox - ox) * ix, (hix - ox) * ix), fminf((loy - oy) * iy, (hiy - oy) * iy)), fmaxf(fminf((loz - oz) * iz, (hiz - oz) * iz), 0.0f));
	if (!(t0 >= 0.0f && t0 < 3.0e38f)) t0 = 0.0f;
	float px = ox + t0 * dx, py = oy + t0 * dy, pz = oz + t0 * dz;
	auto cell = [](float p, float lo, float hi) { float u = (p - lo) / (hi - lo) * 128.0f; int c = (int)u; return (uint32_t)(c < 0 || !(u == u) ? 0 : (c > 127 ? 127 : c)); };
	auto spread = [](uint32_t v) { v = (v | (v << 16)) & 0x030000ffu; v = (v | (v << 8)) & 0x0300f00fu; v = (v | (v << 4)) & 0x030c30c3u; v = (v | (v << 2)) & 0x09249249u; return v; };
	uint32_t m = spread(cell(px, lox, hix)) | (spread(cell(py, loy, hiy)) << 1) | (spread(cell(pz, loz, hiz)) << 2);
	float len = sqrtf(dx * dx + dy * dy + dz * dz);
	float ax = fabsf(dx) / len, ay = fabsf(dy) / len;
	uint32_t qa = (uint32_t)fminf(fmaxf(ax * 16.0f, 0.0f), 15.0f), qb = (uint32_t)fminf(fmaxf(ay * 16.0f, 0.0f), 15.0f);
	uint32_t oct = (dx < 0 ? 1u : 0u) | (dy < 0 ? 2u : 0u) | (dz < 0 ? 4u : 0u);
	keys[i] = (oct << 29) | ((m & 0x1fffffu) << 8) | (qa << 4) | qb;
	idx[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(128) k_trace_octree(OctDev S, int mode, const float* __restrict__ o3, const float* __restrict__ d3, size_t n,
	float tMin, float tMax, float* tOut, int* idOut, const uint32_t* __restrict__ perm) {
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if (perm) i = perm[i];
	V3 o = mk3(o3[3 * i], o3[3 * i + 1], o3[3 * i + 2]), d = mk3(d3[3 * i], d3[3 * i + 1], d3[3 * i + 2]);
	OctHit h = oct_trace<false>(S, mode, o, d, tMin, tMax);
	if (tOut) tOut[i] = h.t;
	if (idOut) idOut[i] = h.id;
}

__global__ void __launch_bounds__(128) k_trace_bvh(BvhDev S, unsigned flags, const float* __restrict__ o3, const float* __restrict__ d3, size_t n,
	float* tOut, int* idOut, const uint32_t* __restrict__ perm) {
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if (perm) i = perm[i];
	V3 o = mk3(o3[3 * i], o3[3 * i + 1], o3[3 * i + 2]), d = mk3(d3[3 * i], d3[3 * i + 1], d3[3 * i + 2]);
	float bestT; int bestPos;
	if (flags & RTO_FLAG_NO_PRUNE) bvh_closest<false>(S, o, d, bestT, bestPos); else bvh_closest<true>(S, o, d, bestT, bestPos);
	if (tOut) tOut[i] = bestT;
	if (idOut) idOut[i] = bestPos >= 0 ? f2i(RTO_LDG(S.tris + 4 * (size_t)bestPos + 2).y) : -1;
}

// BVH::query candidates: pass 1 counts per ray, pass 2 writes ids at the host-computed offsets
__global__ void __launch_bounds__(128) k_bvh_query(BvhDev S, const float* __restrict__ o3, const float* __restrict__ d3, size_t n,
	const long long* __restrict__ offsets, int* counts, int* ids) {
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	V3 o = mk3(o3[3 * i], o3[3 * i + 1], o3[3 * i + 2]), d = mk3(d3[3 * i], d3[3 * i + 1], d3[3 * i + 2]);
	unsigned long long boxes = 0;
	int cnt = 0;
	long long base = offsets ? offsets[i] : 0;
	bvh_replay(S, o, d, boxes, [&](int pos) {
		if (ids) ids[base + cnt] = f2i(RTO_LDG(S.tris + 4 * (size_t)pos + 2).y);
		cnt++;
	});
	if (counts) counts[i] = cnt;
}

// ------------------------------------------------------------------------------------------------
// Work counters of the reference algorithm (B, C, N)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_add(unsigned long long* dst, unsigned long long v) {
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	if ((threadIdx.x & 31) == 0 && v) atomicAdd(dst, v);
}

__global__ void __launch_bounds__(kRenderThreads) k_stats_bvh(BvhDev S, RenderArgs A, unsigned long long* stats) {
	const RtoCamera& cam = A.cam0;
	int px, py; size_t pix;
	bool active = pixel_of_thread(A, cam, px, py, pix);
	unsigned long long B = 0, C = 0, Bs = 0, Cs = 0, nS = 0;
	if (active) {
		Ray ray = gen_ray(cam, px, py);
		bvh_replay(S, ray.o, ray.d, B, [&](int) { C++; });
		if (A.flags & RTO_FLAG_SHADOWS) {
			float bestT; int bestPos;
			bvh_closest<false>(S, ray.o, ray.d, bestT, bestPos);
			if (bestPos >= 0) {
				TriV tri = load_tri(S.tris, bestPos);
				V3 n = normalize3(cross3(tri.e1, tri.e2));
				if (dot3(n, ray.d) > 0.0f) n = -n;
				V3 so = (ray.o + ray.d * bestT) + n * A.shadowBias;
				V3 sd = normalize3(mk3(1.0f, 1.0f, 1.0f));
				nS = 1;
				bvh_replay(S, so, sd, Bs, [&](int) { Cs++; });
			}
		}
	}
	warp_add(stats + 0, B); warp_add(stats + 1, C); warp_add(stats + 2, Bs); warp_add(stats + 3, Cs); warp_add(stats + 4, nS);
}

__global__ void __launch_bounds__(kRenderThreads) k_stats_octree(OctDev S, RenderArgs A, int mode, unsigned long long* stats) {
	const RtoCamera& cam = A.cam0;
	int px, py; size_t pix;
	bool active = pixel_of_thread(A, cam, px, py, pix);
	unsigned long long N = 0;
	if (active) {
		Ray ray = gen_ray(cam, px, py);
		OctHit h = oct_trace<true>(S, mode, ray.o, ray.d, 0.0f, 1e30f);
		N = h.visits;
	}
	warp_add(stats + 0, N);
}

} // namespace rto
