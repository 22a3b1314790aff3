// rto_math.h -- fp32 vector helpers shared by the host builders and the CUDA kernels.
//
// Every helper spells out the operation ORDER of glm 0.9.9.7's scalar paths (the reference's arithmetic,
// thirdparty/glm-0.9.9.7/glm/detail/func_geometric.inl:48-90, type_mat4x4.inl:561-572) with one IEEE
// operation per C++ operator.  Device code is compiled with -fmad=false -prec-div=true -prec-sqrt=true and
// host code with -ffp-contract=off, so no multiply-add is ever fused and CPU and GPU round identically.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define RTO_HD __host__ __device__ __forceinline__
#else
#define RTO_HD inline
#endif

namespace rto {

struct V3 { float x, y, z; };

RTO_HD V3 mk3(float a, float b, float c) { V3 r; r.x = a; r.y = b; r.z = c; return r; }
RTO_HD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
RTO_HD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
RTO_HD V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
RTO_HD V3 operator*(V3 a, V3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
RTO_HD V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
RTO_HD V3 operator*(float s, V3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
RTO_HD V3 operator/(V3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
RTO_HD float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// dot(vec3) = (x + y) + z on the element-wise product
RTO_HD float dot3(V3 a, V3 b) { float px = a.x * b.x, py = a.y * b.y, pz = a.z * b.z; return px + py + pz; }
RTO_HD V3 cross3(V3 a, V3 b) { return mk3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
// normalize(v) = v * (1 / sqrt(dot(v, v)))
RTO_HD V3 normalize3(V3 v) { float s = 1.0f / sqrtf(dot3(v, v)); return v * s; }

// min/max with the operand order of glm::min/max, std::min/max and GLSL min/max (NaN behaviour included)
RTO_HD float minf(float a, float b) { return (b < a) ? b : a; }
RTO_HD float maxf(float a, float b) { return (a < b) ? b : a; }
RTO_HD V3 min3(V3 a, V3 b) { return mk3(minf(a.x, b.x), minf(a.y, b.y), minf(a.z, b.z)); }
RTO_HD V3 max3(V3 a, V3 b) { return mk3(maxf(a.x, b.x), maxf(a.y, b.y), maxf(a.z, b.z)); }

// The fixed light of the reference's shade(): lightDir = normalize(vec3(-1)); colour = (1,.8,.6)*max(0, n.(-L)) + .1
// (GLSL shade, RayTracerBVH.cpp:331-336).
RTO_HD V3 shade_lambert(V3 n) {
	V3 lightDir = normalize3(mk3(-1.0f, -1.0f, -1.0f));
	float ndotl = maxf(0.0f, dot3(n, -lightDir));
	return mk3(1.0f, 0.8f, 0.6f) * ndotl + mk3(0.1f, 0.1f, 0.1f);
}

} // namespace rto
