// rto_shim_math.h -- the few glm types the reference's class signatures mention.
// Define RTO_SHIM_USE_GLM before including a shim header to use the real glm (the reference vendors 0.9.9.7); otherwise
// layout-compatible stand-ins are used (vec3 = 3 floats, mat4 = 16 floats column-major), enough for the signatures.
#pragma once
#ifdef RTO_SHIM_USE_GLM
#include <glm/glm.hpp>
namespace rto_shim { using vec3 = glm::vec3; using mat4 = glm::mat4; }
#else
namespace rto_shim {
struct vec3 {
	float x, y, z;
	vec3() : x(0), y(0), z(0) {}
	explicit vec3(float s) : x(s), y(s), z(s) {}
	vec3(float a, float b, float c) : x(a), y(b), z(c) {}
	float& operator[](int i) { return (&x)[i]; }
	const float& operator[](int i) const { return (&x)[i]; }
};
struct mat4 { float m[16]; float* operator[](int c) { return m + 4 * c; } const float* operator[](int c) const { return m + 4 * c; } };
}
#endif
static_assert(sizeof(rto_shim::vec3) == 12, "vec3 must be 3 packed floats");
