// Camera.h (shim) -- the reference's orbit camera (453-skeleton/Camera.h:5-44, Camera.cpp:8-29) on top of librto.
#pragma once
#include "../../../include/rto_c.h"
#include "rto_shim_math.h"
#include <cstring>

class Camera {
public:
	Camera(float t, float p, float r) : theta(t), phi(p), radius(r), target(0.0f) {}
	rto_shim::mat4 getView() const { RtoCamera c; rto_shim::mat4 v; consts(45.0f, 1.0f, 1, 1, c, &v); return v; }
	rto_shim::vec3 getPos() const { RtoCamera c; consts(45.0f, 1.0f, 1, 1, c, nullptr); return rto_shim::vec3(c.camPos[0], c.camPos[1], c.camPos[2]); }
	void setTarget(const rto_shim::vec3& t) { target = t; }
	const rto_shim::vec3& getTarget() const { return target; }
	float getTheta() const { return theta; }
	float getPhi() const { return phi; }
	float getR() const { return radius; }
	// per-frame constants of the GLSL generateRay (RayTracerBVH.cpp:338-355): inverse(view), tan(fov/2), camera position
	int consts(float fovDeg, float aspect, int w, int h, RtoCamera& out, rto_shim::mat4* view) const {
		float tgt[3] = { target.x, target.y, target.z };
		float v[16];
		int rc = rto_host_camera_orbit(theta, phi, radius, tgt, fovDeg, aspect, w, h, &out, v);
		if (view) std::memcpy(view, v, 64);
		return rc;
	}
	float theta, phi, radius;
	rto_shim::vec3 target;
};
