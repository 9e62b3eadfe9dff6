// settings.h — the compile-time configuration surface of the ray tracer, kept field-for-field compatible with
// the reference's settings.h (/root/reference/project/raytracer/settings.h:4-17): the same enum values, the same
// member names, types and defaults, so a user edits the same lines they would edit there.
// Every field can additionally be overridden at run time through RTDS_* environment variables (main.cpp).
#ifndef RTDS_HOST_SETTINGS_H
#define RTDS_HOST_SETTINGS_H

#include <cstdint>
#include "host_types.h"   // Vec3f, AccType — the reference includes accelerators.h before settings.h for these

// which model createScene_new loads from models/  (settings.h:4)
enum SceneModel { IGEA, ARMADILLO, BUNNY, BUNNIES, TEST, GRASS, BUDDHA, CITY };

struct Settings
{
	uint32_t   width           = 640;               // image width in pixels
	uint32_t   height          = 480;               // image height in pixels
	float      fov             = 90;                // declared by the reference but unused: render() uses 30 (main.cpp:545)
	Vec3f      backgroundColor = Vec3f(1, 1, 1);    // declared but unused: castRay returns (0.6,0.8,1) for the sky
	float      bias            = 0.0001;            // declared but unused: castRay uses 1e-4
	uint32_t   aa_samples      = 1;                 // jittered samples per pixel
	AccType    dataStructure   = BVH;               // BVH, KDTREE, LBVH or NONE
	SceneModel sceneModel      = BUNNY;
};

#endif
