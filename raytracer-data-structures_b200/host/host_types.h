// host_types.h — the few host-side types main.cpp and settings.h need (the reference gets them from
// geometry.h:37-150 and accelerators.h:21-24).
#ifndef RTDS_HOST_TYPES_H
#define RTDS_HOST_TYPES_H

struct Vec3f {
	float x, y, z;
	Vec3f() : x(0), y(0), z(0) {}
	Vec3f(float v) : x(v), y(v), z(v) {}
	Vec3f(float xx, float yy, float zz) : x(xx), y(yy), z(zz) {}
};

// accelerators.h:21 — values are part of the surface (main.cpp prints dataStructure as an integer)
enum AccType { BVH, KDTREE, UNIFORM_GRID, LBVH, NONE };
// accelerators.h:24
enum MaterialType { DIFFUSE_AND_GLOSSY, REFLECTION_AND_REFRACTION, REFLECTION };

#endif
