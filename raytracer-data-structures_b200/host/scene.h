// scene.h — host-side scene creation with the semantics of createScene_new (main.cpp:599-721).
#ifndef RTDS_HOST_SCENE_H
#define RTDS_HOST_SCENE_H

#include <string>
#include <vector>
#include "settings.h"

struct HostScene {
	std::vector<float> cxyz_r;   // n x {cx,cy,cz,radius}, objId = row
	std::vector<float> rgb_mat;  // n x {r,g,b,(float)MaterialType}
	std::vector<float> tris;     // optional: m x 9 (v0,v1,v2) from `f` lines — extension, see load_obj()
	int n() const { return (int)(cxyz_r.size() / 4); }
};

// File name createScene_new picks for a SceneModel (main.cpp:608-649). On case-sensitive file systems the reference's
// "igea.obj" does not match the shipped "Igea.obj"; when the exact name is missing the loader retries the shipped one.
std::string model_file_name(SceneModel m);

// Reads an OBJ the way the reference does (main.cpp:663-698): whitespace-separated tokens; "v" followed by three
// floats adds a vertex; ANY other token ends the file. With parse_faces (extension, off for parity) it instead
// parses a conventional OBJ: `v`, `f` (triangulated fans, 1-based / negative indices), everything else skipped.
bool load_obj(const std::string& path, bool parse_faces, std::vector<float>& vertices, std::vector<int>& faces);

// createScene_new: one sphere of radius 0.05 per vertex at v*100 + 20*clone, y -= 10, z -= 60, colour (0.8,0.7,0),
// followed by the ground sphere (main.cpp:703). Prints the same lines. Returns false when the model cannot be opened
// (the reference then continues with an empty scene and reads out of bounds; we stop instead).
bool create_scene(const Settings& settings, const std::string& models_dir, int number_of_clones, HostScene& scene);

#endif
