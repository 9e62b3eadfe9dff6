// scene.cpp — see scene.h. Float arithmetic follows main.cpp:679-682 operation by operation.
#include "scene.h"
#include <fstream>
#include <iostream>
#include <sstream>

std::string model_file_name(SceneModel m)
{
	switch (m) {
	case IGEA: return "igea.obj";
	case ARMADILLO: return "armadillo.obj";
	case BUNNIES: return "bunnies.obj";
	case CITY: return "city.obj";
	case GRASS: return "grass.obj";
	case BUNNY: return "bunny.obj";
	case BUDDHA: return "buddha.obj";
	default: return "test.obj";   // TEST and anything else (main.cpp:644-648)
	}
}

bool load_obj(const std::string& path, bool parse_faces, std::vector<float>& vertices, std::vector<int>& faces)
{
	std::ifstream file(path);
	if (!file.good()) return false;
	if (!parse_faces) {
		while (true) {
			std::string text;
			file >> text;
			if (text != "v") break;          // main.cpp:694-697: anything but "v" stops the loader
			float x = 0, y = 0, z = 0;
			file >> x; file >> y; file >> z;
			vertices.push_back(x); vertices.push_back(y); vertices.push_back(z);
		}
		return true;
	}
	std::string line;
	while (std::getline(file, line)) {
		std::istringstream ls(line);
		std::string tag;
		ls >> tag;
		if (tag == "v") {
			float x = 0, y = 0, z = 0;
			ls >> x >> y >> z;
			vertices.push_back(x); vertices.push_back(y); vertices.push_back(z);
		} else if (tag == "f") {
			std::vector<int> idx;
			std::string tok;
			while (ls >> tok) {
				int v = std::stoi(tok.substr(0, tok.find('/')));
				int nv = (int)(vertices.size() / 3);
				idx.push_back(v > 0 ? v - 1 : nv + v);
			}
			for (size_t k = 2; k < idx.size(); ++k) { faces.push_back(idx[0]); faces.push_back(idx[k - 1]); faces.push_back(idx[k]); }
		}
	}
	return true;
}

bool create_scene(const Settings& settings, const std::string& models_dir, int number_of_clones, HostScene& scene)
{
	scene = HostScene();
	std::string fileName = models_dir + "/" + model_file_name(settings.sceneModel);
	for (int clone = 0; clone < number_of_clones; clone++) {
		float shift = clone * 20;                                   // main.cpp:652
		std::cout << "Loading file:  " << fileName << " ... " << std::endl;
		std::vector<float> v;
		std::vector<int> f;
		bool ok = load_obj(fileName, false, v, f);
		if (!ok && settings.sceneModel == IGEA) ok = load_obj(models_dir + "/Igea.obj", false, v, f);
		if (!ok) {
			std::cout << "Can't open file " << fileName << std::endl;
			return false;
		}
		for (size_t i = 0; i + 2 < v.size(); i += 3) {
			float cx = v[i] * 100 + shift, cy = v[i + 1] * 100 + shift, cz = v[i + 2] * 100 + shift;   // :680
			cy += -10;                                              // :681
			cz += -60;                                              // :682
			scene.cxyz_r.push_back(cx); scene.cxyz_r.push_back(cy); scene.cxyz_r.push_back(cz);
			scene.cxyz_r.push_back((float)(0.01 * 5));              // :679
			scene.rgb_mat.push_back(0.8f); scene.rgb_mat.push_back(0.7f); scene.rgb_mat.push_back(0.0f);   // :689
			scene.rgb_mat.push_back((float)DIFFUSE_AND_GLOSSY);     // Sphere's ctor hard-sets it (accelerators.h:72)
		}
	}
	unsigned id = (unsigned)scene.n();
	// the ground sphere, main.cpp:703
	scene.cxyz_r.push_back(0.93591022f); scene.cxyz_r.push_back(-105.47120094f); scene.cxyz_r.push_back(-43.2363205f);
	scene.cxyz_r.push_back(100.0f);
	scene.rgb_mat.push_back(0.f); scene.rgb_mat.push_back(0.f); scene.rgb_mat.push_back(0.f); scene.rgb_mat.push_back((float)DIFFUSE_AND_GLOSSY);
	std::cout << "Number of spheres: " << id << std::endl;      // :718
	return true;
}

// C entry points for the Python tests (ctypes): the loader restatement is host logic and is checked on CPU.
extern "C" int rtds_host_scene_from_obj(const char* models_dir, int scene_model, int clones, float* cxyz_r, float* rgb_mat, int cap)
{
	Settings s;
	s.sceneModel = (SceneModel)scene_model;
	HostScene sc;
	std::streambuf* old = std::cout.rdbuf(nullptr);
	bool ok = create_scene(s, models_dir, clones, sc);
	std::cout.rdbuf(old);
	if (!ok) return -1;
	if (sc.n() > cap) return -2;
	for (size_t i = 0; i < sc.cxyz_r.size(); ++i) { cxyz_r[i] = sc.cxyz_r[i]; rgb_mat[i] = sc.rgb_mat[i]; }
	return sc.n();
}
