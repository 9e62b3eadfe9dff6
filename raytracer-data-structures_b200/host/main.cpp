// main.cpp — host driver with the control flow and log lines of the reference's main() (main.cpp:723-853), with
// the hot path (acceleration-structure build, ray generation, traversal, intersection, shading, quantisation)
// running on B200 GPUs through the C ABI of include/rtds.h.  Host code stays C++; there is no CPU fallback.
//
//   settings.h            compile-time surface, as in the reference
//   RTDS_WIDTH/HEIGHT/AA/DS/MODEL/CLONES   run-time overrides of the same fields (+ NUMBER_OF_CLONES, main.cpp:61)
//   RTDS_MODELS_DIR       directory holding the .obj files (default "models", like the reference)
//   RTDS_GPUS             1,2,4,8: interleaved scanline tiles, one context + one host thread per GPU
//   RTDS_EXACT=1          reference traversal (visit every node whose slab test passes) instead of the ordered one
//   RTDS_LBVH_MODE        compat (default: what the reference's LBVH code does) | true (Morton/Karras LBVH)
//   RTDS_PRIMS=triangles  extension: read `f` lines too and render the mesh's triangles (Moller-Trumbore) instead of one
//                         sphere per vertex; vertices get the same transform as the sphere centres (main.cpp:680-682)
//   RTDS_KD_CLOSEST=1     extension: KDTREE frames by closest-hit traversal + shading instead of the reference's any-hit
//                         black/sky frame (main.cpp:362-372)
//   RTDS_OUT              output file (default ./output.ppm)
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rtds.h"
#include "scene.h"
#include "settings.h"

using std::cout;
using std::endl;

static int env_int(const char* name, int dflt)
{
	const char* v = getenv(name);
	return v && *v ? atoi(v) : dflt;
}

static double seconds_since(std::chrono::steady_clock::time_point t0)
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// write_into_file (main.cpp:516-528): the 8-bit quantisation already happened on the device
static void write_ppm(const std::string& path, const Settings& s, const std::vector<uint8_t>& rgb)
{
	std::ofstream ofs(path, std::ios::out | std::ios::binary);
	ofs << "P6\n" << s.width << " " << s.height << "\n255\n";
	ofs.write((const char*)rgb.data(), (std::streamsize)rgb.size());
}

#define CHECK(call)                                                                   \
	do {                                                                              \
		int rc_ = (call);                                                             \
		if (rc_ != RTDS_OK) {                                                         \
			std::cerr << "rtds error " << rc_ << ": " << rtds_last_error() << endl;   \
			exit(2);                                                                  \
		}                                                                             \
	} while (0)

int main(int, char**)
{
	const auto begin_total_time = std::chrono::steady_clock::now();
	const auto begin_time = std::chrono::steady_clock::now();
	Settings settings;
	settings.width = env_int("RTDS_WIDTH", settings.width);
	settings.height = env_int("RTDS_HEIGHT", settings.height);
	settings.aa_samples = env_int("RTDS_AA", settings.aa_samples);
	settings.dataStructure = (AccType)env_int("RTDS_DS", settings.dataStructure);
	settings.sceneModel = (SceneModel)env_int("RTDS_MODEL", settings.sceneModel);
	const int clones = env_int("RTDS_CLONES", 1);
	const int n_gpus = env_int("RTDS_GPUS", 1);
	const int exact = env_int("RTDS_EXACT", 0);
	const char* models_dir = getenv("RTDS_MODELS_DIR") ? getenv("RTDS_MODELS_DIR") : "models";
	const char* out_path = getenv("RTDS_OUT") ? getenv("RTDS_OUT") : "./output.ppm";
	const bool lbvh_true = getenv("RTDS_LBVH_MODE") && !strcmp(getenv("RTDS_LBVH_MODE"), "true");
	// RTDS_TIMING=1: wall-clock stage breakdown of this (cold) run on stderr
	const bool timing = env_int("RTDS_TIMING", 0) != 0;
	auto stage = [&](const char* what) { if (timing) std::cerr << "[rtds_main] " << what << " done at " << seconds_since(begin_total_time) << " s" << endl; };

	// Print settings (main.cpp:730-736)
	cout << "Start rendering .... \n";
	cout << "Settings are.... \n";
	cout << "Height: " << settings.height << endl;
	cout << "Width: " << settings.width << endl;
	cout << "DataStructure: " << settings.dataStructure << endl;
	cout << "Anti-aliasing samples: " << settings.aa_samples << endl;
	cout << "\n====================================================================\n";

	HostScene scene;
	if (!create_scene(settings, models_dir, clones, scene)) return 1;
	const bool triangles = getenv("RTDS_PRIMS") && !strcmp(getenv("RTDS_PRIMS"), "triangles");
	std::vector<float> tri_mat;
	if (triangles) {
		std::vector<float> v;
		std::vector<int> f;
		std::string path = std::string(models_dir) + "/" + model_file_name(settings.sceneModel);
		if (!load_obj(path, true, v, f) || f.empty()) {
			std::cerr << "RTDS_PRIMS=triangles: " << path << " has no `f` lines (the reference's models are vertex-only)" << endl;
			return 1;
		}
		for (size_t k = 0; k + 2 < f.size(); k += 3)
			for (int c = 0; c < 3; ++c) {
				const float* p = &v[3 * (size_t)f[k + c]];
				scene.tris.push_back(p[0] * 100);
				scene.tris.push_back(p[1] * 100 + -10);
				scene.tris.push_back(p[2] * 100 + -60);
			}
		tri_mat.assign(scene.tris.size() / 9 * 4, 0.f);
		for (size_t i = 0; i < tri_mat.size() / 4; ++i) { tri_mat[4 * i] = 0.8f; tri_mat[4 * i + 1] = 0.7f; }
		cout << "Number of triangles: " << scene.tris.size() / 9 << endl;
	}

	stage("scene loaded");
	cout << "Wraping BV for each object .... \n";
	cout << "Done .... \n Time: ";
	cout << float(seconds_since(begin_time)) << "s";
	cout << "\n====================================================================\n";

	std::vector<rtds_ctx*> ctx(n_gpus, nullptr);
	for (int g = 0; g < n_gpus; ++g) {
		CHECK(rtds_create(&ctx[g], g));
		stage("rtds_create (CUDA context + module load)");
		if (triangles) CHECK(rtds_set_triangles(ctx[g], scene.tris.data(), tri_mat.data(), (int)(scene.tris.size() / 9)));
		else CHECK(rtds_set_spheres(ctx[g], scene.cxyz_r.data(), scene.rgb_mat.data(), scene.n()));
	}

	stage("scene uploaded");
	rtds_build_params bp;
	memset(&bp, 0, sizeof bp);
	bp.mode = (settings.dataStructure == LBVH && lbvh_true) ? RTDS_MODE_TRUE : RTDS_MODE_COMPAT;
	std::vector<rtds_build_stats> bs(n_gpus);
	auto build_all = [&]() {   // the structure is replicated: every GPU builds its own copy concurrently
		std::vector<std::thread> th;
		for (int g = 0; g < n_gpus; ++g) th.emplace_back([&, g]() { CHECK(rtds_build(ctx[g], settings.dataStructure, &bp, &bs[g])); });
		for (auto& t : th) t.join();
	};

	switch (settings.dataStructure) {
	case BVH: {
		cout << "<<<<<<< This is BVH >>>>>>" << endl;
		cout << "construct BVH Tree .... \n";
		const auto t0 = std::chrono::steady_clock::now();
		build_all();
		cout << "Done .... Time: ";
		cout << float(seconds_since(t0)) << "s\n";
		cout << "Total number of nodes: " << bs[0].total_nodes << "\n";
		break;
	}
	case KDTREE: {
		cout << "<<<<<<< This is KDTREE >>>>>>";
		cout << "construct KD-Tree .... ";
		const auto t0 = std::chrono::steady_clock::now();
		build_all();
		cout << "Depth is: " << bs[0].max_depth << "\n";
		cout << "Done .... Time: ";
		cout << float(seconds_since(t0)) << "s\n";
		cout << "Number of nodes: " << bs[0].total_nodes << "\n";
		break;
	}
	case LBVH: {
		cout << "<<<<<<< This is LBVH >>>>>>";
		cout << "construct LBVH Tree .... ";
		const auto t0 = std::chrono::steady_clock::now();
		build_all();
		cout << "Total number of nodes: " << bs[0].total_nodes << "\n";
		cout << "Done .... Time: ";
		cout << float(seconds_since(t0)) << "s\n";
		break;
	}
	default:
		cout << "<<<<<<< Warning: No data structure is used, this can take long time! >>>>>>";
		break;
	}

	stage("structure built");
	// render (main.cpp:541-566): every GPU renders its interleaved scanline tiles into the shared host frame
	std::vector<uint8_t> rgb((size_t)settings.width * settings.height * 3);
	std::vector<rtds_render_stats> rs(n_gpus);
	// N > 1: GPU 0 owns the frame and the other GPUs' render kernels store their tiles straight into it (peer access over
	// NVLink, rtds_shared_frame_*); without peer access every GPU renders into its own buffer and copies its rows to the host.
	bool shared = n_gpus > 1 && rtds_shared_frame_create(ctx[0], settings.width, settings.height, n_gpus, nullptr) == RTDS_OK;
	for (int g = 1; shared && g < n_gpus; ++g) shared = rtds_shared_frame_attach(ctx[g], ctx[0], g) == RTDS_OK;
	{
		std::vector<std::thread> th;
		for (int g = 0; g < n_gpus; ++g)
			th.emplace_back([&, g]() {
				rtds_render_params rp;
				memset(&rp, 0, sizeof rp);
				rp.width = settings.width; rp.height = settings.height; rp.aa_samples = settings.aa_samples;
				rp.exact = exact; rp.rank = g; rp.world = n_gpus; rp.tile_rows = 8;
				rp.kd_closest = env_int("RTDS_KD_CLOSEST", 0);   // 0 = the reference's any-hit, unshaded KD frame
				rp.tri_geometric = env_int("RTDS_TRI_GEOMETRIC", 0);   // 1 = the triangle test the reference compiles (main.cpp:163-215)
				if (shared) CHECK(rtds_render_shared(ctx[g], settings.dataStructure, &rp, 1u, &rs[g]));
				else CHECK(rtds_render(ctx[g], settings.dataStructure, &rp, rgb.data(), nullptr, nullptr, &rs[g]));
			});
		for (auto& t : th) t.join();
	}
	if (shared) CHECK(rtds_shared_frame_read(ctx[0], rgb.data()));
	stage("frame rendered and downloaded");
	write_ppm(out_path, settings, rgb);
	stage("output.ppm written");

	unsigned long long tests = 0, rays = 0;
	float traverse_ms = 0;
	for (int g = 0; g < n_gpus; ++g) { tests += rs[g].prim_tests; rays += rs[g].rays; traverse_ms = std::max(traverse_ms, rs[g].ms_kernel); }
	// end-of-run statistics (main.cpp:845-851)
	cout << "\n Number of the tree nodes: " << 0 << " bv node" << endl;
	cout << "\n Tree traverse time spent: " << traverse_ms / 1000.0f << "s" << endl;
	cout << "\n Number of Sphere intersection tests: " << tests << " test" << endl;
	cout << "\n Total time spent: ";
	cout << float(seconds_since(begin_total_time)) << "s" << endl;
	cout << "\n--------- Rendering Completed ---------\n";
	cout << " [rtds] " << rays << " rays on " << n_gpus << " GPU(s), render kernel " << traverse_ms << " ms, "
	     << (rays / 1e6) / (traverse_ms / 1e3) << " Mrays/s" << endl;
	for (int g = 0; g < n_gpus; ++g) rtds_destroy(ctx[g]);
	return 0;
}
