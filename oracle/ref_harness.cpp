// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// oracle/ref_harness.cpp: library-style wrapper around the UNMODIFIED reference
// translation unit (/root/reference/project/raytracer/main.cpp, which #includes
// accelerators.h, geometry.h and settings.h).  Nothing is copied: the reference is
// #included from where it lies, `main` is renamed, and this file only adds
// extern "C" entry points that call the reference's own functions
// (createScene_new, constructBVHNew, constructLBVHTree, constructKDTreeNew,
// boxIntersect, kdtreeIntersect, Sphere::raySphereIntersect, castRay) and dump
// their results.  Built by oracle/Makefile into oracle/_ref/libref_oracle.so
// (git-ignored, travels to the GPU box).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.
//
// The ray loop below mirrors render() (main.cpp:541-566) without the per-pixel
// progress printf and with a generator the harness owns (render()'s is a
// function-local static that cannot be reset); it is the same libstdc++
// std::mt19937 + uniform_real_distribution<double> as main.cpp:503-508.

#include <cstdint>
#include <memory>
#define main ref_main
#include "/root/reference/project/raytracer/main.cpp"
#undef main

#include <unistd.h>
#include <chrono>

namespace {
std::vector<SceneObject> g_scene;          // the vector the builders reorder in place
std::shared_ptr<Node>    g_root;
std::vector<Sphere>      g_lights;
std::vector<Sphere>      g_spheres_unused;
int                      g_total_nodes = 0;
int                      g_acc = NONE;

void reset_kd_globals()
{
    if (::nodes) { free(::nodes); ::nodes = nullptr; }
    nAllocedNodes = 0;
    nextFreeNode = 0;
    totalKdNodes = 0;
    ::bounds = BoxBoundries();
    primBounds.clear();
    kdtreeAllSceneObjects.clear();
    kdtreePrimitiveIndices.clear();
}

void default_lights()
{
    g_lights.clear();
    // main.cpp:775
    g_lights.push_back(Sphere(0, Vec3f(0, 3, 30), 10, Vec3f(1, 1, 1), 0, 0.0, Vec3f(1)));
}

struct DumpRec {           // one pre-order node record (see ref_bvh_dump)
    int32_t  isleaf;
    int32_t  nobjs;
    int32_t  first_obj;    // objs[0] for leaves, -1 otherwise
    int32_t  axis;         // longestAxis for interior nodes, -1 for leaves
    uint32_t box[6];       // min.xyz max.xyz as IEEE bit patterns
};

void dump_preorder(const std::shared_ptr<Node>& n, std::vector<DumpRec>& out, std::vector<int>* leaf_objs)
{
    DumpRec r;
    r.isleaf = n->isleaf ? 1 : 0;
    r.nobjs = (int)n->objs.size();
    r.first_obj = n->objs.empty() ? -1 : (int)n->objs[0];
    r.axis = n->isleaf ? -1 : (int)n->longestAxis;
    float b[6] = { n->boxBoundries.min.x, n->boxBoundries.min.y, n->boxBoundries.min.z,
                   n->boxBoundries.max.x, n->boxBoundries.max.y, n->boxBoundries.max.z };
    memcpy(r.box, b, sizeof b);
    out.push_back(r);
    if (n->isleaf) {
        if (leaf_objs) for (unsigned o : n->objs) leaf_objs->push_back((int)o);
        return;
    }
    dump_preorder(n->leftchild, out, leaf_objs);
    dump_preorder(n->rightchild, out, leaf_objs);
}
} // namespace

extern "C" {

// Scene from the reference's own loader (main.cpp:599-721).  `dir` must contain models/.
int ref_scene_from_obj(const char* dir, int sceneModel, int clones)
{
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) return -1;
    if (chdir(dir) != 0) return -2;
    sceneFixed.clear();
    g_root.reset();
    Settings s;
    s.sceneModel = (SceneModel)sceneModel;
    NUMBER_OF_CLONES = (char)clones;
    std::streambuf* old = std::cout.rdbuf(nullptr);
    g_scene = createScene_new(s);
    std::cout.rdbuf(old);
    NUMBER_OF_CLONES = 1;
    if (chdir(cwd) != 0) return -3;
    default_lights();
    return (int)g_scene.size();
}

// Synthetic scene, filled exactly the way main.cpp:677-692 fills `scene` and `sceneFixed`.
// mat: 0 = DIFFUSE_AND_GLOSSY (what the reference always ends up with), 1 = R&R, 2 = REFLECTION
// (set on the public member after construction — the reference ctor cannot set it).
int ref_scene_from_spheres(const float* cxyz_r, const float* rgb_mat, int n)
{
    sceneFixed.clear();
    g_scene.clear();
    g_root.reset();
    for (int i = 0; i < n; ++i) {
        SceneObject s;
        s.objId = i;
        s.radius = cxyz_r[4 * i + 3];
        s.center = Vec3f(cxyz_r[4 * i], cxyz_r[4 * i + 1], cxyz_r[4 * i + 2]);
        s.position = s.center;
        s.shininess = 64;
        s.isSphere = true;
        Vec3f minPoint = s.center - s.radius;
        Vec3f maxPoint = s.center + s.radius;
        s.boxBoundries = BoxBoundries(minPoint, maxPoint);
        s.sphere = Sphere(i, s.center, s.radius,
                          Vec3f(rgb_mat[4 * i], rgb_mat[4 * i + 1], rgb_mat[4 * i + 2]), 0, 0.0);
        int mat = (int)rgb_mat[4 * i + 3];
        s.sphere.materialType = (MaterialType)mat;
        g_scene.push_back(s);
        sceneFixed.push_back(s);
    }
    default_lights();
    return n;
}

int ref_set_lights(const float* cxyz_r_rgb, int m)
{
    g_lights.clear();
    for (int i = 0; i < m; ++i) {
        const float* l = cxyz_r_rgb + 7 * i;
        g_lights.push_back(Sphere(0, Vec3f(l[0], l[1], l[2]), l[3], Vec3f(1, 1, 1), 0, 0.0,
                                  Vec3f(l[4], l[5], l[6])));
    }
    return m;
}

int ref_scene_size() { return (int)sceneFixed.size(); }

// objId-indexed primitive table (sceneFixed): centre, radius, colour, material.
void ref_scene_get(float* cxyz_r, float* rgb_mat)
{
    for (size_t i = 0; i < sceneFixed.size(); ++i) {
        const SceneObject& s = sceneFixed[i];
        cxyz_r[4 * i] = s.center.x; cxyz_r[4 * i + 1] = s.center.y; cxyz_r[4 * i + 2] = s.center.z;
        cxyz_r[4 * i + 3] = s.radius;
        rgb_mat[4 * i] = s.sphere.surfaceColor.x; rgb_mat[4 * i + 1] = s.sphere.surfaceColor.y;
        rgb_mat[4 * i + 2] = s.sphere.surfaceColor.z; rgb_mat[4 * i + 3] = (float)s.sphere.materialType;
    }
}

// Build with the reference's builders, as main() does (main.cpp:791-843).  Returns the node count the
// reference reports (totalNodes / totalKdNodes), seconds in *secs.  The scene vector is restored to
// objId order first, so repeated builds see the same input main() would.
int ref_build(int accType, double* secs)
{
    g_scene = sceneFixed;
    g_root = std::make_shared<Node>();
    g_acc = accType;
    g_total_nodes = 0;
    std::streambuf* old = std::cout.rdbuf(nullptr);
    auto t0 = std::chrono::steady_clock::now();
    switch (accType) {
    case BVH:
        g_root = constructBVHNew(g_scene, 0, (int)g_scene.size(), &g_total_nodes);
        break;
    case LBVH: {
        std::vector<std::shared_ptr<Node>> nodes_unused;
        // constructLBVHTree keeps its node count local and only prints it; recount below.
        g_root = constructLBVHTree(g_scene, g_root, nodes_unused);
        break;
    }
    case KDTREE:
        reset_kd_globals();
        constructKDTreeNew(g_scene, 80, 1, 0.5f, 1, -1);       // main.cpp:816-820
        g_total_nodes = totalKdNodes;
        break;
    default:
        break;
    }
    auto t1 = std::chrono::steady_clock::now();
    std::cout.rdbuf(old);
    if (secs) *secs = std::chrono::duration<double>(t1 - t0).count();
    if (accType == LBVH) {
        std::vector<DumpRec> recs;
        dump_preorder(g_root, recs, nullptr);
        g_total_nodes = (int)recs.size();
    }
    return g_total_nodes;
}

// Pre-order dump of the pointer tree: records of 10 x 32-bit words (DumpRec), and the objIds the leaves
// hold in DFS order.  Returns the number of nodes; fills at most `cap` records / `cap_objs` ids.
int ref_bvh_dump(void* recs_out, int cap, int* leaf_objs_out, int cap_objs, int* n_leaf_objs)
{
    if (!g_root || (g_acc != BVH && g_acc != LBVH)) return -1;
    std::vector<DumpRec> recs;
    std::vector<int> objs;
    dump_preorder(g_root, recs, &objs);
    int n = (int)recs.size();
    if (recs_out) memcpy(recs_out, recs.data(), sizeof(DumpRec) * std::min(n, cap));
    if (leaf_objs_out) memcpy(leaf_objs_out, objs.data(), sizeof(int) * std::min((int)objs.size(), cap_objs));
    if (n_leaf_objs) *n_leaf_objs = (int)objs.size();
    return n;
}

// The caller's vector after the in-place reorder: scene[i].objId.
void ref_scene_order(int* obj_ids)
{
    for (size_t i = 0; i < g_scene.size(); ++i) obj_ids[i] = g_scene[i].objId;
}

// KD dump: nextFreeNode nodes of 3 words {split|onePrimitive|offset, flags|nPrims|aboveChild, nPrimitivesTest},
// the primitive index list and the tree bounds.
int ref_kd_dump(int32_t* nodes3, int cap, int* prim_indices, int cap_idx, int* n_idx, float* bounds6)
{
    if (g_acc != KDTREE || !::nodes) return -1;
    static_assert(sizeof(KdAccelNode) == 12, "KdAccelNode is 12 bytes");
    int n = nextFreeNode;
    if (nodes3) memcpy(nodes3, ::nodes, 12 * (size_t)std::min(n, cap));
    if (prim_indices) memcpy(prim_indices, kdtreePrimitiveIndices.data(),
                             sizeof(int) * std::min((int)kdtreePrimitiveIndices.size(), cap_idx));
    if (n_idx) *n_idx = (int)kdtreePrimitiveIndices.size();
    if (bounds6) {
        bounds6[0] = ::bounds.min.x; bounds6[1] = ::bounds.min.y; bounds6[2] = ::bounds.min.z;
        bounds6[3] = ::bounds.max.x; bounds6[4] = ::bounds.max.y; bounds6[5] = ::bounds.max.z;
    }
    return n;
}

// Closest-hit probe with the semantics of castRay's candidate loops (main.cpp:343-358 for BVH/LBVH,
// :376-386 for NONE) and kdtreeIntersect for KDTREE (any-hit: hit_id = 1/-1, t = 0).  Calls the
// reference's own boxIntersect / raySphereIntersect / kdtreeIntersect.
void ref_trace(const float* o, const float* d, int nrays, int accType, int* hit_id, float* t_out,
               long long* n_candidates)
{
    long long cand = 0;
    for (int r = 0; r < nrays; ++r) {
        Vec3f ro(o[3 * r], o[3 * r + 1], o[3 * r + 2]), rd(d[3 * r], d[3 * r + 1], d[3 * r + 2]);
        float tnear = INFINITY; int hit = -1;
        float t0, t1;
        if (accType == BVH || accType == LBVH) {
            std::vector<int> boxes;
            boxIntersect(ro, rd, g_root, boxes);
            cand += (long long)boxes.size();
            for (int box : boxes) {
                t0 = INFINITY, t1 = INFINITY;
                if (sceneFixed[box].sphere.raySphereIntersect(ro, rd, t0, t1)) {
                    if (t0 < 0) t0 = t1;
                    if (t0 < tnear) { tnear = t0; hit = sceneFixed[box].sphere.id; }
                }
            }
        } else if (accType == KDTREE) {
            hit = kdtreeIntersect(ro, rd) ? 1 : -1; tnear = 0;
        } else {
            for (unsigned i = 0; i < sceneFixed.size(); ++i) {
                t0 = INFINITY, t1 = INFINITY;
                if (sceneFixed[i].sphere.raySphereIntersect(ro, rd, t0, t1)) {
                    if (t0 < 0) t0 = t1;
                    if (t0 < tnear) { tnear = t0; hit = sceneFixed[i].sphere.id; }
                }
            }
            cand += (long long)sceneFixed.size();
        }
        hit_id[r] = hit;
        t_out[r] = tnear;
    }
    if (n_candidates) *n_candidates = cand;
}

// Brute-force closest hit over triangles with the reference's OWN class Triangle (main.cpp:107-216) and its compiled
// rayTriangleIntersect (the geometric branch; the reference never instantiates the class itself), reduced like the NONE loop
// (main.cpp:376-386: strict <, first candidate wins). tris9: m x {v0,v1,v2}. Also returns, per ray, how many triangles it hit.
void ref_triangle_trace(const float* tris9, int m, const float* o, const float* d, int nrays, int* hit_id, float* t_out, int* n_hits)
{
    std::vector<Triangle> tris;
    tris.reserve(m);
    for (int i = 0; i < m; ++i) {
        const float* v = tris9 + 9 * (size_t)i;
        tris.push_back(Triangle(Vec3f(v[0], v[1], v[2]), Vec3f(v[3], v[4], v[5]), Vec3f(v[6], v[7], v[8]), Vec3f(0.8f, 0.7f, 0.f)));
    }
    for (int r = 0; r < nrays; ++r) {
        Vec3f ro(o[3 * r], o[3 * r + 1], o[3 * r + 2]), rd(d[3 * r], d[3 * r + 1], d[3 * r + 2]);
        float tnear = INFINITY; int hit = -1, cnt = 0;
        for (int i = 0; i < m; ++i) {
            float t = INFINITY;
            if (tris[i].rayTriangleIntersect(ro, rd, t)) {
                ++cnt;
                if (t < tnear) { tnear = t; hit = i; }
            }
        }
        hit_id[r] = hit; t_out[r] = tnear;
        if (n_hits) n_hits[r] = cnt;
    }
}

// castRay on caller-supplied rays (depth 1), colours out.
void ref_cast(const float* o, const float* d, int nrays, int accType, float* rgb)
{
    Settings s; s.dataStructure = (AccType)accType;
    for (int r = 0; r < nrays; ++r) {
        Vec3f c = castRay(Vec3f(o[3 * r], o[3 * r + 1], o[3 * r + 2]), Vec3f(d[3 * r], d[3 * r + 1], d[3 * r + 2]),
                          g_spheres_unused, g_lights, g_scene, g_root, 1, s);
        rgb[3 * r] = c.x; rgb[3 * r + 1] = c.y; rgb[3 * r + 2] = c.z;
    }
}

// Rows [y0,y1) of render() (main.cpp:541-566) + write_into_file's quantisation (:521-523).
// rgb8: (y1-y0)*W*3 bytes; accum (optional): float RGB sums before the divide; dirs (optional):
// the primary ray directions, W*spp*3 floats per row; hit (optional): hit ids via ref_trace semantics are
// NOT computed here (use ref_trace on dirs).  Returns seconds spent in the ray loop.
double ref_render_rows(int width, int height, int spp, int accType, int y0, int y1,
                       uint8_t* rgb8, float* accum, float* dirs)
{
    Settings settings;
    settings.width = width; settings.height = height; settings.aa_samples = spp;
    settings.dataStructure = (AccType)accType;
    std::uniform_real_distribution<double> distribution(0.0, 1.0);
    std::mt19937 generator;
    generator.discard(4ull * (unsigned long long)y0 * width * spp);     // 2 doubles x 2 draws per sample
    float invWidth = 1 / float(settings.width), invHeight = 1 / float(settings.height);
    float fov = 30, aspectratio = settings.width / float(settings.height);
    float angle = tan(M_PI * 0.5 * fov / 180.);
    auto t0c = std::chrono::steady_clock::now();
    size_t k = 0, kd = 0;
    for (unsigned y = (unsigned)y0; y < (unsigned)y1; ++y) {
        for (unsigned x = 0; x < settings.width; ++x, ++k) {
            Vec3f sampled_pixel(0, 0, 0);
            for (unsigned sample = 0; sample < settings.aa_samples; ++sample) {
                float xx = (2 * ((x + distribution(generator)) * invWidth) - 1) * angle * aspectratio;
                float yy = (1 - 2 * ((y + distribution(generator)) * invHeight)) * angle;
                Vec3f raydir(xx, yy, -1);
                raydir.normalize();
                if (dirs) { dirs[kd++] = raydir.x; dirs[kd++] = raydir.y; dirs[kd++] = raydir.z; }
                sampled_pixel += castRay(Vec3f(0), raydir, g_spheres_unused, g_lights, g_scene, g_root, 1, settings);
            }
            if (accum) { accum[3 * k] = sampled_pixel.x; accum[3 * k + 1] = sampled_pixel.y; accum[3 * k + 2] = sampled_pixel.z; }
            if (rgb8) {
                rgb8[3 * k]     = (unsigned char)(std::min(float(1), sampled_pixel.x / settings.aa_samples) * 255);
                rgb8[3 * k + 1] = (unsigned char)(std::min(float(1), sampled_pixel.y / settings.aa_samples) * 255);
                rgb8[3 * k + 2] = (unsigned char)(std::min(float(1), sampled_pixel.z / settings.aa_samples) * 255);
            }
        }
    }
    auto t1c = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1c - t0c).count();
}

// First n doubles of random_double()'s stream (main.cpp:503-508), from a fresh default-seeded generator.
void ref_jitter(double* out, int n)
{
    std::uniform_real_distribution<double> distribution(0.0, 1.0);
    std::mt19937 generator;
    for (int i = 0; i < n; ++i) out[i] = distribution(generator);
}

unsigned ref_morton3D(float x, float y, float z) { return morton3D(x, y, z); }
unsigned ref_expandBits(unsigned v) { return expandBits(v); }
long long ref_sphere_tests() { return spheres_intersections_counter; }
void ref_reset_counters() { spheres_intersections_counter = 0; }

} // extern "C"
