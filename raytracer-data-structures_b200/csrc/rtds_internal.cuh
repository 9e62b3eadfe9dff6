// rtds_internal.cuh — shared declarations of librtds.so (device layouts, context, helpers).
// All device code in this library is compiled for sm_100a only, with -fmad=false (the reference is x86-64
// SSE2 code without FMA contraction; see DESIGN.md "Exact float semantics").
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <functional>
#include "../../include/rtds.h"

// ---------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------
void rtds_set_error(const char* fmt, ...);

#define RTDS_CUDA(call)                                                                            \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            rtds_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));  \
            return RTDS_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

#define RTDS_TRY(call)                                                                             \
    do {                                                                                           \
        int s_ = (call);                                                                           \
        if (s_ != RTDS_OK) return s_;                                                              \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// device layouts
// ---------------------------------------------------------------------------------------------------

// Traversal node: one 64-byte record per INTERIOR node holding both children's AABBs, so one visit is
// four coalescable 16-byte loads and tests two boxes (accelerators.h:131-156 is a 160-byte pointer node).
//   q0 = {lmin.x, lmin.y, lmin.z, lmax.x}   q1 = {lmax.y, lmax.z, rmin.x, rmin.y}
//   q2 = {rmin.z, rmax.x, rmax.y, rmax.z}   q3 = {left, right, axis, parent} (ints)
// Child reference: >= 0 interior node index, < 0 leaf: ~leafpos (position in leaf order).
struct __align__(16) Node64 {
    float lmin[3], lmax[3];
    float rmin[3], rmax[3];
    int   left, right;
    int   axis;        // split axis of this node (exported as LinearBVHNode::axis)
    int   parent;      // interior parent index, -1 for the root
};
static_assert(sizeof(Node64) == 64, "Node64 must be 64 bytes");

// 4-wide traversal node (`wide` option): the binary tree collapsed two levels at a time. wide[i] exists for every interior node i
// at EVEN depth and holds the boxes of i's up to four grandchildren (a child that is a leaf stands for itself), plane by plane
// for the four slots, and their references: >= 0 the grandchild's own index (it is at even depth again), < 0 a leaf (~leafpos),
// WIDE_EMPTY for an unused slot (its box is inverted: min = +inf, max = -inf, no ray enters it). Interior tests only steer
// (DESIGN.md section 7), so any conservative tree over the same leaves returns the same hits.
struct __align__(16) Wide4 {
    float x0[4], x1[4], y0[4], y1[4], z0[4], z1[4];
    int   ref[4];
    int   pad[4];
};
static_assert(sizeof(Wide4) == 128, "Wide4 must be 128 bytes");
constexpr int WIDE_EMPTY = (int)0x80000000;

struct DeviceBvh {
    int      capacity = 0;         // primitives the arrays below were allocated for
    int      n_prims = 0;          // leaves
    int      n_internal = 0;       // n_prims - 1 (0 when n_prims == 1)
    int      root_ref = 0;         // 0 (interior node 0) or ~0 when the tree is a single leaf
    Node64*  nodes = nullptr;      // [n_internal]
    float4*  leaf_sph = nullptr;   // [n_prims] {cx,cy,cz,r} in leaf order (radius2 = r*r is formed at the test, accelerators.h:71)
    float4*  leaf_tri = nullptr;   // [3*n_prims] v0,v1,v2 in leaf order when the scene holds triangles (extension)
    int      tri_capacity = 0;
    int      prim_type = 0;        // 0 spheres, 1 triangles
    int*     prim_order = nullptr; // [n_prims] leafpos -> objId
    int*     leaf_parent = nullptr;// [n_prims] interior parent of each leaf (bit 31 set: right child)
    float    root_box[6] = {0, 0, 0, 0, 0, 0};
    int      tie_by_objid = 0;     // 1: equal-t candidates resolve to the lower objId (NONE order,
                                   // main.cpp:376-386); 0: to the lower leaf position (DFS order of
                                   // boxIntersect, accelerators.h:668-690)
    int      leaf_box_prim = 0;    // 1: sphere leaves and every leaf's box is exactly c -/+ r of its sphere (main.cpp:686-688);
                                   // 0: triangles, or a median-split tree with dropped ranges (a leaf then carries the box of
                                   // its whole range, accelerators.h:321-327)
    int      max_depth = 0;
    Wide4*   wide = nullptr;       // [wide_capacity] indexed like nodes; valid entries at even depth (wide option), else unused
    int      wide_capacity = 0;
    bool     wide_valid = false;
    bool     valid = false;
};

struct DeviceKd {
    int           n_nodes = 0;        // nextFreeNode
    int           total_nodes = 0;    // totalKdNodes as the reference counts them
    int           n_idx = 0;
    rtds_kd_node* nodes = nullptr;
    int*          prim_idx = nullptr;
    float         bounds[6] = {0, 0, 0, 0, 0, 0};
    bool          valid = false;
};

struct RtdsLight { float c[3]; float radius; float le[3]; };
#define RTDS_MAX_LIGHTS 8

// Frame shared by the ranks of a multi-GPU render: lives in the owner's (rank 0's) memory; peers map it (CUDA IPC
// across processes, peer access inside one process) and their render kernels store into it directly.
struct SharedFrame {
    uint8_t*  frame = nullptr;      // height x width x 3, then one 128-byte flag line per rank
    volatile uint32_t* flags = nullptr;
    size_t    bytes = 0;
    int       width = 0, height = 0, world = 0, rank = 0;
    bool      owner = false;
    bool      ipc_mapped = false;   // opened with cudaIpcOpenMemHandle (close with cudaIpcCloseMemHandle)
    uint32_t  seq = 0;
};

#define RTDS_MAX_BANDS 8

// One frame of a device-buffer / shared-frame render as ONE cudaGraphLaunch (frame_graph option): direction kernel and tree
// prefetch side by side -> counters cleared -> render kernel -> [flag raised -> owner waits for every rank's flag] -> counters to
// pinned memory. Built explicitly (no stream capture); kernel arguments are refreshed with cudaGraphExecKernelNodeSetParams only
// when they changed, the graph is rebuilt only when its shape (which nodes, which render kernel) changes.
struct FrameGraph {
    cudaGraph_t     graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaGraphNode_t k_dirs = nullptr, k_pf = nullptr, k_render = nullptr, k_signal = nullptr, k_wait = nullptr;
    const void*     fn_render = nullptr;
    bool            has_dirs = false, has_pf = false, has_signal = false, has_wait = false;
    unsigned        grid_dirs = 0, grid_render = 0;
    std::vector<unsigned char> last_dirs, last_render, last_pf;     // argument bytes of the last launch (skip SetParams when equal)
    uint64_t        launches = 0, rebuilds = 0;
};

// Tuning / test switches of one context. Defaults come from the RTDS_* environment variables, read ONCE in rtds_create
// (never on a frame's path); rtds_set_option() changes them afterwards (tests, A/B tools).
struct RtdsOptions {
    int block_order = 2;     // RTDS_BLOCK_ORDER  0 quadrant-major, 1 row-major, 2 quadrant-major from the centre rows outwards
    int strip = 0;           // RTDS_STRIP        1: fused jitter + render strip kernel (measured slower; parity-tested)
    int bands = 6;           // RTDS_BANDS        row bands of a host-buffer render (each band's download overlaps the rendering of the next ones).
                             //                   Measured end to end on the bench frame (round 2 final kernels): 3 1.898, 4 1.897-1.913, 5 1.888, 6 1.837-1.850, 7 1.868, 8 1.883 ms
    int band_ratio = 100;    // RTDS_BAND_RATIO   each band's share of the one before it, percent
    int packet = 1;          // RTDS_PACKET       0: never the packet kernels
    int wavefront = 1;       // RTDS_WAVEFRONT    frames with shadow rays (aa_samples % 4 == 0) as two kernels (primary packets; shadow rays + shading per sample):
                             //                   0 never, 1 the library times both forms on the first frames of a geometry and keeps the faster, 2 always
    int hull = 1;            // RTDS_HULL         0: interior boxes tested per ray instead of once per packet
    int zerocopy = 0;        // RTDS_ZEROCOPY     1: store the frame straight into pinned host memory (measured slower)
    int trace_frame = 0;     // RTDS_TRACE_FRAME  1: rtds_frame stage timeline on stderr, 2: + per-band events
    int median_small = 0;    // RTDS_MEDIAN_SMALL test hook: sequential/parallel switch-over of the median split (0 = 64)
    int median_coop = 0;     // RTDS_MEDIAN_COOP  ranges above this use the cooperative kernel (0 = 65536)
    int median_debug = 0;    // RTDS_MEDIAN_DEBUG
    int wide = 0;            // RTDS_WIDE         one-ray-per-thread primary rays walk the 4-wide collapse of the tree (built behind every BVH build).
                             //                   Measured: node visits halve, kernels -1 ... -8 %: opt-in
    int node_preorder = 0;   // RTDS_NODE_ORDER=preorder: renumber the nodes in DFS pre-order after the build
    int l2_prefetch = 0;     // RTDS_L2_PREFETCH  stream the tree into L2 on a side stream while the directions are generated
    int lpt_cap = 3;         // RTDS_LPT_CAP      at most lpt_cap x (number of SMs) blocks of a launch count as heavy (and go first, sorted by cost)
    int lpt_bin = 8;         // RTDS_LPT_BIN      a block counts as heavy when its cost is >= lpt_bin / 32 of the frame's largest block cost
    int lpt_split = 128;     // RTDS_LPT_SPLIT    with a learned order in use (lpt): that many of the heaviest 16 x 8 tiles of the packet kernel are rendered by
                             //                   render_heavy_kernel instead, one ray per thread (4 blocks per tile): a straggler's chain of leaf tests is cut in four. 0 = off
    int lpt = 1;             // RTDS_LPT          the previous frame's heaviest blocks are launched first (and, lpt_split, the very heaviest rendered ray-per-thread):
                             //                   1 = the library times the schedules per frame geometry and keeps the fastest, 2 = always, 0 = never
    int frame_graph = 0;     // RTDS_FRAME_GRAPH  device-buffer / shared-frame renders: one CUDA graph launch per frame (measured: no gain)
};
typedef int RtdsOptions::*RtdsOptionField;
struct RtdsOptionName { const char* name; const char* env; RtdsOptionField field; };
extern const RtdsOptionName g_rtds_option_names[];
extern const int g_rtds_n_option_names;

struct rtds_ctx {
    RtdsOptions  opt;
    int          device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // device->host copies of finished row bands
    cudaEvent_t  ev_band = nullptr;
    cudaStream_t band_streams[RTDS_MAX_BANDS] = {};   // descending priority, created on first use (host-buffer renders)
    cudaEvent_t  ev_ready = nullptr, ev_bands[RTDS_MAX_BANDS] = {};
    cudaEvent_t  ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    cudaEvent_t  trace_ev[8] = {};        // rtds_frame's stage timeline (trace_frame option), created on this context's device
    int          sm_count = 148;

    // scene (objId-indexed)
    int     n = 0;
    int     sph_capacity = 0;
    float4* d_sph = nullptr;     // {cx,cy,cz,r}
    float4* d_mat = nullptr;     // {r,g,b,(float)material}
    bool    has_materials = false;   // any primitive with a non-DIFFUSE_AND_GLOSSY material
    bool    materials_pending = false;   // the material upload / flag kernel is still in flight on the copy stream
    int     n_lights = 0;
    RtdsLight lights[RTDS_MAX_LIGHTS];

    // triangles (extension): objId-indexed v0,v1,v2 as 3 float4 per triangle; prim_type selects the primitive table
    int     prim_type = 0;           // 0 spheres (d_sph), 1 triangles (d_tris)
    int     tri_capacity = 0;
    float4* d_tris = nullptr;

    DeviceBvh bvh;               // last BVH/LBVH build
    int       bvh_acc = -1;      // acc type that produced `bvh`
    int       bvh_mode = 0;
    DeviceKd  kd;

    // sorted Morton keys of the last TRUE LBVH build (export / tests)
    uint64_t* d_keys_sorted = nullptr;
    int       n_keys = 0;
    int       keys_capacity = 0;

    // jitter: MT19937 state snapshots (one per RTDS_MT_SNAP_EVERY regenerations) and the expanded words
    uint32_t* d_mt_snap = nullptr;   // [n_snap][624]
    int       n_snap = 0;
    uint32_t* d_jitter = nullptr;    // tempered words
    size_t    jitter_cap_words = 0;
    uint64_t  jitter_first_word = 0; // stream index of d_jitter[0]
    size_t    jitter_n_words = 0;
    float*    d_dirs = nullptr;      // primary ray directions of the last frame, 3 floats per sample
    size_t    dirs_cap_floats = 0;
    uint64_t  dirs_key[6] = {0, 0, 0, 0, 0, 0};   // what d_dirs was generated for (no_jitter_regen reuse)
    bool      dirs_valid = false;
    bool      dirs_pending = false;  // generated ahead of the render on jit_stream (rtds_frame); ev_dirs marks completion
    int       dirs_pending_launches = 0;
    cudaStream_t jit_stream = nullptr;
    cudaEvent_t  ev_dirs = nullptr;

    // scratch
    void*   d_scratch = nullptr;
    size_t  scratch_bytes = 0;
    void*   d_sort_ws = nullptr;     // onesweep workspace (histograms, tile counters, look-back status)
    size_t  sort_ws_bytes = 0;
    uint8_t* d_frame = nullptr;      // rgb8 staging
    size_t   frame_bytes = 0;
    int*     d_hit = nullptr;  size_t hit_bytes = 0;
    float*   d_accum = nullptr; size_t accum_bytes = 0;
    unsigned long long* d_counters = nullptr;  // render counters [8]
    unsigned long long* h_counters = nullptr;  // pinned [16]: [0..7] render counters, [8] material flag, [9..11] LBVH root box, [12] depth
    uint8_t* h_pinned = nullptr;     // pinned host staging for D2H of frames
    SharedFrame shared;
    FrameGraph  fg;
    char*       d_wave = nullptr;         // wavefront form: per-sample primary hits (tnear, leaf)
    size_t      wave_bytes = 0;
    // lpt option: per-block costs of the last frame and the launch order derived from them, band by band (render.cu: block_order_kernel)
    unsigned*   d_block_cost = nullptr;
    int*        d_block_order = nullptr;
    int*        d_heavy_list = nullptr;   // [LPT_SPLIT_MAX] tiles handed to render_heavy_kernel next frame (-1: unused)
    unsigned char* d_block_skip = nullptr; // [block_cap] 1 = the packet kernel leaves this tile to render_heavy_kernel
    int         block_cap = 0;
    uint64_t    block_key[5] = {0, 0, 0, 0, 0};   // geometry + kernel the buffers belong to
    bool        block_order_valid = false;
    bool        lpt_active = false;       // between lpt_frame_begin and lpt_frame_end of the current frame
    int         lpt_phase = 0;            // frames of this geometry timed so far, cycling through the modes (0 launch order, 1 learned order,
                                          // 2 learned order + heaviest tiles ray-per-thread); LPT_TRIAL_FRAMES = decided
    int         lpt_choice = 1;           // the fastest mode for this geometry once decided
    float       lpt_ms[3] = {0.f, 0.f, 0.f};   // best kernel time seen per mode during the trial
    bool        lpt_split_ok = false;     // this frame's kernel and banding allow the ray-per-thread hand-over (plain packet kernel, one band)
    int         lpt_last_mode = 0;        // what the current frame actually ran (mode 1 / 2 need an order from an earlier frame)
    // wavefront = 1: which form of a frame with shadow rays is faster, measured once per frame geometry (render.cu, wave_choose)
    uint64_t    wave_key[4] = {0, 0, 0, 0};
    int         wave_phase = 0;
    bool        wave_use = true, wave_trial = false, wave_this_frame = false;
    float       wave_ms[2] = {0.f, 0.f};
    cudaEvent_t ev_order_go = nullptr, ev_order_done = nullptr;
    cudaStream_t pf_stream = nullptr;   // tree prefetch into L2 beside the direction kernel (l2_prefetch option, non-graph path)
    cudaEvent_t  ev_pf0 = nullptr, ev_pf1 = nullptr;
    size_t   pinned_bytes = 0;
};

// The builders see primitives through this view: a centre (split / Morton key) and an AABB.
//   spheres   : centre = c, box = c -/+ r in float (main.cpp:686-688)
//   triangles : box = min/max of the vertices, centre = (min + max) * 0.5f          (extension)
struct PrimView { int type; const float4* sph; const float4* tri; };
inline PrimView rtds_prim_view(const rtds_ctx* c) { return PrimView{c->prim_type, c->d_sph, c->d_tris}; }
int rtds_alloc_bvh_for(rtds_ctx* ctx, DeviceBvh& b, int n_prims);

int rtds_finish_materials(rtds_ctx* ctx);
int rtds_ensure_scratch(rtds_ctx* ctx, size_t bytes);
template <typename T> int rtds_realloc(T** p, size_t* cap_bytes, size_t need_bytes);

// ---------------------------------------------------------------------------------------------------
// builders / kernels (one translation unit each)
// ---------------------------------------------------------------------------------------------------
void rtds_free_bvh(DeviceBvh& b);
int  rtds_alloc_bvh(DeviceBvh& b, int n_prims);

// lbvh.cu — K1 bounds, K2 Morton, K4 Karras emission, K5 atomic refit
int rtds_build_lbvh_true(rtds_ctx* ctx, const rtds_build_params* p, rtds_build_stats* st);
int rtds_morton30_device(rtds_ctx* ctx, const float* h_xyz, int n, uint32_t* h_codes);
int rtds_bvh_preorder_export(rtds_ctx* ctx, rtds_linear_bvh_node* h_nodes, int cap_nodes, int* n_nodes,
                             int* h_prim_order, int cap_prims, int* n_prims);

// sort.cu — K3 onesweep LSD radix sort of (key, value) pairs; keys 32 or 64 bit
int rtds_onesweep_sort_u32(rtds_ctx* ctx, uint32_t* d_keys, uint32_t* d_vals, uint32_t* d_keys_tmp,
                           uint32_t* d_vals_tmp, int n, int key_bits, int* launches);
int rtds_onesweep_sort_u64(rtds_ctx* ctx, uint64_t* d_keys, uint32_t* d_vals, uint64_t* d_keys_tmp,
                           uint32_t* d_vals_tmp, int n, int key_bits, int* launches);

// median.cu — K6 median-split BVH, bit-exact with constructBVHNew incl. libstdc++ partition/nth_element
int rtds_build_median(rtds_ctx* ctx, int n_use, rtds_build_stats* st);

// sah.cu — K7 binned SAH BVH
int rtds_build_sah(rtds_ctx* ctx, const rtds_build_params* p, rtds_build_stats* st);

// Zero fill by a KERNEL (api.cu). cudaMemsetAsync may be executed by a copy engine; on the rtds_frame path that queues
// the build's clears behind the 17 MB material upload still running on the copy stream (measured: the build stalled until
// the upload had finished). p 4-byte aligned, bytes a multiple of 4.
int rtds_zero_async(void* p, size_t bytes, cudaStream_t s, int* launches = nullptr);

// trace_frame option (profiling aid, api.cu): timing events of rtds_frame's device timeline live in the context
// (rtds_ctx::trace_ev, [0] == nullptr when off); a failing trace-only call never fails the frame
#define RTDS_TRACE_RECORD(ctx, i, stream) do { if ((ctx)->trace_ev[0]) (void)cudaEventRecord((ctx)->trace_ev[i], (stream)); } while (0)

// kd.cu — K8 KD-tree SAH build
int rtds_build_kd(rtds_ctx* ctx, const rtds_build_params* p, rtds_build_stats* st);
void rtds_free_kd(DeviceKd& k);

// render.cu — K10 render/trace, K11 MT19937 jitter stream
int rtds_render_impl(rtds_ctx* ctx, int acc, const rtds_render_params* p, uint8_t* d_rgb_rows, int* d_hit,
                     float* d_accum, rtds_render_stats* st, const std::function<int(int, int, cudaEvent_t)>* on_band = nullptr,
                     bool global_rows = false);
int rtds_shared_frame_signal_wait(rtds_ctx* ctx, uint32_t seq, int* launches);
int rtds_trace_impl(rtds_ctx* ctx, int acc, int exact, const float* h_o, const float* h_d, int nrays,
                    int* h_hit, float* h_t, rtds_render_stats* st);
int rtds_jitter_prepare(rtds_ctx* ctx, uint64_t first_word, size_t n_words, int* launches);
int rtds_prefetch_dirs(rtds_ctx* ctx, const rtds_render_params* p);
int rtds_ensure_band_streams(rtds_ctx* ctx);

// ---------------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
// order-preserving float <-> uint mapping for atomicMin/atomicMax on floats
__device__ __forceinline__ unsigned f2ord(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

__device__ __forceinline__ void prim_fetch(const PrimView& pv, int i, float c[3], float mn[3], float mx[3])
{
    if (pv.type == 0) {
        float4 s = __ldg(pv.sph + i);
        c[0] = s.x; c[1] = s.y; c[2] = s.z;
        mn[0] = s.x - s.w; mn[1] = s.y - s.w; mn[2] = s.z - s.w;
        mx[0] = s.x + s.w; mx[1] = s.y + s.w; mx[2] = s.z + s.w;
    } else {
        float4 a = __ldg(pv.tri + 3 * (size_t)i), b = __ldg(pv.tri + 3 * (size_t)i + 1), d = __ldg(pv.tri + 3 * (size_t)i + 2);
        mn[0] = fminf(fminf(a.x, b.x), d.x); mn[1] = fminf(fminf(a.y, b.y), d.y); mn[2] = fminf(fminf(a.z, b.z), d.z);
        mx[0] = fmaxf(fmaxf(a.x, b.x), d.x); mx[1] = fmaxf(fmaxf(a.y, b.y), d.y); mx[2] = fmaxf(fmaxf(a.z, b.z), d.z);
        c[0] = (mn[0] + mx[0]) * 0.5f; c[1] = (mn[1] + mx[1]) * 0.5f; c[2] = (mn[2] + mx[2]) * 0.5f;
    }
}
// leaf payload in leaf order: spheres {c, r^2}; triangles v0,v1,v2
__device__ __forceinline__ void prim_store_leaf(const PrimView& pv, int prim, int leafpos, float4* leaf_sph, float4* leaf_tri)
{
    if (pv.type == 0) {
        float4 s = __ldg(pv.sph + prim);
        leaf_sph[leafpos] = s;
    } else {
        for (int k = 0; k < 3; ++k) leaf_tri[3 * (size_t)leafpos + k] = __ldg(pv.tri + 3 * (size_t)prim + k);
    }
}
#endif
