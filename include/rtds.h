/* rtds.h — C ABI of librtds.so: the B200-native (sm_100a) acceleration-structure build + ray
 * traversal/intersection path of Alhajras/Raytracer-Data-structures.
 *
 * The reference has no FFI: its hot path is a set of free functions and globals in one translation unit.
 * The seam this ABI replaces is the set of calls main()/castRay() make into accelerators.h (citations are
 * relative to /root/reference/project/raytracer/):
 *
 *   rtds_set_spheres / rtds_set_lights   <- createScene_new() filling `scene` + global `sceneFixed`
 *                                           (main.cpp:599-721, records accelerators.h:54-100,173-188) and the
 *                                           light list of main.cpp:775,788
 *   rtds_build(BVH)                      <- constructBVHNew        (accelerators.h:246-337, call main.cpp:800)
 *   rtds_build(LBVH)                     <- constructLBVHTree      (accelerators.h:570-586, call main.cpp:832)
 *                                           + expandBits/morton3D  (accelerators.h:374-394)
 *   rtds_build(KDTREE)                   <- constructKDTreeNew     (accelerators.h:951-988, call main.cpp:816)
 *   rtds_export_bvh                      <- LinearBVHNode / flattenBVHTree (accelerators.h:231-244; declared,
 *                                           never filled by the reference: its intended flattened export)
 *   rtds_export_kd                       <- KdAccelNode[] + kdtreePrimitiveIndices (accelerators.h:715-773)
 *   rtds_trace                           <- boxIntersect + candidate loop (accelerators.h:668-690,
 *                                           main.cpp:343-358), NONE loop (main.cpp:376-386),
 *                                           kdtreeIntersect (accelerators.h:997-1086)
 *   rtds_render                          <- render() + castRay() + write_into_file's quantisation
 *                                           (main.cpp:541-566, 291-500, 516-528)
 *
 * A per-ray FFI is meaningless for a GPU, so the boundary is cut at scene -> build -> render(frame).
 * Conventions: plain C, no torch types; every call returns 0 on success or a negative rtds_status;
 * rtds_last_error() gives the message of the last failure on the calling thread; the caller owns every
 * host buffer; every call is synchronous on return.  One context drives one GPU; multi-GPU runs use one
 * context per GPU (one process per GPU under torchrun, or several contexts in one process) and the
 * rank/world fields of rtds_render_params.  There is no CPU fallback: without a usable sm_100 device
 * rtds_create fails.
 */
#ifndef RTDS_H
#define RTDS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rtds_ctx rtds_ctx;

typedef enum {
    RTDS_OK               = 0,
    RTDS_ERR_INVALID      = -1,  /* bad argument */
    RTDS_ERR_CUDA         = -2,  /* CUDA runtime failure (message has the cudaError string) */
    RTDS_ERR_NO_DEVICE    = -3,  /* no usable GPU: the library has no CPU path */
    RTDS_ERR_NO_SCENE     = -4,
    RTDS_ERR_NOT_BUILT    = -5,
    RTDS_ERR_DEGENERATE   = -6,  /* input on which the reference builder recurses forever
                                    (std::partition returns endIndex, accelerators.h:311-330) */
    RTDS_ERR_CAPACITY     = -7,  /* caller buffer too small */
    RTDS_ERR_UNSUPPORTED  = -8
} rtds_status;

/* enum AccType of accelerators.h:21, same values. UNIFORM_GRID has no case in the reference and falls
 * into the NONE brute-force loop (main.cpp:373-387); it does the same here. */
typedef enum { RTDS_BVH = 0, RTDS_KDTREE = 1, RTDS_UNIFORM_GRID = 2, RTDS_LBVH = 3, RTDS_NONE = 4 } rtds_acc_type;

/* enum MaterialType of accelerators.h:24. */
typedef enum { RTDS_DIFFUSE_AND_GLOSSY = 0, RTDS_REFLECTION_AND_REFRACTION = 1, RTDS_REFLECTION = 2 } rtds_material;

typedef enum {
    RTDS_MODE_COMPAT = 0, /* what the reference code does: BVH = median split (constructBVHNew), bit-exact
                             topology and primitive order; LBVH = the same split over objects [0,n-1)
                             (the reference drops the last object, accelerators.h:583) */
    RTDS_MODE_TRUE   = 1, /* LBVH as the reference intends it (accelerators.h:371,568): Morton codes over
                             scene-normalised centres, onesweep radix sort, Karras hierarchy, atomic refit */
    RTDS_MODE_SAH    = 2  /* binned-SAH BVH (extension; the reference has no SAH BVH) */
} rtds_build_mode;

typedef struct {
    int   mode;                 /* rtds_build_mode */
    int   morton_bits;          /* 30 (expandBits/morton3D, accelerators.h:374-394) or 63; 0 = 30 */
    int   morton_ref_norm;      /* 1: normalise as (centre+30)/1000 like accelerators.h:577; 0: scene bounds */
    /* KD parameters. 0 = the reference's EFFECTIVE values: the globals of accelerators.h:767-770
       (maxPrims 5, isectCost 80, traversalCost 1, emptyBonus (char)0.5f = 0), not main.cpp:816-820's arguments */
    int   kd_isect_cost;
    int   kd_traversal_cost;
    float kd_empty_bonus;
    int   kd_max_prims;
    int   kd_max_depth;         /* <=0: round(8 + 1.3*floor(log2 n)), accelerators.h:958 */
    int   sah_bins;             /* 0 = 16 */
    int   reserved[6];
} rtds_build_params;

typedef struct {
    int      n_prims;           /* primitives the structure indexes (n-1 for compat LBVH) */
    int      total_nodes;       /* what the reference prints: 2n-1 for BVH/LBVH, totalKdNodes for KD */
    int      alloc_nodes;       /* KD: nextFreeNode (array length); BVH: == total_nodes */
    int      max_depth;
    int      kernel_launches;   /* CUDA kernels launched by this build */
    float    ms;                /* device time of the build, CUDA events on the build stream */
    int      reserved[8];
} rtds_build_stats;

/* 32-byte flattened BVH node, the layout of accelerators.h:231-240 (PBRT's LinearBVHNode):
 * depth-first pre-order, first child at index+1, second child at secondChildOffset. */
typedef struct {
    float    bmin[3], bmax[3];
    int32_t  offset;            /* leaf: primitivesOffset (index into prim_order); interior: secondChildOffset */
    uint16_t nPrimitives;       /* 0 -> interior */
    uint8_t  axis;              /* interior: longest axis used for the split */
    uint8_t  pad;
} rtds_linear_bvh_node;

/* 12-byte KD node, the layout of accelerators.h:715-742 (w0: split | onePrimitive | primitiveIndicesOffset,
 * w1: flags/nPrims/aboveChild packed as value<<2 | axis-or-3, w2: nPrimitivesTest). */
typedef struct { uint32_t w0, w1, w2; } rtds_kd_node;

typedef struct {
    int      width, height, aa_samples;     /* settings.h:9,10,14 */
    float    fov;                           /* <=0 -> 30, what render() hard-codes (main.cpp:545) */
    float    bg[3];                         /* all 0 -> (0.6,0.8,1), castRay's sky (main.cpp:312,318) */
    float    bias;                          /* <=0 -> 1e-4 (main.cpp:404) */
    int      max_depth;                     /* <=0 -> 2 (main.cpp:311) */
    int      shadows;                       /* 0 = reference behaviour (trace_more is a stub, main.cpp:235-245) */
    int      exact;                         /* 1 = reference traversal: every node whose slab test passes, no
                                               pruning; 0 = ordered short-stack traversal with conservative
                                               t-pruning (same hits by construction) */
    int      rank, world;                   /* this context renders row tiles t with t % world == rank */
    int      tile_rows;                     /* <=0 -> 8 */
    uint64_t jitter_offset;                 /* first sample's index in random_double()'s stream (main.cpp:503-508) */
    int      no_jitter_regen;               /* 1: reuse the jitter words generated by the previous call */
    int      kd_closest;                    /* KDTREE only. 0 = reference behaviour: any-hit, hit pixels black
                                               (main.cpp:362-372). 1 = closest-hit KD traversal + castRay's shading
                                               (extension: PBRT's KdTreeAccel::Intersect, whose any-hit sibling
                                               accelerators.h:997-1086 is a port of) */
    int      tri_geometric;                 /* triangle scenes only. 0 = Moeller-Trumbore (the reference's MOLLER_TRUMBORE branch,
                                               main.cpp:138-162, which it never compiles); 1 = the test the reference DOES compile:
                                               the geometric branch of Triangle::rayTriangleIntersect (main.cpp:163-215), bug for
                                               bug (its plane distance is right for rays from the origin only) */
    int      reserved[5];
} rtds_render_params;

typedef struct {
    uint64_t rays;              /* primary + shadow + secondary rays traced */
    uint64_t primary_rays, shadow_rays, secondary_rays;
    uint64_t node_tests;        /* slab tests executed (distinct child boxes fetched and tested) */
    uint64_t prim_tests;        /* ray-primitive tests executed */
    uint64_t node_visits;       /* interior nodes popped */
    float    ms_kernel;         /* device time of the render kernel alone, CUDA events on its stream */
    float    ms_total;          /* device time of the whole call's GPU work (jitter + render + quantise) */
    int      kernel_launches;
    int      rows;              /* rows this rank rendered */
    int      reserved[6];
} rtds_render_stats;

const char* rtds_last_error(void);
const char* rtds_version(void);

/* device: CUDA ordinal. Fails with RTDS_ERR_NO_DEVICE when there is no GPU (no CPU fallback). */
int rtds_create(rtds_ctx** out, int device);
int rtds_destroy(rtds_ctx* ctx);

/* Tuning / test switches of a context (the counterpart of editing settings.h and recompiling, settings.h:7-17, for the
 * knobs the reference does not have). Every switch has an RTDS_<NAME> environment variable that sets its default; the
 * environment is read ONCE, in rtds_create, never while rendering. Names: block_order, strip, bands, band_ratio, packet,
 * wavefront, hull, zerocopy, trace_frame, median_small, median_coop, median_debug, node_preorder, wide, l2_prefetch,
 * frame_graph, lpt, lpt_split, lpt_bin, lpt_cap. Unknown name -> RTDS_ERR_INVALID. No switch changes a result: they select between
 * kernels that are tested to produce identical frames. */
int rtds_set_option(rtds_ctx* ctx, const char* name, int value);
int rtds_get_option(rtds_ctx* ctx, const char* name, int* value);

/* Primitive table, indexed by objId = position (sceneFixed, main.cpp:81). cxyz_r: n x {cx,cy,cz,radius};
 * rgb_mat: n x {r,g,b,(float)rtds_material}, may be NULL (-> (0.8,0.7,0), diffuse: main.cpp:689).
 * AABBs are centre -/+ radius in float, as main.cpp:686-688. Host pointers. */
int rtds_set_spheres(rtds_ctx* ctx, const float* cxyz_r, const float* rgb_mat, int n);
/* The same with both tables already in THIS context's device memory (DEVICE pointers; copied, not adopted; rgb_mat required).
 * For hosts that assemble the scene on the device - e.g. a multi-GPU run in which every rank uploads 1/N of the tables and the
 * ranks exchange the parts over NVLink (NCCL all-gather) instead of pushing N full copies through the host's memory system. */
int rtds_set_spheres_device(rtds_ctx* ctx, const float* d_cxyz_r, const float* d_rgb_mat, int n);
/* Triangle primitives (extension; the reference never instantiates class Triangle, main.cpp:107-216). */
int rtds_set_triangles(rtds_ctx* ctx, const float* v0v1v2, const float* rgb_mat, int n);
/* m x {cx,cy,cz,radius,r,g,b}; default is main.cpp:775's single light (0,3,30), emission (1,1,1). */
int rtds_set_lights(rtds_ctx* ctx, const float* cxyz_r_rgb, int m);

int rtds_build(rtds_ctx* ctx, int acc_type, const rtds_build_params* params, rtds_build_stats* stats);

/* Flattened exports. cap_* are capacities in elements; n_* receive the element counts. */
int rtds_export_bvh(rtds_ctx* ctx, rtds_linear_bvh_node* nodes, int cap_nodes, int* n_nodes,
                    int* prim_order, int cap_prims, int* n_prims);
int rtds_export_kd(rtds_ctx* ctx, rtds_kd_node* nodes, int cap_nodes, int* n_nodes,
                   int* prim_indices, int cap_idx, int* n_idx, float* bounds6);
/* Sorted Morton keys of the last TRUE-mode LBVH build (64-bit container) and the sorted primitive ids. */
int rtds_export_morton(rtds_ctx* ctx, uint64_t* keys, int* prim_ids, int cap, int* n);

/* Parity probe: closest hit of arbitrary rays through the built structure (acc_type NONE: brute force).
 * hit_obj: objId or -1; t: tnear (INFINITY on miss). KDTREE is any-hit like the reference: hit_obj = 1/-1, t = 0;
 * with RTDS_TRACE_KD_CLOSEST or-ed into `exact` it is the closest-hit KD traversal (objId, tnear) of kd_closest.
 * exact as in rtds_render_params. Host pointers, o/d are nrays x 3. */
#define RTDS_TRACE_KD_CLOSEST 2
#define RTDS_TRACE_TRI_GEOMETRIC 4      /* or-ed into `exact`: triangles are tested like rtds_render_params.tri_geometric = 1 */
int rtds_trace(rtds_ctx* ctx, int acc_type, int exact, const float* o_xyz, const float* d_xyz, int nrays,
               int* hit_obj, float* t, rtds_render_stats* stats);

/* One frame. rgb: host buffer of width*height*3 bytes, written for the rows this rank owns (all rows when
 * world == 1). hit_obj (optional, may be NULL): width*height ints, objId hit by the LAST sample of each owned
 * pixel (-1 = sky). accum (optional): width*height*3 floats, the per-pixel sums before the divide. */
int rtds_render(rtds_ctx* ctx, int acc_type, const rtds_render_params* params, uint8_t* rgb,
                int* hit_obj, float* accum, rtds_render_stats* stats);
/* The whole per-run sequence of main() (main.cpp:751,800/816/832,808) in ONE synchronous call: rtds_set_spheres ->
 * rtds_build -> rtds_render with the stages overlapped on the device (the material table is still uploading while the
 * structure is built). Equivalent to the three calls; bst / rst may be NULL. */
int rtds_frame(rtds_ctx* ctx, const float* cxyz_r, const float* rgb_mat, int n, int acc_type, const rtds_build_params* bp,
               const rtds_render_params* rp, uint8_t* rgb, rtds_build_stats* bst, rtds_render_stats* rst);
/* Optional head start for hosts that upload and build with separate calls: begins generating the ray directions of the frame
 * `params` describes (render()'s jitter + ray set-up, main.cpp:554-557) on a side stream and returns at once; the next
 * rtds_render* call with the same parameters uses them instead of generating its own. rtds_frame does this internally. */
int rtds_prepare_frame(rtds_ctx* ctx, const rtds_render_params* params);
/* Same, result left on the device: d_rgb_rows is a DEVICE pointer receiving this rank's rows compactly
 * (local row-tile j = global tile j*world + rank), rtds_rows_for_rank(...)*width*3 bytes. Used by the
 * multi-GPU framebuffer gather (NCCL) and by resident-input timing. */
int rtds_render_device(rtds_ctx* ctx, int acc_type, const rtds_render_params* params, uint8_t* d_rgb_rows,
                       rtds_render_stats* stats);
int rtds_rows_for_rank(int height, int tile_rows, int rank, int world);

/* ---- Multi-GPU frame assembly without a collective -------------------------------------------------------------
 * The scene and structure are replicated, the image is split into interleaved scanline tiles (rank / world of
 * rtds_render_params). Instead of rendering into per-rank buffers and gathering them (rtds_render_device + NCCL),
 * every rank's render kernel stores its tiles STRAIGHT into one frame that lives in rank 0's memory: peer stores
 * over NVLink / NVSwitch, 16 bytes at a time, overlapped with the tracing; a per-rank completion flag follows the
 * tiles and rank 0 waits for all flags on its own stream. Replaces render()'s single `image` buffer (main.cpp:543).
 *   rank 0            : rtds_shared_frame_create (ipc_handle_out: 64 bytes to pass to the other processes)
 *   rank r, other proc: rtds_shared_frame_open(handle, ..., r)        (cudaIpcOpenMemHandle)
 *   rank r, same proc : rtds_shared_frame_attach(ctx_r, ctx_0, r)     (peer access; one context per GPU)
 *   every rank, frame : rtds_render_shared(ctx, acc, params, seq)     seq != 0 and different from the previous frame's;
 *                       (or rtds_frame_shared: upload + build + rtds_render_shared in one call)
 *                       synchronous; on rank 0 it returns when EVERY rank's tiles of frame `seq` have landed
 *   rank 0            : rtds_shared_frame_read (device -> host) or rtds_shared_frame_ptr (device pointer, H*W*3 bytes)
 * The caller keeps frames apart: no rank may start frame k+1 before rank 0 has consumed frame k (a frame-loop barrier). */
#define RTDS_IPC_HANDLE_BYTES 64
int rtds_shared_frame_create(rtds_ctx* owner, int width, int height, int world, void* ipc_handle_out);
int rtds_shared_frame_open(rtds_ctx* ctx, const void* ipc_handle, int width, int height, int world, int rank);
int rtds_shared_frame_attach(rtds_ctx* ctx, rtds_ctx* owner, int rank);
int rtds_render_shared(rtds_ctx* ctx, int acc_type, const rtds_render_params* params, uint32_t frame_seq, rtds_render_stats* stats);
/* rtds_frame for one rank of a multi-GPU run: rtds_set_spheres -> rtds_build -> rtds_render_shared in one synchronous call,
 * stages overlapped as in rtds_frame (what main() does per run, main.cpp:751,800/816/832,808, on every GPU). */
int rtds_frame_shared(rtds_ctx* ctx, const float* cxyz_r, const float* rgb_mat, int n, int acc_type, const rtds_build_params* bp,
                      const rtds_render_params* rp, uint32_t frame_seq, rtds_build_stats* bst, rtds_render_stats* rst);
int rtds_shared_frame_ptr(rtds_ctx* ctx, void** d_frame);
int rtds_shared_frame_read(rtds_ctx* owner, uint8_t* rgb);
int rtds_shared_frame_close(rtds_ctx* ctx);

/* First n doubles of random_double()'s stream starting at double index `first` (main.cpp:503-508), generated
 * by the device MT19937 kernels. Host pointer. */
int rtds_jitter_stream(rtds_ctx* ctx, uint64_t first, int n, double* out);
/* Morton KAT probe: codes of n points with the reference's formula (accelerators.h:385-394), on the device. */
int rtds_morton30(rtds_ctx* ctx, const float* xyz, int n, uint32_t* codes);

#ifdef __cplusplus
}
#endif
#endif /* RTDS_H */
