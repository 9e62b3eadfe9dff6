// api.cu — the C ABI of include/rtds.h: context, scene upload, build dispatch, exports, render/trace glue.
// There is no CPU path in this library: every entry point that computes anything launches CUDA kernels,
// and rtds_create fails when no sm_100 device is usable.
#include "rtds_internal.cuh"
#include <algorithm>
#include <chrono>
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[1024] = "";

// name <-> environment variable <-> field of RtdsOptions (rtds_set_option / rtds_get_option; defaults read once in rtds_create)
const RtdsOptionName g_rtds_option_names[] = {
    {"block_order", "RTDS_BLOCK_ORDER", &RtdsOptions::block_order}, {"strip", "RTDS_STRIP", &RtdsOptions::strip},
    {"bands", "RTDS_BANDS", &RtdsOptions::bands}, {"band_ratio", "RTDS_BAND_RATIO", &RtdsOptions::band_ratio},
    {"packet", "RTDS_PACKET", &RtdsOptions::packet}, {"wavefront", "RTDS_WAVEFRONT", &RtdsOptions::wavefront}, {"hull", "RTDS_HULL", &RtdsOptions::hull},
    {"zerocopy", "RTDS_ZEROCOPY", &RtdsOptions::zerocopy}, {"trace_frame", "RTDS_TRACE_FRAME", &RtdsOptions::trace_frame},
    {"median_small", "RTDS_MEDIAN_SMALL", &RtdsOptions::median_small}, {"median_coop", "RTDS_MEDIAN_COOP", &RtdsOptions::median_coop},
    {"median_debug", "RTDS_MEDIAN_DEBUG", &RtdsOptions::median_debug}, {"node_preorder", "RTDS_NODE_PREORDER", &RtdsOptions::node_preorder}, {"wide", "RTDS_WIDE", &RtdsOptions::wide},
    {"l2_prefetch", "RTDS_L2_PREFETCH", &RtdsOptions::l2_prefetch}, {"frame_graph", "RTDS_FRAME_GRAPH", &RtdsOptions::frame_graph},
    {"lpt", "RTDS_LPT", &RtdsOptions::lpt}, {"lpt_split", "RTDS_LPT_SPLIT", &RtdsOptions::lpt_split}, {"lpt_bin", "RTDS_LPT_BIN", &RtdsOptions::lpt_bin}, {"lpt_cap", "RTDS_LPT_CAP", &RtdsOptions::lpt_cap},
};
const int g_rtds_n_option_names = (int)(sizeof g_rtds_option_names / sizeof g_rtds_option_names[0]);

static __global__ void zero_words_kernel(uint32_t* __restrict__ p, size_t n_words)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride) p[i] = 0u;
}
static __global__ void zero_vec_kernel(uint4* __restrict__ p, size_t n_vec)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) p[i] = make_uint4(0u, 0u, 0u, 0u);
}
int rtds_zero_async(void* p, size_t bytes, cudaStream_t s, int* launches)
{
    if (!bytes) return RTDS_OK;
    if (((uintptr_t)p & 3) || (bytes & 3)) { rtds_set_error("zero fill: unaligned"); return RTDS_ERR_INVALID; }
    if (((uintptr_t)p & 15) == 0 && bytes >= 4096) {
        const size_t n_vec = bytes / 16;
        const unsigned blocks = (unsigned)std::min<size_t>((n_vec + 255) / 256, 148 * 8);
        zero_vec_kernel<<<blocks, 256, 0, s>>>((uint4*)p, n_vec);
        const size_t tail = bytes - n_vec * 16;
        if (tail) zero_words_kernel<<<1, 32, 0, s>>>((uint32_t*)((char*)p + n_vec * 16), tail / 4);
        if (launches) *launches += tail ? 2 : 1;
    } else {
        const size_t n_words = bytes / 4;
        const unsigned blocks = (unsigned)std::min<size_t>((n_words + 255) / 256, 148 * 8);
        zero_words_kernel<<<blocks, 256, 0, s>>>((uint32_t*)p, n_words);
        if (launches) *launches += 1;
    }
    RTDS_CUDA(cudaGetLastError());
    return RTDS_OK;
}

void rtds_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int rtds_ensure_scratch(rtds_ctx* ctx, size_t bytes)
{
    if (ctx->scratch_bytes >= bytes) return RTDS_OK;
    if (ctx->d_scratch) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->d_scratch); }
    ctx->d_scratch = nullptr; ctx->scratch_bytes = 0;
    size_t want = (bytes + bytes / 8 + 8192) & ~(size_t)4095;
    RTDS_CUDA(cudaMalloc(&ctx->d_scratch, want));
    ctx->scratch_bytes = want;
    return RTDS_OK;
}

template <typename T>
static int ensure_buf(T** p, size_t* cap, size_t need)
{
    if (*cap >= need) return RTDS_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    RTDS_CUDA(cudaMalloc((void**)p, need));
    *cap = need;
    return RTDS_OK;
}

void rtds_free_bvh(DeviceBvh& b)
{
    if (b.nodes) cudaFree(b.nodes);
    if (b.leaf_sph) cudaFree(b.leaf_sph);
    if (b.prim_order) cudaFree(b.prim_order);
    if (b.leaf_parent) cudaFree(b.leaf_parent);
    if (b.leaf_tri) cudaFree(b.leaf_tri);
    if (b.wide) cudaFree(b.wide);
    b = DeviceBvh();
}

int rtds_alloc_bvh_for(rtds_ctx* ctx, DeviceBvh& b, int n_prims)
{
    RTDS_TRY(rtds_alloc_bvh(b, n_prims));
    b.prim_type = ctx->prim_type;
    if (ctx->prim_type == 1 && b.tri_capacity < n_prims) {
        if (b.leaf_tri) cudaFree(b.leaf_tri);
        b.leaf_tri = nullptr; b.tri_capacity = 0;
        RTDS_CUDA(cudaMalloc(&b.leaf_tri, sizeof(float4) * 3 * (size_t)n_prims));
        b.tri_capacity = n_prims;
    }
    return RTDS_OK;
}

// (Re)uses the arrays when they are large enough: cudaMalloc/cudaFree are device-wide synchronisation points and
// cost milliseconds once peer access is enabled (NCCL), so a rebuild of the same scene must not touch them.
int rtds_alloc_bvh(DeviceBvh& b, int n_prims)
{
    b.valid = false;
    b.wide_valid = false;
    if (b.capacity >= n_prims && b.nodes) return RTDS_OK;
    rtds_free_bvh(b);
    size_t ni = n_prims > 1 ? (size_t)(n_prims - 1) : 1;
    RTDS_CUDA(cudaMalloc(&b.nodes, sizeof(Node64) * ni));
    RTDS_CUDA(cudaMalloc(&b.leaf_sph, sizeof(float4) * (size_t)n_prims));
    RTDS_CUDA(cudaMalloc(&b.prim_order, sizeof(int) * (size_t)n_prims));
    RTDS_CUDA(cudaMalloc(&b.leaf_parent, sizeof(int) * (size_t)n_prims));
    b.capacity = n_prims;
    return RTDS_OK;
}

void rtds_free_kd(DeviceKd& k)
{
    if (k.nodes) cudaFree(k.nodes);
    if (k.prim_idx) cudaFree(k.prim_idx);
    k = DeviceKd();
}

int rtds_jitter_stream_impl(rtds_ctx* ctx, uint64_t first, int n, double* out);
static size_t shared_frame_bytes(int W, int H, int world, size_t* flags_off)
{
    const size_t fo = ((size_t)W * H * 3 + 255) & ~(size_t)255;
    if (flags_off) *flags_off = fo;
    return fo + 128 * (size_t)world;
}

static void shared_frame_drop(rtds_ctx* c)
{
    SharedFrame& f = c->shared;
    if (f.frame) {
        if (f.owner) cudaFree(f.frame);
        else if (f.ipc_mapped) cudaIpcCloseMemHandle(f.frame);
    }
    f = SharedFrame();
}


// does any primitive carry a material other than DIFFUSE_AND_GLOSSY? (selects the full castRay kernel)
__global__ void material_flag_kernel(const float4* __restrict__ mat, int n, int* flag)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool f = i < n && mat[i].w != 0.0f;
    if (__any_sync(0xffffffffu, f) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

extern "C" int rtds_destroy(rtds_ctx* c);
static int create_resources(rtds_ctx* c)
{
    RTDS_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    RTDS_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    RTDS_CUDA(cudaEventCreateWithFlags(&c->ev_band, cudaEventDisableTiming));
    RTDS_CUDA(cudaStreamCreateWithFlags(&c->jit_stream, cudaStreamNonBlocking));
    RTDS_CUDA(cudaEventCreateWithFlags(&c->ev_dirs, cudaEventDisableTiming));
    RTDS_CUDA(cudaEventCreateWithFlags(&c->ev_order_go, cudaEventDisableTiming));
    RTDS_CUDA(cudaEventCreateWithFlags(&c->ev_order_done, cudaEventDisableTiming));
    RTDS_CUDA(cudaStreamCreateWithFlags(&c->pf_stream, cudaStreamNonBlocking));
    RTDS_CUDA(cudaEventCreateWithFlags(&c->ev_pf0, cudaEventDisableTiming));
    RTDS_CUDA(cudaEventCreateWithFlags(&c->ev_pf1, cudaEventDisableTiming));
    RTDS_CUDA(cudaEventCreate(&c->ev0)); RTDS_CUDA(cudaEventCreate(&c->ev1));
    RTDS_CUDA(cudaEventCreate(&c->ev2)); RTDS_CUDA(cudaEventCreate(&c->ev3));
    RTDS_CUDA(cudaMalloc(&c->d_counters, sizeof(unsigned long long) * 8));
    RTDS_CUDA(cudaMallocHost(&c->h_counters, sizeof(unsigned long long) * 16));    // [8..15]: material flag read-back
    // main.cpp:775: Sphere light2(0, (0,3,30), 10, (1,1,1), 0, 0, emission (1,1,1))
    c->n_lights = 1;
    c->lights[0] = RtdsLight{{0.f, 3.f, 30.f}, 10.f, {1.f, 1.f, 1.f}};
    return RTDS_OK;
}

extern "C" {

const char* rtds_last_error(void) { return g_err; }
const char* rtds_version(void) { return "rtds-b200 0.1 (sm_100a)"; }

int rtds_create(rtds_ctx** out, int device)
{
    if (!out) { rtds_set_error("rtds_create: out is NULL"); return RTDS_ERR_INVALID; }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        rtds_set_error("rtds_create: no CUDA device (%s); librtds has no CPU path", e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
        return RTDS_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) { rtds_set_error("rtds_create: device %d of %d", device, count); return RTDS_ERR_INVALID; }
    RTDS_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RTDS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        rtds_set_error("rtds_create: device %d is sm_%d%d; this library holds sm_100a code only", device, prop.major, prop.minor);
        return RTDS_ERR_NO_DEVICE;
    }
    rtds_ctx* c = new rtds_ctx();
    c->device = device;
    // the RTDS_* switches are read here, once per context, never on a frame's path
    for (int i = 0; i < g_rtds_n_option_names; ++i)
        if (const char* e = getenv(g_rtds_option_names[i].env)) c->opt.*(g_rtds_option_names[i].field) = atoi(e);
    if (const char* e = getenv("RTDS_NODE_ORDER")) c->opt.node_preorder = !strcmp(e, "preorder");
    c->sm_count = prop.multiProcessorCount;
    const int rc = create_resources(c);
    if (rc != RTDS_OK) { rtds_destroy(c); return rc; }      // rtds_destroy copes with a partially built context
    *out = c;
    return RTDS_OK;
}

int rtds_destroy(rtds_ctx* c)
{
    if (!c) return RTDS_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    rtds_free_bvh(c->bvh);
    rtds_free_kd(c->kd);
    void* bufs[] = {c->d_sph, c->d_mat, c->d_tris, c->d_keys_sorted, c->d_mt_snap, c->d_jitter, c->d_dirs, c->d_wave, c->d_block_cost, c->d_block_order, c->d_heavy_list, c->d_block_skip, c->d_scratch,
                    c->d_sort_ws, c->d_frame, c->d_hit, c->d_accum, c->d_counters};
    for (void* b : bufs) if (b) cudaFree(b);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->fg.exec) cudaGraphExecDestroy(c->fg.exec);
    if (c->fg.graph) cudaGraphDestroy(c->fg.graph);
    if (c->pf_stream) cudaStreamDestroy(c->pf_stream);
    for (cudaEvent_t e : {c->ev0, c->ev1, c->ev2, c->ev3, c->ev_band, c->ev_dirs, c->ev_order_go, c->ev_order_done, c->ev_pf0, c->ev_pf1}) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    shared_frame_drop(c);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (int k = 0; k < RTDS_MAX_BANDS; ++k) if (c->band_streams[k]) cudaStreamDestroy(c->band_streams[k]);
    if (c->ev_ready) cudaEventDestroy(c->ev_ready);
    for (int k = 0; k < RTDS_MAX_BANDS; ++k) if (c->ev_bands[k]) cudaEventDestroy(c->ev_bands[k]);
    if (c->jit_stream) cudaStreamDestroy(c->jit_stream);
    for (auto& e : c->trace_ev) if (e) cudaEventDestroy(e);
    delete c;
    return RTDS_OK;
}

}  // extern "C" (internal helpers follow)

// async_mat: the material table goes up on the copy stream (with the material-flag kernel behind it) and is only
// waited for by finish_materials(); the sphere table is complete on return when !async_mat, otherwise the main stream
// is ordered behind it.
static int upload_spheres(rtds_ctx* c, const float* cxyz_r, const float* rgb_mat, int n, bool async_mat, cudaMemcpyKind kind = cudaMemcpyHostToDevice)
{
    if (!c || !cxyz_r || n <= 0) { rtds_set_error("set_spheres: bad arguments"); return RTDS_ERR_INVALID; }
    RTDS_CUDA(cudaSetDevice(c->device));
    RTDS_CUDA(cudaStreamSynchronize(c->stream));
    RTDS_CUDA(cudaStreamSynchronize(c->copy_stream));
    c->n = 0;
    c->prim_type = 0;
    c->bvh.valid = false; c->kd.valid = false; c->bvh_acc = -1;
    if (c->sph_capacity < n) {
        if (c->d_sph) cudaFree(c->d_sph);
        if (c->d_mat) cudaFree(c->d_mat);
        c->d_sph = nullptr; c->d_mat = nullptr; c->sph_capacity = 0;
        RTDS_CUDA(cudaMalloc(&c->d_sph, sizeof(float4) * (size_t)n));
        RTDS_CUDA(cudaMalloc(&c->d_mat, sizeof(float4) * (size_t)n));
        c->sph_capacity = n;
    }
    // async_mat: BOTH tables go up on the copy stream, spheres first, and the main stream only waits for the event
    // behind the sphere table. (Measured: with the sphere copy on the main stream, that stream's next commands - the
    // build's event record and first kernels - were ordered behind the copy engine's queue and did not start before the
    // 17 MB material copy had finished as well: build start 660 us instead of 345 us into the call.)
    const bool split = async_mat && rgb_mat;
    RTDS_CUDA(cudaMemcpyAsync(c->d_sph, cxyz_r, sizeof(float4) * (size_t)n, kind, split ? c->copy_stream : c->stream));
    c->has_materials = false;
    c->materials_pending = false;
    if (rgb_mat) {
        cudaStream_t ms = async_mat ? c->copy_stream : c->stream;
        if (async_mat) {
            RTDS_CUDA(cudaEventRecord(c->ev_band, c->copy_stream));
            RTDS_CUDA(cudaStreamWaitEvent(c->stream, c->ev_band, 0));
        }
        RTDS_CUDA(cudaMemcpyAsync(c->d_mat, rgb_mat, sizeof(float4) * (size_t)n, kind, ms));
        int* d_flag = (int*)(c->d_counters + 6);
        RTDS_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), ms));
        material_flag_kernel<<<(n + 255) / 256, 256, 0, ms>>>(c->d_mat, n, d_flag);
        // the flag comes back into pinned memory behind the kernel: finish_materials() only has to wait for the stream
        RTDS_CUDA(cudaMemcpyAsync(c->h_counters + 8, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ms));
        RTDS_TRACE_RECORD(c, 3, ms);
        c->materials_pending = true;
        if (!async_mat) RTDS_TRY(rtds_finish_materials(c));
    } else {
        std::vector<float> m((size_t)n * 4);
        for (int i = 0; i < n; ++i) { m[4 * i] = 0.8f; m[4 * i + 1] = 0.7f; m[4 * i + 2] = 0.0f; m[4 * i + 3] = 0.f; }  // main.cpp:689
        RTDS_CUDA(cudaMemcpyAsync(c->d_mat, m.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
        RTDS_CUDA(cudaStreamSynchronize(c->stream));
    }
    // split: no host wait here - the main stream is ordered behind the sphere table by the event, and rtds_frame (the
    // only async caller) does not return before rtds_finish_materials has synchronised the copy stream
    if (!split) RTDS_CUDA(cudaStreamSynchronize(c->stream));
    c->n = n;
    return RTDS_OK;
}

// error exit of the overlapped one-call forms: wait for everything they queued (uploads still reading the caller's host
// buffers on the copy stream, the direction kernel on the side stream) and drop the pending state
static void frame_abort(rtds_ctx* c)
{
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamSynchronize(c->jit_stream);
    cudaStreamSynchronize(c->stream);
    c->materials_pending = false;
    c->dirs_pending = false;
    (void)cudaGetLastError();
}

int rtds_finish_materials(rtds_ctx* c)
{
    if (!c->materials_pending) return RTDS_OK;
    RTDS_CUDA(cudaStreamSynchronize(c->copy_stream));
    RTDS_CUDA(cudaStreamSynchronize(c->stream));
    const int flag = *reinterpret_cast<volatile int*>(c->h_counters + 8);
    c->has_materials = flag != 0;
    c->materials_pending = false;
    return RTDS_OK;
}

extern "C" {

int rtds_set_spheres(rtds_ctx* c, const float* cxyz_r, const float* rgb_mat, int n)
{
    return upload_spheres(c, cxyz_r, rgb_mat, n, false);
}

int rtds_set_spheres_device(rtds_ctx* c, const float* d_cxyz_r, const float* d_rgb_mat, int n)
{
    if (!d_rgb_mat) { rtds_set_error("set_spheres_device: the material table is required"); return RTDS_ERR_INVALID; }
    return upload_spheres(c, d_cxyz_r, d_rgb_mat, n, false, cudaMemcpyDeviceToDevice);
}

// What main() does per run, in one synchronous call with the stages overlapped: sphere table up -> build, while the
// material table is still uploading on the copy stream -> render -> frame down.
int rtds_frame(rtds_ctx* c, const float* cxyz_r, const float* rgb_mat, int n, int acc, const rtds_build_params* bp,
               const rtds_render_params* rp, uint8_t* rgb, rtds_build_stats* bst, rtds_render_stats* rst)
{
    if (!c || !rp || !rgb) { rtds_set_error("frame: bad arguments"); return RTDS_ERR_INVALID; }
    // the frame's ray directions depend on the render parameters only: generate them on their own stream while the
    // scene uploads and the structure is built
    RTDS_CUDA(cudaSetDevice(c->device));
    // RTDS_TRACE_FRAME=1: host-clock timeline of the call's stages on stderr (profiling aid)
    const bool trace = c->opt.trace_frame != 0;
    if (trace && !c->trace_ev[0]) {
        for (auto& e : c->trace_ev) if (cudaEventCreate(&e) != cudaSuccess) e = nullptr;
        if (!c->trace_ev[7]) { for (auto& e : c->trace_ev) { if (e) cudaEventDestroy(e); e = nullptr; } (void)cudaGetLastError(); }
    }
    RTDS_TRACE_RECORD(c, 0, c->stream);
    const auto t0 = std::chrono::steady_clock::now();
    auto us = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); };
    // every call is synchronous on return, also a failing one: nothing may still read the caller's buffers or run on the side
    // streams after an error exit
    auto fail = [&](int rc) { frame_abort(c); return rc; };
    int rc = rtds_prefetch_dirs(c, rp);
    if (rc != RTDS_OK) return fail(rc);
    const double t_dirs = us();
    if ((rc = upload_spheres(c, cxyz_r, rgb_mat, n, true)) != RTDS_OK) return fail(rc);
    const double t_up = us();
    if ((rc = rtds_build(c, acc, bp, bst)) != RTDS_OK) return fail(rc);
    const double t_build = us();
    if ((rc = rtds_finish_materials(c)) != RTDS_OK) return fail(rc);
    const double t_mat = us();
    rc = rtds_render(c, acc, rp, rgb, nullptr, nullptr, rst);
    if (rc != RTDS_OK) return fail(rc);
    if (trace && rc == RTDS_OK && c->trace_ev[0]) {
        cudaDeviceSynchronize();
        float d0 = 0, d1 = 0, b0 = 0, b1 = 0, m1 = 0, r0 = 0, r1 = 0;
        cudaEventElapsedTime(&d0, c->trace_ev[0], c->trace_ev[1]); cudaEventElapsedTime(&d1, c->trace_ev[0], c->trace_ev[2]);
        cudaEventElapsedTime(&b0, c->trace_ev[0], c->trace_ev[4]); cudaEventElapsedTime(&b1, c->trace_ev[0], c->trace_ev[5]);
        cudaEventElapsedTime(&m1, c->trace_ev[0], c->trace_ev[3]);
        cudaEventElapsedTime(&r0, c->trace_ev[0], c->ev2); cudaEventElapsedTime(&r1, c->trace_ev[0], c->ev3);
        float c1 = 0;
        cudaEventElapsedTime(&c1, c->trace_ev[0], c->trace_ev[6]);
        if (c->ev_bands[0] && c->opt.trace_frame > 1) {
            fprintf(stderr, "[rtds_frame device] bands end at");
            for (int k = 0; k < RTDS_MAX_BANDS; ++k) { float e = 0; if (cudaEventElapsedTime(&e, c->trace_ev[0], c->ev_bands[k]) == cudaSuccess && e > 0) fprintf(stderr, " %.0f", 1e3 * e); }
            fprintf(stderr, " us\n");
            cudaGetLastError();
        }
        cudaGetLastError();      // an event of a stage that did not run on this path leaves an error behind
        fprintf(stderr, "[rtds_frame device] dirs %.0f..%.0f us | build %.0f..%.0f | materials up %.0f | render %.0f..%.0f | frame down %.0f\n",
                1e3 * d0, 1e3 * d1, 1e3 * b0, 1e3 * b1, 1e3 * m1, 1e3 * r0, 1e3 * r1, 1e3 * c1);
    }
    if (trace)
        fprintf(stderr, "[rtds_frame] dirs enqueued %.0f us | spheres up %.0f | build done %.0f | materials done %.0f | render+download done %.0f\n",
                t_dirs, t_up, t_build, t_mat, us());
    return rc;
}

int rtds_set_triangles(rtds_ctx* c, const float* v0v1v2, const float* rgb_mat, int n)
{
    if (!c || !v0v1v2 || n <= 0) { rtds_set_error("set_triangles: bad arguments"); return RTDS_ERR_INVALID; }
    RTDS_CUDA(cudaSetDevice(c->device));
    RTDS_CUDA(cudaStreamSynchronize(c->stream));
    RTDS_CUDA(cudaStreamSynchronize(c->copy_stream));      // an asynchronous material upload of a previous rtds_frame
    c->materials_pending = false;
    c->n = 0;
    c->bvh.valid = false; c->kd.valid = false; c->bvh_acc = -1;
    if (c->tri_capacity < n) {
        if (c->d_tris) cudaFree(c->d_tris);
        c->d_tris = nullptr; c->tri_capacity = 0;
        RTDS_CUDA(cudaMalloc(&c->d_tris, sizeof(float4) * 3 * (size_t)n));
        c->tri_capacity = n;
    }
    if (c->sph_capacity < n) {   // the material table (and an unused sphere table of the same length) ride along
        if (c->d_sph) cudaFree(c->d_sph);
        if (c->d_mat) cudaFree(c->d_mat);
        c->d_sph = nullptr; c->d_mat = nullptr; c->sph_capacity = 0;
        RTDS_CUDA(cudaMalloc(&c->d_sph, sizeof(float4) * (size_t)n));
        RTDS_CUDA(cudaMalloc(&c->d_mat, sizeof(float4) * (size_t)n));
        c->sph_capacity = n;
    }
    std::vector<float> padded((size_t)n * 12);     // 9 floats -> 3 x float4 (16-byte vertex loads on the device)
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            padded[12 * (size_t)i + 4 * k] = v0v1v2[9 * (size_t)i + 3 * k];
            padded[12 * (size_t)i + 4 * k + 1] = v0v1v2[9 * (size_t)i + 3 * k + 1];
            padded[12 * (size_t)i + 4 * k + 2] = v0v1v2[9 * (size_t)i + 3 * k + 2];
            padded[12 * (size_t)i + 4 * k + 3] = 0.f;
        }
    RTDS_CUDA(cudaMemcpyAsync(c->d_tris, padded.data(), sizeof(float4) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    std::vector<float> m((size_t)n * 4);
    c->has_materials = false;
    for (int i = 0; i < n; ++i) {
        if (rgb_mat) { for (int k = 0; k < 4; ++k) m[4 * (size_t)i + k] = rgb_mat[4 * (size_t)i + k]; if (rgb_mat[4 * (size_t)i + 3] != 0.f) c->has_materials = true; }
        else { m[4 * (size_t)i] = 0.8f; m[4 * (size_t)i + 1] = 0.7f; m[4 * (size_t)i + 2] = 0.f; m[4 * (size_t)i + 3] = 0.f; }
    }
    RTDS_CUDA(cudaMemcpyAsync(c->d_mat, m.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    RTDS_CUDA(cudaStreamSynchronize(c->stream));
    c->prim_type = 1;
    c->n = n;
    return RTDS_OK;
}

int rtds_set_lights(rtds_ctx* c, const float* l, int m)
{
    if (!c || m < 0 || m > RTDS_MAX_LIGHTS || (m > 0 && !l)) { rtds_set_error("set_lights: 0..%d lights", RTDS_MAX_LIGHTS); return RTDS_ERR_INVALID; }
    c->n_lights = m;
    for (int i = 0; i < m; ++i) {
        const float* p = l + 7 * i;
        c->lights[i] = RtdsLight{{p[0], p[1], p[2]}, p[3], {p[4], p[5], p[6]}};
    }
    return RTDS_OK;
}

int rtds_build(rtds_ctx* c, int acc, const rtds_build_params* p, rtds_build_stats* st)
{
    if (!c) { rtds_set_error("build: ctx is NULL"); return RTDS_ERR_INVALID; }
    if (c->n <= 0) { rtds_set_error("build: no scene"); return RTDS_ERR_NO_SCENE; }
    RTDS_CUDA(cudaSetDevice(c->device));
    if (st) memset(st, 0, sizeof *st);
    const int mode = p ? p->mode : RTDS_MODE_COMPAT;
    int rc;
    switch (acc) {
    case RTDS_BVH:
        if (mode == RTDS_MODE_SAH) rc = rtds_build_sah(c, p, st);
        else if (mode == RTDS_MODE_TRUE) rc = rtds_build_lbvh_true(c, p, st);
        else rc = rtds_build_median(c, c->n, st);
        break;
    case RTDS_LBVH:
        if (mode == RTDS_MODE_TRUE) rc = rtds_build_lbvh_true(c, p, st);
        else if (mode == RTDS_MODE_SAH) rc = rtds_build_sah(c, p, st);
        else {
            // accelerators.h:583: constructLBVHNew(objects, codes, 0, size()-1): the last object is dropped
            if (c->n < 2) { rtds_set_error("build: compat LBVH needs >= 2 objects (the reference drops the last one)"); return RTDS_ERR_INVALID; }
            if (c->prim_type != 0) { rtds_set_error("build: compat LBVH (drop the last object) is defined for the reference's sphere scenes only"); return RTDS_ERR_UNSUPPORTED; }
            rc = rtds_build_median(c, c->n - 1, st);
        }
        break;
    case RTDS_KDTREE:
        rc = rtds_build_kd(c, p, st);
        break;
    default:  // NONE / UNIFORM_GRID: nothing to build (main.cpp:838-842)
        if (st) { st->n_prims = c->n; }
        return RTDS_OK;
    }
    if (rc == RTDS_OK && (acc == RTDS_BVH || acc == RTDS_LBVH)) { c->bvh_acc = acc; c->bvh_mode = mode; }
    return rc;
}

int rtds_export_bvh(rtds_ctx* c, rtds_linear_bvh_node* nodes, int cap_nodes, int* n_nodes, int* prim_order, int cap_prims,
                    int* n_prims)
{
    if (!c) { rtds_set_error("export_bvh: ctx is NULL"); return RTDS_ERR_INVALID; }
    RTDS_CUDA(cudaSetDevice(c->device));
    return rtds_bvh_preorder_export(c, nodes, cap_nodes, n_nodes, prim_order, cap_prims, n_prims);
}

int rtds_export_kd(rtds_ctx* c, rtds_kd_node* nodes, int cap_nodes, int* n_nodes, int* prim_indices, int cap_idx, int* n_idx,
                   float* bounds6)
{
    if (!c) { rtds_set_error("export_kd: ctx is NULL"); return RTDS_ERR_INVALID; }
    if (!c->kd.valid) { rtds_set_error("export_kd: no KD-tree has been built"); return RTDS_ERR_NOT_BUILT; }
    RTDS_CUDA(cudaSetDevice(c->device));
    if (n_nodes) *n_nodes = c->kd.n_nodes;
    if (n_idx) *n_idx = c->kd.n_idx;
    if ((nodes && cap_nodes < c->kd.n_nodes) || (prim_indices && cap_idx < c->kd.n_idx)) {
        rtds_set_error("export_kd: need %d nodes / %d indices", c->kd.n_nodes, c->kd.n_idx);
        return RTDS_ERR_CAPACITY;
    }
    if (nodes) RTDS_CUDA(cudaMemcpy(nodes, c->kd.nodes, sizeof(rtds_kd_node) * (size_t)c->kd.n_nodes, cudaMemcpyDeviceToHost));
    if (prim_indices && c->kd.n_idx) RTDS_CUDA(cudaMemcpy(prim_indices, c->kd.prim_idx, sizeof(int) * (size_t)c->kd.n_idx, cudaMemcpyDeviceToHost));
    if (bounds6) memcpy(bounds6, c->kd.bounds, sizeof(float) * 6);
    return RTDS_OK;
}

int rtds_export_morton(rtds_ctx* c, uint64_t* keys, int* prim_ids, int cap, int* n)
{
    if (!c) { rtds_set_error("export_morton: ctx is NULL"); return RTDS_ERR_INVALID; }
    if (!c->bvh.valid || c->bvh_mode != RTDS_MODE_TRUE || !c->d_keys_sorted) {
        rtds_set_error("export_morton: the last build was not a TRUE-mode LBVH");
        return RTDS_ERR_NOT_BUILT;
    }
    RTDS_CUDA(cudaSetDevice(c->device));
    if (n) *n = c->bvh.n_prims;
    if (cap < c->bvh.n_prims) { rtds_set_error("export_morton: need %d", c->bvh.n_prims); return RTDS_ERR_CAPACITY; }
    if (keys) RTDS_CUDA(cudaMemcpy(keys, c->d_keys_sorted, sizeof(uint64_t) * (size_t)c->bvh.n_prims, cudaMemcpyDeviceToHost));
    if (prim_ids) RTDS_CUDA(cudaMemcpy(prim_ids, c->bvh.prim_order, sizeof(int) * (size_t)c->bvh.n_prims, cudaMemcpyDeviceToHost));
    return RTDS_OK;
}

int rtds_trace(rtds_ctx* c, int acc, int exact, const float* o, const float* d, int nrays, int* hit, float* t,
               rtds_render_stats* st)
{
    if (!c || !o || !d || !hit || !t || nrays < 0) { rtds_set_error("trace: bad arguments"); return RTDS_ERR_INVALID; }
    RTDS_CUDA(cudaSetDevice(c->device));
    return rtds_trace_impl(c, acc, exact, o, d, nrays, hit, t, st);
}

int rtds_rows_for_rank(int height, int tile_rows, int rank, int world)
{
    if (tile_rows <= 0) tile_rows = 8;
    if (world <= 0) world = 1;
    int tiles = (height + tile_rows - 1) / tile_rows, rows = 0;
    for (int t = rank; t < tiles; t += world) rows += (t * tile_rows + tile_rows <= height) ? tile_rows : height - t * tile_rows;
    return rows;
}

int rtds_render_device(rtds_ctx* c, int acc, const rtds_render_params* p, uint8_t* d_rgb_rows, rtds_render_stats* st)
{
    if (!c || !p || !d_rgb_rows) { rtds_set_error("render_device: bad arguments"); return RTDS_ERR_INVALID; }
    RTDS_CUDA(cudaSetDevice(c->device));
    RTDS_TRY(rtds_render_impl(c, acc, p, d_rgb_rows, nullptr, nullptr, st));
    RTDS_CUDA(cudaStreamSynchronize(c->stream));
    return RTDS_OK;
}

int rtds_render(rtds_ctx* c, int acc, const rtds_render_params* p, uint8_t* rgb, int* hit_obj, float* accum,
                rtds_render_stats* st)
{
    if (!c || !p || !rgb) { rtds_set_error("render: bad arguments"); return RTDS_ERR_INVALID; }
    RTDS_CUDA(cudaSetDevice(c->device));
    const int W = p->width, H = p->height;
    if (W <= 0 || H <= 0) { rtds_set_error("render: width/height must be positive"); return RTDS_ERR_INVALID; }
    const int world = p->world > 0 ? p->world : 1, tile_rows = p->tile_rows > 0 ? p->tile_rows : 8;
    const int rows = rtds_rows_for_rank(H, tile_rows, p->rank, world);
    const size_t px = (size_t)rows * W;
    RTDS_TRY(ensure_buf(&c->d_frame, &c->frame_bytes, px * 3 + 16));
    if (hit_obj) RTDS_TRY(ensure_buf(&c->d_hit, &c->hit_bytes, px * sizeof(int) + 16));
    if (accum) RTDS_TRY(ensure_buf(&c->d_accum, &c->accum_bytes, px * 3 * sizeof(float) + 16));
    cudaStream_t s = c->stream;
    // RTDS_ZEROCOPY=1 (experiment, MEASURED SLOWER, off by default): when the caller's frame is pinned host memory, the render
    // kernel stores its RGB8 tiles straight into it over PCIe (mapped memory) - no bands, no device->host copy afterwards.
    // Frame identical, but the 8-byte posted writes reach only ~17 GB/s and throttle the kernel: 1.09 -> 1.48 ms, e2e 2.34 vs
    // 2.12 ms with the banded copies.
    if (world == 1 && !hit_obj && !accum && c->opt.zerocopy == 1) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, rgb) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
            RTDS_TRY(rtds_render_impl(c, acc, p, (uint8_t*)at.devicePointer, nullptr, nullptr, st, nullptr, false));
            RTDS_CUDA(cudaStreamSynchronize(s));
            return RTDS_OK;
        }
        cudaGetLastError();
    }
    if (world == 1) {
        // the frame comes back band by band while later bands are still rendering
        std::function<int(int, int, cudaEvent_t)> on_band = [&](int r0, int r1, cudaEvent_t done) -> int {
            const size_t off = (size_t)r0 * W, cnt = (size_t)(r1 - r0) * W;
            RTDS_CUDA(cudaStreamWaitEvent(c->copy_stream, done, 0));
            RTDS_CUDA(cudaMemcpyAsync(rgb + off * 3, c->d_frame + off * 3, cnt * 3, cudaMemcpyDeviceToHost, c->copy_stream));
            if (hit_obj) RTDS_CUDA(cudaMemcpyAsync(hit_obj + off, c->d_hit + off, cnt * sizeof(int), cudaMemcpyDeviceToHost, c->copy_stream));
            if (accum) RTDS_CUDA(cudaMemcpyAsync(accum + off * 3, c->d_accum + off * 3, cnt * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream));
            return RTDS_OK;
        };
        RTDS_TRY(rtds_render_impl(c, acc, p, c->d_frame, hit_obj ? c->d_hit : nullptr, accum ? c->d_accum : nullptr, st, &on_band));
        RTDS_TRACE_RECORD(c, 6, c->copy_stream);
        RTDS_CUDA(cudaStreamSynchronize(s));
        RTDS_CUDA(cudaStreamSynchronize(c->copy_stream));
        return RTDS_OK;
    }
    RTDS_TRY(rtds_render_impl(c, acc, p, c->d_frame, hit_obj ? c->d_hit : nullptr, accum ? c->d_accum : nullptr, st));
    // device -> host: local row tile j is global tile j*world + rank. The RGB8 rows of all whole tiles go down as ONE strided
    // copy (source pitch = one tile, destination pitch = `world` tiles), so N ranks fill one host frame over their own PCIe
    // links side by side - the multi-GPU end-to-end path needs no GPU-side gather when the consumer is the host.
    {
        const size_t tile_bytes = (size_t)tile_rows * W * 3;
        int whole = 0;
        for (int t = p->rank; (t + 1) * tile_rows <= H; t += world) ++whole;
        if (whole > 0)
            RTDS_CUDA(cudaMemcpy2DAsync(rgb + (size_t)p->rank * tile_bytes, (size_t)world * tile_bytes, c->d_frame, tile_bytes, tile_bytes,
                                        (size_t)whole, cudaMemcpyDeviceToHost, s));
        int lrow = 0;
        for (int t = p->rank; t * tile_rows < H; t += world) {
            int r0 = t * tile_rows, nr = (r0 + tile_rows <= H) ? tile_rows : H - r0;
            size_t cnt = (size_t)nr * W;
            if (nr < tile_rows)    // the ragged last tile
                RTDS_CUDA(cudaMemcpyAsync(rgb + (size_t)r0 * W * 3, c->d_frame + (size_t)lrow * W * 3, cnt * 3, cudaMemcpyDeviceToHost, s));
            if (hit_obj) RTDS_CUDA(cudaMemcpyAsync(hit_obj + (size_t)r0 * W, c->d_hit + (size_t)lrow * W, cnt * sizeof(int), cudaMemcpyDeviceToHost, s));
            if (accum) RTDS_CUDA(cudaMemcpyAsync(accum + (size_t)r0 * W * 3, c->d_accum + (size_t)lrow * W * 3, cnt * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
            lrow += nr;
        }
    }
    RTDS_CUDA(cudaStreamSynchronize(s));
    return RTDS_OK;
}

// ---------------------------------------------------------------------------------------------------
// Multi-GPU frame assembly by direct peer stores (no collective): see include/rtds.h
// ---------------------------------------------------------------------------------------------------
int rtds_shared_frame_create(rtds_ctx* c, int width, int height, int world, void* ipc_handle_out)
{
    if (!c || width <= 0 || height <= 0 || world <= 0 || world > 32) { rtds_set_error("shared_frame_create: bad arguments"); return RTDS_ERR_INVALID; }
    RTDS_CUDA(cudaSetDevice(c->device));
    RTDS_CUDA(cudaStreamSynchronize(c->stream));
    shared_frame_drop(c);
    SharedFrame& f = c->shared;
    size_t fo = 0;
    f.bytes = shared_frame_bytes(width, height, world, &fo);
    RTDS_CUDA(cudaMalloc(&f.frame, f.bytes));
    RTDS_CUDA(cudaMemset(f.frame, 0, f.bytes));
    f.flags = reinterpret_cast<volatile uint32_t*>(f.frame + fo);
    f.width = width; f.height = height; f.world = world; f.rank = 0; f.owner = true;
    if (ipc_handle_out) {
        cudaIpcMemHandle_t h;
        RTDS_CUDA(cudaIpcGetMemHandle(&h, f.frame));
        static_assert(sizeof(h) == RTDS_IPC_HANDLE_BYTES, "IPC handle size");
        memcpy(ipc_handle_out, &h, sizeof h);
    }
    return RTDS_OK;
}

int rtds_shared_frame_open(rtds_ctx* c, const void* ipc_handle, int width, int height, int world, int rank)
{
    if (!c || !ipc_handle || width <= 0 || height <= 0 || rank <= 0 || rank >= world || world > 32) {
        rtds_set_error("shared_frame_open: bad arguments (rank 0 is the owner and calls rtds_shared_frame_create)");
        return RTDS_ERR_INVALID;
    }
    RTDS_CUDA(cudaSetDevice(c->device));
    RTDS_CUDA(cudaStreamSynchronize(c->stream));
    shared_frame_drop(c);
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof h);
    void* p = nullptr;
    RTDS_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    SharedFrame& f = c->shared;
    size_t fo = 0;
    f.bytes = shared_frame_bytes(width, height, world, &fo);
    f.frame = (uint8_t*)p;
    f.flags = reinterpret_cast<volatile uint32_t*>(f.frame + fo);
    f.width = width; f.height = height; f.world = world; f.rank = rank; f.owner = false; f.ipc_mapped = true;
    return RTDS_OK;
}

int rtds_shared_frame_attach(rtds_ctx* c, rtds_ctx* owner, int rank)
{
    if (!c || !owner || c == owner || !owner->shared.frame || !owner->shared.owner || rank <= 0 || rank >= owner->shared.world) {
        rtds_set_error("shared_frame_attach: bad arguments");
        return RTDS_ERR_INVALID;
    }
    RTDS_CUDA(cudaSetDevice(c->device));
    RTDS_CUDA(cudaStreamSynchronize(c->stream));
    shared_frame_drop(c);
    if (c->device != owner->device) {
        int can = 0;
        RTDS_CUDA(cudaDeviceCanAccessPeer(&can, c->device, owner->device));
        if (!can) { rtds_set_error("shared_frame_attach: device %d cannot access device %d", c->device, owner->device); return RTDS_ERR_UNSUPPORTED; }
        cudaError_t e = cudaDeviceEnablePeerAccess(owner->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) RTDS_CUDA(e);
        (void)cudaGetLastError();
    }
    c->shared = owner->shared;
    c->shared.owner = false; c->shared.ipc_mapped = false; c->shared.rank = rank;
    return RTDS_OK;
}

int rtds_render_shared(rtds_ctx* c, int acc, const rtds_render_params* p, uint32_t frame_seq, rtds_render_stats* st)
{
    if (!c || !p || frame_seq == 0) { rtds_set_error("render_shared: bad arguments (frame_seq must be non-zero)"); return RTDS_ERR_INVALID; }
    SharedFrame& f = c->shared;
    if (!f.frame) { rtds_set_error("render_shared: no shared frame (rtds_shared_frame_create / _open / _attach first)"); return RTDS_ERR_INVALID; }
    const int world = p->world > 0 ? p->world : 1;
    if (p->width != f.width || p->height != f.height || world != f.world || p->rank != f.rank) {
        rtds_set_error("render_shared: params (%dx%d rank %d/%d) do not match the shared frame (%dx%d rank %d/%d)", p->width, p->height,
                       p->rank, world, f.width, f.height, f.rank, f.world);
        return RTDS_ERR_INVALID;
    }
    RTDS_CUDA(cudaSetDevice(c->device));
    f.seq = frame_seq;
    rtds_render_stats local;
    RTDS_TRY(rtds_render_impl(c, acc, p, f.frame, nullptr, nullptr, st ? st : &local, nullptr, true));   // synchronous (reads the counters)
    const int status = (st ? st : &local)->reserved[0];      // frame_wait_kernel's verdict, read back with the counters
    if (f.owner && status != 0) {
        rtds_set_error("render_shared: timed out waiting for rank %d's tiles of frame %u", status - 1, frame_seq);
        return RTDS_ERR_CUDA;
    }
    return RTDS_OK;
}

// rtds_frame for one rank of a multi-GPU run: upload + build + rtds_render_shared with the stages overlapped exactly as in
// rtds_frame (ray directions on a side stream and the material table on the copy stream while the sphere table uploads and
// the structure is built).
int rtds_frame_shared(rtds_ctx* c, const float* cxyz_r, const float* rgb_mat, int n, int acc, const rtds_build_params* bp,
                      const rtds_render_params* rp, uint32_t frame_seq, rtds_build_stats* bst, rtds_render_stats* rst)
{
    if (!c || !rp) { rtds_set_error("frame_shared: bad arguments"); return RTDS_ERR_INVALID; }
    RTDS_CUDA(cudaSetDevice(c->device));
    int rc = rtds_prefetch_dirs(c, rp);
    if (rc == RTDS_OK) rc = upload_spheres(c, cxyz_r, rgb_mat, n, true);
    if (rc == RTDS_OK) rc = rtds_build(c, acc, bp, bst);
    if (rc == RTDS_OK) rc = rtds_finish_materials(c);
    if (rc == RTDS_OK) rc = rtds_render_shared(c, acc, rp, frame_seq, rst);
    if (rc != RTDS_OK) frame_abort(c);       // synchronous on return, also when failing
    return rc;
}

int rtds_shared_frame_ptr(rtds_ctx* c, void** d_frame)
{
    if (!c || !d_frame || !c->shared.frame) { rtds_set_error("shared_frame_ptr: no shared frame"); return RTDS_ERR_INVALID; }
    *d_frame = c->shared.frame;
    return RTDS_OK;
}

int rtds_shared_frame_read(rtds_ctx* c, uint8_t* rgb)
{
    if (!c || !rgb || !c->shared.frame) { rtds_set_error("shared_frame_read: no shared frame"); return RTDS_ERR_INVALID; }
    RTDS_CUDA(cudaSetDevice(c->device));
    RTDS_CUDA(cudaMemcpyAsync(rgb, c->shared.frame, (size_t)c->shared.width * c->shared.height * 3, cudaMemcpyDeviceToHost, c->stream));
    RTDS_CUDA(cudaStreamSynchronize(c->stream));
    return RTDS_OK;
}

int rtds_shared_frame_close(rtds_ctx* c)
{
    if (!c) return RTDS_ERR_INVALID;
    RTDS_CUDA(cudaSetDevice(c->device));
    RTDS_CUDA(cudaStreamSynchronize(c->stream));
    shared_frame_drop(c);
    return RTDS_OK;
}

int rtds_jitter_stream(rtds_ctx* c, uint64_t first, int n, double* out)
{
    if (!c || !out || n < 0) { rtds_set_error("jitter_stream: bad arguments"); return RTDS_ERR_INVALID; }
    RTDS_CUDA(cudaSetDevice(c->device));
    return rtds_jitter_stream_impl(c, first, n, out);
}

int rtds_prepare_frame(rtds_ctx* c, const rtds_render_params* p)
{
    if (!c || !p) { rtds_set_error("prepare_frame: bad arguments"); return RTDS_ERR_INVALID; }
    RTDS_CUDA(cudaSetDevice(c->device));
    return rtds_prefetch_dirs(c, p);
}

int rtds_set_option(rtds_ctx* c, const char* name, int value)
{
    if (!c || !name) { rtds_set_error("set_option: bad arguments"); return RTDS_ERR_INVALID; }
    for (int i = 0; i < g_rtds_n_option_names; ++i)
        if (!strcmp(name, g_rtds_option_names[i].name)) { c->opt.*(g_rtds_option_names[i].field) = value; return RTDS_OK; }
    rtds_set_error("set_option: unknown option '%s'", name);
    return RTDS_ERR_INVALID;
}

int rtds_get_option(rtds_ctx* c, const char* name, int* value)
{
    if (!c || !name || !value) { rtds_set_error("get_option: bad arguments"); return RTDS_ERR_INVALID; }
    for (int i = 0; i < g_rtds_n_option_names; ++i)
        if (!strcmp(name, g_rtds_option_names[i].name)) { *value = c->opt.*(g_rtds_option_names[i].field); return RTDS_OK; }
    rtds_set_error("get_option: unknown option '%s'", name);
    return RTDS_ERR_INVALID;
}

int rtds_morton30(rtds_ctx* c, const float* xyz, int n, uint32_t* codes)
{
    if (!c || !xyz || !codes || n < 0) { rtds_set_error("morton30: bad arguments"); return RTDS_ERR_INVALID; }
    RTDS_CUDA(cudaSetDevice(c->device));
    return rtds_morton30_device(c, xyz, n, codes);
}

}  // extern "C"
