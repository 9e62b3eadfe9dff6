// sah.cu — K7: binned-SAH BVH builder (extension: the reference's only BVH split is the object median,
// accelerators.h:22,246-337; SAH exists there only in the KD-tree). Same output as the other builders: Node64
// records, one primitive per leaf, leaves in DFS order, exportable as LinearBVHNode[] (accelerators.h:231-240).
//
// Deterministic definition (restated sequentially in oracle/oracle.cpp, orc_build_sah; trees are compared bit for bit):
//   per node over the contiguous range [s,e) of the current order:
//     cb   = min/max of the centres; axis = longest extent of cb (ties as GeMaximumAxis, accelerators.h:190-199)
//     bin  = min(B-1, (int)(B * ((c[axis] - cb.min) / (cb.max - cb.min))))            B = 16 bins
//     cost(i) = nL(i) * SA(union of bins 0..i) + nR(i) * SA(union of bins i+1..B-1)   SA as accelerators.h:122-125
//     split = first i with minimal cost among those with nL > 0 and nR > 0;  left = bin <= i, STABLE partition
//     no such i (all centres in one bin / zero extent): left = the first (e-s)/2 primitives of the range
// Level-synchronous: bounds and bins are accumulated with exact, order-independent atomics (min/max on ordered
// uints, integer adds), the stable partition is one global scan per level, boxes come from one atomic refit.
#include "rtds_internal.cuh"
#include "scan.cuh"
#include <math.h>
#include <algorithm>

int rtds_bvh_refit(rtds_ctx* ctx, DeviceBvh& b, const uint32_t* d_ids, int n, unsigned* d_counters, float* d_root_box, int* d_depth);  // lbvh.cu
int rtds_bvh_compute_depth(rtds_ctx* ctx, DeviceBvh& b, int* depth_out);
int rtds_bvh_reorder_preorder(rtds_ctx* ctx, DeviceBvh& b, void* scratch, int* launches);

namespace {

constexpr int MAXB = 32;
// Tasks of at most SAH_SMALL primitives - all tasks of the lower half of the levels - never touch the global bin records (896
// bytes per task: resetting, filling and reading them was 80 % of a 7 M-primitive build): a warp (or, for tiny tasks, a thread)
// builds the task's bins in shared (local) memory and decides the split with the same code. Same definition, same tree.
constexpr int SAH_SMALL = 512;     // ... at most this many primitives: one WARP per task (sah_warp_tasks), bins in shared memory
constexpr int SAH_TINY = 32;       // ... at most this many: one THREAD per task (sah_small_tasks), bins in local memory

// The primitive at position p of the current order. Sphere scenes carry the sphere records ALONG with the permutation (psph:
// position-ordered copies, scattered together with perm by sah_scatter): every per-level pass then reads them as a sequential
// stream. Gathering sph[perm[p]] instead - 16 scattered bytes per position, two or three times per level - was what the level
// kernels spent their time on (7 M primitives: 0.33 ms per pass). Triangle scenes (extension) keep the gather.
__device__ __forceinline__ void pos_fetch(const PrimView& pv, const int* __restrict__ perm, const float4* __restrict__ psph, int p, float c[3],
                                          float mn[3], float mx[3])
{
    if (psph) {
        const float4 s = __ldg(psph + p);
        c[0] = s.x; c[1] = s.y; c[2] = s.z;
        mn[0] = s.x - s.w; mn[1] = s.y - s.w; mn[2] = s.z - s.w;       // main.cpp:686-688, as prim_fetch
        mx[0] = s.x + s.w; mx[1] = s.y + s.w; mx[2] = s.z + s.w;
    } else prim_fetch(pv, perm[p], c, mn, mx);
}

struct SahTask {
    int s, e, parent_enc;     // range, parent*2+side (-1 root)
    int node;                 // Node64 index
    int axis, split_bin, nL;  // decision
    int child_base;           // index of the first child task in the next level (or -1)
    unsigned cb[6];           // centre bounds as ordered uints (min x,y,z, max x,y,z)
};

struct SahBins {              // per task
    unsigned cnt[MAXB];
    unsigned box[MAXB][6];    // ordered uints
};

__global__ void sah_reset(SahTask* tasks, SahBins* bins, int n_tasks, int B)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    if (tasks[t].e - tasks[t].s <= SAH_SMALL) return;
    for (int a = 0; a < 3; ++a) { tasks[t].cb[a] = 0xffffffffu; tasks[t].cb[3 + a] = 0u; }
    for (int b = 0; b < B; ++b) {
        bins[t].cnt[b] = 0;
        for (int a = 0; a < 3; ++a) { bins[t].box[b][a] = 0xffffffffu; bins[t].box[b][3 + a] = 0u; }
    }
}

// The kernels of the LARGE tasks (more than SAH_SMALL >= 2 x 256 primitives). A block walks a contiguous SPAN of positions in
// chunks of 256 and keeps ONE accumulator - for the large task it is currently inside - in shared memory; it touches that task's
// global words only when the walk leaves the task (or the span ends). A chunk meets at most two large tasks (each covers more
// than 512 consecutive positions): the one that continues from the chunk before (the smaller id is not required - only that the
// two are distinct) and the one that starts in it, in this order. With one block per 256 positions the 27,000 blocks of a 7 M
// build all sent their 112 bin words to the SAME few addresses at the top levels: 0.3 ms per pass of pure same-address atomic
// traffic. Exact either way (min / max / integer adds are order-independent).
constexpr int SAH_SPAN_BLOCKS_PER_SM = 8;
__device__ __forceinline__ void sah_span(int n, int& p_begin, int& p_end)
{
    const int chunks = (n + 255) / 256, per = (chunks + gridDim.x - 1) / gridDim.x;
    p_begin = min((long long)n, (long long)blockIdx.x * per * 256);
    p_end = min(n, p_begin + per * 256);
}
// the large task at the chunk's first active position (it may continue from the chunk before) and the other large task of the
// chunk, if any (-1 otherwise). All 256 threads call.
__device__ __forceinline__ void sah_chunk_tasks(bool act, int t, int& first, int& other)
{
    __shared__ int s_minpos, s_first, s_other;
    if (threadIdx.x == 0) { s_minpos = 0x7fffffff; s_first = -1; s_other = -1; }
    __syncthreads();
    if (act) atomicMin(&s_minpos, (int)threadIdx.x);
    __syncthreads();
    if ((int)threadIdx.x == s_minpos) s_first = t;
    __syncthreads();
    if (act && t != s_first) s_other = t;            // every writer writes the same id
    __syncthreads();
    first = s_first; other = s_other;
    __syncthreads();                                 // the three words are reused by the next chunk
}

__global__ void __launch_bounds__(256) sah_centre_bounds(const int* __restrict__ perm, const int* __restrict__ owner, int n, const PrimView pv,
                                                         const float4* __restrict__ psph, SahTask* tasks)
{
    __shared__ int s_cur;
    __shared__ unsigned s_cb[6];
    int p0, p1;
    sah_span(n, p0, p1);
    if (threadIdx.x == 0) s_cur = -1;
    if (threadIdx.x < 6) s_cb[threadIdx.x] = threadIdx.x < 3 ? 0xffffffffu : 0u;
    __syncthreads();
    auto flush = [&]() {        // all threads; leaves the accumulator empty
        if (threadIdx.x < 6 && s_cur >= 0) {
            const unsigned v = s_cb[threadIdx.x];
            if (threadIdx.x < 3) { if (v != 0xffffffffu) atomicMin(&tasks[s_cur].cb[threadIdx.x], v); }
            else atomicMax(&tasks[s_cur].cb[threadIdx.x], v);
            s_cb[threadIdx.x] = threadIdx.x < 3 ? 0xffffffffu : 0u;
        }
        __syncthreads();
    };
    for (int base = p0; base < p1; base += 256) {
        const int p = base + threadIdx.x;
        int t = p < p1 ? owner[p] : -2;
        if (t >= 0 && tasks[t].e - tasks[t].s <= SAH_SMALL) t = -2;      // small tasks are handled by sah_warp_tasks / sah_small_tasks
        const bool act = t >= 0;
        int first, other;
        sah_chunk_tasks(act, t, first, other);
        for (int phase = 0; phase < 2; ++phase) {
            const int task = phase ? other : first;
            if (task < 0) break;
            if (s_cur != task) { flush(); if (threadIdx.x == 0) s_cur = task; __syncthreads(); }
            if (act && t == task) {
                float c[3], mn[3], mx[3];
                pos_fetch(pv, perm, psph, p, c, mn, mx);
                for (int a = 0; a < 3; ++a) { const unsigned o = f2ord(c[a]); atomicMin(&s_cb[a], o); atomicMax(&s_cb[3 + a], o); }
            }
            __syncthreads();
        }
    }
    flush();
}

__device__ __forceinline__ int sah_axis(const unsigned cb[6], float& lo, float& hi)
{
    float mn[3] = {ord2f(cb[0]), ord2f(cb[1]), ord2f(cb[2])}, mx[3] = {ord2f(cb[3]), ord2f(cb[4]), ord2f(cb[5])};
    float ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
    int axis = (ex > ey && ex > ez) ? 0 : (ey > ez ? 1 : 2);
    lo = mn[axis]; hi = mx[axis];
    return axis;
}

__global__ void __launch_bounds__(256) sah_binning(const int* __restrict__ perm, const int* __restrict__ owner, int n, const PrimView pv,
                                                   const float4* __restrict__ psph, const SahTask* __restrict__ tasks, SahBins* bins,
                                                   int* __restrict__ bin_of, int B)
{
    __shared__ int s_cur;
    __shared__ unsigned s_cnt[MAXB];
    __shared__ unsigned s_box[MAXB][6];
    int p0, p1;
    sah_span(n, p0, p1);
    if (threadIdx.x == 0) s_cur = -1;
    if (threadIdx.x < MAXB) { s_cnt[threadIdx.x] = 0; for (int a = 0; a < 3; ++a) { s_box[threadIdx.x][a] = 0xffffffffu; s_box[threadIdx.x][3 + a] = 0u; } }
    __syncthreads();
    auto flush = [&]() {        // all threads; leaves the accumulator empty
        if (threadIdx.x < B && s_cur >= 0 && s_cnt[threadIdx.x]) {
            SahBins& g = bins[s_cur];
            const int bb = threadIdx.x;
            atomicAdd(&g.cnt[bb], s_cnt[bb]);
            for (int a = 0; a < 3; ++a) { atomicMin(&g.box[bb][a], s_box[bb][a]); atomicMax(&g.box[bb][3 + a], s_box[bb][3 + a]); }
        }
        __syncthreads();
        if (threadIdx.x < MAXB) { s_cnt[threadIdx.x] = 0; for (int a = 0; a < 3; ++a) { s_box[threadIdx.x][a] = 0xffffffffu; s_box[threadIdx.x][3 + a] = 0u; } }
        __syncthreads();
    };
    for (int base = p0; base < p1; base += 256) {
        const int p = base + threadIdx.x;
        int t = p < p1 ? owner[p] : -2;
        if (t >= 0 && tasks[t].e - tasks[t].s <= SAH_SMALL) t = -2;      // small tasks are handled by sah_warp_tasks / sah_small_tasks
        const bool act = t >= 0;
        int first, other;
        sah_chunk_tasks(act, t, first, other);
        if (first < 0) continue;
        int b = 0;
        float c[3], mn[3], mx[3];
        if (act) {
            pos_fetch(pv, perm, psph, p, c, mn, mx);
            float lo, hi;
            int axis = sah_axis(tasks[t].cb, lo, hi);
            if (hi > lo) {
                float k = c[axis];
                b = (int)((float)B * ((k - lo) / (hi - lo)));
                if (b > B - 1) b = B - 1;
            }
            bin_of[p] = b;
        }
        for (int phase = 0; phase < 2; ++phase) {
            const int task = phase ? other : first;
            if (task < 0) break;
            if (s_cur != task) { flush(); if (threadIdx.x == 0) s_cur = task; __syncthreads(); }
            if (act && t == task) {
                atomicAdd(&s_cnt[b], 1u);
                unsigned* bx = s_box[b];
                atomicMin(&bx[0], f2ord(mn[0])); atomicMin(&bx[1], f2ord(mn[1])); atomicMin(&bx[2], f2ord(mn[2]));
                atomicMax(&bx[3], f2ord(mx[0])); atomicMax(&bx[4], f2ord(mx[1])); atomicMax(&bx[5], f2ord(mx[2]));
            }
            __syncthreads();
        }
    }
    flush();
}

__device__ __forceinline__ float box_area(const float mn[3], const float mx[3])   // BoxBoundries::SurfaceArea, accelerators.h:122-125
{
    float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
    return 2 * (dx * dy + dx * dz + dy * dz);
}

// the split decision from a task's bins (counts + boxes as ordered uints): first minimum of the binned SAH cost.
// Backward pass: area and count of the union of bins b..B-1 (32 words of local memory instead of the suffix boxes themselves);
// forward pass: prefix box in registers, cost(i) = nL * SA(bins 0..i) + nR * SA(bins i+1..B-1).
__device__ __forceinline__ void sah_decide_core(const unsigned* cnt, const unsigned (*box)[6], int B, int m, int& split_bin, int& nL_out)
{
    float rarea[MAXB];
    unsigned rcnt[MAXB];
    float amn[3] = {INFINITY, INFINITY, INFINITY}, amx[3] = {-INFINITY, -INFINITY, -INFINITY};
    unsigned acc = 0;
    for (int b = B - 1; b >= 1; --b) {
        if (cnt[b]) for (int a = 0; a < 3; ++a) { amn[a] = fminf(amn[a], ord2f(box[b][a])); amx[a] = fmaxf(amx[a], ord2f(box[b][3 + a])); }
        acc += cnt[b];
        rarea[b] = box_area(amn, amx);
        rcnt[b] = acc;
    }
    float lmn[3] = {INFINITY, INFINITY, INFINITY}, lmx[3] = {-INFINITY, -INFINITY, -INFINITY};
    unsigned nL = 0;
    float best = INFINITY;
    int best_i = -1, best_nL = 0;
    for (int i = 0; i < B - 1; ++i) {
        if (cnt[i]) for (int a = 0; a < 3; ++a) { lmn[a] = fminf(lmn[a], ord2f(box[i][a])); lmx[a] = fmaxf(lmx[a], ord2f(box[i][3 + a])); }
        nL += cnt[i];
        unsigned nR = rcnt[i + 1];
        if (nL == 0 || nR == 0) continue;
        float cost = (float)nL * box_area(lmn, lmx) + (float)nR * rarea[i + 1];
        if (cost < best) { best = cost; best_i = i; best_nL = (int)nL; }
    }
    if (best_i < 0) { split_bin = -1; nL_out = m / 2; }     // all centres in one bin: positional median of the range
    else { split_bin = best_i; nL_out = best_nL; }
}

// one thread per LARGE task: evaluate the B-1 split planes; has_child[2t], has_child[2t+1] = child is a task (size >= 2)
__global__ void sah_decide(SahTask* tasks, const SahBins* __restrict__ bins, int n_tasks, int B, int node_base, int* __restrict__ child_flags,
                           int* __restrict__ n_large_next)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    SahTask& T = tasks[t];
    const int m = T.e - T.s;
    if (m <= SAH_SMALL) return;
    const SahBins& bn = bins[t];
    float lo, hi;
    T.axis = sah_axis(T.cb, lo, hi);
    T.node = node_base + t;
    sah_decide_core(bn.cnt, bn.box, B, m, T.split_bin, T.nL);
    child_flags[2 * t] = T.nL >= 2 ? 1 : 0;
    child_flags[2 * t + 1] = (m - T.nL) >= 2 ? 1 : 0;
    // the next level's large and medium tasks (n_large_next[0], [1]); tiny children are not counted
    const int nl = T.nL, nr = m - T.nL;
    const int large = (nl > SAH_SMALL ? 1 : 0) + (nr > SAH_SMALL ? 1 : 0);
    const int medium = ((nl > SAH_TINY && nl <= SAH_SMALL) ? 1 : 0) + ((nr > SAH_TINY && nr <= SAH_SMALL) ? 1 : 0);
    if (large) atomicAdd(n_large_next, large);
    if (medium) atomicAdd(n_large_next + 1, medium);
}

// one thread per TINY task (<= SAH_TINY primitives): centre bounds, bins and decision from the task's own primitives - what
// sah_reset + sah_centre_bounds + sah_binning + sah_decide do for the large ones through global memory
__global__ void __launch_bounds__(128) sah_small_tasks(SahTask* tasks, int n_tasks, int B, int node_base, const int* __restrict__ perm, const PrimView pv,
                                                       const float4* __restrict__ psph, int* __restrict__ bin_of, int* __restrict__ child_flags)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    SahTask& T = tasks[t];
    const int s0 = T.s, m = T.e - T.s;
    if (m > SAH_TINY) return;
    unsigned cb[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    for (int p = s0; p < s0 + m; ++p) {
        float c[3], mn[3], mx[3];
        pos_fetch(pv, perm, psph, p, c, mn, mx);
        for (int a = 0; a < 3; ++a) { const unsigned o = f2ord(c[a]); cb[a] = min(cb[a], o); cb[3 + a] = max(cb[3 + a], o); }
    }
    for (int a = 0; a < 6; ++a) T.cb[a] = cb[a];
    float lo, hi;
    const int axis = sah_axis(cb, lo, hi);
    unsigned cnt[MAXB];
    unsigned box[MAXB][6];
    for (int b = 0; b < B; ++b) { cnt[b] = 0u; for (int a = 0; a < 3; ++a) { box[b][a] = 0xffffffffu; box[b][3 + a] = 0u; } }
    for (int p = s0; p < s0 + m; ++p) {
        float c[3], mn[3], mx[3];
        pos_fetch(pv, perm, psph, p, c, mn, mx);
        int b = 0;
        if (hi > lo) {
            float k = c[axis];
            b = (int)((float)B * ((k - lo) / (hi - lo)));
            if (b > B - 1) b = B - 1;
        }
        bin_of[p] = b;
        cnt[b] += 1u;
        for (int a = 0; a < 3; ++a) { box[b][a] = min(box[b][a], f2ord(mn[a])); box[b][3 + a] = max(box[b][3 + a], f2ord(mx[a])); }
    }
    T.axis = axis;
    T.node = node_base + t;
    sah_decide_core(cnt, box, B, m, T.split_bin, T.nL);
    child_flags[2 * t] = T.nL >= 2 ? 1 : 0;
    child_flags[2 * t + 1] = (m - T.nL) >= 2 ? 1 : 0;
}

// one WARP per task of SAH_TINY < m <= SAH_SMALL primitives: the lanes stride over the task's primitives, centre bounds by warp
// reductions, bins by shared-memory atomics (exact: min / max / integer adds), lane 0 decides with the same code
__global__ void __launch_bounds__(128) sah_warp_tasks(SahTask* tasks, int n_tasks, int B, int node_base, const int* __restrict__ perm, const PrimView pv,
                                                      const float4* __restrict__ psph, int* __restrict__ bin_of, int* __restrict__ child_flags,
                                                      int* __restrict__ n_next)
{
    __shared__ unsigned s_cnt[4][MAXB];
    __shared__ unsigned s_box[4][MAXB][6];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * 4 + warp;
    if (t >= n_tasks) return;
    SahTask& T = tasks[t];
    const int s0 = T.s, m = T.e - T.s;
    if (m <= SAH_TINY || m > SAH_SMALL) return;
    unsigned cb[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    for (int p = s0 + lane; p < s0 + m; p += 32) {
        float c[3], mn[3], mx[3];
        pos_fetch(pv, perm, psph, p, c, mn, mx);
        for (int a = 0; a < 3; ++a) { const unsigned o = f2ord(c[a]); cb[a] = min(cb[a], o); cb[3 + a] = max(cb[3 + a], o); }
    }
    for (int a = 0; a < 3; ++a) { cb[a] = __reduce_min_sync(0xffffffffu, cb[a]); cb[3 + a] = __reduce_max_sync(0xffffffffu, cb[3 + a]); }
    float lo, hi;
    const int axis = sah_axis(cb, lo, hi);
    for (int b = lane; b < B; b += 32) { s_cnt[warp][b] = 0u; for (int a = 0; a < 3; ++a) { s_box[warp][b][a] = 0xffffffffu; s_box[warp][b][3 + a] = 0u; } }
    __syncwarp();
    for (int p = s0 + lane; p < s0 + m; p += 32) {
        float c[3], mn[3], mx[3];
        pos_fetch(pv, perm, psph, p, c, mn, mx);
        int b = 0;
        if (hi > lo) {
            float k = c[axis];
            b = (int)((float)B * ((k - lo) / (hi - lo)));
            if (b > B - 1) b = B - 1;
        }
        bin_of[p] = b;
        atomicAdd(&s_cnt[warp][b], 1u);
        for (int a = 0; a < 3; ++a) { atomicMin(&s_box[warp][b][a], f2ord(mn[a])); atomicMax(&s_box[warp][b][3 + a], f2ord(mx[a])); }
    }
    __syncwarp();
    if (lane == 0) {
        for (int a = 0; a < 6; ++a) T.cb[a] = cb[a];
        T.axis = axis;
        T.node = node_base + t;
        sah_decide_core(s_cnt[warp], s_box[warp], B, m, T.split_bin, T.nL);
        child_flags[2 * t] = T.nL >= 2 ? 1 : 0;
        child_flags[2 * t + 1] = (m - T.nL) >= 2 ? 1 : 0;
        const int medium = ((T.nL > SAH_TINY) ? 1 : 0) + ((m - T.nL > SAH_TINY) ? 1 : 0);      // its children are at most SAH_SMALL
        if (medium) atomicAdd(n_next + 1, medium);
    }
}

__global__ void sah_left_flags(const int* __restrict__ owner, const int* __restrict__ bin_of, int n, const SahTask* __restrict__ tasks, int* __restrict__ flags)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int t = owner[p];
    int f = 0;
    if (t >= 0) {
        const SahTask& T = tasks[t];
        f = T.split_bin >= 0 ? (bin_of[p] <= T.split_bin) : ((p - T.s) < T.nL);
    }
    flags[p] = f;
}

// stable partition of every active range + creation of nodes / child tasks / leaf records
__global__ void sah_scatter(const int* __restrict__ perm, const int* __restrict__ owner, const int* __restrict__ flags, const int* __restrict__ scan,
                            int n, const SahTask* __restrict__ tasks, const int* __restrict__ child_scan, int* __restrict__ perm_next,
                            int* __restrict__ owner_next, int* __restrict__ leaf_info, const float4* __restrict__ psph, float4* __restrict__ psph_next)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int t = owner[p];
    if (t < 0) { perm_next[p] = perm[p]; owner_next[p] = -1; if (psph) psph_next[p] = psph[p]; return; }
    const SahTask& T = tasks[t];
    const int lefts_before = scan[p] - scan[T.s];
    const bool left = flags[p] != 0;
    const int dst = left ? T.s + lefts_before : T.s + T.nL + ((p - T.s) - lefts_before);
    perm_next[dst] = perm[p];
    if (psph) psph_next[dst] = psph[p];
    const int m = T.e - T.s;
    const int csize = left ? T.nL : m - T.nL;
    if (csize >= 2) owner_next[dst] = child_scan[2 * t + (left ? 0 : 1)];
    else { owner_next[dst] = -1; leaf_info[dst] = T.node * 2 + (left ? 0 : 1) + 2; }
}

__global__ void sah_make_children(const SahTask* __restrict__ tasks, int n_tasks, const int* __restrict__ child_flags, const int* __restrict__ child_scan,
                                  SahTask* __restrict__ next, Node64* __restrict__ nodes, int next_node_base)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    const SahTask& T = tasks[t];
    Node64& nd = nodes[T.node];
    nd.axis = T.axis;
    nd.parent = T.parent_enc;
    if (T.parent_enc >= 0) { if (T.parent_enc & 1) nodes[T.parent_enc >> 1].right = T.node; else nodes[T.parent_enc >> 1].left = T.node; }
    for (int side = 0; side < 2; ++side) {
        if (!child_flags[2 * t + side]) continue;
        SahTask c;
        c.s = side ? T.s + T.nL : T.s;
        c.e = side ? T.e : T.s + T.nL;
        c.parent_enc = T.node * 2 + side;
        c.node = next_node_base + child_scan[2 * t + side];
        c.axis = 0; c.split_bin = -1; c.nL = 0; c.child_base = -1;
        next[child_scan[2 * t + side]] = c;
    }
}

__global__ void sah_finalize_leaves(const int* __restrict__ leaf_info, int n, Node64* nodes, int* __restrict__ leaf_parent)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int pe = leaf_info[p] - 2;
    if (pe < 0) { leaf_parent[p] = 0; return; }
    leaf_parent[p] = (pe >> 1) | ((pe & 1) ? 0x80000000 : 0);
    if (pe & 1) nodes[pe >> 1].right = ~p; else nodes[pe >> 1].left = ~p;
}

__global__ void sah_init(int* perm, int* owner, int n, const float4* __restrict__ sph, float4* __restrict__ psph)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { perm[i] = i; owner[i] = n >= 2 ? 0 : -1; if (psph) psph[i] = sph[i]; }
}

template <typename T> T* carve(char*& p, size_t count)
{
    T* r = (T*)p;
    p += (sizeof(T) * count + 255) & ~(size_t)255;
    return r;
}

}  // namespace

int rtds_build_sah(rtds_ctx* ctx, const rtds_build_params* bp, rtds_build_stats* st)
{
    const int n = ctx->n;
    const int B = (bp && bp->sah_bins > 1) ? std::min(bp->sah_bins, MAXB) : 16;
    DeviceBvh& b = ctx->bvh;
    RTDS_TRY(rtds_alloc_bvh_for(ctx, b, n));
    const PrimView pv = rtds_prim_view(ctx);
    const size_t max_tasks = (size_t)n / 2 + 2;
    const int tiles = (int)(((size_t)std::max<size_t>(n, 2 * max_tasks) + rtds_scan::SC_TILE - 1) / rtds_scan::SC_TILE) + 1;
    size_t bytes = 6 * (((size_t)n * 4 + 255) & ~(size_t)255) + 2 * ((sizeof(SahTask) * max_tasks + 255) & ~(size_t)255) +
                   ((sizeof(SahBins) * max_tasks + 255) & ~(size_t)255) + 4 * ((8 * max_tasks + 255) & ~(size_t)255) + ((size_t)tiles * 4 + 255) +
                   (((size_t)n * 4 + 255) & ~(size_t)255) + 4096;
    const bool carry = ctx->prim_type == 0;      // sphere records travel with the permutation (pos_fetch)
    if (carry) bytes += 2 * (((size_t)n * 16 + 255) & ~(size_t)255);
    RTDS_TRY(rtds_ensure_scratch(ctx, bytes));
    char* p = (char*)ctx->d_scratch;
    int* perm[2] = {carve<int>(p, n), carve<int>(p, n)};
    int* owner[2] = {carve<int>(p, n), carve<int>(p, n)};
    int* bin_of = carve<int>(p, n);
    int* flags = carve<int>(p, n);
    int* scan = carve<int>(p, n);
    SahTask* tasks[2] = {carve<SahTask>(p, max_tasks), carve<SahTask>(p, max_tasks)};
    SahBins* bins = carve<SahBins>(p, max_tasks);
    int* child_flags = carve<int>(p, 2 * max_tasks);
    int* child_scan = carve<int>(p, 2 * max_tasks);
    int* sums = carve<int>(p, tiles);
    int* d_small = carve<int>(p, 64);
    cudaStream_t s = ctx->stream;
    int launches = 0;
    const int T = 256;
    auto G = [&](long long m) { return (unsigned)((m + T - 1) / T); };
    const unsigned span_grid = std::max(1u, std::min(G(n), (unsigned)(ctx->sm_count * SAH_SPAN_BLOCKS_PER_SM)));      // sah_span: contiguous spans of positions
    rtds_scan::Scanner scanner{ctx, sums, tiles, d_small};
    auto xscan = [&](const int* in, int* out, int m) -> int { return scanner.run(in, out, m, &launches); };
    int* leaf_arr = carve<int>(p, n);   // scene position -> parent*2+side+2 once the position holds a finished leaf
    float4* psph[2] = {nullptr, nullptr};
    if (carry) { psph[0] = carve<float4>(p, n); psph[1] = carve<float4>(p, n); }
    RTDS_CUDA(cudaEventRecord(ctx->ev0, s));
    RTDS_CUDA(cudaMemsetAsync(leaf_arr, 0, (size_t)n * 4, s));
    sah_init<<<G(n), T, 0, s>>>(perm[0], owner[0], n, ctx->d_sph, psph[0]);
    ++launches;
    int cur = 0, n_tasks = n >= 2 ? 1 : 0, node_base = 0, levels = 0;
    int n_large = n > SAH_SMALL ? 1 : 0;        // tasks of the current level above SAH_SMALL primitives ...
    int n_medium = (n > SAH_TINY && n <= SAH_SMALL) ? 1 : 0;      // ... and of SAH_TINY < m <= SAH_SMALL
    int* d_large = d_small + 40;                // both counted for the next level by the deciding kernels
    if (n_tasks) {
        SahTask root;
        memset(&root, 0, sizeof root);
        root.s = 0; root.e = n; root.parent_enc = -1; root.node = 0; root.split_bin = -1; root.child_base = -1;
        RTDS_CUDA(cudaMemcpyAsync(tasks[0], &root, sizeof root, cudaMemcpyHostToDevice, s));
    } else {
        const int two = 1;   // single primitive: leaf_info = root leaf
        RTDS_CUDA(cudaMemcpyAsync(leaf_arr, &two, sizeof two, cudaMemcpyHostToDevice, s));
    }
    while (n_tasks > 0) {
        SahTask* tk = tasks[cur];
        RTDS_CUDA(cudaMemsetAsync(d_large, 0, 2 * sizeof(int), s));
        if (n_large > 0) {       // tasks above SAH_SMALL primitives: bins in global memory, filled by all their positions
            sah_reset<<<G(n_tasks), T, 0, s>>>(tk, bins, n_tasks, B);
            sah_centre_bounds<<<span_grid, T, 0, s>>>(perm[cur], owner[cur], n, pv, psph[cur], tk);
            sah_binning<<<span_grid, T, 0, s>>>(perm[cur], owner[cur], n, pv, psph[cur], tk, bins, bin_of, B);
            sah_decide<<<G(n_tasks), T, 0, s>>>(tk, bins, n_tasks, B, node_base, child_flags, d_large);
            launches += 4;
        }
        if (n_medium > 0) {      // up to SAH_SMALL primitives: one warp each, bins in shared memory
            sah_warp_tasks<<<(n_tasks + 3) / 4, 128, 0, s>>>(tk, n_tasks, B, node_base, perm[cur], pv, psph[cur], bin_of, child_flags, d_large);
            launches += 1;
        }
        if (n_tasks > n_large + n_medium) {  // up to SAH_TINY primitives: one thread each
            sah_small_tasks<<<(n_tasks + 127) / 128, 128, 0, s>>>(tk, n_tasks, B, node_base, perm[cur], pv, psph[cur], bin_of, child_flags);
            launches += 1;
        }
        RTDS_TRY(xscan(child_flags, child_scan, 2 * n_tasks));
        int next_tasks = 0, next_large[2] = {0, 0};
        RTDS_CUDA(cudaMemcpyAsync(&next_tasks, d_small, sizeof(int), cudaMemcpyDeviceToHost, s));
        RTDS_CUDA(cudaMemcpyAsync(next_large, d_large, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
        sah_left_flags<<<G(n), T, 0, s>>>(owner[cur], bin_of, n, tk, flags);
        ++launches;
        RTDS_TRY(xscan(flags, scan, n));
        sah_scatter<<<G(n), T, 0, s>>>(perm[cur], owner[cur], flags, scan, n, tk, child_scan, perm[cur ^ 1], owner[cur ^ 1], leaf_arr, psph[cur], psph[cur ^ 1]);
        RTDS_CUDA(cudaStreamSynchronize(s));
        sah_make_children<<<G(n_tasks), T, 0, s>>>(tk, n_tasks, child_flags, child_scan, tasks[cur ^ 1], b.nodes, node_base + n_tasks);
        launches += 2;
        node_base += n_tasks;
        n_tasks = next_tasks;
        n_large = next_large[0];
        n_medium = next_large[1];
        cur ^= 1;
        ++levels;
        if (levels > 4096) { rtds_set_error("sah build: runaway depth"); return RTDS_ERR_CUDA; }
    }
    sah_finalize_leaves<<<G(n), T, 0, s>>>(leaf_arr, n, b.nodes, b.leaf_parent);
    ++launches;
    unsigned* d_counters = (unsigned*)scan;
    float* d_root = (float*)(d_small + 16);
    int* d_depth = (int*)(d_root + 8);      // the refit carries subtree heights up: the root's height is the deepest leaf's depth
    RTDS_CUDA(cudaMemsetAsync(d_depth, 0, sizeof(int), s));
    RTDS_TRY(rtds_bvh_refit(ctx, b, (const uint32_t*)perm[cur], n, d_counters, d_root, d_depth));
    ++launches;
    RTDS_CUDA(cudaGetLastError());
    int depth = 0;
    RTDS_CUDA(cudaMemcpyAsync(b.root_box, d_root, sizeof(float) * 6, cudaMemcpyDeviceToHost, s));
    RTDS_CUDA(cudaMemcpyAsync(&depth, d_depth, sizeof(int), cudaMemcpyDeviceToHost, s));
    RTDS_CUDA(cudaStreamSynchronize(s));
    b.n_prims = n;
    b.n_internal = n - 1;
    b.root_ref = n > 1 ? 0 : ~0;
    b.tie_by_objid = 1;
    b.leaf_box_prim = ctx->prim_type == 0;
    RTDS_TRY(rtds_bvh_reorder_preorder(ctx, b, ctx->d_scratch, &launches));   // the level-loop scratch is free now
    RTDS_CUDA(cudaEventRecord(ctx->ev1, s));
    RTDS_CUDA(cudaStreamSynchronize(s));
    b.max_depth = depth;
    b.valid = true;
    float ms = 0;
    RTDS_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (st) {
        st->n_prims = n; st->total_nodes = 2 * n - 1; st->alloc_nodes = 2 * n - 1; st->max_depth = depth;
        st->kernel_launches = launches; st->ms = ms;
    }
    return RTDS_OK;
}
