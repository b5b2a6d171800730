// psi_b200.cu -- CUDA kernels (sm_100a) and the C ABI of libpsi_b200.so.
//
// K0 background_kernel : J0/J1 of a dust distribution              (psi_dispersion.py:74-84, 183-187)
// K1 disp_eval_kernel  : D(omega; Kx, Kz, alpha), one warp per point  (psi_dispersion.py:375-456 +
//                        tanhsinh.py:72-128)
// Persistent CTAs (one per SM) stage the level-major node table in shared memory once and their
// warps pull points from a device-side atomic queue, so unevenly expensive points (1-4 quadrature
// sub-intervals, different early-exit levels) balance themselves.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/psi_b200.h"
#include "psi_internal.hpp"
#include "psi_mode.cuh"
#include "psi_table.hpp"

using namespace psi;

// ------------------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
int psi_internal_fail(int code, const std::string& msg) { return fail(code, msg); }
#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(PSI_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)

// ------------------------------------------------------------------------------------------------
// warp policy on the device
// ------------------------------------------------------------------------------------------------
struct WarpCuda {
    static constexpr int width = 32;
    static __device__ __forceinline__ int lane() { return threadIdx.x & 31; }
    static __device__ __forceinline__ double sum(double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    static __device__ __forceinline__ double max(double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    // four complex sums at once: the eight butterflies are independent, one rolled loop over the rounds
    static __device__ __noinline__ void sum4c(double (&re)[4], double (&im)[4]) {
#pragma unroll 1
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                re[c] += __shfl_xor_sync(0xffffffffu, re[c], o);
                im[c] += __shfl_xor_sync(0xffffffffu, im[c], o);
            }
        }
    }
    static __device__ __forceinline__ void sync() { __syncwarp(); }
    static __device__ __forceinline__ double bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
    static __device__ __forceinline__ unsigned ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
    static __device__ __forceinline__ int popc(unsigned v) { return __popc(v); }
    static __device__ __forceinline__ bool any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
    // largest v over the lanes, ties to the smallest index (idx < 0: lane has no candidate)
    static __device__ __forceinline__ void argmax(double& v, int& idx) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (oi >= 0 && (idx < 0 || ov > v || (ov == v && oi < idx))) { v = ov; idx = oi; }
        }
    }
};

// Sub-warp policy: G lanes (8, 16 or 32) cooperate on one item, 32/G items share a warp -- and with it one
// instruction stream.  The analysis stage of K4 (psi_aaa.cuh) works on matrices of 18 x 10 (20 samples) to about
// 90 x 32 entries: a whole warp per search leaves most lanes idle and is bound by instruction fetch and by the
// latency of its dependent load/shuffle chains, so several searches per warp multiply the useful work per issued
// instruction.  All collectives name only the lanes of their own group (groups of one warp may diverge).
template <int G>
struct GroupCuda {
    static constexpr int width = G;
    static __device__ __forceinline__ int lane() { return threadIdx.x & (G - 1); }
    static __device__ __forceinline__ unsigned shift() { return threadIdx.x & 31 & ~(G - 1); }
    static __device__ __forceinline__ unsigned mask() {
        return G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << shift());
    }
    static __device__ __forceinline__ double sum(double v) {
        const unsigned mk = mask();
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mk, v, o);
        return v;
    }
    static __device__ __forceinline__ double max(double v) {
        const unsigned mk = mask();
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(mk, v, o));
        return v;
    }
    static __device__ __noinline__ void sum4c(double (&re)[4], double (&im)[4]) {
        const unsigned mk = mask();
#pragma unroll 1
        for (int o = G / 2; o > 0; o >>= 1) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                re[c] += __shfl_xor_sync(mk, re[c], o);
                im[c] += __shfl_xor_sync(mk, im[c], o);
            }
        }
    }
    static __device__ __forceinline__ void sync() { __syncwarp(mask()); }
    static __device__ __forceinline__ double bcast(double v, int src) { return __shfl_sync(mask(), v, src, G); }
    static __device__ __forceinline__ unsigned ballot(bool p) {
        const unsigned b = __ballot_sync(mask(), p) >> shift();
        return G == 32 ? b : (b & ((1u << (G & 31)) - 1u));
    }
    static __device__ __forceinline__ int popc(unsigned v) { return __popc(v); }
    static __device__ __forceinline__ bool any(bool p) { return ballot(p) != 0u; }
    static __device__ __forceinline__ void argmax(double& v, int& idx) {
        const unsigned mk = mask();
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(mk, v, o);
            const int oi = __shfl_xor_sync(mk, idx, o);
            if (oi >= 0 && (idx < 0 || ov > v || (ov == v && oi < idx))) { v = ov; idx = oi; }
        }
    }
};

// Thread policy: one thread = one evaluation, no cooperation (fast path of K1, below).
struct ThreadCuda {
    static constexpr int width = 1;
    static __device__ __forceinline__ int lane() { return 0; }
    static __device__ __forceinline__ double sum(double v) { return v; }
    static __device__ __forceinline__ double max(double v) { return v; }
    static __device__ __forceinline__ void sync() {}
    static __device__ __forceinline__ bool any(bool p) { return p; }
};

// CTA policy: ALL threads of the block cooperate on one evaluation (the quadrature code only sees lane(),
// width and the reductions).  Used by K3 when a launch has too few root searches to fill the GPU with one
// warp each: a secant polish is a serial chain of D evaluations, and the slowest chain sets the launch time.
// Reductions: butterfly inside each warp, one shared-memory slot per warp, one barrier, then every thread
// adds the slots in the same order (all threads get bit-identical sums; two slot sets alternate so that one
// barrier per reduction is enough).
template <int THREADS>
struct CtaCuda {
    static constexpr int width = THREADS;
    static constexpr int kWarps = THREADS / 32;
    static __device__ __forceinline__ int lane() { return threadIdx.x; }
    static __device__ __forceinline__ double* slots() {
        __shared__ double s[2][kWarps];
        return &s[0][0];
    }
    static __device__ __forceinline__ unsigned& phase() {
        __shared__ unsigned ph[THREADS];      // per-thread toggle (threads stay in lock step through the barriers)
        return ph[threadIdx.x];
    }
    template <class OP>
    static __device__ __forceinline__ double reduce(double v, OP op) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
        unsigned& ph = phase();
        double* s = slots() + (ph & 1u) * kWarps;
        ph ^= 1u;
        if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
        __syncthreads();
        double r = s[0];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) r = op(r, s[w]);
        return r;
    }
    static __device__ __forceinline__ double sum(double v) {
        return reduce(v, [](double a, double b) { return a + b; });
    }
    static __device__ __forceinline__ double max(double v) {
        return reduce(v, [](double a, double b) { return fmax(a, b); });
    }
    static __device__ __forceinline__ void sync() { __syncthreads(); }
    static __device__ __forceinline__ bool any(bool p) { return __syncthreads_or(p ? 1 : 0) != 0; }
};

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
constexpr int kMaxDists = 4096;
constexpr int kContourCap = 4096;          // points per contour buffer of K2 (the reference's runs stay < 1000)
#ifndef PSI_EVAL_THREADS
#define PSI_EVAL_THREADS 384
#endif
constexpr int kEvalThreads = PSI_EVAL_THREADS;   // warps*32 per persistent CTA (one CTA per SM)

struct psi_ctx {
    int device = 0;
    int sm_count = 0;
    int smem_optin = 0;
    cudaStream_t stream = nullptr;
    // node table
    node_t* d_nodes = nullptr;
    int* d_level_off = nullptr;
    int n_nodes = 0, max_m = 0;
    double eps = 0;
    size_t table_smem = 0;                 // bytes needed to stage the table in shared memory
    bool table_in_smem = false;
    // distributions
    DistParams* d_dists = nullptr;
    std::vector<DistParams> h_dists;
    // work queue + statistics
    unsigned long long* d_counter = nullptr;     // [0] queue head, [1] node evals (host entry points)
    // staging for the *_host entry points
    void* d_stage = nullptr;
    size_t stage_bytes = 0;
    // per-warp ping-pong point buffers of the contour kernel (K2)
    void* d_contour_scratch = nullptr;
    // per-warp AAA scratch of the mode kernel (K4), grown on demand
    void* d_mode_scratch = nullptr;
    size_t mode_scratch_bytes = 0;
    long long launches = 0;
    int predict_exit = 1;                  // quadrature exit rule of K1-K4 (psi_ctx_set_exit_rule)
    int fast_cap = 7;                      // K1 fast path: last level a thread-per-point evaluation may use (0 = off);
                                           // measured on config 2: cap 6 / 7 / 8 / 9 = 4.36 / 3.47 / 3.59 / 4.72 ms per step
    long long* d_hard = nullptr;           // indices of the points the fast path handed to the warp kernel
    size_t hard_cap = 0;
    std::vector<double*> d_tables;         // tabulated size densities (psi_dist_add_table)
    // K1 launches of one context share its work-queue head, hard-point list and counters: launches given
    // DIFFERENT streams (psi_disp_eval_dev) are therefore chained through this event -- a launch starts only after
    // the context's previous one has finished, whatever stream that ran on
    cudaEvent_t ev_k1 = nullptr;
    bool ev_k1_recorded = false;
    // pinned bounce buffers for psi_disp_eval_host (pageable caller memory), grow-only
    void *h_in = nullptr, *h_out = nullptr;
    size_t h_in_bytes = 0, h_out_bytes = 0;
    // early read-back of psi_disp_eval_host: results leave on a copy stream once the thread-per-point stage is done,
    // the points of the leftover stage follow as (index, value) patches
    static constexpr long long kPatchCap = 8192;
    cudaStream_t stream_copy = nullptr;
    cudaEvent_t ev_thread = nullptr, ev_copied = nullptr;
    char* d_patch = nullptr;        // kPatchCap x (long long index, double2 value)
    char* h_patch = nullptr;        // pinned mirror + the count
};

extern "C" { static int ensure_stage(psi_ctx* c, size_t bytes); }

static NodeTableView table_view(const psi_ctx* c) {
    NodeTableView t;
    t.nodes = c->d_nodes; t.level_off = c->d_level_off; t.n_nodes = c->n_nodes; t.max_m = c->max_m;
    t.eps = c->eps;
    t.predict_exit = c->predict_exit;
    t.level_cap = 0;
    return t;
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
__global__ void background_kernel(NodeTableView t, DistParams* dists, int id) {
    DistParams d = dists[id];
    unsigned long long nodes = 0;
    t.predict_exit = 0;                    // the background integrals always run the reference's literal rule
    background_warp<WarpCuda>(t, d, nodes);
    if (threadIdx.x == 0) {
        dists[id].j0 = d.j0; dists[id].j1 = d.j1; dists[id].denom = d.denom;
        dists[id].vgx = d.vgx; dists[id].vgy = d.vgy;
    }
}

// Copy the level-major node table into shared memory (once per persistent CTA) and return a view of it.
// The 204 KB of (x, w) pairs move with ONE bulk asynchronous copy (TMA, cp.async.bulk) issued by thread
// 0 and tracked by an mbarrier; the 56 bytes of level offsets go through ordinary loads.
__device__ __forceinline__ NodeTableView stage_table(const NodeTableView& gt, bool in_smem) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    NodeTableView t = gt;
    if (in_smem) {
        __shared__ __align__(8) unsigned long long bar;
        node_t* s_nodes = reinterpret_cast<node_t*>(smem_raw);
        int* s_off = reinterpret_cast<int*>(smem_raw + sizeof(node_t) * (size_t)gt.n_nodes);
        const unsigned bytes = (unsigned)(sizeof(node_t) * (size_t)gt.n_nodes);          // multiple of 16
        const unsigned bar_addr = (unsigned)__cvta_generic_to_shared(&bar);
        const unsigned dst_addr = (unsigned)__cvta_generic_to_shared(s_nodes);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_addr),
                "l"(gt.nodes), "r"(bytes), "r"(bar_addr)
                : "memory");
        }
        for (int i = threadIdx.x; i < gt.max_m + 2; i += blockDim.x) s_off[i] = gt.level_off[i];
        // every thread waits for the bulk copy to land (phase 0 of the barrier)
        unsigned done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar_addr)
                : "memory");
        }
        __syncthreads();
        t.nodes = s_nodes;
        t.level_off = s_off;
    }
    return t;
}

// the same for a whole block (thread 0 pulls, everybody reads it after the barrier)
__device__ __forceinline__ bool next_item_cta(unsigned long long* queue, long long n, long long& idx) {
    __shared__ unsigned long long pulled;
    __syncthreads();                                    // everybody is done with the previous item
    if (threadIdx.x == 0) pulled = atomicAdd(queue, 1ull);
    __syncthreads();
    idx = (long long)pulled;
    return pulled < (unsigned long long)n;
}

__device__ __forceinline__ bool next_item(unsigned long long* queue, long long n, long long& idx) {
    unsigned long long v = 0;
    if ((threadIdx.x & 31) == 0) v = atomicAdd(queue, 1ull);
    v = __shfl_sync(0xffffffffu, v, 0);
    idx = (long long)v;
    return v < (unsigned long long)n;
}

template <class W>
__device__ __forceinline__ bool next_item_group(unsigned long long* queue, long long n, long long& idx) {
    double v = 0.0;
    if (W::lane() == 0) v = (double)atomicAdd(queue, 1ull);          // item counts stay far below 2^53
    v = W::bcast(v, 0);
    idx = (long long)v;
    return idx < n;
}

// K1.  One specialisation per work class (NK, VISC); a launch handles the points of ITS class and
// skips the others (a batch normally holds a single class, see launch_disp_eval).
template <int NK, bool VISC>
__global__ void __launch_bounds__(kEvalThreads, 1)
disp_eval_kernel(NodeTableView gt, int table_in_smem, const DistParams* __restrict__ dists, long long n,
                 int broadcast,
                 const int* __restrict__ dist, const double* __restrict__ kx,
                 const double* __restrict__ kz, const double* __restrict__ alpha,
                 const int* __restrict__ pidx, const double2* __restrict__ w, double2* __restrict__ D,
                 double2* __restrict__ M, unsigned long long* __restrict__ queue,
                 unsigned long long* __restrict__ node_evals, const long long* __restrict__ elem,
                 int elem_cap, const long long* __restrict__ remap,
                 const unsigned long long* __restrict__ n_dev, unsigned long long min_work) {
    // second stage of the fast path: the work list is the `remap` array, its length sits in device memory; with
    // fewer than min_work points on it the block-per-point kernel below has taken them
    const unsigned long long n_work = n_dev != nullptr ? *n_dev : (unsigned long long)n;
    if (n_dev != nullptr && n_work < min_work) return;
    const NodeTableView t = stage_table(gt, table_in_smem != 0);
    const int lane = threadIdx.x & 31;
    unsigned long long nodes = 0;
    while (true) {
        unsigned long long idx = 0;
        if (lane == 0) idx = atomicAdd(queue, 1ull);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= n_work) break;
        if (remap != nullptr) idx = (unsigned long long)remap[idx];
        // element mode (mode pipeline): point = entry e of the per-item sample buffers, item = e / cap
        const long long e = elem != nullptr ? elem[idx] : (long long)idx;
        const long long pi = elem != nullptr ? e / elem_cap
                             : (pidx != nullptr ? (long long)pidx[idx] : (broadcast ? 0 : (long long)idx));
        const DistParams& d = dists[dist[pi]];
        const double al = alpha[pi];
        if (point_class(d, al) != NK * 2 + (VISC ? 1 : 0)) continue;
        const double2 wv = w[e];
        cx m[7];
        const cx val = disp_eval_warp<WarpCuda, NK, VISC>(t, d, cx(wv.x, wv.y), kx[pi], kz[pi], al,
                                                          M != nullptr ? m : nullptr, nodes);
        if (lane == 0) D[e] = make_double2(val.re, val.im);
        if (M != nullptr && lane < 7) {
            cx mv = m[0];
#pragma unroll
            for (int k = 1; k < 7; ++k) if (lane == k) mv = m[k];
            M[idx * 7 + lane] = make_double2(mv.re, mv.im);
        }
    }
    if (node_evals != nullptr && lane == 0 && nodes) atomicAdd(node_evals, nodes);
}

typedef void (*eval_kernel_t)(NodeTableView, int, const DistParams*, long long, int, const int*, const double*,
                              const double*, const double*, const int*, const double2*, double2*, double2*,
                              unsigned long long*, unsigned long long*, const long long*, int, const long long*,
                              const unsigned long long*, unsigned long long);

// Left-over stage of the fast path when only a HANDFUL of points is left (with the split at tau* = -1/Im w that is the
// rule: ~0.02 % of BASELINE config 2's grid, each needing all twelve levels -- 30 k nodes): one BLOCK per point, the
// nodes of a level spread over its 384 threads (the block policy of K3's block mode), instead of one warp grinding
// through a point for a millisecond while 99 % of the machine idles.  Runs only when the list holds fewer than
// `max_work` points; the warp kernel takes longer lists.
template <int NK, bool VISC>
__global__ void __launch_bounds__(kEvalThreads, 1)
disp_eval_cta_kernel(NodeTableView gt, int table_in_smem, const DistParams* __restrict__ dists, int broadcast,
                     const int* __restrict__ dist, const double* __restrict__ kx, const double* __restrict__ kz,
                     const double* __restrict__ alpha, const int* __restrict__ pidx, const double2* __restrict__ w,
                     double2* __restrict__ D, unsigned long long* __restrict__ queue,
                     unsigned long long* __restrict__ node_evals, const long long* __restrict__ elem, int elem_cap,
                     const long long* __restrict__ remap, const unsigned long long* __restrict__ n_dev,
                     unsigned long long max_work) {
    const unsigned long long n_work = *n_dev;
    if (n_work == 0 || n_work >= max_work) return;
    const NodeTableView t = stage_table(gt, table_in_smem != 0);
    typedef CtaCuda<kEvalThreads> W;
    W::phase() = 0u;
    unsigned long long nodes = 0;
    long long k;
    while (next_item_cta(queue, (long long)n_work, k)) {
        const long long idx = remap[k];
        const long long e = elem != nullptr ? elem[idx] : idx;
        const long long pi = elem != nullptr ? e / elem_cap
                             : (pidx != nullptr ? (long long)pidx[idx] : (broadcast ? 0 : idx));
        const DistParams& d = dists[dist[pi]];
        const double al = alpha[pi];
        if (point_class(d, al) != NK * 2 + (VISC ? 1 : 0)) continue;
        const double2 wv = w[e];
        const cx val = disp_eval_warp<W, NK, VISC>(t, d, cx(wv.x, wv.y), kx[pi], kz[pi], al, nullptr, nodes);
        if (threadIdx.x == 0) D[e] = make_double2(val.re, val.im);
    }
    if (node_evals != nullptr && threadIdx.x == 0 && nodes) atomicAdd(node_evals, nodes);
}
typedef decltype(&disp_eval_cta_kernel<0, false>) eval_cta_kernel_t;

// K1 fast path.  Most evaluations converge within the first few levels of the rule (a few hundred nodes: the
// exit predictor of psi_quad.cuh); for them a whole warp per point spends most of its time in per-level
// reductions and exit tests over a handful of nodes per lane.  Here ONE THREAD evaluates one point -- the same
// quadrature code instantiated with the one-lane policy, all partial sums in registers, no shuffles -- using
// levels 0..cap only; a point that has not converged by then (|Im w| tiny next to a resonance) is appended to
// `hard_idx` and evaluated from scratch by the warp-per-point kernel above.  Warps take 32 consecutive points,
// so the lanes walk the node table together (shared-memory broadcasts) and finish at similar levels.
template <int NK, bool VISC>
__global__ void __launch_bounds__(kEvalThreads, 1)
disp_eval_thread_kernel(NodeTableView gt, int table_in_smem, const DistParams* __restrict__ dists, long long n,
                        int broadcast, const int* __restrict__ dist, const double* __restrict__ kx,
                        const double* __restrict__ kz, const double* __restrict__ alpha,
                        const int* __restrict__ pidx, const double2* __restrict__ w, double2* __restrict__ D,
                        unsigned long long* __restrict__ queue, unsigned long long* __restrict__ node_evals,
                        const long long* __restrict__ elem, int elem_cap, int level_cap,
                        long long* __restrict__ hard_idx, unsigned long long* __restrict__ hard_count) {
    NodeTableView t = stage_table(gt, table_in_smem != 0);
    t.predict_exit = 2;
    t.level_cap = level_cap;
    const int lane = threadIdx.x & 31;
    unsigned long long nodes = 0;
    while (true) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(queue, 32ull);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= (unsigned long long)n) break;
        const unsigned long long idx = base + lane;
        if (idx >= (unsigned long long)n) continue;
        const long long e = elem != nullptr ? elem[idx] : (long long)idx;
        const long long pi = elem != nullptr ? e / elem_cap
                             : (pidx != nullptr ? (long long)pidx[idx] : (broadcast ? 0 : (long long)idx));
        const DistParams& d = dists[dist[pi]];
        const double al = alpha[pi];
        if (point_class(d, al) != NK * 2 + (VISC ? 1 : 0)) continue;
        const double2 wv = w[e];
        const cx val = disp_eval_warp<ThreadCuda, NK, VISC>(t, d, cx(wv.x, wv.y), kx[pi], kz[pi], al, nullptr, nodes);
        if (val.re != val.re || val.im != val.im) {          // not converged within the cap (or genuinely NaN)
            hard_idx[atomicAdd(hard_count, 1ull)] = (long long)idx;
        } else {
            D[e] = make_double2(val.re, val.im);
        }
    }
    if (node_evals != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nodes += __shfl_xor_sync(0xffffffffu, nodes, o);
        if (lane == 0 && nodes) atomicAdd(node_evals, nodes);
    }
}

typedef void (*thread_kernel_t)(NodeTableView, int, const DistParams*, long long, int, const int*, const double*,
                                const double*, const double*, const int*, const double2*, double2*,
                                unsigned long long*, unsigned long long*, const long long*, int, int, long long*,
                                unsigned long long*);
static thread_kernel_t thread_kernel_for(int cls) {
    switch (cls) {
    case 0: return disp_eval_thread_kernel<0, false>;
    case 1: return disp_eval_thread_kernel<0, true>;
    case 2: return disp_eval_thread_kernel<1, false>;
    case 3: return disp_eval_thread_kernel<1, true>;
    case 4: return disp_eval_thread_kernel<2, false>;
    case 5: return disp_eval_thread_kernel<2, true>;
    case 6: return disp_eval_thread_kernel<3, false>;
    case 7: return disp_eval_thread_kernel<3, true>;
    case 8: return disp_eval_thread_kernel<4, false>;
    default: return disp_eval_thread_kernel<4, true>;
    }
}

static eval_kernel_t eval_kernel_for(int cls) {
    switch (cls) {
    case 0: return disp_eval_kernel<0, false>;
    case 1: return disp_eval_kernel<0, true>;
    case 2: return disp_eval_kernel<1, false>;
    case 3: return disp_eval_kernel<1, true>;
    case 4: return disp_eval_kernel<2, false>;
    case 5: return disp_eval_kernel<2, true>;
    case 6: return disp_eval_kernel<3, false>;
    case 7: return disp_eval_kernel<3, true>;
    case 8: return disp_eval_kernel<4, false>;
    default: return disp_eval_kernel<4, true>;
    }
}

// phases of an item in the PSIMode pipeline (K4, below)
enum ModePhase { kPhaseSample = 0, kPhaseAnalyze = 1, kPhasePolish = 2, kPhaseDecide = 3, kPhaseZoom = 4,
                 kPhaseDone = 5 };

// K3 polish_kernel: one warp per root search; the whole two-stage secant polish of all its start
// values runs on the device, every D evaluated inline (psi_batch.cuh, polish_run).
// CTA = true: the whole block takes one root search at a time (every D evaluation spread over all its
// threads, the scalar secant logic executed redundantly by all of them).
template <int NK, bool VISC, bool CTA = false>
__global__ void __launch_bounds__(kEvalThreads, 1)
polish_kernel(NodeTableView gt, int table_in_smem, const DistParams* __restrict__ dists, int n_items,
              const int* __restrict__ pidx, const int* __restrict__ dist, const double* __restrict__ kx,
              const double* __restrict__ kz, const double* __restrict__ alpha,
              const double* __restrict__ dom, const int* __restrict__ off, double2* w0,
              const int* __restrict__ max_iter, double2* out, int* n_calls, int* max_conv, int* status,
              unsigned long long* queue, unsigned long long* counters, int stride, const int* __restrict__ cnt,
              const int* __restrict__ phase) {
    NodeTableView t = stage_table(gt, table_in_smem != 0);
    if (t.predict_exit == 1) t.predict_exit = 2;      // the looser predictor of K1's fast path (psi_quad.cuh)
    const int lane = CTA ? (int)threadIdx.x : (int)(threadIdx.x & 31);
    if (CTA) CtaCuda<kEvalThreads>::phase() = 0u;
    unsigned long long nodes = 0, evals = 0;
    long long idx;
    while (CTA ? next_item_cta(queue, n_items, idx) : next_item(queue, n_items, idx)) {
        if (phase != nullptr && phase[idx] != kPhasePolish) continue;       // mode pipeline: not this item's turn
        const int pi = pidx != nullptr ? pidx[idx] : (int)idx;
        const DistParams& d = dists[dist[pi]];
        const double al = alpha[pi];
        if (point_class(d, al) != NK * 2 + (VISC ? 1 : 0)) continue;
        typedef typename std::conditional<CTA, CtaCuda<kEvalThreads>, WarpCuda>::type Policy;
        WarpEvalClass<Policy, NK, VISC> f;
        f.t = &t; f.d = &d; f.kx = kx[pi]; f.kz = kz[pi]; f.alpha = al; f.nodes = 0; f.evals = 0;
        const RectDom rd{dom[4 * idx], dom[4 * idx + 1], dom[4 * idx + 2], dom[4 * idx + 3]};
        // start values: packed with offsets (batch API) or strided per item (mode pipeline)
        const long long o = off != nullptr ? (long long)off[idx] : idx * (long long)stride;
        const int n = off != nullptr ? off[idx + 1] - off[idx] : cnt[idx];
        int conv = 0, st = 0;
        const int calls = polish_run(f, rd, reinterpret_cast<cx*>(w0 + o), n, max_iter[idx],
                                     reinterpret_cast<cx*>(out + o), &conv, &st);
        if (lane == 0) { n_calls[idx] = calls; max_conv[idx] = conv; status[idx] = st; }
        nodes += f.nodes; evals += (unsigned long long)f.evals;
    }
    if (counters != nullptr && lane == 0) {
        if (nodes) atomicAdd(counters, nodes);
        if (evals) atomicAdd(counters + 1, evals);
    }
}

// K2 contour_kernel: one warp per closed contour; adaptive refinement passes and the winding sum on
// the device (psi_batch.cuh, contour_run).  scratch: 4*cap complex per warp of the grid.
template <int NK, bool VISC>
__global__ void __launch_bounds__(kEvalThreads, 1)
contour_kernel(NodeTableView gt, int table_in_smem, const DistParams* __restrict__ dists, int n_items,
               const int* __restrict__ dist, const double* __restrict__ kx, const double* __restrict__ kz,
               const double* __restrict__ alpha, const int* __restrict__ z_off,
               const double2* __restrict__ z_list, const double* __restrict__ tol,
               const double* __restrict__ max_step, const int* __restrict__ max_iter, double2* scratch,
               int cap, int* counts, int* status, int* n_points, unsigned long long* queue,
               unsigned long long* counters) {
    NodeTableView t = stage_table(gt, table_in_smem != 0);
    if (t.predict_exit == 1) t.predict_exit = 2;      // the looser predictor of K1's fast path (psi_quad.cuh)
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    cx* base = reinterpret_cast<cx*>(scratch) + (size_t)warp_global * 4 * cap;
    unsigned long long nodes = 0, evals = 0;
    long long idx;
    while (next_item(queue, n_items, idx)) {
        const DistParams& d = dists[dist[idx]];
        const double al = alpha[idx];
        if (point_class(d, al) != NK * 2 + (VISC ? 1 : 0)) continue;
        WarpEvalClass<WarpCuda, NK, VISC> f;
        f.t = &t; f.d = &d; f.kx = kx[idx]; f.kz = kz[idx]; f.alpha = al; f.nodes = 0; f.evals = 0;
        const int o = z_off[idx], n0 = z_off[idx + 1] - o;
        int cnt = -666, np = n0, st = kContourCapacity;
        if (n0 <= cap) {
            for (int i = 0; i < n0; ++i) base[i] = cx(z_list[o + i].x, z_list[o + i].y);
            st = contour_run(f, base, base + cap, base + 2 * cap, base + 3 * cap, n0, cap, tol[idx],
                             max_step[idx], max_iter[idx], &cnt, &np);
        }
        if (lane == 0) { counts[idx] = cnt; status[idx] = st; n_points[idx] = np; }
        nodes += f.nodes; evals += (unsigned long long)f.evals;
    }
    if (counters != nullptr && lane == 0) {
        if (nodes) atomicAdd(counters, nodes);
        if (evals) atomicAdd(counters + 1, evals);
    }
}

// ---- K4: the PSIMode.calculate pipeline.  Per-item state lives in device memory; every stage is a
// homogeneous kernel over all items (sampling -> K1 -> AAA analysis -> K3 polish -> decision), repeated
// once per zoom level.  Homogeneous stages keep the SMs' instruction caches and FP64 pipes busy; the
// single-warp form of the same search (mode_run in psi_mode.cuh) is what the CPU harness exercises.

struct ModeDev {            // device pointers of the per-item state (n items, cap samples, P start values)
    cx *Z, *F, *w0, *pol, *zoom, *roots;
    double* box_dx;
    int *M, *level, *upos, *phase, *n0, *max_iter, *n_zoom, *calls, *status, *n_out, *pcalls, *pconv, *pstat;
    int *conv_max;                      // largest pconv over the levels of a search
    int *mask_out;                      // [n * cap] support mask of the last analysis (null: not wanted)
    long long* elem;                    // request list for K1 (element indices into Z/F)
    unsigned long long* counters;       // [0] number of requests, [1] items still active
    int cap, P, max_roots;
};

// thread per item: draw the next batch of sample points (initial domain + guess domains, or the zoom
// boxes of the finished level) from the item's random stream and queue them for K1
__global__ void mode_sample_kernel(ModeDev m, int n_items, const double* __restrict__ domain,
                                   const int* __restrict__ iparams, const double* __restrict__ uni,
                                   const int* __restrict__ uni_off, const int* __restrict__ guess_off,
                                   const double2* __restrict__ guesses) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const int ph = m.phase[i];
    if (ph != kPhaseSample && ph != kPhaseZoom) return;
    const RectDom dom{domain[4 * i], domain[4 * i + 1], domain[4 * i + 2], domain[4 * i + 3]};
    const int ns = iparams[3 * i];
    const double* u = uni + uni_off[i];
    cx* Z = m.Z + (size_t)i * m.cap;
    int M = m.M[i], upos = m.upos[i];
    const int first = M;
    auto box = [&](double x0, double x1, double y0, double y1) {
        for (int k = 0; k < ns; ++k)
            Z[M + k] = cx(add_rn(x0, mul_rn(x1 - x0, u[upos + k])), add_rn(y0, mul_rn(y1 - y0, u[upos + ns + k])));
        M += ns; upos += 2 * ns;
    };
    auto clamped = [&](double sx, double sy, cx c) {
        double re = c.re, im = c.im;
        if (re < dom.xmin + 0.5 * sx) re = dom.xmin + 0.5 * sx;
        if (re > dom.xmax - 0.5 * sx) re = dom.xmax - 0.5 * sx;
        if (im < dom.ymin + 0.5 * sy) im = dom.ymin + 0.5 * sy;
        if (im > dom.ymax - 0.5 * sy) im = dom.ymax - 0.5 * sy;
        box(re - 0.5 * sx, re + 0.5 * sx, im - 0.5 * sy, im + 0.5 * sy);
    };
    if (ph == kPhaseSample) {
        box(dom.xmin, dom.xmax, dom.ymin, dom.ymax);
        for (int g = guess_off[i]; g < guess_off[i + 1]; ++g) {
            const double side = fmin(1.0e-3, fabs(guesses[g].y) * 4.0);
            clamped(side, side, cx(guesses[g].x, guesses[g].y));
        }
    } else {
        const double dx = m.box_dx[i];
        for (int q = 0; q < m.n_zoom[i]; ++q) {
            const cx c = m.zoom[4 * (size_t)i + q];
            clamped(dx, fmin(0.1, fabs(c.im)), c);
        }
        m.level[i] += 1;
    }
    const int n_new = M - first;
    m.calls[i] += n_new;
    m.M[i] = M; m.upos[i] = upos;
    m.phase[i] = kPhaseAnalyze;
    const unsigned long long pos = atomicAdd(m.counters, (unsigned long long)n_new);
    for (int k = 0; k < n_new; ++k) m.elem[pos + k] = (long long)i * m.cap + first + k;
}

// Size classes of the analysis stage: a search with M samples is analysed by 8 lanes (M <= 32: the first level of a
// plain map, 18 x 10 Loewner matrices), 16 lanes (M <= kAnClass1; unused by default) or a whole warp -- a property of
// the search at this zoom level alone, never of what else is in the batch (results do not depend on how a batch is
// cut).  Measured (1 B200): 8 lanes for M = 20 take the first level of the 256^2 map from 65 to 27 ms; 16 lanes for
// M = 40..100 gain 10 % on its second level but lose 35 % on the refinement sweeps of config 4, whose searches carry
// guess domains and end with more support points than 16 lanes hold in registers (profiles/k4_groups_r2.txt).
// Thread per search: the searches waiting for their analysis are appended to their class's list.
constexpr int kAnClass0 = 32, kAnClass1 = 32;
// (Appended warp by warp, the 32 searches of a warp in index order: neighbouring searches -- neighbouring pixels of a
// map, which behave alike -- stay neighbours on the list and end up in the same analysis warp.)
// With the 16-lane class off (class1 <= class0, the default) its list holds the LARGEST 32-lane searches (more than
// `huge` samples) instead; the 32-lane kernel takes them first, so that the longest analyses of a level start at
// once and the short ones fill in behind them instead of the other way round.
__global__ void mode_classify_kernel(ModeDev m, int n_items, int* lists, unsigned long long* counts, int class0,
                                     int class1, int huge) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int cls = -1;
    if (i < n_items && m.phase[i] == kPhaseAnalyze) {
        const int M = m.M[i];
        cls = M <= class0 ? 0 : (class1 > class0 ? (M <= class1 ? 1 : 2) : (M > huge ? 1 : 2));
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const unsigned members = __ballot_sync(0xffffffffu, cls == c);
        if (members == 0u) continue;
        unsigned long long base = 0;
        if (lane == __ffs(members) - 1) base = atomicAdd(counts + c, (unsigned long long)__popc(members));
        base = __shfl_sync(0xffffffffu, base, __ffs(members) - 1);
        if (cls == c) lists[(size_t)c * n_items + base + __popc(members & ((1u << lane) - 1u))] = i;
    }
}

// Analysis stage of K4.  G lanes per search (32/G searches per warp): AAA analysis of the current samples
// (mode_analyze, psi_mode.cuh) -> polish start values + zoom plan.  (Tried in round 2 and dropped: one kernel per
// phase of the state machine -- control / fits / poles / zeros -- in lock-step rounds over all searches, so that
// an SM runs one compact piece of code at a time.  0.208 s instead of 0.146 s on the 256^2 map: the fits are bound
// by the traffic of their trailing-matrix updates, not by instruction fetch, each round waits for its slowest
// search, and the per-search state no longer stays in L2; profiles/k4_phase_kernels_r2_launches.csv.)
template <int G>
__global__ void __launch_bounds__(256, 3)
mode_analyze_kernel(ModeDev m, int n_items, const double* __restrict__ domain, const int* __restrict__ iparams,
                    const double* __restrict__ dparams, const int* __restrict__ guess_off,
                    unsigned char* scratch, size_t scratch_per_group, int scratch_cap,
                    const int* __restrict__ list, const unsigned long long* __restrict__ n_list,
                    unsigned long long* queue, unsigned long long* n_polish, const int* __restrict__ list_first,
                    const unsigned long long* __restrict__ n_first) {
    typedef GroupCuda<G> W;
    const int lane = W::lane();
    const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    AAAScratch sc = aaa_carve(scratch + (size_t)group * scratch_per_group, scratch_cap);
    const long long n_head = list_first != nullptr ? (long long)*n_first : 0;     // the largest searches, taken first
    const long long n_work = n_head + (long long)*n_list; // searches of this launch's size class (mode_classify_kernel)
    long long k;
    while (next_item_group<W>(queue, n_work, k)) {
        const long long idx = k < n_head ? list_first[k] : list[k - n_head];
        ModeCfg cfg;
        cfg.dom = RectDom{domain[4 * idx], domain[4 * idx + 1], domain[4 * idx + 2], domain[4 * idx + 3]};
        cfg.n_sample = iparams[3 * idx]; cfg.max_zoom = iparams[3 * idx + 1]; cfg.max_secant = iparams[3 * idx + 2];
        cfg.tol = dparams[2 * idx]; cfg.clean_tol = dparams[2 * idx + 1];
        sc.Z = m.Z + (size_t)idx * m.cap;                 // analyse the item's own sample buffers in place
        sc.F = m.F + (size_t)idx * m.cap;
        sc.w0 = m.w0 + (size_t)idx * m.P;
        AAAState st;
        st.M = m.M[idx]; st.m = 0; st.r = 0; st.fmax = 0.0; st.tol = cfg.tol; st.clean_tol = cfg.clean_tol;
        st.dom = cfg.dom; st.memo_next = 0; st.slot = 0; st.v_m = 0;
        ModePlan plan;
        PROF_T0(t_an);
        const bool ok = mode_analyze<W>(cfg, m.level[idx], guess_off[idx + 1] - guess_off[idx], sc, st, plan);
        W::sync();
        PROF_ADD(12, t_an);
        PROF_MAX(22, ((unsigned long long)(clock64() - t_an) << 22) | (unsigned long long)(idx & 0x3fffff));
        if (m.mask_out != nullptr && ok)
            for (int i = lane; i < st.M; i += W::width) m.mask_out[(size_t)idx * m.cap + i] = sc.mask[i];
        if (lane == 0) {
            if (!ok) {
                m.status[idx] = kModeBadSamples; m.phase[idx] = kPhaseDone; m.n_out[idx] = 0;
            } else {
                m.n0[idx] = plan.n0; m.max_iter[idx] = plan.max_iter; m.n_zoom[idx] = plan.n_zoom;
                m.box_dx[idx] = plan.box_dx;
                for (int q = 0; q < plan.n_zoom; ++q) m.zoom[4 * (size_t)idx + q] = plan.zoom[q];
                m.pcalls[idx] = 0; m.pconv[idx] = 0; m.pstat[idx] = 0;
                m.phase[idx] = plan.n0 > 0 ? kPhasePolish : kPhaseDecide;
                if (plan.n0 > 0) atomicAdd(n_polish, 1ull);
            }
        }
    }
}

// thread per item: collect the polished roots, then finish or schedule the zoom boxes (psi_mode.py:312-331)
__global__ void mode_decide_kernel(ModeDev m, int n_items, const double* __restrict__ domain,
                                   const int* __restrict__ iparams) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const int ph = m.phase[i];
    if (ph != kPhasePolish && ph != kPhaseDecide) return;
    const RectDom dom{domain[4 * i], domain[4 * i + 1], domain[4 * i + 2], domain[4 * i + 3]};
    if (m.pstat[i] != 0) { m.status[i] = kModeZeroTol; m.phase[i] = kPhaseDone; m.n_out[i] = 0; return; }
    m.calls[i] += m.pcalls[i];
    if (ph == kPhasePolish && m.pconv[i] > m.conv_max[i]) m.conv_max[i] = m.pconv[i];
    int n_out = 0;
    bool grow = false;
    const int n0 = ph == kPhasePolish ? m.n0[i] : 0;
    for (int k = 0; k < n0; ++k) {
        const cx z = m.pol[(size_t)i * m.P + k];
        if (!dom.is_in(z)) continue;
        if (n_out < m.max_roots) m.roots[(size_t)i * m.max_roots + n_out] = z;
        ++n_out;
        if (z.im > 0.0) grow = true;
    }
    m.n_out[i] = n_out;
    if (grow || m.level[i] == iparams[3 * i + 1] || m.n_zoom[i] == 0) {
        m.phase[i] = kPhaseDone;
        m.status[i] = n_out > m.max_roots ? kModeTruncated : kModeOk;
    } else {
        m.phase[i] = kPhaseZoom;
        atomicAdd(m.counters + 1, 1ull);
    }
}

// thread per row: packed result records {n_roots, status, n_function_call, 0, roots...} (rows >= n_items zeroed)
__global__ void mode_pack_kernel(ModeDev m, int n_items, int rows, double* rec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const int width = 4 + 2 * m.max_roots;
    double* r = rec + (size_t)i * width;
    if (i >= n_items) { for (int k = 0; k < width; ++k) r[k] = 0.0; return; }
    const int n = m.n_out[i];
    r[0] = n; r[1] = m.status[i]; r[2] = m.calls[i]; r[3] = 0.0;
    for (int k = 0; k < m.max_roots; ++k) {
        const cx z = k < n ? m.roots[(size_t)i * m.max_roots + k] : cx(0.0, 0.0);
        r[4 + 2 * k] = z.re; r[5 + 2 * k] = z.im;
    }
}

#define PSI_CLASS_TABLE(NAME)                                                                      \
    {NAME<0, false>, NAME<0, true>, NAME<1, false>, NAME<1, true>, NAME<2, false>, NAME<2, true>,    \
     NAME<3, false>, NAME<3, true>, NAME<4, false>, NAME<4, true>}
typedef decltype(&polish_kernel<0, false, false>) polish_kernel_t;
static const eval_cta_kernel_t kEvalCtaKernels[kNumClasses] = PSI_CLASS_TABLE(disp_eval_cta_kernel);
#define PSI_CLASS_TABLE_CTA(NAME)                                                                          \
    {NAME<0, false, true>, NAME<0, true, true>, NAME<1, false, true>, NAME<1, true, true>, NAME<2, false, true>, \
     NAME<2, true, true>, NAME<3, false, true>, NAME<3, true, true>, NAME<4, false, true>, NAME<4, true, true>}
typedef decltype(&contour_kernel<0, false>) contour_kernel_t;
static const polish_kernel_t kPolishKernels[kNumClasses] = PSI_CLASS_TABLE(polish_kernel);
static const polish_kernel_t kPolishKernelsCta[kNumClasses] = PSI_CLASS_TABLE_CTA(polish_kernel);
// K3 takes one block per root search when a launch has at most this many searches (PSI_K3_CTA_ITEMS)
static long long k3_cta_items() {
    const char* v = getenv("PSI_K3_CTA_ITEMS");       // read per call: tests switch it
    return v ? atoll(v) : 4096;
}
static const contour_kernel_t kContourKernels[kNumClasses] = PSI_CLASS_TABLE(contour_kernel);

// DFMA-pipe peak: 8 independent FMA chains per thread, all in registers.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5,
           a6 = seed + 6, a7 = seed + 7;
    const double m = 1.0 + 1e-9 * threadIdx.x, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* psi_version(void) { return "psi_b200 0.1 (sm_100a)"; }
const char* psi_last_error(void) { return g_err.c_str(); }

int psi_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int psi_ctx_create(psi_ctx** out, int device, const double* xj, const double* wj, int n_nodes,
                   int max_level, double eps) {
    if (!out || !xj || !wj || n_nodes <= 0 || max_level < 1 || max_level > 20 || !(eps > 0))
        return fail(PSI_ERR_ARG, "psi_ctx_create: bad argument");
    int ndev = psi_device_count();
    if (ndev <= 0)
        return fail(PSI_ERR_NODEVICE, "no CUDA device visible; libpsi_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(PSI_ERR_ARG, "psi_ctx_create: bad device index");
    CK(cudaSetDevice(device));
    psi_ctx* c = new psi_ctx();
    c->device = device;
    CK(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    CK(cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_k1, cudaEventDisableTiming));
    c->n_nodes = n_nodes; c->max_m = max_level; c->eps = eps;
    {
        const char* rule = getenv("PSI_B200_EXIT_RULE");          // "reference" = the literal rule of tanhsinh.py
        c->predict_exit = (rule && std::string(rule) == "reference") ? 0 : 1;
        const char* cap = getenv("PSI_K1_FAST_CAP");               // 0 switches the thread-per-point fast path off
        if (cap) c->fast_cap = atoi(cap);
    }
    // level-major reordering of the caller's table
    std::vector<node_t> lv;
    std::vector<int> off;
    if (!build_level_major(xj, wj, n_nodes, max_level, lv, off)) {
        delete c;
        return fail(PSI_ERR_ARG, "node table reorder failed");
    }
    CK(cudaMalloc(&c->d_nodes, sizeof(node_t) * n_nodes));
    CK(cudaMalloc(&c->d_level_off, sizeof(int) * off.size()));
    CK(cudaMemcpy(c->d_nodes, lv.data(), sizeof(node_t) * n_nodes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_level_off, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice));
    c->table_smem = sizeof(node_t) * (size_t)n_nodes + sizeof(int) * off.size();
    c->table_in_smem = c->table_smem + 4096 <= (size_t)c->smem_optin;
    if (c->table_in_smem) {
        for (int cls = 0; cls < kNumClasses; ++cls)
            CK(cudaFuncSetAttribute(eval_kernel_for(cls), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)c->table_smem));
        for (int cls = 0; cls < kNumClasses; ++cls) {
            CK(cudaFuncSetAttribute(thread_kernel_for(cls), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)c->table_smem));
            CK(cudaFuncSetAttribute(kEvalCtaKernels[cls], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)c->table_smem));
        }
        for (int cls = 0; cls < kNumClasses; ++cls) {
            CK(cudaFuncSetAttribute(kPolishKernels[cls], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->table_smem));
            CK(cudaFuncSetAttribute(kPolishKernelsCta[cls], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->table_smem));
            CK(cudaFuncSetAttribute(kContourKernels[cls], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->table_smem));
        }
    }
    CK(cudaMalloc(&c->d_dists, sizeof(DistParams) * kMaxDists));
    CK(cudaMalloc(&c->d_counter, sizeof(unsigned long long) * 16));
    CK(cudaMemset(c->d_counter, 0, sizeof(unsigned long long) * 16));
    *out = c;
    return PSI_OK;
}

void psi_ctx_destroy(psi_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->d_nodes); cudaFree(c->d_level_off); cudaFree(c->d_dists); cudaFree(c->d_counter);
    cudaFree(c->d_stage); cudaFree(c->d_contour_scratch); cudaFree(c->d_mode_scratch); cudaFree(c->d_hard);
    for (double* t : c->d_tables) cudaFree(t);
    cudaStreamDestroy(c->stream);
    if (c->ev_k1) cudaEventDestroy(c->ev_k1);
    if (c->h_in) cudaFreeHost(c->h_in);
    if (c->h_out) cudaFreeHost(c->h_out);
    if (c->stream_copy) cudaStreamDestroy(c->stream_copy);
    if (c->ev_thread) cudaEventDestroy(c->ev_thread);
    if (c->ev_copied) cudaEventDestroy(c->ev_copied);
    if (c->d_patch) cudaFree(c->d_patch);
    if (c->h_patch) cudaFreeHost(c->h_patch);
    delete c;
}

int psi_ctx_sm_count(psi_ctx* c) { return c ? c->sm_count : 0; }
int psi_ctx_set_exit_rule(psi_ctx* c, int predict) {
    if (!c) return fail(PSI_ERR_ARG, "null context");
    c->predict_exit = predict ? 1 : 0;
    return PSI_OK;
}
int psi_ctx_get_exit_rule(psi_ctx* c) { return c ? c->predict_exit : 0; }
long long psi_ctx_launch_count(psi_ctx* c) { return c ? c->launches : 0; }

int psi_ctx_synchronize(psi_ctx* c) {
    if (!c) return fail(PSI_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return PSI_OK;
}

static int dist_add_common(psi_ctx* c, const psi_dist_desc* s, const psi_table_desc* tb);

int psi_dist_add(psi_ctx* c, const psi_dist_desc* s) {
    if (!c || !s) return fail(PSI_ERR_ARG, "psi_dist_add: null argument");
    if (s->kind < 0 || s->kind > 2) return fail(PSI_ERR_ARG, "psi_dist_add: unknown kind");
    return dist_add_common(c, s, nullptr);
}

int psi_dist_add_table(psi_ctx* c, const psi_dist_desc* s, const psi_table_desc* tb) {
    if (!c || !s || !tb || !tb->coef) return fail(PSI_ERR_ARG, "psi_dist_add_table: null argument");
    if (s->kind != PSI_KIND_TABLE) return fail(PSI_ERR_ARG, "psi_dist_add_table: desc->kind must be PSI_KIND_TABLE");
    if (tb->n_seg < 1 || tb->n_seg > PSI_MAX_TABLE_SEGS) return fail(PSI_ERR_ARG, "psi_dist_add_table: bad n_seg");
    for (int k = 0; k < tb->n_seg; ++k)
        if (tb->n_cells[k] < 1 || !(tb->brk[k + 1] > tb->brk[k]))
            return fail(PSI_ERR_ARG, "psi_dist_add_table: pieces must be non-empty and ascending");
    return dist_add_common(c, s, tb);
}

static int dist_add_common(psi_ctx* c, const psi_dist_desc* s, const psi_table_desc* tb) {
    if (s->n_poles < 0 || s->n_poles > PSI_MAX_POLES)
        return fail(PSI_ERR_CAPACITY, "psi_dist_add: too many size-density poles");
    if ((int)c->h_dists.size() >= kMaxDists) return fail(PSI_ERR_CAPACITY, "too many distributions");
    if (!(s->tau_min > 0) || !(s->tau_max > 0) || !(s->rhod != 0))
        return fail(PSI_ERR_ARG, "psi_dist_add: bad Stokes range or rhod");
    CK(cudaSetDevice(c->device));
    DistParams d = dist_from_desc(*s);
    if (tb) {
        size_t cells = 0;
        d.tab_nseg = tb->n_seg;
        for (int k = 0; k < tb->n_seg; ++k) {
            d.tab_off[k] = (int)cells; d.tab_n[k] = tb->n_cells[k]; d.tab_log[k] = tb->log_mode[k];
            d.tab_brk[k] = tb->brk[k];
            d.tab_invh[k] = tb->n_cells[k] / (tb->brk[k + 1] - tb->brk[k]);
            cells += (size_t)tb->n_cells[k];
        }
        d.tab_brk[tb->n_seg] = tb->brk[tb->n_seg];
        double* dev = nullptr;
        CK(cudaMalloc(&dev, 32 * cells));
        CK(cudaMemcpy(dev, tb->coef, 32 * cells, cudaMemcpyHostToDevice));
        c->d_tables.push_back(dev);
        d.tab = dev;
    }
    const int id = (int)c->h_dists.size();
    CK(cudaMemcpyAsync(c->d_dists + id, &d, sizeof(d), cudaMemcpyHostToDevice, c->stream));
    background_kernel<<<1, 32, 0, c->stream>>>(table_view(c), c->d_dists, id);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&d, c->d_dists + id, sizeof(d), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->h_dists.push_back(d);
    return id;
}

int psi_point_class(psi_ctx* c, int dist, double alpha) {
    if (!c || dist < 0 || dist >= (int)c->h_dists.size())
        return fail(PSI_ERR_ARG, "psi_point_class: bad argument");
    return point_class(c->h_dists[dist], alpha);
}

int psi_dist_get_background(psi_ctx* c, int dist, double out[5]) {
    if (!c || !out || dist < 0 || dist >= (int)c->h_dists.size())
        return fail(PSI_ERR_ARG, "psi_dist_get_background: bad argument");
    const DistParams& d = c->h_dists[dist];
    out[0] = d.j0; out[1] = d.j1; out[2] = d.denom; out[3] = d.vgx; out[4] = d.vgy;
    return PSI_OK;
}

// class_mask: bit c set = the batch may contain points of work class c (see point_class()).
static int launch_disp_eval(psi_ctx* c, long long n, int broadcast, const int* dist, const double* kx,
                            const double* kz, const double* alpha, const double* w, double* D,
                            double* M, unsigned long long* node_evals_dev, unsigned class_mask,
                            cudaStream_t st, const int* pidx = nullptr, const long long* elem = nullptr,
                            int elem_cap = 1, cudaEvent_t after_thread_stage = nullptr) {
    if (n == 0) return PSI_OK;
    const int warps = kEvalThreads / 32;
    long long want = (n + warps - 1) / warps;
    int grid = (int)(want < c->sm_count ? want : c->sm_count);
    NodeTableView t = table_view(c);
    if ((class_mask & 0x3ffu) == 0) class_mask = 0x3ffu;
    if (c->ev_k1_recorded) CK(cudaStreamWaitEvent(st, c->ev_k1, 0));
    // fast path: thread per point first, the warp kernel only for what that leaves (needs the predictor rule,
    // no M output, and enough points to fill the machine with one thread each)
    const bool fast = c->predict_exit != 0 && c->fast_cap > 0 && M == nullptr && n >= 4096;
    if (fast && (size_t)n > c->hard_cap) {
        if (c->d_hard) CK(cudaFree(c->d_hard));
        c->d_hard = nullptr; c->hard_cap = 0;
        CK(cudaMalloc(&c->d_hard, sizeof(long long) * (size_t)n));
        c->hard_cap = (size_t)n;
    }
    for (int cls = 0; cls < kNumClasses; ++cls) {
        if (!(class_mask & (1u << cls))) continue;
        CK(cudaMemsetAsync(c->d_counter, 0, sizeof(unsigned long long), st));
        eval_kernel_t k = eval_kernel_for(cls);
        if (fast) {
            CK(cudaMemsetAsync(c->d_counter + 9, 0, sizeof(unsigned long long), st));
            const long long want_t = (n + kEvalThreads - 1) / kEvalThreads;
            const int grid_t = (int)(want_t < c->sm_count ? want_t : c->sm_count);
            thread_kernel_for(cls)<<<grid_t, kEvalThreads, c->table_in_smem ? c->table_smem : 0, st>>>(
                t, c->table_in_smem ? 1 : 0, c->d_dists, n, broadcast, dist, kx, kz, alpha, pidx, (const double2*)w,
                (double2*)D, c->d_counter, node_evals_dev, elem, elem_cap, c->fast_cap, c->d_hard, c->d_counter + 9);
            c->launches++;
            CK(cudaGetLastError());
            if (after_thread_stage) CK(cudaEventRecord(after_thread_stage, st));
            // what the thread kernel left: a block per point when it is a handful, a warp per point otherwise
            static const long long few_sms = getenv("PSI_K1_CTA_LIST") ? atoll(getenv("PSI_K1_CTA_LIST")) : 4;
            const unsigned long long few = (unsigned long long)few_sms * (unsigned long long)c->sm_count;
            CK(cudaMemsetAsync(c->d_counter, 0, sizeof(unsigned long long), st));
            kEvalCtaKernels[cls]<<<c->sm_count, kEvalThreads, c->table_in_smem ? c->table_smem : 0, st>>>(
                t, c->table_in_smem ? 1 : 0, c->d_dists, broadcast, dist, kx, kz, alpha, pidx, (const double2*)w,
                (double2*)D, c->d_counter, node_evals_dev, elem, elem_cap, c->d_hard, c->d_counter + 9, few);
            CK(cudaMemsetAsync(c->d_counter, 0, sizeof(unsigned long long), st));
            k<<<c->sm_count, kEvalThreads, c->table_in_smem ? c->table_smem : 0, st>>>(
                t, c->table_in_smem ? 1 : 0, c->d_dists, n, broadcast, dist, kx, kz, alpha, pidx, (const double2*)w,
                (double2*)D, nullptr, c->d_counter, node_evals_dev, elem, elem_cap, c->d_hard, c->d_counter + 9, few);
            c->launches += 2;
            CK(cudaGetLastError());
            continue;
        }
        k<<<grid, kEvalThreads, c->table_in_smem ? c->table_smem : 0, st>>>(
            t, c->table_in_smem ? 1 : 0, c->d_dists, n, broadcast, dist, kx, kz, alpha, pidx, (const double2*)w, (double2*)D,
            (double2*)M, c->d_counter, node_evals_dev, elem, elem_cap, nullptr, nullptr, 0ull);
        c->launches++;
        CK(cudaGetLastError());
    }
    if (after_thread_stage && !fast) CK(cudaEventRecord(after_thread_stage, st));     // (no thread stage: after everything)
    CK(cudaEventRecord(c->ev_k1, st));
    c->ev_k1_recorded = true;
    return PSI_OK;
}

int psi_disp_eval_dev(psi_ctx* c, long long n, int broadcast, const int* dist, const double* kx,
                      const double* kz, const double* alpha, const double* w, double* D, double* M,
                      unsigned long long* node_evals_dev, unsigned class_mask, void* stream) {
    if (!c || n < 0 || !dist || !kx || !kz || !alpha || (n > 0 && (!w || !D)))
        return fail(PSI_ERR_ARG, "psi_disp_eval_dev: bad argument");
    CK(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    return launch_disp_eval(c, n, broadcast, dist, kx, kz, alpha, w, D, M, node_evals_dev, class_mask,
                            st);
}

static int ensure_stage(psi_ctx* c, size_t bytes) {
    if (bytes <= c->stage_bytes) return PSI_OK;
    if (c->d_stage) CK(cudaFree(c->d_stage));
    c->d_stage = nullptr; c->stage_bytes = 0;
    CK(cudaMalloc(&c->d_stage, bytes));
    c->stage_bytes = bytes;
    return PSI_OK;
}

// (index, value) of the points the leftover stage evaluated, for the early read-back of psi_disp_eval_host
__global__ void gather_hard_kernel(const long long* __restrict__ hard, const unsigned long long* __restrict__ n_hard,
                                   const double2* __restrict__ D, long long cap, long long* __restrict__ idx_out,
                                   double2* __restrict__ val_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long n = *n_hard;
    if (i < cap && (unsigned long long)i < n) { idx_out[i] = hard[i]; val_out[i] = D[hard[i]]; }
}

// memcpy on all host cores (a single core moves ~10 GB/s, the 16 cores of a GPU box 40+)
static void copy_parallel(void* dst, const void* src, size_t bytes) {
    const size_t chunk = (size_t)1 << 20;
    const long long nchunk = (long long)((bytes + chunk - 1) / chunk);
#pragma omp parallel for schedule(static) if (nchunk >= 4)
    for (long long k = 0; k < nchunk; ++k) {
        const size_t lo = (size_t)k * chunk, len = std::min(chunk, bytes - lo);
        memcpy((char*)dst + lo, (const char*)src + lo, len);
    }
}

static bool is_pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

static int ensure_pinned(void** buf, size_t* have, size_t need) {
    if (need <= *have) return PSI_OK;
    if (*buf) cudaFreeHost(*buf);
    *buf = nullptr; *have = 0;
    if (cudaHostAlloc(buf, need, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return PSI_ERR_CUDA; }
    *have = need;
    return PSI_OK;
}

int psi_disp_eval_host(psi_ctx* c, long long n, int broadcast, const int* dist, const double* kx,
                       const double* kz, const double* alpha, const double* w, double* D, double* M,
                       unsigned long long* node_evals) {
    if (!c || n < 0 || !dist || !kx || !kz || !alpha || (n > 0 && (!w || !D)))
        return fail(PSI_ERR_ARG, "psi_disp_eval_host: bad argument");
    if (node_evals) *node_evals = 0;
    if (n == 0) return PSI_OK;
    CK(cudaSetDevice(c->device));
    const long long np = broadcast ? 1 : n;
    unsigned class_mask = 0;
    for (long long i = 0; i < np; ++i) {
        if (dist[i] < 0 || dist[i] >= (int)c->h_dists.size())
            return fail(PSI_ERR_ARG, "psi_disp_eval_host: distribution id out of range");
        class_mask |= 1u << point_class(c->h_dists[dist[i]], alpha[i]);
    }
    // staging layout: w | D | M | kx | kz | alpha | dist   (all 16-byte aligned)
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_w = 0, o_D = al(o_w + 16 * n), o_M = al(o_D + 16 * n),
                 o_kx = al(o_M + (M ? 112 * n : 0)), o_kz = al(o_kx + 8 * np),
                 o_al = al(o_kz + 8 * np), o_ds = al(o_al + 8 * np), total = al(o_ds + 4 * np);
    int rc = ensure_stage(c, total);
    if (rc != PSI_OK) return rc;
    char* base = (char*)c->d_stage;
    cudaStream_t st = c->stream;
    // Pageable caller memory (numpy arrays): the driver's own staged copies run at 6-8 GB/s, i.e. 4 ms for the
    // 2 x 16.8 MB of a 1,048,576-point call whose kernels take 3.5 ms.  Large transfers go through pinned bounce
    // buffers of the context instead, filled and emptied by all host cores.  (Splitting the call into pieces so that
    // copies overlap the kernels was measured and dropped: every K1 launch has ~0.25 ms of fixed cost -- its tail of
    // hard points -- and 4 / 8 pieces took e2e from 258 M to 225 M / 177 M evals/s.)
    static const size_t bounce_min = getenv("PSI_BOUNCE_MIN") ? (size_t)atoll(getenv("PSI_BOUNCE_MIN")) : ((size_t)1 << 20);
    const bool bounce_in = 16 * (size_t)n >= bounce_min && !is_pinned_host(w) &&
                           ensure_pinned(&c->h_in, &c->h_in_bytes, 16 * (size_t)n) == PSI_OK;
    const bool bounce_out = 16 * (size_t)n >= bounce_min && !is_pinned_host(D) &&
                            ensure_pinned(&c->h_out, &c->h_out_bytes, 16 * (size_t)n) == PSI_OK;
    if (bounce_in || bounce_out) CK(cudaStreamSynchronize(st));         // an earlier call may still own the buffers
    if (bounce_in) copy_parallel(c->h_in, w, 16 * (size_t)n);
    const double* w_src = bounce_in ? (const double*)c->h_in : w;
    CK(cudaMemcpyAsync(base + o_w, w_src, 16 * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_kx, kx, 8 * np, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_kz, kz, 8 * np, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_al, alpha, 8 * np, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_ds, dist, 4 * np, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(c->d_counter + 1, 0, sizeof(unsigned long long), st));
    // Early read-back (one work class, fast path, a large batch): D leaves on a copy stream as soon as the
    // thread-per-point stage is done -- under the leftover stage, which rewrites a few hundred entries -- and those
    // entries follow as (index, value) patches.  Hides the 0.3 ms of the 16.8 MB copy of a 1,048,576-point call.
    static const long long early_min = getenv("PSI_K1_EARLY_COPY") ? atoll(getenv("PSI_K1_EARLY_COPY")) : 262144;
    const bool early = early_min > 0 && n >= early_min && M == nullptr && c->predict_exit != 0 && c->fast_cap > 0 &&
                       (class_mask & (class_mask - 1)) == 0 && (bounce_out || is_pinned_host(D));
    if (early && !c->stream_copy) {
        CK(cudaStreamCreateWithFlags(&c->stream_copy, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c->ev_thread, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_copied, cudaEventDisableTiming));
        CK(cudaMalloc(&c->d_patch, 24 * (size_t)psi_ctx::kPatchCap));
        CK(cudaHostAlloc(&c->h_patch, 24 * (size_t)psi_ctx::kPatchCap + 8, cudaHostAllocDefault));
    }
    rc = launch_disp_eval(c, n, broadcast, (const int*)(base + o_ds), (const double*)(base + o_kx),
                          (const double*)(base + o_kz), (const double*)(base + o_al),
                          (const double*)(base + o_w), (double*)(base + o_D),
                          M ? (double*)(base + o_M) : nullptr, c->d_counter + 1, class_mask, st, nullptr, nullptr, 1,
                          early ? c->ev_thread : nullptr);
    if (rc != PSI_OK) return rc;
    if (early) {
        double* dst = bounce_out ? (double*)c->h_out : D;
        CK(cudaStreamWaitEvent(c->stream_copy, c->ev_thread, 0));
        CK(cudaMemcpyAsync(dst, base + o_D, 16 * n, cudaMemcpyDeviceToHost, c->stream_copy));
        CK(cudaEventRecord(c->ev_copied, c->stream_copy));
        const long long cap = psi_ctx::kPatchCap;
        long long* d_idx = (long long*)c->d_patch;
        double2* d_val = (double2*)(c->d_patch + 8 * cap);
        gather_hard_kernel<<<(int)((cap + 255) / 256), 256, 0, st>>>(c->d_hard, c->d_counter + 9,
                                                                    (const double2*)(base + o_D), cap, d_idx, d_val);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(c->h_patch, c->d_patch, 24 * (size_t)cap, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(c->h_patch + 24 * cap, c->d_counter + 9, 8, cudaMemcpyDeviceToHost, st));
        unsigned long long ne = 0;
        CK(cudaMemcpyAsync(&ne, c->d_counter + 1, sizeof(ne), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaStreamSynchronize(c->stream_copy));
        const unsigned long long n_hard = *(const unsigned long long*)(c->h_patch + 24 * cap);
        if (n_hard > (unsigned long long)cap) {            // more leftover points than patches: the plain copy after all
            CK(cudaMemcpyAsync(dst, base + o_D, 16 * n, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
        } else {
            const long long* h_idx = (const long long*)c->h_patch;
            const double* h_val = (const double*)(c->h_patch + 8 * cap);
            for (unsigned long long k = 0; k < n_hard; ++k) {
                dst[2 * h_idx[k]] = h_val[2 * k];
                dst[2 * h_idx[k] + 1] = h_val[2 * k + 1];
            }
        }
        if (bounce_out) copy_parallel(D, c->h_out, 16 * (size_t)n);
        if (node_evals) *node_evals = ne;
        return PSI_OK;
    }
    CK(cudaMemcpyAsync(bounce_out ? (double*)c->h_out : D, base + o_D, 16 * n, cudaMemcpyDeviceToHost, st));
    if (M) CK(cudaMemcpyAsync(M, base + o_M, 112 * n, cudaMemcpyDeviceToHost, st));
    unsigned long long ne = 0;
    CK(cudaMemcpyAsync(&ne, c->d_counter + 1, sizeof(ne), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (bounce_out) copy_parallel(D, c->h_out, 16 * (size_t)n);
    if (node_evals) *node_evals = ne;
    return PSI_OK;
}

}  // extern "C"

int psi_internal_eval_indexed(psi_ctx* c, long long n, const int* pidx, int n_params, const int* dist,
                              const double* kx, const double* kz, const double* alpha, const double* w,
                              double* D) {
    if (n == 0) return PSI_OK;
    CK(cudaSetDevice(c->device));
    unsigned class_mask = 0;
    for (int i = 0; i < n_params; ++i) {
        if (dist[i] < 0 || dist[i] >= (int)c->h_dists.size())
            return fail(PSI_ERR_ARG, "distribution id out of range");
        class_mask |= 1u << point_class(c->h_dists[dist[i]], alpha[i]);
    }
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t np = (size_t)n_params;
    const size_t o_w = 0, o_D = al(16 * (size_t)n), o_px = al(o_D + 16 * (size_t)n), o_kx = al(o_px + 4 * (size_t)n),
                 o_kz = al(o_kx + 8 * np), o_al = al(o_kz + 8 * np), o_ds = al(o_al + 8 * np),
                 total = al(o_ds + 4 * np);
    int rc = ensure_stage(c, total);
    if (rc != PSI_OK) return rc;
    char* base = (char*)c->d_stage;
    cudaStream_t st = c->stream;
    CK(cudaMemcpyAsync(base + o_w, w, 16 * (size_t)n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_px, pidx, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_kx, kx, 8 * np, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_kz, kz, 8 * np, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_al, alpha, 8 * np, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_ds, dist, 4 * np, cudaMemcpyHostToDevice, st));
    rc = launch_disp_eval(c, n, 0, (const int*)(base + o_ds), (const double*)(base + o_kx),
                          (const double*)(base + o_kz), (const double*)(base + o_al),
                          (const double*)(base + o_w), (double*)(base + o_D), nullptr, c->d_counter + 2,
                          class_mask, st, (const int*)(base + o_px));
    if (rc != PSI_OK) return rc;
    CK(cudaMemcpyAsync(D, base + o_D, 16 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PSI_OK;
}

// stage a list of host arrays contiguously on the device; returns device pointers in the same order
struct StageItem { const void* src; size_t bytes; };
static int stage_upload(psi_ctx* c, const StageItem* items, int n, size_t extra, char** dptr, char** extra_ptr) {
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t total = 0;
    for (int i = 0; i < n; ++i) total += al(items[i].bytes);
    total += al(extra);
    int rc = ensure_stage(c, total);
    if (rc != PSI_OK) return rc;
    char* p = (char*)c->d_stage;
    for (int i = 0; i < n; ++i) {
        dptr[i] = p;
        if (items[i].src && items[i].bytes)
            CK(cudaMemcpyAsync(p, items[i].src, items[i].bytes, cudaMemcpyHostToDevice, c->stream));
        p += al(items[i].bytes);
    }
    if (extra_ptr) *extra_ptr = p;
    return PSI_OK;
}

static unsigned class_mask_of(psi_ctx* c, int n_params, const int* dist, const double* alpha, bool* ok) {
    unsigned m = 0;
    *ok = true;
    for (int i = 0; i < n_params; ++i) {
        if (dist[i] < 0 || dist[i] >= (int)c->h_dists.size()) { *ok = false; return 0; }
        m |= 1u << point_class(c->h_dists[dist[i]], alpha[i]);
    }
    return m;
}

int psi_internal_polish(psi_ctx* c, int n_items, const int* pidx, int n_params, const int* dist,
                        const double* kx, const double* kz, const double* alpha, const double* dom,
                        const int* off, double* w0, const int* max_iter, double* out, int* n_calls,
                        int* max_conv, int* status, long long* evals) {
    if (evals) *evals = 0;
    if (n_items == 0) return PSI_OK;
    CK(cudaSetDevice(c->device));
    bool ok;
    const unsigned mask = class_mask_of(c, n_params, dist, alpha, &ok);
    if (!ok) return fail(PSI_ERR_ARG, "distribution id out of range");
    const size_t nw = (size_t)off[n_items], ni = (size_t)n_items, np = (size_t)n_params;
    StageItem it[] = {{pidx, 4 * ni}, {dist, 4 * np}, {kx, 8 * np}, {kz, 8 * np}, {alpha, 8 * np},
                      {dom, 32 * ni}, {off, 4 * (ni + 1)}, {w0, 16 * nw}, {max_iter, 4 * ni},
                      {nullptr, 16 * nw}, {nullptr, 4 * ni}, {nullptr, 4 * ni}, {nullptr, 4 * ni}};
    char* d[13];
    int rc = stage_upload(c, it, 13, 0, d, nullptr);
    if (rc != PSI_OK) return rc;
    cudaStream_t st = c->stream;
    const int warps = kEvalThreads / 32;
    const int grid = std::min(c->sm_count, (n_items + warps - 1) / warps);
    CK(cudaMemsetAsync(c->d_counter + 4, 0, 2 * sizeof(unsigned long long), st));
    NodeTableView t = table_view(c);
    for (int cls = 0; cls < kNumClasses; ++cls) {
    if (!(mask & (1u << cls))) continue;
    CK(cudaMemsetAsync(c->d_counter, 0, sizeof(unsigned long long), st));
    const bool k3_cta = n_items <= k3_cta_items();
    (k3_cta ? kPolishKernelsCta : kPolishKernels)[cls]<<<k3_cta ? std::min(c->sm_count, n_items) : grid, kEvalThreads,
                                                         c->table_in_smem ? c->table_smem : 0, st>>>(
        t, c->table_in_smem ? 1 : 0, c->d_dists, n_items, (const int*)d[0], (const int*)d[1],
        (const double*)d[2], (const double*)d[3], (const double*)d[4], (const double*)d[5], (const int*)d[6],
        (double2*)d[7], (const int*)d[8], (double2*)d[9], (int*)d[10], (int*)d[11], (int*)d[12], c->d_counter,
        c->d_counter + 4, 0, nullptr, nullptr);
    c->launches++;
    CK(cudaGetLastError());
    }
    unsigned long long cnt[2] = {0, 0};
    CK(cudaMemcpyAsync(out, d[9], 16 * nw, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(n_calls, d[10], 4 * ni, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(max_conv, d[11], 4 * ni, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(status, d[12], 4 * ni, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(cnt, c->d_counter + 4, sizeof(cnt), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (evals) *evals = (long long)cnt[1];
    return PSI_OK;
}

int psi_internal_mode_device(psi_ctx* c, int n_items, const int* dist, const double* kx, const double* kz,
                             const double* alpha, const double* domain, const int* iparams,
                             const double* dparams, const double* uni, long long n_uni, const int* uni_off,
                             const int* guess_off, const double* guesses, int cap, int max_roots,
                             double* roots, int* n_roots, int* n_calls, int* status, long long* stats,
                             const PsiModeExtra* extra) {
    const auto wall0 = std::chrono::steady_clock::now();
    auto wall_us = [&]() {
        return (long long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - wall0).count();
    };
    long long us_setup = 0;
    if (n_items == 0) {
        if (extra && extra->rec_dev) {                      // a rank without work still contributes (zero) records
            CK(cudaSetDevice(c->device));
            const size_t bytes = 8 * (size_t)std::max(extra->rec_rows, 1) * (4 + 2 * max_roots);
            int rc0 = ensure_stage(c, bytes);
            if (rc0 != PSI_OK) return rc0;
            CK(cudaMemsetAsync(c->d_stage, 0, bytes, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            *extra->rec_dev = (double*)c->d_stage;
        }
        return PSI_OK;
    }
    CK(cudaSetDevice(c->device));
    bool ok;
    const unsigned mask = class_mask_of(c, n_items, dist, alpha, &ok);
    if (!ok) return fail(PSI_ERR_ARG, "distribution id out of range");
    int max_zoom = 0, max_req = 0;
    for (int i = 0; i < n_items; ++i) {
        const int ns = iparams[3 * i], ng = guess_off[i + 1] - guess_off[i];
        if (ns * (1 + ng + 4 * iparams[3 * i + 1]) > cap) return fail(PSI_ERR_CAPACITY, "sample capacity too small");
        max_zoom = std::max(max_zoom, iparams[3 * i + 1]);
        max_req = std::max(max_req, ns * std::max(1 + ng, 4));
    }
    const size_t ni = (size_t)n_items, ng = (size_t)guess_off[n_items];
    const int P = cap / 2 + 2;
    // analysis stage: 8 / 16 / 32 lanes per search by its sample count (mode_classify_kernel); one scratch region
    // per class, areas carved for the largest search of the class
    static const int an_ctas = getenv("PSI_AN_CTAS") ? atoi(getenv("PSI_AN_CTAS")) : 3;
    static const int an_class0 = getenv("PSI_AN_CLASS0") ? atoi(getenv("PSI_AN_CLASS0")) : kAnClass0;
    static const int an_class1 = getenv("PSI_AN_CLASS1") ? atoi(getenv("PSI_AN_CLASS1")) : kAnClass1;
    static const int an_huge = getenv("PSI_AN_HUGE") ? atoi(getenv("PSI_AN_HUGE")) : 64;   // A/B: profiles/k4_groups_r2.txt
    const int an_threads = 256;
    const int an_g[3] = {8, 16, 32};
    const int an_cap[3] = {std::min(cap, an_class0), std::min(cap, an_class1), cap};
    size_t an_per_group[3], an_off[3], need = 0;
    for (int k = 0; k < 3; ++k) {
        const bool used = k == 0 || cap > (k == 1 ? an_class0 : an_class1);
        an_per_group[k] = (aaa_scratch_bytes(an_cap[k]) + 255) & ~(size_t)255;
        an_off[k] = need;
        if (used) need += an_per_group[k] * (size_t)c->sm_count * an_ctas * (an_threads / an_g[k]);
    }
    if (need > c->mode_scratch_bytes) {
        if (c->d_mode_scratch) CK(cudaFree(c->d_mode_scratch));
        c->d_mode_scratch = nullptr; c->mode_scratch_bytes = 0;
        CK(cudaMalloc(&c->d_mode_scratch, need));
        c->mode_scratch_bytes = need;
    }
    // inputs + per-item state in one staging allocation
    const size_t nreq = ni * (size_t)max_req;
    const bool want_mask = extra && extra->mask;
    const int rec_rows = (extra && extra->rec_dev) ? std::max(extra->rec_rows, n_items) : 0;
    StageItem it[] = {
        {dist, 4 * ni}, {kx, 8 * ni}, {kz, 8 * ni}, {alpha, 8 * ni}, {domain, 32 * ni},            // 0-4
        {iparams, 12 * ni}, {dparams, 16 * ni}, {uni, 8 * (size_t)n_uni}, {uni_off, 4 * ni},       // 5-8
        {guess_off, 4 * (ni + 1)}, {guesses, 16 * std::max<size_t>(ng, 1)},                        // 9-10
        {nullptr, 16 * ni * cap}, {nullptr, 16 * ni * cap},                                        // 11 Z, 12 F
        {nullptr, 16 * ni * P}, {nullptr, 16 * ni * P}, {nullptr, 64 * ni},                        // 13 w0, 14 pol, 15 zoom
        {nullptr, 16 * ni * max_roots}, {nullptr, 8 * ni},                                         // 16 roots, 17 box_dx
        {nullptr, 4 * ni * 14}, {nullptr, 8 * nreq},                                               // 18 ints, 19 elem
        {nullptr, want_mask ? 4 * ni * cap : 0},                                                   // 20 final masks
        {nullptr, rec_rows ? 8 * (size_t)rec_rows * (4 + 2 * max_roots) : 0},                      // 21 records
        {nullptr, 12 * ni}, {nullptr, 64}};                                                        // 22-23 size classes
    char* d[24];
    int rc = stage_upload(c, it, 24, 0, d, nullptr);
    if (rc != PSI_OK) return rc;
    cudaStream_t st = c->stream;
    ModeDev m;
    m.Z = (cx*)d[11]; m.F = (cx*)d[12]; m.w0 = (cx*)d[13]; m.pol = (cx*)d[14]; m.zoom = (cx*)d[15];
    m.roots = (cx*)d[16]; m.box_dx = (double*)d[17];
    int* ints = (int*)d[18];
    m.M = ints; m.level = ints + ni; m.upos = ints + 2 * ni; m.phase = ints + 3 * ni; m.n0 = ints + 4 * ni;
    m.max_iter = ints + 5 * ni; m.n_zoom = ints + 6 * ni; m.calls = ints + 7 * ni; m.status = ints + 8 * ni;
    m.n_out = ints + 9 * ni; m.pcalls = ints + 10 * ni; m.pconv = ints + 11 * ni; m.pstat = ints + 12 * ni;
    m.conv_max = ints + 13 * ni;
    m.mask_out = want_mask ? (int*)d[20] : nullptr;
    m.elem = (long long*)d[19];
    m.counters = c->d_counter + 6;
    m.cap = cap; m.P = P; m.max_roots = max_roots;
    CK(cudaMemsetAsync(ints, 0, 4 * ni * 14, st));
    if (want_mask) CK(cudaMemsetAsync(d[20], 0, 4 * ni * cap, st));                         // phase 0 = kPhaseSample everywhere
    CK(cudaMemsetAsync(c->d_counter + 4, 0, 2 * sizeof(unsigned long long), st));
    NodeTableView t = table_view(c);
    const int tgrid = (n_items + 255) / 256;
    const int warps = kEvalThreads / 32;
    const int pgrid = std::min(c->sm_count, (n_items + warps - 1) / warps);
    int launches = 0;
    long long evals = 0;
    // stage timing (sample, K1, analysis, K3, decide) with CUDA events on the context's stream
    struct Events {                        // destroyed on every return path
        cudaEvent_t e[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        ~Events() { for (auto& x : e) if (x) cudaEventDestroy(x); }
    } events;
    cudaEvent_t* ev = events.e;
    for (int k = 0; k < 6; ++k) CK(cudaEventCreate(&ev[k]));
    double stage_us[5] = {0, 0, 0, 0, 0};
    auto collect = [&](int upto) {
        for (int k = 0; k < upto; ++k) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ev[k], ev[k + 1]) == cudaSuccess) stage_us[k] += 1e3 * ms;
        }
    };
    CK(cudaStreamSynchronize(st));
    us_setup = wall_us();                  // allocations + upload of the inputs
    for (int round = 0; round <= max_zoom; ++round) {
        CK(cudaMemsetAsync(m.counters, 0, 2 * sizeof(unsigned long long), st));
        CK(cudaEventRecord(ev[0], st));
        mode_sample_kernel<<<tgrid, 256, 0, st>>>(m, n_items, (const double*)d[4], (const int*)d[5],
                                                  (const double*)d[7], (const int*)d[8], (const int*)d[9],
                                                  (const double2*)d[10]);
        CK(cudaGetLastError());
        unsigned long long cnt[2] = {0, 0};
        CK(cudaEventRecord(ev[1], st));
        CK(cudaMemcpyAsync(cnt, m.counters, sizeof(cnt), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const long long n_req = (long long)cnt[0];
        if (n_req == 0) { collect(1); break; }
        evals += n_req;
        rc = launch_disp_eval(c, n_req, 0, (const int*)d[0], (const double*)d[1], (const double*)d[2],
                              (const double*)d[3], (const double*)m.Z, (double*)m.F, nullptr, c->d_counter + 4,
                              mask, st, nullptr, m.elem, cap);
        if (rc != PSI_OK) return rc;
        CK(cudaMemsetAsync(c->d_counter, 0, sizeof(unsigned long long), st));
        CK(cudaMemsetAsync(c->d_counter + 8, 0, sizeof(unsigned long long), st));
        CK(cudaEventRecord(ev[2], st));
        {
            int* an_lists = (int*)d[22];
            unsigned long long* an_counts = (unsigned long long*)d[23];        // [0..2] class sizes, [4..6] queue heads
            CK(cudaMemsetAsync(an_counts, 0, 8 * sizeof(unsigned long long), st));
            mode_classify_kernel<<<tgrid, 256, 0, st>>>(m, n_items, an_lists, an_counts, an_class0, an_class1, an_huge);
            const bool head_list = an_class1 <= an_class0;                  // list 1 = largest 32-lane searches
            for (int k = 0; k < 3; ++k) {
                if (k > 0 && cap <= (k == 1 ? an_class0 : an_class1)) continue;       // no search can be that large
                if (k == 1 && head_list) continue;
                const int per_cta = an_threads / an_g[k];
                const int agrid = std::min(c->sm_count * an_ctas, (n_items + per_cta - 1) / per_cta);
                auto kern = k == 0 ? mode_analyze_kernel<8> : k == 1 ? mode_analyze_kernel<16> : mode_analyze_kernel<32>;
                kern<<<agrid, an_threads, 0, st>>>(m, n_items, (const double*)d[4], (const int*)d[5],
                                                   (const double*)d[6], (const int*)d[9],
                                                   (unsigned char*)c->d_mode_scratch + an_off[k], an_per_group[k],
                                                   an_cap[k], an_lists + (size_t)k * n_items, an_counts + k,
                                                   an_counts + 4 + k, c->d_counter + 8,
                                                   k == 2 && head_list ? an_lists + (size_t)n_items : nullptr,
                                                   an_counts + 1);
                ++launches;
            }
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(ev[3], st));
        // few root searches to polish: one BLOCK per search instead of one warp (the slowest serial secant
        // chain sets the launch time; spreading each D evaluation over the block shortens it ~5x)
        unsigned long long n_polish = 0;
        CK(cudaMemcpyAsync(&n_polish, c->d_counter + 8, sizeof(n_polish), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const bool k3_cta = n_polish > 0 && (long long)n_polish <= k3_cta_items();
        for (int cls = 0; cls < kNumClasses; ++cls) {
            if (!(mask & (1u << cls))) continue;
            CK(cudaMemsetAsync(c->d_counter, 0, sizeof(unsigned long long), st));
            (k3_cta ? kPolishKernelsCta : kPolishKernels)[cls]<<<k3_cta ? c->sm_count : pgrid, kEvalThreads,
                                                                 c->table_in_smem ? c->table_smem : 0, st>>>(
                t, c->table_in_smem ? 1 : 0, c->d_dists, n_items, nullptr, (const int*)d[0],
                (const double*)d[1], (const double*)d[2], (const double*)d[3], (const double*)d[4], nullptr,
                (double2*)m.w0, m.max_iter, (double2*)m.pol, m.pcalls, m.pconv, m.pstat, c->d_counter,
                c->d_counter + 4, P, m.n0, m.phase);
            CK(cudaGetLastError());
            ++launches;
        }
        CK(cudaMemsetAsync(m.counters, 0, 2 * sizeof(unsigned long long), st));
        CK(cudaEventRecord(ev[4], st));
        mode_decide_kernel<<<tgrid, 256, 0, st>>>(m, n_items, (const double*)d[4], (const int*)d[5]);
        CK(cudaGetLastError());
        CK(cudaEventRecord(ev[5], st));
        launches += 3;
        c->launches += 4;
        CK(cudaMemcpyAsync(cnt, m.counters, sizeof(cnt), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        collect(5);
        if (cnt[1] == 0) break;                                          // nobody wants to zoom
    }
    unsigned long long cnt2[2] = {0, 0};
    if (rec_rows) {
        mode_pack_kernel<<<(rec_rows + 255) / 256, 256, 0, st>>>(m, n_items, rec_rows, (double*)d[21]);
        CK(cudaGetLastError());
        *extra->rec_dev = (double*)d[21];
        ++launches;
    }
    if (roots) CK(cudaMemcpyAsync(roots, m.roots, 16 * ni * max_roots, cudaMemcpyDeviceToHost, st));
    if (n_roots) CK(cudaMemcpyAsync(n_roots, m.n_out, 4 * ni, cudaMemcpyDeviceToHost, st));
    if (n_calls) CK(cudaMemcpyAsync(n_calls, m.calls, 4 * ni, cudaMemcpyDeviceToHost, st));
    if (status) CK(cudaMemcpyAsync(status, m.status, 4 * ni, cudaMemcpyDeviceToHost, st));
    if (extra) {
        if (extra->Z) CK(cudaMemcpyAsync(extra->Z, m.Z, 16 * ni * cap, cudaMemcpyDeviceToHost, st));
        if (extra->F) CK(cudaMemcpyAsync(extra->F, m.F, 16 * ni * cap, cudaMemcpyDeviceToHost, st));
        if (extra->M) CK(cudaMemcpyAsync(extra->M, m.M, 4 * ni, cudaMemcpyDeviceToHost, st));
        if (extra->mask) CK(cudaMemcpyAsync(extra->mask, m.mask_out, 4 * ni * cap, cudaMemcpyDeviceToHost, st));
        if (extra->upos) CK(cudaMemcpyAsync(extra->upos, m.upos, 4 * ni, cudaMemcpyDeviceToHost, st));
        if (extra->conv) CK(cudaMemcpyAsync(extra->conv, m.conv_max, 4 * ni, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaMemcpyAsync(cnt2, c->d_counter + 4, sizeof(cnt2), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (stats) {
        stats[0] = launches; stats[1] = evals + (long long)cnt2[1]; stats[2] = (long long)cnt2[0];
        stats[3] = us_setup; stats[4] = wall_us();                             // host wall clock: setup, whole call
        for (int k = 0; k < 5; ++k) stats[8 + k] = (long long)stage_us[k];     // stats has 16 slots
    }
    return PSI_OK;
}

extern "C" {

int psi_polish_batch_host(psi_ctx* c, int n_items, const int* dist, const double* kx, const double* kz,
                          const double* alpha, const double* domain, const int* w0_off, const double* w0,
                          const int* max_iter, double* roots, int* n_calls, int* max_conv, int* status) {
    if (!c || n_items < 0 || !dist || !kx || !kz || !alpha || !domain || !w0_off || !w0 || !max_iter ||
        !roots || !n_calls || !max_conv || !status)
        return fail(PSI_ERR_ARG, "psi_polish_batch_host: bad argument");
    std::vector<int> pidx(n_items);
    for (int i = 0; i < n_items; ++i) pidx[i] = i;
    std::vector<double> w0copy(w0, w0 + 2 * (size_t)w0_off[n_items]);
    return psi_internal_polish(c, n_items, pidx.data(), n_items, dist, kx, kz, alpha, domain, w0_off,
                               w0copy.data(), max_iter, roots, n_calls, max_conv, status, nullptr);
}

int psi_contour_batch_host(psi_ctx* c, int n_items, const int* dist, const double* kx, const double* kz,
                           const double* alpha, const int* z_off, const double* z_list, const double* tol,
                           const double* max_step, const int* max_iter, int* counts, int* status,
                           int* n_points, long long* stats) {
    if (!c || n_items < 0 || !dist || !kx || !kz || !alpha || !z_off || !z_list || !tol || !max_step ||
        !max_iter || !counts || !status || !n_points)
        return fail(PSI_ERR_ARG, "psi_contour_batch_host: bad argument");
    if (stats) { stats[0] = 0; stats[1] = 0; }
    if (n_items == 0) return PSI_OK;
    CK(cudaSetDevice(c->device));
    bool ok;
    const unsigned mask = class_mask_of(c, n_items, dist, alpha, &ok);
    if (!ok) return fail(PSI_ERR_ARG, "distribution id out of range");
    const int cap = kContourCap;
    const int warps = kEvalThreads / 32;
    const int grid = std::min(c->sm_count, (n_items + warps - 1) / warps);
    if (!c->d_contour_scratch)
        CK(cudaMalloc(&c->d_contour_scratch, (size_t)c->sm_count * warps * 4 * cap * sizeof(double2)));
    const size_t ni = (size_t)n_items, nz = (size_t)z_off[n_items];
    StageItem it[] = {{dist, 4 * ni}, {kx, 8 * ni}, {kz, 8 * ni}, {alpha, 8 * ni}, {z_off, 4 * (ni + 1)},
                      {z_list, 16 * nz}, {tol, 8 * ni}, {max_step, 8 * ni}, {max_iter, 4 * ni},
                      {nullptr, 4 * ni}, {nullptr, 4 * ni}, {nullptr, 4 * ni}};
    char* d[12];
    int rc = stage_upload(c, it, 12, 0, d, nullptr);
    if (rc != PSI_OK) return rc;
    cudaStream_t st = c->stream;
    CK(cudaMemsetAsync(c->d_counter + 4, 0, 2 * sizeof(unsigned long long), st));
    NodeTableView t = table_view(c);
    for (int cls = 0; cls < kNumClasses; ++cls) {
    if (!(mask & (1u << cls))) continue;
    CK(cudaMemsetAsync(c->d_counter, 0, sizeof(unsigned long long), st));
    kContourKernels[cls]<<<grid, kEvalThreads, c->table_in_smem ? c->table_smem : 0, st>>>(
        t, c->table_in_smem ? 1 : 0, c->d_dists, n_items, (const int*)d[0], (const double*)d[1],
        (const double*)d[2], (const double*)d[3], (const int*)d[4], (const double2*)d[5], (const double*)d[6],
        (const double*)d[7], (const int*)d[8], (double2*)c->d_contour_scratch, cap, (int*)d[9], (int*)d[10],
        (int*)d[11], c->d_counter, c->d_counter + 4);
    c->launches++;
    CK(cudaGetLastError());
    }
    unsigned long long cnt[2] = {0, 0};
    CK(cudaMemcpyAsync(counts, d[9], 4 * ni, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(status, d[10], 4 * ni, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(n_points, d[11], 4 * ni, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(cnt, c->d_counter + 4, sizeof(cnt), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (stats) { stats[0] = 1; stats[1] = (long long)cnt[1]; }
    return PSI_OK;
}

#if defined(PSI_PROF)
// profiling build only (tools/prof_analysis.py): read and clear the per-phase cycle counters of the analysis stage
int psi_prof_read(unsigned long long* out) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out, g_psi_prof, sizeof(unsigned long long) * 32) != cudaSuccess) return -1;
    unsigned long long zero[32] = {0};
    cudaMemcpyToSymbol(g_psi_prof, zero, sizeof(zero));
    return 0;
}
#endif

int psi_measure_fp64_peak(psi_ctx* c, int reps, double* tflops) {
    if (!c || !tflops || reps < 1) return fail(PSI_ERR_ARG, "psi_measure_fp64_peak: bad argument");
    CK(cudaSetDevice(c->device));
    const int threads = 256, blocks = c->sm_count * 8, iters = 4096;
    double* buf = nullptr;
    CK(cudaMalloc(&buf, sizeof(double) * threads * blocks));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double best = 0;
    for (int r = 0; r < reps + 2; ++r) {
        CK(cudaEventRecord(e0, c->stream));
        dfma_peak_kernel<<<blocks, threads, 0, c->stream>>>(buf, iters, 1.0 + r);
        c->launches++;
        CK(cudaEventRecord(e1, c->stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8 * 16 * (double)iters * threads * blocks;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (r >= 2 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
    *tflops = best;
    return PSI_OK;
}

}  // extern "C"
