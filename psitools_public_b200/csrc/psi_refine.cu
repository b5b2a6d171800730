// psi_refine.cu -- device-resident root map of PSIGridRefiner and the sweep kernels (C ABI: psi_refine_*).
//
// The refiner's cell queue -- which pixels run next, with which guesses -- is built on the GPU from a mirror
// of the RootStore: kernel `refine_count_kernel` (thread per pixel) decides and counts, two CUB scans turn
// the counts into a compact row-major work list, `refine_emit_kernel` writes it.  The host only moves the
// compact list (pixel, guesses) into the K4 pipeline and the new roots back into the mirror.
#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "psi_internal.hpp"
#include "psi_refine.cuh"

using namespace psi;

#define RCK(call)                                                                              \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return psi_internal_fail(PSI_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct psi_refine {
    int device = 0, nz = 0, nx = 0, width = 0;
    long long key_capacity = 0;
    cudaStream_t stream = nullptr;
    int *runs = nullptr, *nguess = nullptr, *cnt = nullptr, *nk = nullptr, *overflow = nullptr;
    cx* mat = nullptr;                      // packed pool of roots (pool_used of pool_cap entries in use)
    long long* off = nullptr;               // [key_capacity] start of a run id's roots in the pool
    long long pool_cap = 0, pool_used = 0;
    long long *flag_scan = nullptr, *guess_scan = nullptr;     // exclusive scans over the pixels
    void* cub_tmp = nullptr;
    size_t cub_bytes = 0;
    long long n_run = 0, n_guess = 0;
};

__global__ void __launch_bounds__(128)
refine_count_kernel(RefineView g, int* nk, long long* flag, long long* n_guess, int* overflow) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (long long)g.nz * g.nx) return;
    cx near[kRefineMaxNear], kept[kRefineMaxNear];
    const int n = refine_pixel(g, (int)(p / g.nx), (int)(p % g.nx), near, kept);
    nk[p] = n;
    if (n < 0) atomicExch(overflow, 1);  // more neighbour roots than the buffer holds: the host sweeps instead
    flag[p] = n > 0 ? 1 : 0;             // scanned in place into the pixel's position in the work list
    n_guess[p] = n > 0 ? n : 0;          // ... and into the start of its guesses
}

// second pass: recompute the pruned guesses of the running pixels and write them at their scanned places;
// the pixel's guess count becomes its new `nguess`
__global__ void __launch_bounds__(128)
refine_emit_kernel(RefineView g, const int* nk, const long long* flag_scan, const long long* guess_scan,
                   int* nguess, int* pix_out, int* nkept_out, cx* guess_out) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (long long)g.nz * g.nx || nk[p] <= 0) return;
    cx near[kRefineMaxNear], kept[kRefineMaxNear];
    const int n = refine_pixel(g, (int)(p / g.nx), (int)(p % g.nx), near, kept);
    const long long j = flag_scan[p], o = guess_scan[p];
    pix_out[j] = (int)p;
    nkept_out[j] = n;
    for (int k = 0; k < n; ++k) guess_out[o + k] = kept[k];
    nguess[p] = n;                         // read by this thread only (refine_pixel above)
}

// store the roots of n_changed run ids (keys == nullptr: ids 0..n_changed-1) from dense rows of `width` slots at
// start[i] of the pool; set the run id of n_pix pixels
__global__ void refine_update_kernel(int* cnt, long long* off, cx* pool, int width, long long n_changed,
                                     const long long* keys, const int* new_cnt, const long long* start,
                                     const cx* rows, int* runs, long long n_pix, const int* pix, const int* run_ids) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_changed) {
        const long long key = keys ? keys[i] : i;
        const int c = new_cnt[i];
        cnt[key] = c;
        off[key] = start[i];
        for (int t = 0; t < c; ++t) pool[start[i] + t] = rows[i * width + t];
    }
    if (i < n_pix) runs[pix[i]] = run_ids[i];
}

// host half of an update: pool positions of the new rows, pool grown (x2) when they do not fit
static cudaError_t refine_store(psi_refine* r, long long n_changed, const long long* keys, const int* cnt,
                                const double* rows, long long n_pix, const int* pix, const int* run_ids) {
    std::vector<long long> start((size_t)std::max<long long>(n_changed, 1));
    long long used = r->pool_used;
    for (long long i = 0; i < n_changed; ++i) {
        start[i] = used;
        if (cnt[i] > r->width) return cudaErrorInvalidValue;
        used += cnt[i] > 0 ? cnt[i] : 0;
    }
    cudaError_t e = cudaSuccess;
    if (used > r->pool_cap) {
        const long long cap = std::max<long long>(2 * used, 4096);
        cx* bigger = nullptr;
        e = cudaMallocAsync(&bigger, 16 * (size_t)cap, r->stream);
        if (e == cudaSuccess && r->pool_used > 0)
            e = cudaMemcpyAsync(bigger, r->mat, 16 * (size_t)r->pool_used, cudaMemcpyDeviceToDevice, r->stream);
        if (e == cudaSuccess && r->mat) e = cudaFreeAsync(r->mat, r->stream);
        if (e != cudaSuccess) return e;
        r->mat = bigger; r->pool_cap = cap;
    }
    long long *d_keys = nullptr, *d_start = nullptr;
    int *d_cnt = nullptr, *d_pix = nullptr, *d_ids = nullptr;
    cx* d_rows = nullptr;
    const size_t nc = (size_t)std::max<long long>(n_changed, 1), npx = (size_t)std::max<long long>(n_pix, 1);
    if (keys) e = cudaMallocAsync(&d_keys, 8 * nc, r->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&d_start, 8 * nc, r->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&d_cnt, 4 * nc, r->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&d_rows, 16 * nc * r->width, r->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&d_pix, 4 * npx, r->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&d_ids, 4 * npx, r->stream);
    if (e == cudaSuccess && n_changed > 0) {
        if (keys) e = cudaMemcpyAsync(d_keys, keys, 8 * (size_t)n_changed, cudaMemcpyHostToDevice, r->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_start, start.data(), 8 * (size_t)n_changed, cudaMemcpyHostToDevice, r->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_cnt, cnt, 4 * (size_t)n_changed, cudaMemcpyHostToDevice, r->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(d_rows, rows, 16 * (size_t)n_changed * r->width, cudaMemcpyHostToDevice, r->stream);
    }
    if (e == cudaSuccess && n_pix > 0) {
        e = cudaMemcpyAsync(d_pix, pix, 4 * (size_t)n_pix, cudaMemcpyHostToDevice, r->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_ids, run_ids, 4 * (size_t)n_pix, cudaMemcpyHostToDevice, r->stream);
    }
    const long long n = std::max(n_changed, n_pix);
    if (e == cudaSuccess && n > 0) {
        refine_update_kernel<<<(int)((n + 255) / 256), 256, 0, r->stream>>>(r->cnt, r->off, r->mat, r->width, n_changed,
                                                                            d_keys, d_cnt, d_start, d_rows, r->runs,
                                                                            n_pix, d_pix, d_ids);
        e = cudaGetLastError();
    }
    void* tmp[] = {d_keys, d_start, d_cnt, d_rows, d_pix, d_ids};
    for (void* q : tmp) if (q) cudaFreeAsync(q, r->stream);
    const cudaError_t e2 = cudaStreamSynchronize(r->stream);       // (start[] is read by the copy until here)
    if (e == cudaSuccess) e = e2;
    if (e == cudaSuccess) r->pool_used = used;
    return e;
}

extern "C" {

void psi_refine_close(psi_refine* r) {
    if (!r) return;
    cudaSetDevice(r->device);
    // back into the stream-ordered pool (a campaign opens one map per refinement level: cudaFree of the previous
    // level's arrays, which unmaps them and synchronises the device, cost 0.02-0.1 s per level)
    if (r->stream) {
        void* p[] = {r->runs, r->nguess, r->cnt, r->nk, r->mat, r->off, r->flag_scan, r->guess_scan, r->cub_tmp, r->overflow};
        for (void* q : p) if (q) cudaFreeAsync(q, r->stream);
        cudaStreamSynchronize(r->stream);
        cudaStreamDestroy(r->stream);
    }
    delete r;
}

int psi_refine_open(psi_refine** out, int device, int nz, int nx, const int* runs, const int* nguess,
                    long long n_keys, long long key_capacity, int width, const double* mat, const int* cnt) {
    if (!out || nz <= 0 || nx <= 0 || !runs || !nguess || n_keys < 0 || key_capacity < n_keys || width <= 0 ||
        (n_keys > 0 && (!mat || !cnt)))
        return psi_internal_fail(PSI_ERR_ARG, "psi_refine_open: bad argument");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev)
        return psi_internal_fail(PSI_ERR_NODEVICE, "psi_refine_open: no such CUDA device");
    RCK(cudaSetDevice(device));
    {   // the per-sweep temporaries come from the stream-ordered pool: keep freed blocks in the pool instead of
        // returning them to the driver at every synchronisation (the default) and re-mapping them the next sweep
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    psi_refine* r = new psi_refine();
    r->device = device; r->nz = nz; r->nx = nx; r->width = width; r->key_capacity = key_capacity;
    const size_t np = (size_t)nz * nx;
    const bool trace = getenv("PSI_REFINE_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_a = now();
    cudaError_t e = cudaStreamCreate(&r->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&r->runs, 4 * np, r->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&r->nguess, 4 * np, r->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&r->nk, 4 * np, r->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&r->overflow, 4, r->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&r->flag_scan, 8 * np, r->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&r->guess_scan, 8 * np, r->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&r->cnt, 4 * (size_t)std::max<long long>(key_capacity, 1), r->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&r->off, 8 * (size_t)std::max<long long>(key_capacity, 1), r->stream);
    if (e == cudaSuccess)
        e = cub::DeviceScan::ExclusiveSum(nullptr, r->cub_bytes, r->flag_scan, r->flag_scan, (long long)np, r->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&r->cub_tmp, r->cub_bytes + 256, r->stream);
    if (trace) cudaStreamSynchronize(r->stream);
    const double t_b = now();
    if (e == cudaSuccess) e = cudaMemcpyAsync(r->runs, runs, 4 * np, cudaMemcpyHostToDevice, r->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(r->nguess, nguess, 4 * np, cudaMemcpyHostToDevice, r->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(r->cnt, 0xff, 4 * (size_t)std::max<long long>(key_capacity, 1), r->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(r->off, 0, 8 * (size_t)std::max<long long>(key_capacity, 1), r->stream);
    if (e == cudaSuccess && n_keys > 0) e = refine_store(r, n_keys, nullptr, cnt, mat, 0, nullptr, nullptr);
    if (e == cudaSuccess) e = cudaStreamSynchronize(r->stream);
    if (trace)
        fprintf(stderr, "psi_refine_open %dx%d keys %lld cap %lld width %d: alloc %.4f s, upload %.4f s\n", nz, nx, n_keys,
                key_capacity, width, t_b - t_a, now() - t_b);
    if (e != cudaSuccess) {
        psi_refine_close(r);
        return psi_internal_fail(PSI_ERR_CUDA, std::string("psi_refine_open: ") + cudaGetErrorString(e));
    }
    *out = r;
    return PSI_OK;
}

int psi_refine_sweep(psi_refine* r, long long* n_run, long long* n_guess) {
    if (!r || !n_run || !n_guess) return psi_internal_fail(PSI_ERR_ARG, "psi_refine_sweep: bad argument");
    RCK(cudaSetDevice(r->device));
    const long long np = (long long)r->nz * r->nx;
    RefineView g{r->nz, r->nx, r->width, r->runs, r->cnt, r->mat, r->nguess, r->off};
    const int grid = (int)((np + 127) / 128);
    RCK(cudaMemsetAsync(r->overflow, 0, 4, r->stream));
    refine_count_kernel<<<grid, 128, 0, r->stream>>>(g, r->nk, r->flag_scan, r->guess_scan, r->overflow);
    RCK(cudaGetLastError());
    size_t bytes = r->cub_bytes + 256;
    RCK(cub::DeviceScan::ExclusiveSum(r->cub_tmp, bytes, r->flag_scan, r->flag_scan, np, r->stream));
    bytes = r->cub_bytes + 256;
    RCK(cub::DeviceScan::ExclusiveSum(r->cub_tmp, bytes, r->guess_scan, r->guess_scan, np, r->stream));
    long long last[2] = {0, 0};
    int nk_last = 0;
    RCK(cudaMemcpyAsync(&last[0], r->flag_scan + np - 1, 8, cudaMemcpyDeviceToHost, r->stream));
    RCK(cudaMemcpyAsync(&last[1], r->guess_scan + np - 1, 8, cudaMemcpyDeviceToHost, r->stream));
    RCK(cudaMemcpyAsync(&nk_last, r->nk + np - 1, 4, cudaMemcpyDeviceToHost, r->stream));
    int overflow = 0;
    RCK(cudaMemcpyAsync(&overflow, r->overflow, 4, cudaMemcpyDeviceToHost, r->stream));
    RCK(cudaStreamSynchronize(r->stream));
    if (overflow) {                            // a pixel collected more than kRefineMaxNear neighbour roots
        r->n_run = 0; r->n_guess = 0;
        *n_run = -1; *n_guess = 0;
        return PSI_OK;
    }
    r->n_run = last[0] + (nk_last > 0 ? 1 : 0);
    r->n_guess = last[1] + (nk_last > 0 ? nk_last : 0);
    *n_run = r->n_run;
    *n_guess = r->n_guess;
    return PSI_OK;
}

int psi_refine_fetch(psi_refine* r, int* pix, int* n_kept, double* guesses) {
    if (!r || (r->n_run > 0 && (!pix || !n_kept || !guesses)))
        return psi_internal_fail(PSI_ERR_ARG, "psi_refine_fetch: bad argument");
    if (r->n_run == 0) return PSI_OK;
    RCK(cudaSetDevice(r->device));
    const long long np = (long long)r->nz * r->nx;
    int *d_pix = nullptr, *d_nk = nullptr;
    cx* d_g = nullptr;
    RCK(cudaMallocAsync(&d_pix, 4 * (size_t)r->n_run, r->stream));
    RCK(cudaMallocAsync(&d_nk, 4 * (size_t)r->n_run, r->stream));
    RCK(cudaMallocAsync(&d_g, 16 * (size_t)std::max<long long>(r->n_guess, 1), r->stream));
    RefineView g{r->nz, r->nx, r->width, r->runs, r->cnt, r->mat, r->nguess, r->off};
    refine_emit_kernel<<<(int)((np + 127) / 128), 128, 0, r->stream>>>(g, r->nk, r->flag_scan, r->guess_scan,
                                                                        r->nguess, d_pix, d_nk, d_g);
    RCK(cudaGetLastError());
    RCK(cudaMemcpyAsync(pix, d_pix, 4 * (size_t)r->n_run, cudaMemcpyDeviceToHost, r->stream));
    RCK(cudaMemcpyAsync(n_kept, d_nk, 4 * (size_t)r->n_run, cudaMemcpyDeviceToHost, r->stream));
    RCK(cudaMemcpyAsync(guesses, d_g, 16 * (size_t)r->n_guess, cudaMemcpyDeviceToHost, r->stream));
    RCK(cudaFreeAsync(d_pix, r->stream));
    RCK(cudaFreeAsync(d_nk, r->stream));
    RCK(cudaFreeAsync(d_g, r->stream));
    RCK(cudaStreamSynchronize(r->stream));
    return PSI_OK;
}

int psi_refine_update(psi_refine* r, long long n_changed, const long long* keys, const int* cnt,
                      const double* rows, long long n_pix, const int* pix, const int* run_ids) {
    if (!r || n_changed < 0 || n_pix < 0 || (n_changed > 0 && (!keys || !cnt || !rows)) ||
        (n_pix > 0 && (!pix || !run_ids)))
        return psi_internal_fail(PSI_ERR_ARG, "psi_refine_update: bad argument");
    for (long long i = 0; i < n_changed; ++i)
        if (keys[i] < 0 || keys[i] >= r->key_capacity)
            return psi_internal_fail(PSI_ERR_CAPACITY, "psi_refine_update: run id beyond the mirror's capacity");
    if (std::max(n_changed, n_pix) == 0) return PSI_OK;
    for (long long i = 0; i < n_changed; ++i)
        if (cnt[i] > r->width) return psi_internal_fail(PSI_ERR_ARG, "psi_refine_update: more roots than row slots");
    RCK(cudaSetDevice(r->device));
    RCK(refine_store(r, n_changed, keys, cnt, rows, n_pix, pix, run_ids));
    return PSI_OK;
}

}  // extern "C"
