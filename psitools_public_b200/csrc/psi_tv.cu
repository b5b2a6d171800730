// psi_tv.cu -- psi_tv_roots: the terminal-velocity PSI mode on a map of wave-number pairs, one thread per pair
// (psi_tv.cuh).  Guess generator for map sweeps; no node table, no context.
#include <cuda_runtime.h>

#include <string>

#include "psi_internal.hpp"
#include "psi_tv.cuh"

using namespace psi;

__global__ void __launch_bounds__(128)
tv_roots_kernel(TvParams P, long long n, const double* __restrict__ kx, const double* __restrict__ kz,
                double2* __restrict__ omega, int* __restrict__ status) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cx w;
    status[i] = tv_find_root(P, kx[i], kz[i], &w);
    omega[i] = make_double2(w.re, w.im);
}

#define TCK(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            cudaFree(d_kx); cudaFree(d_kz); cudaFree(d_w); cudaFree(d_st);                               \
            return psi_internal_fail(PSI_ERR_CUDA, std::string("psi_tv_roots: ") + cudaGetErrorString(e_)); \
        }                                                                                                \
    } while (0)

extern "C" int psi_tv_roots(int device, double dust_fraction, double tau_max, double b, double tau_min,
                            double diffusion, int max_it, long long n, const double* kx, const double* kz,
                            double* omega, int* status) {
    if (n < 0 || (n > 0 && (!kx || !kz || !omega || !status)) || !(tau_max > 0.0) || !(tau_min > 0.0) || max_it < 1)
        return psi_internal_fail(PSI_ERR_ARG, "psi_tv_roots: bad argument");
    if (!tv_supported(b))
        return psi_internal_fail(PSI_ERR_ARG, "psi_tv_roots: no closed form for this size-distribution exponent");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev)
        return psi_internal_fail(PSI_ERR_NODEVICE, "psi_tv_roots: no such CUDA device");
    if (n == 0) return PSI_OK;
    double *d_kx = nullptr, *d_kz = nullptr;
    double2* d_w = nullptr;
    int* d_st = nullptr;
    TCK(cudaSetDevice(device));
    TCK(cudaMalloc(&d_kx, 8 * (size_t)n));
    TCK(cudaMalloc(&d_kz, 8 * (size_t)n));
    TCK(cudaMalloc(&d_w, 16 * (size_t)n));
    TCK(cudaMalloc(&d_st, 4 * (size_t)n));
    TCK(cudaMemcpy(d_kx, kx, 8 * (size_t)n, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(d_kz, kz, 8 * (size_t)n, cudaMemcpyHostToDevice));
    TvParams P{dust_fraction, tau_max, b, tau_min / tau_max, diffusion, max_it};
    tv_roots_kernel<<<(unsigned)((n + 127) / 128), 128>>>(P, n, d_kx, d_kz, d_w, d_st);
    TCK(cudaGetLastError());
    TCK(cudaMemcpy(omega, d_w, 16 * (size_t)n, cudaMemcpyDeviceToHost));
    TCK(cudaMemcpy(status, d_st, 4 * (size_t)n, cudaMemcpyDeviceToHost));
    cudaFree(d_kx); cudaFree(d_kz); cudaFree(d_w); cudaFree(d_st);
    return PSI_OK;
}
