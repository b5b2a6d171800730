// psi_internal.hpp -- internal C++ interface between the translation units of libpsi_b200.so.
#pragma once
#include <string>

#include "../../include/psi_b200.h"

// record an error message for psi_last_error() and return `code`
int psi_internal_fail(int code, const std::string& msg);

// K1 for n HOST points w[i] whose parameters are row pidx[i] of the HOST arrays dist/kx/kz/alpha
// (n_params rows).  Synchronous; D receives n complex values.
int psi_internal_eval_indexed(psi_ctx* ctx, long long n, const int* pidx, int n_params, const int* dist,
                              const double* kx, const double* kz, const double* alpha, const double* w,
                              double* D);

// K3 for n_items polish jobs (HOST arrays): item k polishes the start values w0[off[k]..off[k+1]) of the
// dispersion relation of parameter row pidx[k] inside the rectangle dom[4k..4k+4) with at most
// max_iter[k] secant iterations; out gets one complex per start value (root or the sentinel -1j),
// n_calls / max_conv / status one int per item; *evals the number of D evaluations performed.
int psi_internal_polish(psi_ctx* ctx, int n_items, const int* pidx, int n_params, const int* dist,
                        const double* kx, const double* kz, const double* alpha, const double* dom,
                        const int* off, double* w0, const int* max_iter, double* out, int* n_calls,
                        int* max_conv, int* status, long long* evals);

// Optional inputs/outputs of psi_internal_mode_device (any pointer may be null).
struct PsiModeExtra {
    // packed result records that STAY ON THE DEVICE (for a collective straight out of the library's buffer):
    // rec_rows rows (>= n_items, the excess zeroed) of 4 + 2*max_roots doubles
    //   {n_roots, status, n_function_call, 0, Re root_0, Im root_0, ...};
    // *rec_dev receives the device pointer, valid until the next call on the context.  With records the host
    // outputs roots / n_roots / n_calls / status may be null.
    double** rec_dev = nullptr;
    int rec_rows = 0;
    // per-search detail (host arrays, cap entries per item where an array): sample points and values, number of
    // samples, support mask of the final rational approximation, random numbers consumed, largest number of
    // secant iterations of a converged polish
    double* Z = nullptr; double* F = nullptr; int* M = nullptr; int* mask = nullptr; int* upos = nullptr;
    int* conv = nullptr;
};

// K4 for n_items whole PSIMode.calculate searches (HOST arrays; see psi_mode_batch_device in
// include/psi_b200.h).  uni: concatenated U[0,1) streams, uni_off[i]: start of item i's stream; cap:
// largest number of samples any item can reach.  stats = {launches, D evaluations, node evaluations}.
int psi_internal_mode_device(psi_ctx* ctx, int n_items, const int* dist, const double* kx, const double* kz,
                             const double* alpha, const double* domain, const int* iparams,
                             const double* dparams, const double* uni, long long n_uni, const int* uni_off,
                             const int* guess_off, const double* guesses, int cap, int max_roots,
                             double* roots, int* n_roots, int* n_calls, int* status, long long* stats,
                             const PsiModeExtra* extra = nullptr);
