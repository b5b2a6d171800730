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
