// psi_refine.cuh -- one sweep of the adaptive (Kx, Kz) refinement as per-pixel device logic.
//
// Replaces the body of PSIGridRefiner.sweep_last_grid (psi_grid_refine.py:218-290 of the reference): a
// pixel WITHOUT roots collects the roots of its eight neighbours (reference order: (iz-1,ix-1), (iz,ix-1),
// (iz+1,ix-1), (iz-1,ix), (iz+1,ix), (iz-1,ix+1), (iz,ix+1), (iz+1,ix+1); each run's roots in stored
// order), prunes them with prune_eps (:75-84: greedy, epsilon = mean(Im)*1e-4) and launches a run when it
// has more guesses than the last time it was run (nguess).  The root map: run id per pixel, count per run id and
// the roots of a run id either in `width` dense slots (the numpy RootStore; CPU harness) or, on the device, at
// off[run id] in one packed pool (most runs of a growth-rate map have no root: dense slots for every possible
// run id of a 1089^2 map would be 4 GB, the pool is 20 MB).
// PSI_HD like the other headers: the CPU harness runs the same code against the numpy sweep.
#pragma once
#include "psi_batch.cuh"

namespace psi {

constexpr int kRefineMaxNear = 128;        // most neighbour roots one pixel can collect (8 * width)

struct RefineView {
    int nz, nx, width;
    const int* runs;        // [nz*nx] run id or -1
    const int* cnt;         // [keys]  roots of a run id, -1 = no such run
    const cx* mat;          // [keys*width], or the packed pool when off is set
    const int* nguess;      // [nz*nx]
    const long long* off = nullptr;    // [keys] start of a run id's roots in the pool (packed layout)
    PSI_HD const cx* roots_of(int key) const { return off ? mat + off[key] : mat + (size_t)key * width; }
};

PSI_HD bool refine_has(const RefineView& g, int iz, int ix) {
    if (iz < 0 || iz >= g.nz || ix < 0 || ix >= g.nx) return false;
    const int key = g.runs[iz * g.nx + ix];
    return key >= 0 && g.cnt[key] > 0;
}

// numpy's pairwise summation of a[0..n).im for n <= 128 (numpy/_core/src/umath/loops_utils.h.src): a plain
// loop below 8 elements, eight interleaved accumulators above -- so that epsilon is bit-identical to
// np.array(near_roots).imag.mean() of the reference.
PSI_HD double np_sum_imag(const cx* a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i].im;
        return res;
    }
    double r[8];
    for (int k = 0; k < 8; ++k) r[k] = a[k].im;
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int k = 0; k < 8; ++k) r[k] += a[i + k].im;
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i].im;
    return res;
}

// Number of guesses the pixel (iz, ix) would be run with, 0 if it does not run this sweep; -1 if it collects
// more than kRefineMaxNear neighbour roots (the caller falls back to the host sweep).  kept[] (room for
// kRefineMaxNear values) receives the pruned guesses.
PSI_HD int refine_pixel(const RefineView& g, int iz, int ix, cx* near, cx* kept) {
    if (refine_has(g, iz, ix)) return 0;                   // never re-run a pixel that has roots
    const int dz[8] = {-1, 0, 1, -1, 1, -1, 0, 1}, dx[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
    int n = 0;
    for (int q = 0; q < 8; ++q) {
        const int jz = iz + dz[q], jx = ix + dx[q];
        if (jz < 0 || jz >= g.nz || jx < 0 || jx >= g.nx) continue;
        const int key = g.runs[jz * g.nx + jx];
        if (key < 0) continue;
        const int c = g.cnt[key];
        const cx* rk = g.roots_of(key);
        for (int t = 0; t < c; ++t) {
            if (n >= kRefineMaxNear) return -1;
            near[n++] = rk[t];
        }
    }
    if (n == 0) return 0;
    int nk = 0;
    if (n == 1) {
        kept[nk++] = near[0];                              // prune_eps returns a single value untouched
    } else {
        const double eps = np_sum_imag(near, n) / (double)n * 1e-4;
        for (int i = 0; i < n; ++i) {
            bool close = false;
            for (int k = 0; k < nk && !close; ++k) close = abs_c(near[i] - kept[k]) < eps;
            if (!close) kept[nk++] = near[i];
        }
    }
    return nk > g.nguess[iz * g.nx + ix] ? nk : 0;
}

}  // namespace psi
