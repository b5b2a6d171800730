// hostsim.cpp -- TEST-ONLY lane-emulation harness.
//
// Compiles the device headers (psi_math.cuh / psi_quad.cuh / ...) with g++ and a ONE-LANE warp
// policy so that the kernel logic can be exercised in the GPU-less build container against the
// oracle.  It is built by tests/conftest.py into tests/hostsim/_build/, is never loaded by the
// psitools_public_b200 package, and is not a fallback: the product library fails with
// PSI_ERR_NODEVICE when no GPU is present.
#include <vector>

#include "../../psitools_public_b200/csrc/psi_table.hpp"

using namespace psi;

struct WarpOne {
    static constexpr int width = 1;
    static int lane() { return 0; }
    static double sum(double v) { return v; }
};

struct Table {
    std::vector<node_t> lv;
    std::vector<int> off;
    NodeTableView view;
};

static Table make_table(const double* xj, const double* wj, int n, int max_m, double eps) {
    Table t;
    build_level_major(xj, wj, n, max_m, t.lv, t.off);
    t.view.nodes = t.lv.data(); t.view.level_off = t.off.data();
    t.view.n_nodes = n; t.view.max_m = max_m; t.view.eps = eps;
    return t;
}

extern "C" {

// out = {J0, J1, denom, vgx, vgy}
int hs_background(const psi_dist_desc* desc, const double* xj, const double* wj, int n, int max_m,
                  double eps, double* out) {
    Table t = make_table(xj, wj, n, max_m, eps);
    t.view.nodes = t.lv.data(); t.view.level_off = t.off.data();
    DistParams d = dist_from_desc(*desc);
    unsigned long long nodes = 0;
    background_warp<WarpOne>(t.view, d, nodes);
    out[0] = d.j0; out[1] = d.j1; out[2] = d.denom; out[3] = d.vgx; out[4] = d.vgy;
    return 0;
}

int hs_disp_eval(const psi_dist_desc* desc, const double* xj, const double* wj, int n, int max_m,
                 double eps, long long npts, const double* kx, const double* kz, const double* alpha,
                 const double* w, double* D, double* M, unsigned long long* node_evals) {
    Table t = make_table(xj, wj, n, max_m, eps);
    t.view.nodes = t.lv.data(); t.view.level_off = t.off.data();
    DistParams d = dist_from_desc(*desc);
    unsigned long long nodes = 0;
    background_warp<WarpOne>(t.view, d, nodes);
    nodes = 0;
    for (long long i = 0; i < npts; ++i) {
        cx m[7];
        cx v = disp_eval_any<WarpOne>(t.view, d, cx(w[2 * i], w[2 * i + 1]), kx[i], kz[i], alpha[i], m,
                                       nodes);
        D[2 * i] = v.re; D[2 * i + 1] = v.im;
        if (M) for (int k = 0; k < 7; ++k) { M[14 * i + 2 * k] = m[k].re; M[14 * i + 2 * k + 1] = m[k].im; }
    }
    if (node_evals) *node_evals = nodes;
    return 0;
}

}  // extern "C"
