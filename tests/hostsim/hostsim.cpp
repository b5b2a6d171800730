// hostsim.cpp -- TEST-ONLY lane-emulation harness.
//
// Compiles the device headers (psi_math.cuh / psi_quad.cuh) with g++ and a ONE-LANE warp policy, and
// the native lock-step engine (psi_engine.cu, host code) against a CPU stand-in for kernel K1, so that
// the kernel arithmetic and the engine logic can be exercised in the GPU-less build container against
// the oracle.  Built by tests/hostsim_util.py into tests/hostsim/_build/, never loaded by the
// psitools_public_b200 package, and not a fallback: the product library fails with PSI_ERR_NODEVICE
// when no GPU is present.
#include <string>
#include <vector>

#include "../../psitools_public_b200/csrc/psi_table.hpp"

using namespace psi;

struct WarpOne {
    static constexpr int width = 1;
    static int lane() { return 0; }
    static double sum(double v) { return v; }
};

// test stand-in for the library's context
struct psi_ctx {
    std::vector<node_t> lv;
    std::vector<int> off;
    NodeTableView view;
    std::vector<DistParams> dists;
    unsigned long long nodes = 0;
};

static std::string g_err;
int psi_internal_fail(int code, const std::string& msg) { g_err = msg; return code; }

static void eval_points(psi_ctx* c, long long n, const int* pidx, const int* dist, const double* kx,
                        const double* kz, const double* alpha, const double* w, double* D, double* M) {
    unsigned long long total = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : total)
    for (long long i = 0; i < n; ++i) {
        const long long pi = pidx ? pidx[i] : i;
        unsigned long long nodes = 0;
        cx m[7];
        cx v = disp_eval_any<WarpOne>(c->view, c->dists[dist[pi]], cx(w[2 * i], w[2 * i + 1]), kx[pi], kz[pi],
                                      alpha[pi], m, nodes);
        D[2 * i] = v.re; D[2 * i + 1] = v.im;
        if (M) for (int k = 0; k < 7; ++k) { M[14 * i + 2 * k] = m[k].re; M[14 * i + 2 * k + 1] = m[k].im; }
        total += nodes;
    }
    c->nodes += total;
}

int psi_internal_eval_indexed(psi_ctx* c, long long n, const int* pidx, int n_params, const int* dist,
                              const double* kx, const double* kz, const double* alpha, const double* w,
                              double* D) {
    (void)n_params;
    eval_points(c, n, pidx, dist, kx, kz, alpha, w, D, nullptr);
    return PSI_OK;
}

#include "../../psitools_public_b200/csrc/psi_engine.cu"

extern "C" {

const char* hs_last_error(void) { return g_err.c_str(); }

psi_ctx* hs_ctx_create(const double* xj, const double* wj, int n, int max_m, double eps) {
    psi_ctx* c = new psi_ctx();
    build_level_major(xj, wj, n, max_m, c->lv, c->off);
    c->view.nodes = c->lv.data(); c->view.level_off = c->off.data();
    c->view.n_nodes = n; c->view.max_m = max_m; c->view.eps = eps;
    return c;
}
void hs_ctx_destroy(psi_ctx* c) { delete c; }

int hs_dist_add(psi_ctx* c, const psi_dist_desc* desc) {
    DistParams d = dist_from_desc(*desc);
    unsigned long long nodes = 0;
    background_warp<WarpOne>(c->view, d, nodes);
    c->dists.push_back(d);
    return (int)c->dists.size() - 1;
}

// out = {J0, J1, denom, vgx, vgy}
int hs_dist_background(psi_ctx* c, int id, double* out) {
    const DistParams& d = c->dists[id];
    out[0] = d.j0; out[1] = d.j1; out[2] = d.denom; out[3] = d.vgx; out[4] = d.vgy;
    return 0;
}

int hs_disp_eval(psi_ctx* c, long long n, const int* dist, const double* kx, const double* kz,
                 const double* alpha, const double* w, double* D, double* M, unsigned long long* node_evals) {
    c->nodes = 0;
    eval_points(c, n, nullptr, dist, kx, kz, alpha, w, D, M);
    if (node_evals) *node_evals = c->nodes;
    return 0;
}

}  // extern "C"
