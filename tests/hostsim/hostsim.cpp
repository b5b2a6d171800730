// hostsim.cpp -- TEST-ONLY lane-emulation harness.
//
// Compiles the device headers (psi_math.cuh / psi_quad.cuh) with g++ and a ONE-LANE warp policy, and
// the native lock-step engine (psi_engine.cu, host code) against a CPU stand-in for kernel K1, so that
// the kernel arithmetic and the engine logic can be exercised in the GPU-less build container against
// the oracle.  Built by tests/hostsim_util.py into tests/hostsim/_build/, never loaded by the
// psitools_public_b200 package, and not a fallback: the product library fails with PSI_ERR_NODEVICE
// when no GPU is present.
#include <pthread.h>

#include <string>
#include <thread>
#include <vector>

#define PSI_AAA_STATS 1

#include "../../psitools_public_b200/csrc/psi_mode.cuh"
#include "../../psitools_public_b200/csrc/psi_refine.cuh"
#include "../../psitools_public_b200/csrc/psi_tv.cuh"
#include "../../psitools_public_b200/csrc/psi_internal.hpp"
#include "../../psitools_public_b200/csrc/psi_table.hpp"

using namespace psi;

struct WarpOne {
    static constexpr int width = 1;
    static int lane() { return 0; }
    static double sum(double v) { return v; }
    static double max(double v) { return v; }
    static void sync() {}
    static double bcast(double v, int) { return v; }
    static unsigned ballot(bool p) { return p ? 1u : 0u; }
    static int popc(unsigned v) { return __builtin_popcount(v); }
    static void sum4c(double (&)[4], double (&)[4]) {}
    static bool any(bool p) { return p; }
    static void argmax(double&, int&) {}
};

// 32-lane emulation of the device's warp policy (WarpCuda in psi_b200.cu): every lane is a host thread, the
// shuffles and reductions exchange through a slot array between two barriers, in the butterfly order of the
// device code.  Slow (a barrier per exchange) but it runs the multi-lane paths of the warp kernels -- lane-strided
// loops, shuffle broadcasts, ballots -- in the GPU-less container.  Used for single AAA fits only.
struct WarpEmu {
    static constexpr int width = 32;
    static inline thread_local int tl_lane = 0;
    static inline pthread_barrier_t bar;
    static inline double slot[32][8];
    static inline int islot[32];
    static int lane() { return tl_lane; }
    static void sync() { pthread_barrier_wait(&bar); }
    static double bcast(double v, int src) {
        slot[tl_lane][0] = v; sync();
        const double r = slot[src][0]; sync();
        return r;
    }
    static double sum(double v) {
        for (int o = 16; o > 0; o >>= 1) { slot[tl_lane][0] = v; sync(); v += slot[tl_lane ^ o][0]; sync(); }
        return v;
    }
    static double max(double v) {
        for (int o = 16; o > 0; o >>= 1) { slot[tl_lane][0] = v; sync(); v = fmax(v, slot[tl_lane ^ o][0]); sync(); }
        return v;
    }
    static void sum4c(double (&re)[4], double (&im)[4]) {
        for (int o = 16; o > 0; o >>= 1) {
            for (int c = 0; c < 4; ++c) { slot[tl_lane][c] = re[c]; slot[tl_lane][4 + c] = im[c]; }
            sync();
            for (int c = 0; c < 4; ++c) { re[c] += slot[tl_lane ^ o][c]; im[c] += slot[tl_lane ^ o][4 + c]; }
            sync();
        }
    }
    static unsigned ballot(bool p) {
        islot[tl_lane] = p ? 1 : 0; sync();
        unsigned m = 0;
        for (int i = 0; i < 32; ++i) m |= (unsigned)islot[i] << i;
        sync();
        return m;
    }
    static int popc(unsigned v) { return __builtin_popcount(v); }
    static bool any(bool p) { return ballot(p) != 0; }
    static void argmax(double& v, int& idx) {
        for (int o = 16; o > 0; o >>= 1) {
            slot[tl_lane][0] = v; islot[tl_lane] = idx; sync();
            const double ov = slot[tl_lane ^ o][0];
            const int oi = islot[tl_lane ^ o];
            sync();
            if (oi >= 0 && (idx < 0 || ov > v || (ov == v && oi < idx))) { v = ov; idx = oi; }
        }
    }
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

// >1: hs_disp_eval runs every point on 32 emulated lanes (the warp form of K1) instead of one lane
static int g_eval_lanes = 1;

static void eval_points(psi_ctx* c, long long n, const int* pidx, const int* dist, const double* kx,
                        const double* kz, const double* alpha, const double* w, double* D, double* M) {
    unsigned long long total = 0;
    if (g_eval_lanes > 1) {
        for (long long i = 0; i < n; ++i) {
            const long long pi = pidx ? pidx[i] : i;
            pthread_barrier_init(&WarpEmu::bar, nullptr, 32);
            std::vector<std::thread> lanes;
            for (int l = 0; l < 32; ++l)
                lanes.emplace_back([&, l]() {
                    WarpEmu::tl_lane = l;
                    unsigned long long nodes = 0;
                    cx m[7];
                    const cx v = disp_eval_any<WarpEmu>(c->view, c->dists[dist[pi]], cx(w[2 * i], w[2 * i + 1]),
                                                        kx[pi], kz[pi], alpha[pi], m, nodes);
                    WarpEmu::sync();
                    if (l == 0) {
                        D[2 * i] = v.re; D[2 * i + 1] = v.im;
                        if (M) for (int k = 0; k < 7; ++k) { M[14 * i + 2 * k] = m[k].re; M[14 * i + 2 * k + 1] = m[k].im; }
                        total += nodes;
                    }
                });
            for (auto& t : lanes) t.join();
            pthread_barrier_destroy(&WarpEmu::bar);
        }
        c->nodes += total;
        return;
    }
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

// CPU stand-in for kernels K2-K4: run `body` with the evaluator of one item
template <class F>
static void with_eval(psi_ctx* c, int dist, double kx, double kz, double alpha, F&& body) {
    WarpEval<WarpOne> ev;
    ev.t = &c->view; ev.d = &c->dists[dist]; ev.kx = kx; ev.kz = kz; ev.alpha = alpha; ev.nodes = 0; ev.evals = 0;
    body(ev);
}

int psi_internal_polish(psi_ctx* c, int n_items, const int* pidx, int n_params, const int* dist,
                        const double* kx, const double* kz, const double* alpha, const double* dom,
                        const int* off, double* w0, const int* max_iter, double* out, int* n_calls,
                        int* max_conv, int* status, long long* evals) {
    (void)n_params;
    long long total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int i = 0; i < n_items; ++i) {
        const int pi = pidx[i];
        with_eval(c, dist[pi], kx[pi], kz[pi], alpha[pi], [&](auto& ev) {
            RectDom rd{dom[4 * i], dom[4 * i + 1], dom[4 * i + 2], dom[4 * i + 3]};
            int conv = 0, st = 0;
            n_calls[i] = polish_run(ev, rd, reinterpret_cast<cx*>(w0) + off[i], off[i + 1] - off[i], max_iter[i],
                                    reinterpret_cast<cx*>(out) + off[i], &conv, &st);
            max_conv[i] = conv; status[i] = st;
            total += ev.evals;
        });
    }
    if (evals) *evals = total;
    return PSI_OK;
}

// >1: psi_internal_mode_device runs every search on 32 emulated lanes (WarpEmu) instead of one lane
static int g_mode_lanes = 1;

static int mode_run_emulated(psi_ctx* c, int dist, double kx, double kz, double alpha, const ModeCfg& cfg,
                             const cx* guesses, int ng, const double* uni, int cap, cx* roots, int max_roots,
                             int* nr_out, int* nc_out, long long* evals) {
    std::vector<unsigned char> buf(aaa_scratch_bytes(cap));
    const AAAScratch sc = aaa_carve(buf.data(), cap);
    pthread_barrier_init(&WarpEmu::bar, nullptr, 32);
    int status = 0;
    std::vector<std::thread> lanes;
    std::vector<cx> lane_roots((size_t)32 * max_roots);       // every lane keeps its own copy of the outputs
    for (int l = 0; l < 32; ++l)
        lanes.emplace_back([&, l]() {
            WarpEmu::tl_lane = l;
            WarpEval<WarpEmu> ev;
            ev.t = &c->view; ev.d = &c->dists[dist]; ev.kx = kx; ev.kz = kz; ev.alpha = alpha; ev.nodes = 0; ev.evals = 0;
            int nr = 0, nc = 0;
            const int st = mode_run<WarpEmu>(ev, cfg, guesses, ng, uni, sc, lane_roots.data() + (size_t)l * max_roots,
                                             max_roots, &nr, &nc);
            WarpEmu::sync();
            if (l == 0) { status = st; *nr_out = nr; *nc_out = nc; *evals = ev.evals; }
        });
    for (auto& t : lanes) t.join();
    pthread_barrier_destroy(&WarpEmu::bar);
    for (int k = 0; k < max_roots; ++k) roots[k] = lane_roots[k];
    return status;
}

int psi_internal_mode_device(psi_ctx* c, int n_items, const int* dist, const double* kx, const double* kz,
                             const double* alpha, const double* domain, const int* iparams,
                             const double* dparams, const double* uni, long long n_uni, const int* uni_off,
                             const int* guess_off, const double* guesses, int cap, int max_roots,
                             double* roots, int* n_roots, int* n_calls, int* status, long long* stats,
                             const PsiModeExtra* extra) {
    (void)n_uni;
    if (extra && extra->rec_dev) return psi_internal_fail(PSI_ERR_ARG, "no device records in the CPU harness");
    long long total = 0;
    if (g_mode_lanes > 1) {
        for (int i = 0; i < n_items; ++i) {
            ModeCfg cfg;
            cfg.dom = RectDom{domain[4 * i], domain[4 * i + 1], domain[4 * i + 2], domain[4 * i + 3]};
            cfg.n_sample = iparams[3 * i]; cfg.max_zoom = iparams[3 * i + 1]; cfg.max_secant = iparams[3 * i + 2];
            cfg.tol = dparams[2 * i]; cfg.clean_tol = dparams[2 * i + 1];
            const int g0 = guess_off[i], ng = guess_off[i + 1] - g0;
            long long ev_n = 0;
            status[i] = mode_run_emulated(c, dist[i], kx[i], kz[i], alpha[i], cfg,
                                          reinterpret_cast<const cx*>(guesses) + g0, ng, uni + uni_off[i], cap,
                                          reinterpret_cast<cx*>(roots) + (size_t)i * max_roots, max_roots,
                                          &n_roots[i], &n_calls[i], &ev_n);
            total += ev_n;
        }
        if (stats) { stats[0] = 1; stats[1] = total; stats[2] = 0; }
        return PSI_OK;
    }
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int i = 0; i < n_items; ++i) {
        with_eval(c, dist[i], kx[i], kz[i], alpha[i], [&](auto& ev) {
            std::vector<unsigned char> buf(aaa_scratch_bytes(cap));
            const AAAScratch sc = aaa_carve(buf.data(), cap);
            ModeCfg cfg;
            cfg.dom = RectDom{domain[4 * i], domain[4 * i + 1], domain[4 * i + 2], domain[4 * i + 3]};
            cfg.n_sample = iparams[3 * i]; cfg.max_zoom = iparams[3 * i + 1]; cfg.max_secant = iparams[3 * i + 2];
            cfg.tol = dparams[2 * i]; cfg.clean_tol = dparams[2 * i + 1];
            const int g0 = guess_off[i], ng = guess_off[i + 1] - g0;
            int nr = 0, nc = 0;
            ModeDetail det;
            status[i] = mode_run<WarpOne>(ev, cfg, reinterpret_cast<const cx*>(guesses) + g0, ng, uni + uni_off[i], sc,
                                          reinterpret_cast<cx*>(roots) + (size_t)i * max_roots, max_roots, &nr, &nc,
                                          &det);
            n_roots[i] = nr; n_calls[i] = nc;
            if (extra) {
                if (extra->M) extra->M[i] = det.M;
                if (extra->upos) extra->upos[i] = det.upos;
                if (extra->conv) extra->conv[i] = det.conv;
                for (int k = 0; k < det.M; ++k) {
                    if (extra->Z) { extra->Z[2 * ((size_t)i * cap + k)] = sc.Z[k].re; extra->Z[2 * ((size_t)i * cap + k) + 1] = sc.Z[k].im; }
                    if (extra->F) { extra->F[2 * ((size_t)i * cap + k)] = sc.F[k].re; extra->F[2 * ((size_t)i * cap + k) + 1] = sc.F[k].im; }
                    if (extra->mask) extra->mask[(size_t)i * cap + k] = sc.mask[k];
                }
            }
            total += ev.evals;
        });
    }
    if (stats) { stats[0] = 1; stats[1] = total; stats[2] = 0; }
    return PSI_OK;
}

#include "../../psitools_public_b200/csrc/psi_engine.cu"

extern "C" {

const char* hs_last_error(void) { return g_err.c_str(); }
// Loewner fit of psi_aaa.cuh for a given support mask: weights (m values, support points in sample order)
// and the returned residual ratio; lets the CPU suite check the QR + inverse-iteration weights against
// numpy's SVD.  Returns m.
int hs_aaa_fit(int M, const double* Z, const double* F, const int* mask, double* w_out, double* resid) {
    std::vector<unsigned char> buf(aaa_scratch_bytes(M));
    const AAAScratch sc = aaa_carve(buf.data(), M);
    AAAState st;
    st.M = M; st.tol = 1e-13; st.clean_tol = 1e-4; st.dom = RectDom{-1e300, 1e300, -1e300, 1e300};
    st.memo_next = 0; st.slot = 0; st.v_m = 0;
    double fm = 0.0;
    for (int i = 0; i < M; ++i) {
        sc.Z[i] = cx(Z[2 * i], Z[2 * i + 1]); sc.F[i] = cx(F[2 * i], F[2 * i + 1]); sc.mask[i] = mask[i];
        fm = fmax(fm, abs_c(sc.F[i]));
    }
    st.fmax = fm;
    aaa_memo_clear<WarpOne>(sc, st);
    *resid = aaa_fit<WarpOne>(sc, st);
    for (int j = 0; j < st.m; ++j) { w_out[2 * j] = sc.w[j].re; w_out[2 * j + 1] = sc.w[j].im; }
    return st.m;
}

void hs_set_mode_lanes(int lanes) { g_mode_lanes = lanes; }
void hs_set_eval_lanes(int lanes) { g_eval_lanes = lanes; }

// The same fit run by 32 emulated lanes (WarpEmu); general_solves != 0 forces the memory form of the inverse
// iteration's triangular solves, which must give bit-identical weights to the register form.
int hs_aaa_fit_warp(int M, const double* Z, const double* F, const int* mask, double* w_out, double* resid,
                    int general_solves) {
    std::vector<unsigned char> buf(aaa_scratch_bytes(M));
    const AAAScratch sc = aaa_carve(buf.data(), M);
    double fm = 0.0;
    for (int i = 0; i < M; ++i) {
        sc.Z[i] = cx(Z[2 * i], Z[2 * i + 1]); sc.F[i] = cx(F[2 * i], F[2 * i + 1]); sc.mask[i] = mask[i];
        fm = fmax(fm, abs_c(sc.F[i]));
    }
    g_aaa_general_solves = general_solves != 0;
    pthread_barrier_init(&WarpEmu::bar, nullptr, 32);
    int m_out = 0;
    double res_out = 0.0;
    std::vector<std::thread> lanes;
    for (int l = 0; l < 32; ++l)
        lanes.emplace_back([&, l]() {
            WarpEmu::tl_lane = l;
            AAAState st;                      // per lane, like the registers of a device thread
            st.M = M; st.m = 0; st.r = 0; st.tol = 1e-13; st.clean_tol = 1e-4;
            st.dom = RectDom{-1e300, 1e300, -1e300, 1e300};
            st.memo_next = 0; st.slot = 0; st.v_m = 0; st.fmax = fm;
            aaa_memo_clear<WarpEmu>(sc, st);
            WarpEmu::sync();
            const double r = aaa_fit<WarpEmu>(sc, st);
            WarpEmu::sync();
            if (l == 0) { m_out = st.m; res_out = r; }
        });
    for (auto& t : lanes) t.join();
    pthread_barrier_destroy(&WarpEmu::bar);
    g_aaa_general_solves = false;
    *resid = res_out;
    for (int j = 0; j < m_out; ++j) { w_out[2 * j] = sc.w[j].re; w_out[2 * j + 1] = sc.w[j].im; }
    return m_out;
}

// the per-pixel sweep logic of csrc/psi_refine.cuh over a whole map: nk[p] = guesses pixel p runs with (0 = it
// does not run), kept[p*128 ..] = those guesses
void hs_refine_count(int nz, int nx, const int* runs, const int* nguess, int width, const double* mat,
                     const int* cnt, int* nk, double* kept) {
    RefineView g{nz, nx, width, runs, cnt, reinterpret_cast<const cx*>(mat), nguess};
    std::vector<cx> near(kRefineMaxNear);
    for (int p = 0; p < nz * nx; ++p)
        nk[p] = refine_pixel(g, p / nx, p % nx, near.data(), reinterpret_cast<cx*>(kept) + (size_t)p * kRefineMaxNear);
}

// csrc/psi_tv.cuh on the host: what one thread of tv_roots_kernel runs, pair by pair
void hs_tv_roots(double fd, double tau_max, double b, double tau_min, double diffusion, int max_it, long long n,
                 const double* kx, const double* kz, double* omega, int* status) {
    TvParams P{fd, tau_max, b, tau_min / tau_max, diffusion, max_it};
    for (long long i = 0; i < n; ++i) {
        cx w;
        status[i] = tv_find_root(P, kx[i], kz[i], &w);
        omega[2 * i] = w.re; omega[2 * i + 1] = w.im;
    }
}

void hs_aaa_stats(long long* out) { for (int i = 0; i < 16; ++i) { out[i] = g_aaa_stats[i]; g_aaa_stats[i] = 0; } }

int psi_polish_batch_host(psi_ctx* c, int n_items, const int* dist, const double* kx, const double* kz,
                          const double* alpha, const double* domain, const int* w0_off, const double* w0,
                          const int* max_iter, double* roots, int* n_calls, int* max_conv, int* status) {
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
    const int cap = 4096;
    long long total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int i = 0; i < n_items; ++i) {
        with_eval(c, dist[i], kx[i], kz[i], alpha[i], [&](auto& ev) {
            std::vector<cx> buf(4 * (size_t)cap);
            const int o = z_off[i], n0 = z_off[i + 1] - o;
            for (int k = 0; k < n0; ++k) buf[k] = cx(z_list[2 * (o + k)], z_list[2 * (o + k) + 1]);
            int cnt = -666, np = n0;
            status[i] = contour_run(ev, buf.data(), buf.data() + cap, buf.data() + 2 * cap, buf.data() + 3 * cap,
                                    n0, cap, tol[i], max_step[i], max_iter[i], &cnt, &np);
            counts[i] = cnt; n_points[i] = np;
            total += ev.evals;
        });
    }
    if (stats) { stats[0] = 1; stats[1] = total; }
    return PSI_OK;
}

psi_ctx* hs_ctx_create(const double* xj, const double* wj, int n, int max_m, double eps) {
    psi_ctx* c = new psi_ctx();
    build_level_major(xj, wj, n, max_m, c->lv, c->off);
    c->view.nodes = c->lv.data(); c->view.level_off = c->off.data();
    c->view.n_nodes = n; c->view.max_m = max_m; c->view.eps = eps; c->view.predict_exit = 0; c->view.level_cap = 0;
    return c;
}
void hs_ctx_destroy(psi_ctx* c) { delete c; }
void hs_set_exit_rule(psi_ctx* c, int predict) { c->view.predict_exit = predict; }      // 0, 1 or 2 (fast path)
void hs_set_level_cap(psi_ctx* c, int cap) { c->view.level_cap = cap; }

static std::vector<std::vector<double>> g_tables;      // tabulated size densities stay alive with the process

int hs_dist_add_table(psi_ctx* c, const psi_dist_desc* desc, const psi_table_desc* tb);

int hs_dist_add(psi_ctx* c, const psi_dist_desc* desc) { return hs_dist_add_table(c, desc, nullptr); }

int hs_dist_add_table(psi_ctx* c, const psi_dist_desc* desc, const psi_table_desc* tb) {
    DistParams d = dist_from_desc(*desc);
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
        g_tables.emplace_back(tb->coef, tb->coef + 4 * cells);
        d.tab = g_tables.back().data();
    }
    unsigned long long nodes = 0;
    NodeTableView literal = c->view;
    literal.predict_exit = 0;                 // like the product: background integrals use the literal rule
    background_warp<WarpOne>(literal, d, nodes);
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
