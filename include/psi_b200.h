/* psi_b200.h -- C ABI of libpsi_b200.so: the B200 (sm_100a) implementation of the psitools
 * D(omega; Kx, Kz) hot path.
 *
 * The reference (psitools, pure Python) has no FFI boundary; its boundary is the Python class API
 * (SURVEY.md section 8b).  The entry points below are what a ctypes binding placed under those
 * classes calls; each one names the reference interface it replaces (file:line relative to the
 * psitools tree).  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C types only; complex128 arrays are interleaved (re, im) doubles;
 *   - every function returns PSI_OK (0) or a negative PSI_ERR_* code; psi_last_error() gives the text;
 *     nothing throws, nothing falls back to the CPU: without a CUDA device psi_ctx_create fails;
 *   - "*_host" entry points take HOST pointers and perform the H2D/D2H copies themselves
 *     (synchronous on return); "*_dev" entry points take DEVICE pointers, enqueue on `stream`
 *     (a cudaStream_t passed as void*; NULL = the context's own stream) and return without
 *     synchronising;
 *   - a context is bound to one device and may be used from one host thread at a time.
 */
#ifndef PSI_B200_H
#define PSI_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PSI_OK 0
#define PSI_ERR_CUDA (-1)      /* a CUDA runtime call failed (text in psi_last_error) */
#define PSI_ERR_ARG (-2)       /* invalid argument */
#define PSI_ERR_NODEVICE (-3)  /* no usable CUDA device: there is NO CPU fallback */
#define PSI_ERR_CAPACITY (-4)  /* a fixed capacity (distributions, poles, contour points) exceeded */

#define PSI_KIND_POWERLAW 0       /* psi_dispersion.py:57-66 (default MRN-like size density) */
#define PSI_KIND_POWERBUMP 1      /* power_bump.py:38-130  PowerBump.sigma0 */
#define PSI_KIND_POWERBUMPTAIL 2  /* power_bump.py:133-212 PowerBumpTail.sigma0 */
#define PSI_MAX_POLES 4

typedef struct psi_ctx psi_ctx;

/* One PSIDispersion instance (psi_dispersion.py:43-84): dust-to-gas ratio, Stokes range, sound
 * speed, single-size flag and an analytic size density.  rhod is SizeDensity.rhod
 * (sizedensity.py:37-40), poles is SizeDensity.poles (sizedensity.py:44), unscaled. */
typedef struct psi_dist_desc {
    int kind;
    int single_size;
    double mu, tau_min, tau_max, sound_speed;
    double rhod;
    double q;                                   /* POWERLAW: sigma(a) = a^q, q = 3 - power */
    double beta, norm_pl, amp, curv, aL, aP, aBT, scale;   /* POWERBUMP / POWERBUMPTAIL */
    int n_poles;
    double poles[PSI_MAX_POLES];
} psi_dist_desc;

/* Tabulated size density (kind PSI_KIND_TABLE): for ANY callable sigma(a) the reference accepts
 * (SizeDensity(sigma, size_range), sizedensity.py:25-48, consumed at psi_dispersion.py:57-67, 139-145) the host
 * tabulates the quadrature weight w(u) = e^u F(e^u), u = ln s, s = tau/tau_max in [tau_min/tau_max, 1], once per
 * distribution: per smooth piece between the size density's kinks (SizeDensity.poles / tau_max) a uniform grid of
 * cells, on each cell the cubic v(t) = c0 + t (c1 + t (c2 + t c3)), t in [0, 1), that interpolates w -- or ln w where
 * w > 0 -- at four points of the cell.  Four doubles per cell, pieces one after the other: the device reads a cell
 * with two 16-byte loads.  brk: n_seg + 1 break points (brk[0] = ln(tau_min/tau_max), brk[n_seg] = 0). */
#define PSI_KIND_TABLE 3
#define PSI_MAX_TABLE_SEGS (PSI_MAX_POLES + 1)
typedef struct psi_table_desc {
    int n_seg;
    double brk[PSI_MAX_TABLE_SEGS + 1];
    int n_cells[PSI_MAX_TABLE_SEGS];
    int log_mode[PSI_MAX_TABLE_SEGS];
    const double* coef;                         /* host pointer, 4 * sum(n_cells) doubles */
} psi_table_desc;

/* Library / device info. */
const char* psi_version(void);
const char* psi_last_error(void);
int psi_device_count(void);

/* Context = one device + one tanh-sinh node table.  Replaces TanhSinh.__init__'s tables
 * (tanhsinh.py:34-66): the caller uploads the exact xj/wj it built on the host (n_nodes entries,
 * original order), with max_level and eps = 10^-precision_digit. */
int psi_ctx_create(psi_ctx** out, int device, const double* xj, const double* wj, int n_nodes,
                   int max_level, double eps);
void psi_ctx_destroy(psi_ctx* ctx);
int psi_ctx_sm_count(psi_ctx* ctx);
/* Exit rule of the tanh-sinh level loop in K1-K4 (tanhsinh.py:120-126: stop when an increment is below
 * 1e-15 ABSOLUTE, discarding it).  predict = 0: that rule, literally.  predict = 1 (default; environment
 * PSI_B200_EXIT_RULE=reference selects 0): additionally stop as soon as the last two increments show the
 * quadratic collapse of the rule and the extrapolated next increment is 1e4 times below the tolerance --
 * the levels skipped only resolve rounding noise (csrc/psi_quad.cuh). */
int psi_ctx_set_exit_rule(psi_ctx* ctx, int predict);
int psi_ctx_get_exit_rule(psi_ctx* ctx);
int psi_ctx_synchronize(psi_ctx* ctx);
/* kernels launched by this context since creation (for bench.py's gpu_launches) */
long long psi_ctx_launch_count(psi_ctx* ctx);

/* Register a distribution and compute its background state on the device (kernel K0):
 * replaces PSIDispersion.__init__ (psi_dispersion.py:43-84) + int_j (:183-187).
 * Returns the distribution id (>= 0) or an error code. */
int psi_dist_add(psi_ctx* ctx, const psi_dist_desc* desc);
/* The same for a tabulated size density: desc->kind = PSI_KIND_TABLE with mu, tau range, sound speed, rhod,
 * single_size and poles filled in (the analytic-family fields are ignored); the table is copied to the device. */
int psi_dist_add_table(psi_ctx* ctx, const psi_dist_desc* desc, const psi_table_desc* table);
/* out = {J0, J1, denom, vgx, vgy}  (psi_dispersion.py:74-79) */
int psi_dist_get_background(psi_ctx* ctx, int dist, double out[5]);

/* Kernel K1 -- replaces PSIDispersion.calculate (psi_dispersion.py:375-456) for a batch of n
 * points (one warp per point).  dist/kx/kz/alpha have n entries, or ONE entry when broadcast != 0.
 * w and D are n interleaved complex128; M (optional, may be NULL) receives 7 complex per point
 * {M11,M12,M13,M21,M22,M23,M33} (psi_dispersion.py:363-373); node_evals (optional) receives the
 * number of joint integrand evaluations performed for the whole batch. */
int psi_disp_eval_host(psi_ctx* ctx, long long n, int broadcast, const int* dist, const double* kx,
                       const double* kz, const double* alpha, const double* w, double* D, double* M,
                       unsigned long long* node_evals);
/* Same with device pointers; node_evals_dev (optional) is a device counter that is INCREMENTED.
 * K1 is compiled once per work class (size-density family x viscous/inviscid, psi_point_class);
 * class_mask has bit c set when the batch may contain points of class c (0 = all classes, i.e. one
 * launch per class, each skipping foreign points).  stream: a cudaStream_t (NULL = the context's own); the
 * launches of ONE context share its work queue and counters, so calls given different streams are chained
 * (each waits for the context's previous K1 launch): use one context per stream for concurrent batches. */
int psi_disp_eval_dev(psi_ctx* ctx, long long n, int broadcast, const int* dist, const double* kx,
                      const double* kz, const double* alpha, const double* w, double* D, double* M,
                      unsigned long long* node_evals_dev, unsigned class_mask, void* stream);
/* Work class (0..9: node generator {MRN sqrt, general power law, PowerBump, PowerBumpTail, table} x {inviscid,
 * viscous}) of a point of distribution `dist` at viscosity parameter alpha. */
int psi_point_class(psi_ctx* ctx, int dist, double alpha);

/* ---- native lock-step engine (csrc/psi_engine.cu): whole batches of root searches / contour counts,
 * one K1 launch per pass over the union of all points the items need.  Host pointers throughout. ---- */

/* Bind LAPACK (zgesdd for the AAA Loewner SVD, zggev for the arrowhead pencils) from the OpenBLAS
 * shared objects that numpy (ILP64, symbol scipy_zgesdd_64_) and scipy (LP64, scipy_zggev_) load, so
 * that the engine reproduces numpy.linalg.svd / scipy.linalg.eig of the same interpreter. */
int psi_engine_bind_lapack(const char* numpy_openblas64, const char* scipy_openblas32);

/* Host threads the engine uses for the per-item logic between kernel launches (AAA, state machines).
 * n <= 0: OpenMP default.  Under torchrun (OMP_NUM_THREADS=1) the Python layer sets
 * cpu_count / LOCAL_WORLD_SIZE. */
void psi_engine_set_threads(int n);

/* profiling aid: {zgesdd calls, zgesdd ns, zggev calls, zggev ns} accumulated since load */
void psi_engine_counters(long long out[4]);

/* n_items PSIMode.calculate searches (psi_mode.py:106-333) in lock-step.
 *   dist, kx, kz, alpha : per item (PSIMode(...).dispersion of the item, wave numbers, viscous_alpha)
 *   domain   : 4 doubles per item {real_range[0], real_range[1], imag_range[0], imag_range[1]}
 *   iparams  : 3 ints per item {n_sample, max_zoom_domains, max_secant_iterations}
 *   dparams  : 2 doubles per item {tol, clean_tol}
 *   seeds    : per item random_seed (numpy legacy MT19937 stream, psi_mode_mpi.py:164-167)
 *   guess_off: n_items+1 offsets into guesses (interleaved complex) = guess_roots of each item
 *   roots    : max_roots complex slots per item; n_roots = roots found (may exceed max_roots ->
 *              status 2, truncated); n_calls = PSIMode.n_function_call; status 0 ok / 1 failed
 *   stats    : 7 values {rounds, D evaluations, failed items, host-logic us, device us, kernel
 *              launches, D evaluations inside K3} (optional) */
int psi_mode_batch_host(psi_ctx* ctx, int n_items, const int* dist, const double* kx, const double* kz,
                        const double* alpha, const double* domain, const int* iparams, const double* dparams,
                        const unsigned* seeds, const int* guess_off, const double* guesses, int max_roots,
                        double* roots, int* n_roots, int* n_calls, int* status, long long* stats);

/* Kernel K4 -- the same batch (same arguments, same outputs) with every search running ENTIRELY on the
 * device, as a pipeline of homogeneous kernels over all searches, repeated once per zoom level:
 * sampling (thread per search) -> K1 on the searches' own sample buffers -> AAA analysis (8/16/32 lanes per
 * search: greedy Loewner fits by Householder QR + inverse iteration, Froissart cleanup and zeros by
 * Ehrlich-Aberth, clean_tol halving, zoom plan; csrc/psi_aaa.cuh, psi_mode.cuh) -> K3 secant polish -> decision.
 * Host traffic per level: two 16-byte read-backs (request count, "anyone zooming?").  Replaces
 * PSIMode.calculate (psi_mode.py:106-333) + RationalApproximation (complex_roots.py:38-368) for whole maps.
 * status: 0 ok, 1 non-finite samples (ValueError in the reference), 2 sample capacity, 3 start value on the
 * real axis (ValueError), 4 more roots than max_roots.  stats (16 slots) = {launches, D evaluations, node
 * evaluations, -, -, -, -, -, then microseconds on the stream of the five stages sample, K1, AAA analysis,
 * K3 polish, decide}.  At most PSI_MODE_MAX_SAMPLES samples per search. */
#define PSI_MODE_MAX_SAMPLES 512
int psi_mode_batch_device(psi_ctx* ctx, int n_items, const int* dist, const double* kx, const double* kz,
                          const double* alpha, const double* domain, const int* iparams, const double* dparams,
                          const unsigned* seeds, const int* guess_off, const double* guesses, int max_roots,
                          double* roots, int* n_roots, int* n_calls, int* status, long long* stats);

/* The same batch with the results left ON THE DEVICE: *rec_dev receives a device pointer to rec_rows
 * (>= n_items, the excess zeroed) packed records of 4 + 2*max_roots doubles
 *     {n_roots, status, n_function_call, 0, Re root_0, Im root_0, ..., root_{max_roots-1}}
 * valid until the next call on the context -- what one rank of a multi-GPU map hands to ncclAllGather
 * (torch.distributed.all_gather_into_tensor on a tensor wrapping the pointer) without a detour through host
 * memory; replaces the pickled point-to-point result traffic of psi_mode_mpi.py:57-160 and the per-sweep
 * bcast of psi_grid_refine.py:325-326. */
int psi_mode_batch_device_rec(psi_ctx* ctx, int n_items, const int* dist, const double* kx, const double* kz,
                              const double* alpha, const double* domain, const int* iparams, const double* dparams,
                              const unsigned* seeds, const int* guess_off, const double* guesses, int max_roots,
                              int rec_rows, double** rec_dev, long long* stats);

/* ONE search (PSIMode.calculate, psi_mode.py:106-333) through the same pipeline, with the caller's own stream
 * of U[0,1) numbers (numpy's global legacy stream: n_uni >= 2*n_sample*(1 + n_guess + 4*max_zoom)) and
 * everything PSIMode exposes afterwards: roots, n_function_call, z_sample / f_sample (cap =
 * n_sample*(1 + n_guess + 4*max_zoom) complex slots each, *n_samples used), ra.maskF (cap ints), the count
 * of random numbers consumed (so that the caller can advance its generator exactly as the reference would)
 * and max_convergence_iterations.  domain = {xmin, xmax, ymin, ymax}, iparams = {n_sample, max_zoom_domains,
 * max_secant_iterations}, dparams = {tol, clean_tol}. */
int psi_mode_search_detail(psi_ctx* ctx, int dist, double kx, double kz, double alpha, const double* domain,
                           const int* iparams, const double* dparams, const double* uni, int n_uni,
                           const double* guesses, int n_guess, int max_roots, double* roots, int* n_roots,
                           int* n_calls, int* status, double* z_sample, double* f_sample, int* n_samples,
                           int* mask, int* n_uni_used, int* max_conv);

/* Kernel K2 -- n_items ClosedPath(f, z_list).count_roots(max_iter, tol, max_step)
 * (complex_roots.py:517-625) in ONE launch: one warp per contour runs every refinement pass and the
 * winding sum on the device.
 *   z_off : n_items+1 offsets into z_list (interleaved complex contour corners, counter-clockwise)
 *   counts: winding number, or -666 when it did not converge (complex_roots_mpi.py:171-174)
 *   status: 0 ok, 1 no integer winding number, 2 max_iter exceeded (both RuntimeError in the
 *           reference), 3 non-finite D (ValueError), 4 more than 4096 contour points
 *   stats : {launches, D evaluations} (optional) */
int psi_contour_batch_host(psi_ctx* ctx, int n_items, const int* dist, const double* kx, const double* kz,
                           const double* alpha, const int* z_off, const double* z_list, const double* tol,
                           const double* max_step, const int* max_iter, int* counts, int* status,
                           int* n_points, long long* stats);

/* Kernel K3 -- n_items PSIMode.find_dispersion_roots(zeros, max_iter) (psi_mode.py:340-352, 411-491) in
 * ONE launch: one warp per item polishes its start values one after the other with the two-stage secant
 * iteration (scipy newton's rule), every D evaluated on the device.
 *   domain : 4 doubles per item; w0_off: n_items+1 offsets into w0 (interleaved complex start values)
 *   roots  : one complex per start value: the root, or the sentinel -1j (failed / left the domain)
 *   n_calls: D evaluations as PSIMode.n_function_call counts them; max_conv: largest iteration count of
 *            a converged start; status 1 = a start value on the real axis (ValueError in the reference) */
int psi_polish_batch_host(psi_ctx* ctx, int n_items, const int* dist, const double* kx, const double* kz,
                          const double* alpha, const double* domain, const int* w0_off, const double* w0,
                          const int* max_iter, double* roots, int* n_calls, int* max_conv, int* status);

/* FP64 FMA-pipe peak of the device (roofline denominator; MEASURED_PEAKS.json has no FP64 entry):
 * runs a register-resident DFMA chain kernel `reps` times and returns the best TFLOP/s. */
int psi_measure_fp64_peak(psi_ctx* ctx, int reps, double* tflops);

/* ---- grid refinement (csrc/psi_refine.cu): the cell queue of PSIGridRefiner.sweep_last_grid
 * (psi_grid_refine.py:218-290) built on the device from a mirror of the root map.
 *   runs    : nz*nx run id per pixel (-1 = never run), row-major [iz][ix]
 *   nguess  : nz*nx number of guesses the pixel was last run with
 *   cnt/mat : per run id 0..n_keys-1 the number of roots (-1 = no such run) and `width` complex root slots
 *   key_capacity >= n_keys: room for run ids added by psi_refine_update
 * psi_refine_sweep decides for every root-less pixel that touches a pixel with roots whether it runs (more
 * pruned neighbour roots than nguess) and returns the number of running pixels and of their guesses
 * (n_run = -1: a pixel collected more than 128 neighbour roots, nothing was changed, sweep on the host);
 * psi_refine_fetch writes the compact row-major work list {pixel index, guess count, guesses} (and makes
 * the guess count the pixel's new nguess); psi_refine_update stores the roots of finished runs (`width`
 * slots per changed run id) and re-points pixels to their new run ids. */
typedef struct psi_refine psi_refine;
int psi_refine_open(psi_refine** out, int device, int nz, int nx, const int* runs, const int* nguess,
                    long long n_keys, long long key_capacity, int width, const double* mat, const int* cnt);
int psi_refine_sweep(psi_refine* map, long long* n_run, long long* n_guess);
int psi_refine_fetch(psi_refine* map, int* pix, int* n_kept, double* guesses);
int psi_refine_update(psi_refine* map, long long n_changed, const long long* keys, const int* cnt,
                      const double* rows, long long n_pix, const int* pix, const int* run_ids);
void psi_refine_close(psi_refine* map);

/* ---- guess generator: terminal-velocity PSI mode on a map of wave-number pairs --------------------------
 * Replaces a Python loop over PSIModeTV.find_roots (reference terminalvelocitysolver.py:368-418, one wave-number
 * pair per call) for ONE minimum Stokes number: one thread per pair follows the growing root of the closed-form
 * terminal-velocity relation from the single-size limit tau_min = tau_max down to tau_min (step-size controlled
 * continuation, scipy's secant rule; csrc/psi_tv.cuh).  dust_fraction = mu/(1+mu); b = -(3 + power-law exponent),
 * one of 0, 0.5, 1.5, 2.5 (PSI_ERR_ARG otherwise); diffusion = viscous_alpha * c_over_eta^2; max_it = PSIModeTV's
 * maximum_iterations.  omega: n complex mode frequencies; status[i] = 0 mode found, 1 gave up (omega is then the
 * value the reference returns for z = 0).  Needs no psi_ctx. */
int psi_tv_roots(int device, double dust_fraction, double tau_max, double b, double tau_min, double diffusion,
                 int max_it, long long n, const double* kx, const double* kz, double* omega, int* status);

#ifdef __cplusplus
}
#endif
#endif /* PSI_B200_H */
