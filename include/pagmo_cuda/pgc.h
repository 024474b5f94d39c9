/* pgc.h - C ABI of the B200-native pagmo batch-evaluation / generation engine (libpgc.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, `extern "C"`, no C++/torch types.
 * Every entry point names the reference interface it stands in for (file:line under esa/pagmo2 v2.19.1).
 * A maintainer binds these from pagmo's C++ plugin layer - see include/pagmo_cuda/cuda_bfe.hpp for the
 * header-only adapters (UDBFE `cuda_bfe`, CUDA-backed UDPs with `batch_fitness`) and INTEGRATION.md.
 *
 * Conventions
 *   - every function returns PGC_OK (0) or a negative pgc_status; the message of the last failure on the
 *     calling thread is available from pgc_last_error().  There is NO CPU fallback: without a usable
 *     CUDA device every compute entry point fails with PGC_ERR_CUDA.
 *   - decision vectors are flat row-major [n x nx] doubles, fitness vectors flat row-major [n x nf]
 *     (same layout as pagmo::bfe, reference include/pagmo/bfe.hpp:320, src/problem.cpp:383-410).
 *   - `stream` arguments are a cudaStream_t passed as void* (NULL = the context's own stream).
 */
#ifndef PAGMO_CUDA_PGC_H
#define PAGMO_CUDA_PGC_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__) || defined(__clang__)
#define PGC_API __attribute__((visibility("default")))
#else
#define PGC_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pgc_status {
    PGC_OK = 0,
    PGC_ERR_INVALID_ARGUMENT = -1, /* maps to std::invalid_argument in the C++ adapters (pagmo_throw) */
    PGC_ERR_UNSUPPORTED = -2,      /* problem/operator not implemented on the device: adapters throw, no fallback */
    PGC_ERR_CUDA = -3,             /* CUDA runtime failure / no device: std::runtime_error */
    PGC_ERR_OUT_OF_MEMORY = -4
} pgc_status;

/* Problem families = the reference's UDPs on the hot path. */
typedef enum pgc_family {
    PGC_RASTRIGIN = 1,     /* src/problems/rastrigin.cpp:62-72 */
    PGC_ACKLEY = 2,        /* src/problems/ackley.cpp:61-76 */
    PGC_GRIEWANK = 3,      /* src/problems/griewank.cpp:60-75 */
    PGC_SCHWEFEL = 4,      /* src/problems/schwefel.cpp:60-69 */
    PGC_ROSENBROCK = 5,    /* src/problems/rosenbrock.cpp:59-66 */
    PGC_CEC2014 = 6,       /* src/problems/cec2014.cpp:119-247 */
    PGC_CEC2013 = 7,       /* src/problems/cec2013.cpp:78-197 */
    PGC_ZDT = 8,           /* src/problems/zdt.cpp:233-356 */
    PGC_DTLZ = 9,          /* src/problems/dtlz.cpp:258-409 */
    PGC_WFG = 10,          /* src/problems/wfg.cpp:304-1066 */
    PGC_LENNARD_JONES = 11, /* src/problems/lennard_jones.cpp:72-92 */
    /* meta-problems: created by pgc_problem_translate / pgc_problem_decompose only */
    PGC_TRANSLATE = 12, /* src/problems/translate.cpp:100-153 */
    PGC_DECOMPOSE = 13, /* src/problems/decompose.cpp:139-154 */
    /* constrained UDPs: fitness rows are [objective | nec equality | nic inequality constraints] (problem.cpp:383-410) */
    PGC_HOCK_SCHITTKOWSKI_71 = 14, /* src/problems/hock_schittkowski_71.cpp:48-55: nx 4, 1 equality, 1 inequality */
    PGC_LUKSAN_VLCEK1 = 15,        /* src/problems/luksan_vlcek1.cpp:60-77: dim >= 3, dim - 2 equalities */
    PGC_UNCONSTRAIN = 16           /* src/problems/unconstrain.cpp:136-223; created by pgc_problem_unconstrain only */
} pgc_family;

/* unconstrain methods (src/problems/unconstrain.cpp:92-97): "death penalty", "kuri", "weighted", "ignore_c", "ignore_o" */
typedef enum pgc_unconstrain_method {
    PGC_UNCONSTRAIN_DEATH = 0, PGC_UNCONSTRAIN_KURI = 1, PGC_UNCONSTRAIN_WEIGHTED = 2, PGC_UNCONSTRAIN_IGNORE_C = 3, PGC_UNCONSTRAIN_IGNORE_O = 4
} pgc_unconstrain_method;

/* decompose_objectives methods (src/utils/multi_objective.cpp:603-632) */
typedef enum pgc_decompose_method { PGC_DECOMPOSE_WEIGHTED = 0, PGC_DECOMPOSE_TCHEBYCHEFF = 1, PGC_DECOMPOSE_BI = 2 } pgc_decompose_method;

/* POD description of a UDP.  Mirrors the constructor arguments of the reference UDPs plus, for the CEC
 * suites, the data tables the reference keeps in private members (cec2014.hpp:227-236).
 *   prob_id : cec2014 1..30, cec2013 1..28, zdt 1..6, dtlz 1..7, wfg 1..9; ignored otherwise
 *   dim     : nx (cec/simple), zdt `param`, dtlz/wfg number of decision variables, lennard_jones `atoms`
 *   nobj    : dtlz `fdim`, wfg `dim_obj`; ignored otherwise
 *   param   : dtlz `alpha`, wfg `dim_k`; ignored otherwise
 *   rotation: cec2014 `m_rotation_matrix` (>= k*dim*dim row-major, component i at i*dim*dim, cec2014.cpp:1046)
 *             cec2013 `m_rotation_matrix` = MD[dim] (10*dim*dim)
 *   shift   : cec2014 `m_origin_shift` AFTER the ctor's compaction (component i at i*dim, cec2014.cpp:76-86,1046)
 *             cec2013 `m_origin_shift` flat table (component i at i*dim, cec2013.cpp:878)
 *   shuffle : cec2014 `m_shuffle`, 1-based (component i at i*dim, cec2014.cpp:807-809,1189)
 * The tables are copied to the device at creation; the caller keeps ownership of the host arrays. */
typedef struct pgc_problem_desc {
    int32_t family;
    uint32_t prob_id;
    uint32_t dim;
    uint32_t nobj;
    uint32_t param;
    const double *rotation;
    size_t rotation_len;
    const double *shift;
    size_t shift_len;
    const int32_t *shuffle;
    size_t shuffle_len;
} pgc_problem_desc;

typedef struct pgc_ctx pgc_ctx;
typedef struct pgc_problem pgc_problem;

/* ---- library / context ------------------------------------------------------------------------------ */
PGC_API const char *pgc_version(void);
PGC_API const char *pgc_last_error(void);
PGC_API int pgc_device_count(int *count);
/* One context pins one device and owns one stream plus pinned/device staging buffers.  Distinct contexts may
 * be used concurrently from different threads (thread_safety::basic, reference threading.hpp:42). */
PGC_API int pgc_ctx_create(int device, pgc_ctx **out);
PGC_API int pgc_ctx_destroy(pgc_ctx *ctx);
/* A hint for the launch shapes of latency-bound (island-sized) batches: `n` contexts evaluate concurrently on this context's device
 * (the islands of an archipelago that share one GPU, one context / stream each).  Alone, a small batch is spread thinly over all
 * SMs; with sharers each context keeps fuller tiles and the sharers fill the device together.  Results do not depend on it. */
PGC_API int pgc_ctx_set_sharers(pgc_ctx *ctx, int n);
PGC_API int pgc_ctx_device(const pgc_ctx *ctx, int *device);
PGC_API int pgc_ctx_stream(const pgc_ctx *ctx, void **stream);
PGC_API int pgc_ctx_synchronize(pgc_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py's `gpu_launches`). */
PGC_API int pgc_ctx_launch_count(const pgc_ctx *ctx, uint64_t *count);

/* ---- problems (stand-in for pagmo::problem{UDP}, reference src/problem.cpp:154-242) ------------------ */
PGC_API int pgc_problem_create(pgc_ctx *ctx, const pgc_problem_desc *desc, pgc_problem **out);
PGC_API int pgc_problem_destroy(pgc_problem *prob);
PGC_API int pgc_problem_nx(const pgc_problem *prob, size_t *nx);     /* problem::get_nx  */
PGC_API int pgc_problem_nix(const pgc_problem *prob, size_t *nix);   /* problem::get_nix: integer genes at the end (zdt5) */
PGC_API int pgc_problem_nobj(const pgc_problem *prob, size_t *nobj); /* problem::get_nobj */
PGC_API int pgc_problem_nf(const pgc_problem *prob, size_t *nf);     /* problem::get_nf = nobj + nec + nic: the width of a fitness row */
PGC_API int pgc_problem_nec(const pgc_problem *prob, size_t *nec);   /* problem::get_nec (hock_schittkowski_71: 1, luksan_vlcek1: dim - 2) */
PGC_API int pgc_problem_nic(const pgc_problem *prob, size_t *nic);   /* problem::get_nic (hock_schittkowski_71: 1) */
/* problem::set_c_tol / get_c_tol (src/problem.cpp:620-644): the nec + nic constraint tolerances (default 0) that feasibility and the
 * unconstrain meta-problem use; wrong length, NaN or negative entries fail with the reference's messages.  A wrapper created by
 * pgc_problem_unconstrain copies the inner problem's tolerances at creation, as unconstrain copies its inner problem. */
PGC_API int pgc_problem_set_c_tol(pgc_problem *prob, const double *c_tol, size_t len);
PGC_API int pgc_problem_c_tol(const pgc_problem *prob, double *c_tol);
/* problem::feasibility_f per row (src/problem.cpp:709-721) on device rows [n x nf]: d_feasible[i] = 1 if every constraint of row i is
 * satisfied within the tolerances, else 0. */
PGC_API int pgc_feasibility_device(pgc_problem *prob, const double *d_f, size_t n, uint8_t *d_feasible, void *stream);
PGC_API int pgc_problem_bounds(const pgc_problem *prob, double *lb, double *ub); /* UDP::get_bounds */
PGC_API int pgc_problem_name(const pgc_problem *prob, char *buf, size_t buflen); /* UDP::get_name */
/* cec2013 only.  on != 0: every rotation accumulates one product at a time in the reference's order (rotatefunc,
 * cec2013.cpp:1046-1051) instead of on the FP64 tensor path, so the rotated vectors are bit-identical to the reference's.  Several
 * times slower; meant for checking the functions whose sin / cos / pow of large arguments amplify the last bits of the rotation
 * (f7, f8, f20, f28) against the reference at the 1e-12 tolerance.  Default off. */
PGC_API int pgc_problem_set_strict(pgc_problem *prob, int on);
/* FP64 add/mul/fma(=2) per evaluation and libm calls per evaluation, as tabulated in DESIGN.md
 * (roofline bookkeeping for bench.py). */
PGC_API int pgc_problem_work(const pgc_problem *prob, double *flops_per_eval, double *transcendentals_per_eval,
                     double *bytes_per_eval);

/* Meta-problems over an existing device problem.  `inner` is borrowed and must outlive the wrapper; the
 * wrapper evaluates through every entry point that takes a pgc_problem (pgc_eval_*, pgc_*_evolve_*, ...).
 *   translate (src/problems/translate.cpp:83-87,100-153,175-181): fitness(x) = inner.fitness(x - translation); the bounds
 *     are the inner bounds + translation; len must equal inner nx ("Length of shift vector is: ...").
 *   decompose (src/problems/decompose.cpp:66-124,139-154; decompose_objectives, src/utils/multi_objective.cpp:582-638):
 *     one objective = weighted / tchebycheff / bi decomposition of the inner objectives; the constructor's checks (>= 2
 *     objectives, sizes, finite values, weights >= 0 summing to 1 within 1e-8) fail with PGC_ERR_INVALID_ARGUMENT and the
 *     reference's messages.  adapt_ideal != 0 is refused (PGC_ERR_UNSUPPORTED): the reference adapts z call by call.
 *   unconstrain (src/problems/unconstrain.cpp:66-97,136-223,269-276): the constraints of the inner problem folded into its
 *     objectives by one of the five methods of pgc_unconstrain_method (one objective, the norm of the violation, for ignore_o); the
 *     constructor's checks (inner problem constrained, weights only with - and exactly nec + nic of them for - "weighted") fail with
 *     PGC_ERR_INVALID_ARGUMENT and the reference's messages.  The result has no constraints, so every algorithm entry point takes it. */
PGC_API int pgc_problem_translate(pgc_problem *inner, const double *translation, size_t len, pgc_problem **out);
PGC_API int pgc_problem_unconstrain(pgc_problem *inner, int method /* pgc_unconstrain_method */, const double *weights, size_t len,
                                    pgc_problem **out);
PGC_API int pgc_problem_decompose(pgc_problem *inner, const double *weight, const double *z, size_t len, int method /* pgc_decompose_method */,
                                  int adapt_ideal, pgc_problem **out);

/* ---- batch fitness evaluation ------------------------------------------------------------------------
 * Stand-in for `UDBFE::operator()(const problem&, const vector_double& dvs)` (bfe.hpp:119, bfe.cpp:91-110,
 * thread_bfe.cpp:64-144) and for `UDP::batch_fitness` (problem.cpp:413-427, member_bfe.cpp:40-45). */
/* device-resident: d_dvs [n x nx], d_fvs [n x nf] are device pointers on the context's device; asynchronous
 * on `stream`. */
PGC_API int pgc_eval_device(pgc_problem *prob, const double *d_dvs, size_t n, double *d_fvs, void *stream);
/* host vectors (the pagmo::bfe contract): pageable or pinned host memory; chunked H2D -> kernel -> D2H
 * through the context's pinned staging ring; returns when fvs is complete. */
PGC_API int pgc_eval_host(pgc_problem *prob, const double *dvs, size_t n, double *fvs);

/* ---- multi-objective utilities (reference src/utils/multi_objective.cpp) ---------------------------------------------
 * f is flat row-major [n x m] (a std::vector<vector_double> of objective vectors, flattened).  Index results are the
 * reference's pop_size_t values: size_t in the host entry points, uint32_t on the device.
 *
 * fast_non_dominated_sorting (:200-257): rank[n] = non_dom_rank, dom_count[n] (may be NULL), the fronts concatenated in
 * front_idx[n] - same order inside every front as the reference - with front k = front_idx[front_off[k] .. front_off[k+1]).
 * The O(N^2) dom_list (tuple element 1, only consumed by tests) is not materialised.  Needs n >= 2, m <= 8. */
PGC_API int pgc_fnds_host(pgc_ctx *ctx, const double *f, size_t n, size_t m, size_t *rank, size_t *dom_count,
                          size_t *front_idx, size_t *front_off /* n+1 */, size_t *nfronts);
PGC_API int pgc_fnds_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t m, uint32_t *d_rank, uint32_t *d_dom_count,
                            uint32_t *d_front_idx, uint32_t *d_front_off /* n+1 */, uint32_t *nfronts, void *stream);
/* crowding_distance (:280-315) of ONE front given as its fitness list (n >= 2, m >= 2). */
PGC_API int pgc_crowding_distance_host(pgc_ctx *ctx, const double *f, size_t n, size_t m, double *out);
/* crowding distance of every front of an FNDS result at once, written at the points' own indices.
 * small_front_rule: 0 = none (fronts of size < 2 keep 0), 1 = nsga2.cpp:188-198 (size 1 or 2 -> +inf),
 * 2 = sort_population_mo :441-443 (size 1 -> 0). */
PGC_API int pgc_crowding_fronts_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t m, const uint32_t *d_front_idx,
                                       const uint32_t *d_front_off, uint32_t nfronts, int small_front_rule, double *d_cd,
                                       void *stream);
/* select_best_N_mo (:344-396): whole fronts while they fit, then the cut front by crowding distance descending (stable:
 * equal distances keep front order; the reference's std::sort leaves the order of ties unspecified). */
PGC_API int pgc_select_best_N_mo_host(pgc_ctx *ctx, const double *f, size_t n, size_t m, size_t N, size_t *out, size_t *nout);
PGC_API int pgc_select_best_N_mo_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t m, size_t N, uint32_t *d_out,
                                        uint32_t *nout, void *stream);
/* sort_population_mo (:425-465): indices by (rank ascending, crowding distance descending), stable. */
PGC_API int pgc_sort_population_mo_host(pgc_ctx *ctx, const double *f, size_t n, size_t m, size_t *out);

/* ---- generation operators on a counter-based Philox stream --------------------------------------------------------
 * Every random draw is u01(seed, tag, generation, index, slot) = Philox4x32-10 with counter {slot / 2, index, generation, tag}
 * and key {seed}; u64 = word1:word0 for an even slot, word3:word2 for the odd slot after it; u01 = (u64 >> 11) * 2^-53.
 * Tags: 1/2 = the two NSGA-II shuffles, 3 = NSGA-II variation (index = group of 4), 4 = DE family, 5 = PSO, 6 = sga,
 * 7 = DE self-adaptation init, 8 = cmaes normals, 9 = migration decisions, 10 = population init.  A CPU restatement consuming the same values reproduces the
 * device results (tests/: "parity on injected draws"). */
PGC_API int pgc_philox_u01(uint64_t seed, uint32_t tag, uint32_t generation, uint32_t index, uint32_t slot, double *out);
/* permutation of 0..n-1: stable argsort of the keys u64(seed, tag, generation, i, 0) (stands in for std::shuffle,
 * nsga2.cpp:180-181) */
PGC_API int pgc_philox_permutation_device(pgc_ctx *ctx, size_t n, uint64_t seed, uint32_t tag, uint32_t generation,
                                          uint32_t *d_perm, void *stream);
/* nsga2.cpp:215-239: per group of 4, tournament x2 + SBX + polynomial mutation x2 on each shuffle; children [NP x nx].
 * (genetic_operators.cpp:71-144,148-197,200-211).  Continuous decision variables only. */
PGC_API int pgc_nsga2_variation_device(pgc_ctx *ctx, const double *d_x, const uint32_t *d_rank, const double *d_cd, size_t NP,
                                       size_t nx, const double *d_lb, const double *d_ub, const uint32_t *d_shuffle1,
                                       const uint32_t *d_shuffle2, double cr, double eta_c, double m, double eta_m, uint64_t seed,
                                       uint32_t generation, double *d_children, void *stream);
/* nsga2::evolve (nsga2.cpp:91-307) on a device-resident population: `gens` generations of shuffle, FNDS + crowding,
 * variation, batch evaluation, select_best_N_mo; d_x [NP x nx] and d_f [NP x nobj] are updated in place. */
PGC_API int pgc_nsga2_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t NP, unsigned gens, double cr, double eta_c,
                                    double m, double eta_m, uint64_t seed, uint32_t first_generation, void *stream);

/* pso_gen::evolve (pso_gen.cpp:120-530) on a device-resident swarm: variants 1-5; topologies 1 gbest, 2 lbest ring, 3 von Neumann
 * lattice (:719-744), 4 adaptive random graph with out-degree neighb_param, re-drawn after a generation without a new best (:772-796).
 * In: d_x [n x nx] positions, d_f [n] fitness, d_v velocities or NULL (then drawn as pso_gen.cpp:187-196).
 * Out: d_x / d_f = the particles' best positions / fitness (what evolve() writes back, :524-527), d_v = final velocities,
 * d_xcur (optional) = final current positions.  Reference defaults: omega 0.7298, eta1 = eta2 = 2.05, max_vel 0.5,
 * variant 5, neighb_type 2, neighb_param 4. */
PGC_API int pgc_pso_evolve_device(pgc_problem *prob, double *d_x, double *d_f, double *d_v, double *d_xcur, size_t n, unsigned gens,
                                  double omega, double eta1, double eta2, double max_vel, unsigned variant, unsigned neighb_type,
                                  unsigned neighb_param, uint64_t seed, uint32_t first_generation, void *stream);

/* One generation of ONE SHARD of a pso_gen swarm with the lbest ring topology (pso_gen.cpp:679-698), for swarms spread over several
 * GPUs.  The shard holds n_loc consecutive particles, global indices index_offset .. index_offset + n_loc - 1.  d_X / d_V: current
 * positions / velocities [n_loc x nx]; d_lbX_ext [(n_loc + 2 radius) x nx] and d_lbfit_ext [n_loc + 2 radius]: best positions and
 * fitness with `radius` HALO rows at each end (radius = neighb_param / 2), the shard's own rows in the middle.  The caller fills the
 * halos from the neighbouring shards before every step (pagmo2_b200/swarm.py does it with one all_gather).  init_velocity = 1:
 * only draw the initial velocities (pso_gen.cpp:187-196).  Draws are addressed by the global particle index, so a sharded swarm
 * moves exactly like pgc_pso_evolve_device on one device. */
PGC_API int pgc_pso_shard_step_device(pgc_problem *prob, double *d_X, double *d_V, double *d_lbX_ext, double *d_lbfit_ext, size_t n_loc,
                                      unsigned radius, unsigned index_offset, double omega, double eta1, double eta2, double max_vel,
                                      unsigned variant, uint64_t seed, uint32_t generation, int init_velocity, void *stream);

/* The same for the gbest topology (pso_gen.cpp:644-677: every particle's best neighbour is the swarm's best particle; tracking rule
 * :452-457).  The extended arrays carry ONE row at each end: row 0 = the swarm's current best position / fitness, kept up to date by
 * the caller; rows 1 .. n_loc = the shard's own particles; the last row is unused.  After the step d_cand (device, 2 doubles) holds the
 * shard's candidate for the next best: [fitness, local index] of the particle that improved to the smallest fitness (last index on
 * ties), or [+inf, -1].  The caller reduces the candidates of all shards (smallest fitness, largest GLOBAL index on ties, accepted if
 * <= the current best) and writes the winner's row into row 0 everywhere: pagmo2_b200/swarm.py (GbestSwarm) does it with one
 * all_gather of nx + 2 doubles per shard and generation. */
PGC_API int pgc_pso_shard_step_gbest_device(pgc_problem *prob, double *d_X, double *d_V, double *d_lbX_ext, double *d_lbfit_ext, size_t n_loc,
                                            unsigned index_offset, double omega, double eta1, double eta2, double max_vel, unsigned variant,
                                            uint64_t seed, uint32_t generation, int init_velocity, double *d_cand, void *stream);

/* Differential evolution family as a generational device loop (trial vectors for all individuals -> one batch evaluation ->
 * selection): algo 0 = de (de.cpp:76-345; variant 1..10, F, CR), 1 = sade (sade.cpp:78-560; variant 1..18, variant_adptv 1 = jDE,
 * 2 = iDE), 2 = de1220 (de1220.cpp:80-600; allowed_variants, variant_adptv).  d_x [NP x nx], d_f [NP] updated in place; d_F / d_CR /
 * d_variant = per-individual self-adaptation memory (NULL: initialised as the reference does without memory).  Stops early on
 * the reference's xtol / ftol exits; *gens_done = generations run.  Reference defaults: de(F 0.8, CR 0.9, variant 2), sade(variant 2,
 * adptv 1), de1220(allowed {2,3,7,10,13,14,15,16}, adptv 1), ftol = xtol = 1e-6. */
PGC_API int pgc_de_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t NP, unsigned gens, unsigned algo, unsigned variant,
                                 unsigned variant_adptv, double F, double CR, const uint32_t *allowed_variants, unsigned n_allowed,
                                 double ftol, double xtol, double *d_F, double *d_CR, uint32_t *d_variant, uint64_t seed,
                                 uint32_t first_generation, unsigned *gens_done, void *stream);

/* ---- hypervolume (src/utils/hypervolume.cpp:196-330; hv2d hv_hv2d.cpp:59-148, hv3d / HyCon3D hv_hv3d.cpp:107-343) ----------------
 * points [n x m] row-major, m = 2 or 3 (the dimensions the reference serves with hv2d / hv3d), r_point[m] on the host.
 * compute: *hv = hypervolume(points).compute(r_point); contributions: out[n] = exclusive contribution of every point.
 * An invalid reference point (hv_algorithm.cpp:226-258) gives PGC_ERR_INVALID_ARGUMENT. */
PGC_API int pgc_hv_compute_host(pgc_ctx *ctx, const double *points, size_t n, size_t m, const double *r_point, double *hv);
PGC_API int pgc_hv_contributions_host(pgc_ctx *ctx, const double *points, size_t n, size_t m, const double *r_point, double *out);
/* device-resident points; d_out: n doubles (compute writes the hypervolume to d_out[0]) */
PGC_API int pgc_hv_device(pgc_ctx *ctx, const double *d_points, size_t n, size_t m, const double *r_point, int compute, double *d_out,
                          void *stream);
/* bf_fpras::compute (hv_bf_fpras.cpp:91-146): (eps, delta) approximation of the hypervolume by the Karp-Luby estimator; the trial
 * budget T = 12 log2(1 / delta) n / eps^2 is split over device threads that run whole rounds on their own Philox substreams (the same
 * estimator T V / (n M), not the same number as the reference's mt19937 run).  Reference defaults: eps 1e-2, delta 1e-2. */
PGC_API int pgc_hv_fpras_host(pgc_ctx *ctx, const double *points, size_t n, size_t m, const double *r_point, double eps, double delta,
                              uint64_t seed, double *hv);
/* bf_approx::least_contributor / greatest_contributor (hv_bf_approx.cpp:98-112, :337-470): the Bringmann-Friedrich approximation of
 * the extreme contributor - rounds of Monte-Carlo sampling inside every point's bounding box (all outstanding samples of a round in
 * one launch, one CTA per point) with the reference's elimination and stopping rules, and its switch to the exact exclusive volume
 * for small / expensive boxes (use_exact).  Reference defaults: use_exact 1, trivial_subcase_size 1, eps 1e-2, delta 1e-6,
 * delta_multiplier 0.775, alpha 0.2, initial_delta_coeff 0.1, gamma 0.25. */
PGC_API int pgc_hv_approx_extreme_host(pgc_ctx *ctx, const double *points, size_t n, size_t m, const double *r_point, int greatest,
                                       int use_exact, unsigned trivial_subcase_size, double eps, double delta, double delta_multiplier,
                                       double alpha, double initial_delta_coeff, double gamma, uint64_t seed, size_t *idx);

/* ---- dense contractions of CMA-ES / xNES (cmaes.cpp:246-253,362-380; xnes.cpp:302-305) ------------------------------------------
 * sampling: x_i = mean + sigma * BD * z_i, i < lambda, z ~ N(0, I) from Philox (seed, tag 8, generation, i, .); BD = B*D row-major
 * [D x D] (the caller's eigendecomposition, cmaes.cpp:386-401, stays on the host); d_z optional (xnes keeps z). */
PGC_API int pgc_cmaes_sample_device(pgc_ctx *ctx, const double *d_mean, const double *d_bd, double sigma, size_t lambda, size_t D,
                                    uint64_t seed, uint32_t generation, double *d_z, double *d_x, void *stream);
/* out[D x D] = sum_{i<k} w_i (r_i - c)(r_i - c)^T / scale_div, r_i = rows[idx ? idx[i] : i] (rank-mu: rows = population, idx =
 * the mu best, c = old mean, scale_div = sigma^2; xnes: rows = z, idx = s_idx, c = NULL, w = u, scale_div = 1). */
PGC_API int pgc_weighted_gram_device(pgc_ctx *ctx, const double *d_rows, const uint32_t *d_idx, const double *d_center, const double *d_w,
                                     size_t k, size_t D, double scale_div, double *d_out, void *stream);
/* out[D] = sum_{i<k} w_i r_i in the reference's order (cmaes.cpp:363-366) */
PGC_API int pgc_weighted_mean_device(pgc_ctx *ctx, const double *d_rows, const uint32_t *d_idx, const double *d_w, size_t k, size_t D,
                                     double *d_out, void *stream);

/* cmaes::evolve (cmaes.cpp:111-407, memory = false) on a device-resident population of lambda individuals: normal draws, sampling,
 * batch evaluation, recombination and the rank-mu matrix on the device; evolution paths, the combination of C and its
 * eigendecomposition (a cyclic Jacobi solver where the reference calls Eigen's SelfAdjointEigenSolver) on the host.  cc, cs, c1,
 * cmu = -1: the automatic values of :169-183.  d_x / d_f end as the LAST generation sampled (the reference replaces the
 * population every generation); *sigma_out (optional) = final step size.  Reference defaults: sigma0 0.5, ftol = xtol = 1e-6. */
/* nspso::evolve (src/algorithms/nspso.cpp:84-411) on a device-resident swarm: d_x [n x nx], d_f [n x nobj] in place (the moved
 * swarm of the last generation, as the reference returns it).  diversity: 0 "crowding distance", 1 "niche count", 2 "max min".
 * d_vel [n x nx], d_best_x [n x nx], d_best_f [n x nobj]: the algorithm's memory (velocities and the archive m_best_dvs / m_best_fit),
 * read and updated when given (memory = true), NULL = memory-less start (velocities drawn, archive = the population, :127-152). */
PGC_API int pgc_nspso_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t n, unsigned gens, double omega, double c1, double c2,
                                    double chi, double v_coeff, unsigned leader_selection_range, unsigned diversity, uint64_t seed,
                                    uint32_t first_generation, double *d_vel, double *d_best_x, double *d_best_f, void *stream);
/* moead_gen::evolve (src/algorithms/moead_gen.cpp:128-345), the reference's generational MOEA/D, on a device-resident population
 * d_x [n x nx], d_f [n x nobj] (in place).  weights [n x nobj] and neigh [n x T] (HOST arrays): the weight vectors of the n
 * sub-problems (pagmo::decomposition_weights) and the indices of each one's T nearest weight vectors (pagmo::kNN) - utilities
 * outside evolve() that the caller computes.  decomposition: 0 "weighted", 1 "tchebycheff", 2 "bi". */
PGC_API int pgc_moead_gen_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t n, unsigned gens, const double *weights,
                                        const uint32_t *neigh, unsigned T, int decomposition, double CR, double F, double eta_m, double realb,
                                        unsigned limit, int preserve_diversity, uint64_t seed, uint32_t first_generation, void *stream);
/* gaco::evolve (src/algorithms/gaco.cpp:104-445), extended ant colony optimisation, on a device-resident population of an
 * unconstrained single-objective problem (in place); the problem's last nix variables are sampled as integers.
 * Constructor arguments as gaco.hpp:104-107 (reference defaults: ker 63, q 1.0, oracle 0, acc 0.01, threshold 1, n_gen_mark 7,
 * impstop 100000, evalstop 100000, focus 0).  The reference keeps m_oracle, m_q and its stopping counters in the algorithm object
 * between evolve() calls: they travel in pgc_gaco_state (initialized == 0: taken from q / oracle, counters at 1; updated on return).
 * *gens_done (optional): generations run before a stopping criterion returned. */
typedef struct pgc_gaco_state {
    double oracle, q;
    uint32_t n_evalstop, n_impstop, gen_mark, initialized;
    uint64_t fevals;
    /* memory = true (gaco.cpp:106-108,223-250,732-752,778-784): the solution archive and a call counter survive between evolve() calls,
     * the kernel weights follow the counter, and the archive is NOT written back into the population.  h_archive: HOST array of
     * ker * (nx + 2) doubles owned by the caller. */
    uint32_t memory, counter;
    double *h_archive;
    size_t h_archive_len;
    /* the champion of the POPULATION (population.cpp:209-246, the best individual it ever held) is what the evalstop counter watches
     * (gaco.cpp:338-347); has_champion == 0: the best of the population handed in.  Updated on return. */
    uint32_t has_champion, reserved_;
    double champion_f;
} pgc_gaco_state;
PGC_API int pgc_gaco_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t n, unsigned gens, unsigned ker, double q, double oracle,
                                   double acc, unsigned threshold, unsigned n_gen_mark, unsigned impstop, unsigned evalstop, double focus,
                                   uint64_t seed, uint32_t first_generation, pgc_gaco_state *state, unsigned *gens_done, void *stream);
/* maco::evolve (src/algorithms/maco.cpp:88-533), multi-objective hypervolume-based ant colony optimisation, on a device-resident
 * population (in place), memory = false.  Constructor arguments as maco.hpp:107-109 (reference defaults: ker 63, q 1.0, threshold 1,
 * n_gen_mark 7, evalstop 100000, focus 0).  m_q, m_n_evalstop and m_gen_mark of the algorithm object travel in pgc_maco_state
 * (initialized == 0: q from the argument, counters as a fresh maco has them). */
typedef struct pgc_maco_state {
    double q;
    uint32_t n_evalstop, gen_mark, initialized, reserved_;
} pgc_maco_state;
PGC_API int pgc_maco_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t n, unsigned gens, unsigned ker, double q,
                                   unsigned threshold, unsigned n_gen_mark, unsigned evalstop, double focus, uint64_t seed,
                                   uint32_t first_generation, pgc_maco_state *state, unsigned *gens_done, void *stream);
PGC_API int pgc_cmaes_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t lambda, unsigned gens, double cc, double cs, double c1,
                                    double cmu, double sigma0, double ftol, double xtol, int force_bounds, uint64_t seed,
                                    uint32_t first_generation, unsigned *gens_done, double *sigma_out, void *stream);

/* xnes::evolve (src/algorithms/xnes.cpp:96-303), exponential natural evolution strategies, on a device-resident population (in place),
 * memory = false: sampling, evaluation and the natural-gradient contractions on the device, the D x D updates on the host (exp of the
 * symmetric d_A through a Jacobi eigendecomposition where the reference uses Eigen's matrix exponential).  Constructor arguments as
 * xnes.hpp:107-108 (-1: automatic).  *sigma_out (optional): the step size the run ends with. */
PGC_API int pgc_xnes_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t lambda, unsigned gens, double eta_mu, double eta_sigma,
                                   double eta_b, double sigma0, double ftol, double xtol, int force_bounds, uint64_t seed,
                                   uint32_t first_generation, unsigned *gens_done, double *sigma_out, void *stream);

/* sga::evolve (sga.cpp:184-292) on a device-resident single-objective population, in place (the population comes back sorted by
 * fitness, as the reference's reinsertion leaves it).  crossover: 0 exponential, 1 binomial, 2 single, 3 sbx; mutation: 0 gaussian,
 * 1 uniform, 2 polynomial; selection: 0 tournament (param_s <= 16), 1 truncated.  Reference defaults: cr 0.9, eta_c 1, m 0.02,
 * param_m 1, param_s 2, exponential, polynomial, tournament. */
PGC_API int pgc_sga_evolve_device(pgc_problem *prob, double *d_x, double *d_f, size_t NP, unsigned gens, double cr, double eta_c, double m,
                                  double param_m, unsigned param_s, unsigned crossover, unsigned mutation, unsigned selection,
                                  uint64_t seed, uint32_t first_generation, void *stream);

/* ---- algorithms behind one descriptor (pagmo::algorithm::evolve(pop), src/algorithm.cpp) ------------------------------- */
typedef enum pgc_algo {
    PGC_ALGO_DE = 1,      /* src/algorithms/de.cpp:76-345 */
    PGC_ALGO_SADE = 2,    /* src/algorithms/sade.cpp:78-560 */
    PGC_ALGO_DE1220 = 3,  /* src/algorithms/de1220.cpp:80-600 */
    PGC_ALGO_PSO_GEN = 4, /* src/algorithms/pso_gen.cpp:120-590 */
    PGC_ALGO_NSGA2 = 5,   /* src/algorithms/nsga2.cpp:91-307 */
    PGC_ALGO_SGA = 6,     /* src/algorithms/sga.cpp:184-292 */
    PGC_ALGO_CMAES = 7,   /* src/algorithms/cmaes.cpp:111-407 */
    PGC_ALGO_NSPSO = 8,   /* src/algorithms/nspso.cpp:84-411 */
    PGC_ALGO_XNES = 9     /* src/algorithms/xnes.cpp:96-303: eta_mu, eta_sigma, eta_b travel in cma_cc, cma_cs, cma_c1 (-1: automatic) */
} pgc_algo;

/* Constructor arguments of the reference UDAs; pgc_algo_defaults() fills in the reference's default values
 * (de.hpp:119, sade.hpp:138, de1220.hpp:158, pso_gen.hpp:127, nsga2.hpp:103, sga.hpp:166). */
typedef struct pgc_algo_desc {
    int32_t algo;
    uint32_t gens;
    uint32_t variant, variant_adptv;      /* de family; pso_gen variant */
    uint32_t neighb_type, neighb_param;   /* pso_gen */
    uint32_t n_allowed;
    uint32_t allowed_variants[18];        /* de1220 */
    double F, CR, ftol, xtol;             /* de family */
    double omega, eta1, eta2, max_vel;    /* pso_gen */
    double cr, eta_c, m, eta_m;           /* nsga2; sga uses cr, eta_c, m */
    uint64_t seed;
    double param_m;                       /* sga */
    uint32_t param_s, crossover, mutation, selection; /* sga: see pgc_sga_evolve_device */
    double cma_cc, cma_cs, cma_c1, cma_cmu, sigma0;   /* cmaes (-1: automatic), cmaes.hpp:110 */
    uint32_t force_bounds;
    uint32_t memory;                      /* sade / de1220 / pso_gen / nspso / cmaes / xnes: keep the state between evolve() calls.  Honoured where
                                           * somebody owns that state: pgc_island_evolve (resident in the island) and the C++ adapters; through
                                           * pgc_algo_evolve_device a call is memory-less, pgc_algo_evolve_memory_device takes the state explicitly */
    double nspso_c1, nspso_c2, nspso_chi, nspso_v_coeff;       /* nspso (omega is shared with pso_gen), nspso.hpp:59-62 */
    uint32_t leader_selection_range, diversity;               /* nspso: diversity 0 crowding distance, 1 niche count, 2 max min */
} pgc_algo_desc;

PGC_API int pgc_algo_defaults(int algo, unsigned gens, uint64_t seed, pgc_algo_desc *out);
/* algorithm::evolve on a device-resident population (d_x [n x nx], d_f [n x nobj], in place).  `first_generation` offsets the
 * Philox generation counter so that successive calls continue the random stream; *gens_done (optional) = generations run. */
PGC_API int pgc_algo_evolve_device(pgc_problem *prob, const pgc_algo_desc *algo, double *d_x, double *d_f, size_t n,
                                   uint32_t first_generation, unsigned *gens_done, void *stream);

/* The state an algorithm constructed with memory = true keeps between evolve() calls (device arrays owned by the caller):
 *   sade / de1220: a = F [n], b = CR [n], u = mutation variant [n] (de1220 only)      (sade.cpp:137-156, de1220.cpp:147-165)
 *   pso_gen      : a = velocities [n x nx]                                            (pso_gen.cpp:193-201)
 *   nspso        : a = velocities [n x nx], b = archive decision vectors [n x nx], c = archive fitness [n x nobj] (nspso.cpp:127-152)
 *   cmaes / xnes : h_state = HOST array of pgc_es_state_len() doubles: sigma, mean, the evolution paths and B / D / C / C^-1/2 of cmaes
 *                  (cmaes.cpp:201-228), sigma, mean and A of xnes (xnes.cpp:163-175); kept while the dimension (and, for cmaes, the
 *                  population size) stays the same, as in the reference
 * initialized == 0: the arrays hold nothing yet; pgc_algo_evolve_memory_device first fills them as the reference's first evolve() with
 * memory does (drawn from the Philox streams of `first_generation`) and sets it to 1.  Other algorithms have no such state.
 * pgc_island_evolve keeps this state in HBM inside the island when algo->memory != 0 (re-drawn if the algorithm changes, as a new
 * reference UDA object would). */
typedef struct pgc_algo_memory {
    double *a, *b, *c;
    uint32_t *u;
    int32_t initialized, reserved_;
    double *h_state;      /* cmaes / xnes only (host) */
    size_t h_state_len;
} pgc_algo_memory;
PGC_API int pgc_es_state_len(int algo, size_t nx, size_t *len);
PGC_API int pgc_algo_evolve_memory_device(pgc_problem *prob, const pgc_algo_desc *algo, double *d_x, double *d_f, size_t n,
                                          uint32_t first_generation, unsigned *gens_done, pgc_algo_memory *memory, void *stream);

/* algorithm::set_verbosity(v) + get_log() (de.hpp:160-199 and the like): evolve() that also records the log lines the reference UDA
 * records - one every `verbosity` generations (generations 1, 1 + v, 1 + 2v, ...; every generation for v = 1), computed on the
 * device from the resident population.  Row layouts (doubles; pgc_algo_log_row_len gives the length):
 *   de      gen, fevals, best, dx, df                        (de.cpp:324-347)
 *   sade    gen, fevals, best, F, CR, dx, df                 (sade.cpp:556-580)
 *   de1220  gen, fevals, best, F, CR, variant, dx, df        (de1220.cpp:570-595)
 *   pso_gen gen, fevals, gbest, mean velocity, mean lbest, average distance   (pso_gen.cpp:464-518)
 *   sga     gen, fevals, best, improvement                   (sga.cpp:252-274; verbosity 1 logs only the generations that improve)
 *   cmaes   gen, fevals, best, dx, df, sigma                 (cmaes.cpp:276-296); xnes the same (xnes.cpp:238-256)
 *   nsga2   gen, fevals, ideal point [nobj]                  (nsga2.cpp:144-173; logged BEFORE the generation, as the reference does)
 *   nspso   gen, fevals, ideal point of the archive [nobj]   (nspso.cpp:163-192; fevals counted from this call's start)
 * log_rows: HOST array [max_rows x row_len]; *n_rows = lines written (the DE family stops logging at the generation whose exit test
 * fires).  memory: NULL, or the state of an algorithm built with memory = true (pgc_algo_evolve_memory_device). */
PGC_API int pgc_algo_log_row_len(const pgc_problem *prob, int algo, size_t *row_len);
PGC_API int pgc_algo_evolve_logged_device(pgc_problem *prob, const pgc_algo_desc *algo, double *d_x, double *d_f, size_t n,
                                          uint32_t first_generation, unsigned *gens_done, pgc_algo_memory *memory, unsigned verbosity,
                                          double *log_rows, size_t max_rows, size_t *n_rows, void *stream);

/* The same log capture for the algorithms that have their own entry point instead of a pgc_algo_desc: between pgc_log_capture_begin
 * and pgc_log_capture_end (same thread) a pgc_gaco_ / pgc_maco_ / pgc_moead_gen_evolve_device call records the reference's lines:
 *   gaco      gen, fevals, best, kernel, oracle, dx, dp        (gaco.cpp:254-287 inside the loop - not for the last generation -
 *                                                               and :405-445 after it; fevals as the reference counts them, before
 *                                                               the generation's evaluations)
 *   maco      gen, fevals, ideal point of the archive [nobj]    (maco.cpp:415-463)
 *   moead_gen gen, fevals, ADF, ideal point [nobj]              (moead_gen.cpp:180-211; ADF = sum of the decomposed fitnesses)
 * row_len must be 7, 2 + nobj, 3 + nobj.  pgc_log_capture_end copies at most max_rows rows to rows_out (HOST) and ends the capture. */
PGC_API int pgc_log_capture_begin(pgc_ctx *ctx, unsigned verbosity, size_t max_rows, size_t row_len);
PGC_API int pgc_log_capture_end(pgc_ctx *ctx, double *rows_out, size_t *n_rows);

/* ---- populations and migration (island.cpp:428-652) ------------------------------------------------------------------- */
/* population(prob, bfe, n, seed) (population.cpp:82-103, generic.hpp:326-389): n uniform random decision vectors in the bounds,
 * one batch evaluation, random 64-bit IDs; the last nix genes are drawn as integers in [lb, ub] (generic.hpp:289-295).  d_f and
 * d_ids may be NULL.  The generation operators (pgc_*_evolve_device) refuse problems with integer genes (PGC_ERR_UNSUPPORTED), except pgc_nsga2_evolve_device,
 * which applies the reference's integer operators to the last nix genes. */
PGC_API int pgc_population_init_device(pgc_problem *prob, size_t n, uint64_t seed, double *d_x, double *d_f, uint64_t *d_ids,
                                       void *stream);
/* select_best::select (select_best.cpp:63-171): the best `rate` individuals (absolute count, or a fraction of n when
 * rate_is_frac) of the group (ids, x, f); outputs sized for n rows; *n_out = rows written.  Unconstrained problems. */
PGC_API int pgc_select_best_device(pgc_ctx *ctx, const uint64_t *d_ids, const double *d_x, const double *d_f, size_t n, size_t nx,
                                   size_t nobj, int rate_is_frac, double rate, uint64_t *d_ids_out, double *d_x_out, double *d_f_out,
                                   size_t *n_out, void *stream);
/* fair_replace::replace (fair_replace.cpp:63-221): merge the best min(rate, nm) migrants into the group and keep its best n, IN
 * PLACE and in sorted order, as the reference returns it.  Unconstrained problems. */
PGC_API int pgc_fair_replace_device(pgc_ctx *ctx, uint64_t *d_ids, double *d_x, double *d_f, size_t n, size_t nx, size_t nobj,
                                    int rate_is_frac, double rate, const uint64_t *d_mids, const double *d_mx, const double *d_mf,
                                    size_t nm, void *stream);
/* The single-objective CONSTRAINED branches of the two policies (select_best.cpp:137-152, fair_replace.cpp:158-188) and the order they
 * use, sort_population_con (constrained.cpp:180-202): rows of f are [objective | nec equality | nic inequality constraints], tol =
 * HOST array of the nec + nic tolerances (problem::get_c_tol).  Order: more satisfied constraints first; among the feasible the smaller
 * objective; among equally infeasible ones the smaller Euclidean norm of the violations (compare_fc measures its left argument with the
 * sum of the equality and inequality norms instead, constrained.cpp:96-106: the same order whenever an individual violates
 * constraints of one kind only); ties keep the input order.  d_order: device, n entries. */
PGC_API int pgc_sort_population_con_device(pgc_ctx *ctx, const double *d_f, size_t n, size_t nec, size_t nic, const double *tol,
                                           uint32_t *d_order, void *stream);
PGC_API int pgc_select_best_con_device(pgc_ctx *ctx, const uint64_t *d_ids, const double *d_x, const double *d_f, size_t n, size_t nx,
                                       size_t nec, size_t nic, const double *tol, int rate_is_frac, double rate, uint64_t *d_ids_out,
                                       double *d_x_out, double *d_f_out, size_t *n_out, void *stream);
PGC_API int pgc_fair_replace_con_device(pgc_ctx *ctx, uint64_t *d_ids, double *d_x, double *d_f, size_t n, size_t nx, size_t nec, size_t nic,
                                        const double *tol, int rate_is_frac, double rate, const uint64_t *d_mids, const double *d_mx,
                                        const double *d_mf, size_t nm, void *stream);
/* topology::get_connections(i) for kind 0 = unconnected, 1 = ring (ring.cpp:74-116), 2 = fully_connected
 * (fully_connected.cpp:86-115) with n vertices: sources of the edges into i and their weights; buffers sized n. */
PGC_API int pgc_topology_connections(int kind, size_t n, size_t i, double weight, size_t *idx_out, double *w_out, size_t *count);

/* ---- device-resident islands (island.cpp:428-652, thread_island.cpp:79-132) -----------------------------------------------
 * A pgc_island keeps one island's population (ids | x | f) in HBM between calls, so that successive evolve() calls, selection
 * and replacement never cross PCIe.  It also owns a migration OUTBOX (what select_best published last: the island's entry in the
 * archipelago's migrants database, archipelago.cpp:658-714) and `max_in_edges` INBOX slots (what arrived along in-edges), each a
 * packed group of at most `max_migrants` individuals.  Calls on one island are ordered on its context's stream.  `prob` is
 * borrowed and must outlive the island. */
typedef struct pgc_island pgc_island;
PGC_API int pgc_island_create(pgc_problem *prob, size_t n, size_t max_migrants, size_t max_in_edges, pgc_island **out);
PGC_API int pgc_island_destroy(pgc_island *isl);
PGC_API int pgc_island_size(const pgc_island *isl, size_t *n, size_t *nx, size_t *nf);
PGC_API int pgc_island_pointers(pgc_island *isl, uint64_t **d_ids, double **d_x, double **d_f);
/* island::set_population / get_population: host <-> device (the only PCIe traffic of an island) */
PGC_API int pgc_island_upload(pgc_island *isl, const uint64_t *ids, const double *x, const double *f);
PGC_API int pgc_island_download(pgc_island *isl, uint64_t *ids, double *x, double *f);
/* population(prob, bfe, n, seed) directly on the device (pgc_population_init_device) */
PGC_API int pgc_island_init(pgc_island *isl, uint64_t seed);
/* algorithm::evolve on the resident population; the island keeps the Philox generation counter (successive calls continue the
 * random stream; pgc_island_generation / pgc_island_set_generation read and set it) */
PGC_API int pgc_island_evolve(pgc_island *isl, const pgc_algo_desc *algo, unsigned *gens_done);
PGC_API int pgc_island_generation(const pgc_island *isl, uint32_t *generation);
PGC_API int pgc_island_set_generation(pgc_island *isl, uint32_t generation);
/* s_policy: select_best::select (select_best.cpp:63-171) into the outbox; *k = individuals published */
PGC_API int pgc_island_select(pgc_island *isl, int rate_is_frac, double rate, size_t *k);
PGC_API int pgc_island_clear_outbox(pgc_island *isl); /* archipelago::extract_migrants (migrant_handling::evict) */
PGC_API int pgc_island_outbox_download(pgc_island *isl, uint64_t *ids, double *x, double *f, size_t *k);
/* host migrants into inbox slot `slot` (a pagmo archipelago's host-side migrants database feeding a GPU island) */
PGC_API int pgc_island_inbox_upload(pgc_island *isl, size_t slot, const uint64_t *ids, const double *x, const double *f, size_t k);
/* r_policy: fair_replace::replace (fair_replace.cpp:63-221) of the population with the rows of inbox slots [0, n_slots).
 * accepted_ids / accepted_slot (optional, max_migrants * max_in_edges entries): the immigrants that are in the population
 * afterwards and the slot each came through - the rows of archipelago::get_migration_log() (island.cpp:525-536). */
PGC_API int pgc_island_replace(pgc_island *isl, int rate_is_frac, double rate, size_t n_slots, uint64_t *accepted_ids,
                               uint32_t *accepted_slot, size_t *n_accepted);
/* The same in two halves for drivers that must not block between two evolve() calls.  enqueue: asynchronous on the island's stream;
 * `counts` = rows in each of the n_slots inbox slots (known to a driver that replays the senders' policies), want_log != 0 also copies
 * the acceptance flags to pinned memory behind an event.  collect: waits for that event only and returns the migration-log rows of the
 * last enqueue (nothing when none is pending); it must be called before the next enqueue with want_log. */
PGC_API int pgc_island_replace_enqueue(pgc_island *isl, int rate_is_frac, double rate, size_t n_slots, const size_t *counts, int want_log);
PGC_API int pgc_island_replace_collect(pgc_island *isl, uint64_t *accepted_ids, uint32_t *accepted_slot, size_t *n_accepted);
/* population::champion_x / champion_f (single objective) */
PGC_API int pgc_island_champion(pgc_island *isl, double *x, double *f);

/* ---- migration over NCCL (the exchange step of SURVEY.md 8e: migrants travel device to device along topology edges) --------
 * NCCL is bound at run time (libnccl.so.2, or $PGC_NCCL_LIBRARY); without it these return PGC_ERR_UNSUPPORTED.
 *   pgc_comm_init        one process driving several GPUs: rank g = devices[g]                      (ncclCommInitAll)
 *   pgc_comm_unique_id / pgc_comm_init_rank   one process per GPU: rank 0 creates the 128-byte id, the host application
 *                        distributes it (MPI, torch.distributed, a file), every rank joins            (ncclCommInitRank) */
typedef struct pgc_comm pgc_comm;
PGC_API int pgc_comm_nccl_version(int *version);
PGC_API int pgc_comm_init(int ndev, const int *devices, pgc_comm **out);
PGC_API int pgc_comm_unique_id(void *id, size_t len /* >= 128 */);
PGC_API int pgc_comm_init_rank(int device, int nranks, int rank, const void *id, size_t len, pgc_comm **out);
PGC_API int pgc_comm_destroy(pgc_comm *comm);
PGC_API int pgc_comm_size(const pgc_comm *comm, int *nranks, int *nlocal);
/* One migration step (island.cpp:461-620 decides the edges, this moves the rows): for every edge e the OUTBOX of island
 * edge_src[e] is delivered into inbox slot edge_slot[e] of island edge_dst[e] - ncclSend / ncclRecv in one group between GPUs, a
 * device-to-device copy inside one GPU.  islands[i] is NULL for an island owned by another process, owner_rank[i] is the
 * communicator rank of the GPU holding island i; every process passes the same edge list in the same order.  Asynchronous on
 * the islands' streams (after the source's select, before the destination's replace).  comm may be NULL when no edge leaves a GPU. */
PGC_API int pgc_migrate(pgc_comm *comm, pgc_island *const *islands, const int *owner_rank, size_t n_islands, const uint32_t *edge_src,
                        const uint32_t *edge_dst, const uint32_t *edge_slot, size_t n_edges);

/* Debug/profiling aid for the CEC2014 stage kernel: same evaluation with clock64() phase counters, summed over all
 * warp-tiles and stages.  out7 = {load, weight pass, token wait, GEMM, z store, epilogue} cycles, warp-tiles. */
PGC_API int pgc_debug_cec2014_phase_cycles(pgc_problem *prob, const double *d_dvs, size_t n, double *d_fvs, uint64_t *out7);

/* Debug / measurement aid for the tensor-core rotation (rot_i8.cu): z = M * ((x - os) * rate) for n device-resident rows of D (even,
 * <= 128) doubles through tcgen05.mma kind::i8 on exact digit planes (Ozaki scheme), FP64-accurate.  coef == NULL: d_out = z [n x D];
 * otherwise d_out = f [n] = sum_j coef[j] * z_j^2 + fbias.  M (D x D row-major), os, coef are host arrays.  reps > 1: the kernel is
 * launched reps + 1 times and *ms_per_launch (optional) receives the CUDA-event time of one launch. */
PGC_API int pgc_debug_rot_i8(pgc_ctx *ctx, const double *M, size_t D, const double *os, const double *coef, double rate, double fbias,
                             const double *d_x, size_t n, double *d_out, int reps, float *ms_per_launch, void *stream);

/* Design probe: cycles per tcgen05.mma kind::i8 (M = 128, K = 32) for N = 32, 64, 128, 256 and four issue patterns (one accumulator;
 * seven accumulators in turn; the same with collector::a reuse; fresh operand tiles); out32 = [4][4][2] doubles (issue, completion). */
PGC_API int pgc_debug_mma_i8_probe(pgc_ctx *ctx, double *out32);

/* ---- device memory helpers (so a host-language binding needs no CUDA runtime of its own) ------------- */
PGC_API int pgc_malloc_device(pgc_ctx *ctx, size_t bytes, void **out);
PGC_API int pgc_free_device(pgc_ctx *ctx, void *ptr);
PGC_API int pgc_malloc_pinned(pgc_ctx *ctx, size_t bytes, void **out);
PGC_API int pgc_free_pinned(pgc_ctx *ctx, void *ptr);
PGC_API int pgc_memcpy_h2d(pgc_ctx *ctx, void *dst, const void *src, size_t bytes);
PGC_API int pgc_memcpy_d2h(pgc_ctx *ctx, void *dst, const void *src, size_t bytes);

/* ---- measurement helpers ------------------------------------------------------------------------------
 * Dependent-chain-free DFMA loop over the whole chip: the measured FP64-pipe ceiling that bench.py uses as
 * the roofline denominator for the rotation-bound kernels (MEASURED_PEAKS.json has no FP64 figure). */
PGC_API int pgc_measure_fp64_peak(pgc_ctx *ctx, int iters, double *tflops);
/* The same probe through mma.sync m8n8k4 f64 (DMMA), to decide SIMT-vs-DMMA with numbers. */
PGC_API int pgc_measure_fp64_mma_peak(pgc_ctx *ctx, int iters, double *tflops);

/* Design probe: one CTA of `total_warps` warps per SM; the first `dmma_warps` run 26 independent DMMA chains each
 * (the rotation kernel's accumulator count), the others 8 independent DFMA chains.  tflops2[0] / [1] = the two groups'
 * throughput in the same launch - answers "can one warp per sub-partition saturate DMMA" and "do DMMA and DFMA share
 * a pipe". */
PGC_API int pgc_debug_fp64_mix_probe(pgc_ctx *ctx, int iters, int total_warps, int dmma_warps, double *tflops2);

#ifdef __cplusplus
}
#endif
#endif /* PAGMO_CUDA_PGC_H */
