/* oracle/oracle.h - C API of the CPU restatement (liboracle.so).  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's algorithms on the hot path, one function per reference routine,
 * each citing the reference file:line it follows.  It exists to CHECK the CUDA path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  It is itself pinned
 * bit-exactly against the unmodified reference (oracle/_ref/libpagmo_ref.so) and against the golden vectors in
 * tests/golden/ (see tests/test_oracle_*.py).
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_RASTRIGIN = 1, ORACLE_ACKLEY = 2, ORACLE_GRIEWANK = 3, ORACLE_SCHWEFEL = 4, ORACLE_ROSENBROCK = 5 };

/* ---- simple UDPs (restate_simple.c) ---- */
int oracle_simple_fitness(int family, size_t dim, const double *x, double *f);
int oracle_simple_batch(int family, size_t dim, const double *xs, size_t n, double *fs);
int oracle_simple_bounds(int family, double *lo, double *hi);

/* ---- CEC2014 (restate_cec2014.c) ----
 * Mr: rotation table (component i at i*dim*dim), Os: COMPACTED shift (component i at i*dim, i.e. what
 * cec2014::get_origin_shift() returns, cec2014.cpp:76-86), S: 1-based shuffle (component i at i*dim). */
int oracle_cec2014_fitness(unsigned func, unsigned dim, const double *Mr, const double *Os, const int *S,
                           const double *x, double *f);
/* nthreads<=1: sequential.  Static contiguous partition over individuals (thread_bfe.cpp:94-137 restated). */
int oracle_cec2014_batch(unsigned func, unsigned dim, const double *Mr, const double *Os, const int *S,
                         const double *xs, size_t n, double *fs, int nthreads);
/* ---- CEC2013 (restate_cec2013.c): Mr = MD[dim] (10 matrices), Os = shift_data (10 lines of 100, addressed at i*dim) ---- */
int oracle_cec2013_fitness(unsigned func, unsigned dim, const double *Mr, const double *Os, const double *x, double *f);
int oracle_cec2013_batch(unsigned func, unsigned dim, const double *Mr, const double *Os, const double *xs, size_t n, double *fs);
/* compaction performed by the cec2014 constructor (cec2014.cpp:76-86): keep the first dim of every 100 */
size_t oracle_cec2014_compact_shift(const double *lines, size_t nlines, unsigned dim, double *out);

/* ---- multi-objective UDPs (restate_mo.c) ---- */
int oracle_zdt_fitness(unsigned id, const double *x, size_t N, double *f);
int oracle_dtlz_fitness(unsigned id, const double *x, size_t N, size_t M, unsigned alpha, double *f);
int oracle_zdt_batch(unsigned id, const double *xs, size_t n, size_t N, double *fs);
int oracle_dtlz_batch(unsigned id, const double *xs, size_t n, size_t N, size_t M, unsigned alpha, double *fs);

/* ---- meta-problems (restate_meta.c): translate.cpp:137-150, multi_objective.cpp:582-638 (method 0 weighted, 1 tchebycheff, 2 bi) ---- */
int oracle_translate_rows(const double *xs, size_t n, size_t nx, const double *t, double *out);
int oracle_decompose_objectives(const double *f, size_t m, const double *weight, const double *ref_point, int method, double *out);
int oracle_decompose_rows(const double *fs, size_t n, size_t m, const double *weight, const double *ref_point, int method, double *out);

/* ---- constrained UDPs and the unconstrain meta-problem (restate_constrained.c): hock_schittkowski_71.cpp:48-55,
 * luksan_vlcek1.cpp:60-77, unconstrain.cpp:136-223 (method 0 death penalty, 1 kuri, 2 weighted, 3 ignore_c, 4 ignore_o) ---- */
int oracle_hs71_batch(const double *xs, size_t n, double *fs);
int oracle_luksan_vlcek1_batch(size_t dim, const double *xs, size_t n, double *fs);
int oracle_unconstrain_rows(const double *fs, size_t n, size_t nobj, size_t nec, size_t nic, const double *c_tol, int method,
                            const double *weights, double *out);

/* ---- Lennard-Jones (restate_lj.c) ---- */
int oracle_lj_fitness(unsigned atoms, const double *x, double *f);
int oracle_lj_batch(unsigned atoms, const double *xs, size_t n, double *fs);

/* ---- multi-objective utilities (restate_mo_utils.c), f flat row-major [n x m] ---- */
int oracle_pareto_dominance(const double *a, const double *b, size_t m);
int oracle_fnds(const double *f, size_t n, size_t m, size_t *rank, size_t *dom_count, size_t *front_idx, size_t *front_off,
                size_t *nfronts);
/* same results as oracle_fnds, O(n) memory and OpenMP-parallel dominance tests: the full-size checker */
int oracle_fnds_nolist(const double *f, size_t n, size_t m, size_t *rank, size_t *dom_count, size_t *front_idx, size_t *front_off,
                       size_t *nfronts);
int oracle_crowding_distance(const double *f, size_t n, size_t m, double *out);
int oracle_select_best_N_mo(const double *f, size_t n, size_t m, size_t N, size_t *out, size_t *nout);
int oracle_sort_population_mo(const double *f, size_t n, size_t m, size_t *out);

/* ---- generic problem handle for the restated algorithms (restate_pso.c, restate_de.c) ----
 * family ids = pgc_family (1-5 simple, 6 cec2014, 8 zdt, 9 dtlz, 11 lennard_jones: dim = atoms) */
typedef struct oracle_problem {
    int family;
    unsigned prob_id, dim, nobj, param;
    const double *rotation, *shift;
    const int *shuffle;
} oracle_problem;
int oracle_problem_eval(const oracle_problem *p, const double *xs, size_t n, double *fs);
int oracle_pso_evolve(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, double *v, double *xcur,
                      size_t n, size_t dim, unsigned gens, double omega, double eta1, double eta2, double max_vel, unsigned variant,
                      unsigned neighb_type, unsigned neighb_param, uint64_t seed, uint32_t first_generation);

int oracle_de_evolve(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t NP, size_t dim,
                     unsigned gens, unsigned algo, unsigned variant, unsigned variant_adptv, double F, double CR,
                     const unsigned *allowed, unsigned n_allowed, double ftol, double xtol, uint64_t seed, uint32_t first_generation,
                     unsigned *gens_done, double *F_state, double *CR_state, unsigned *variant_state);

/* ---- WFG1..9 (restate_wfg.c): n = dim_dvs, M = dim_obj, k = dim_k ---- */
int oracle_wfg_check(unsigned prob_id, size_t n, size_t M, size_t k);
int oracle_wfg_fitness(unsigned prob_id, size_t n, size_t M, size_t k, const double *x, double *f);
int oracle_wfg_batch(unsigned prob_id, size_t n, size_t M, size_t k, const double *xs, size_t count, double *fs);

/* ---- exact hypervolume for m = 2, 3 (restate_hv.c) ---- */
int oracle_hv_check(const double *f, size_t n, size_t m, const double *r);
int oracle_hv_compute(const double *f, size_t n, size_t m, const double *r, double *out);
int oracle_hv_contributions(const double *f, size_t n, size_t m, const double *r, double *out);

/* ---- CMA-ES / xNES contractions (restate_cmaes.c) ---- */
int oracle_weighted_mean(const double *rows, const uint32_t *idx, const double *w, size_t k, size_t D, double *out);
int oracle_weighted_gram(const double *rows, const uint32_t *idx, const double *center, const double *w, size_t k, size_t D, double scale_div,
                         double *out);
int oracle_cmaes_sample(const double *mean, const double *bd, double sigma, size_t lambda, size_t D, uint64_t seed, uint32_t generation,
                        double *z, double *x);

int oracle_cmaes_evolve(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t lam, size_t D,
                        unsigned gens, double cc, double cs, double c1, double cmu, double sigma0, double ftol, double xtol, int force_bounds,
                        uint64_t seed, uint32_t first_generation, unsigned *gens_done, double *sigma_out);
/* xnes::evolve (xnes.cpp:96-303, memory = false) on the Philox normals; -1 = automatic for the etas and sigma0 */
int oracle_xnes_evolve(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t lam, size_t D, unsigned gens,
                       double eta_mu, double eta_sigma, double eta_b, double sigma0, double ftol, double xtol, int force_bounds, uint64_t seed,
                       uint32_t first_generation, unsigned *gens_done, double *sigma_out);

/* ---- migration (restate_migration.c): select_best / fair_replace on flat groups, topology in-edge lists ---- */
int oracle_select_best(const uint64_t *ids, const double *x, const double *f, size_t n, size_t nx, size_t nobj, int rate_is_frac, double rate,
                       uint64_t *ids_out, double *x_out, double *f_out, size_t *n_out);
int oracle_fair_replace(const uint64_t *ids, const double *x, const double *f, size_t n, size_t nx, size_t nobj, int rate_is_frac, double rate,
                        const uint64_t *mids, const double *mx, const double *mf, size_t nm, uint64_t *ids_out, double *x_out, double *f_out);
/* single-objective constrained groups: rows [f | nec eq | nic ineq], tol [nec + nic] (sort_population_con, constrained.cpp:180-202) */
int oracle_sort_population_con(const double *f, size_t n, size_t nec, size_t nic, const double *tol, size_t *out);
int oracle_select_best_con(const uint64_t *ids, const double *x, const double *f, size_t n, size_t nx, size_t nec, size_t nic, const double *tol,
                           int rate_is_frac, double rate, uint64_t *ids_out, double *x_out, double *f_out, size_t *n_out);
int oracle_fair_replace_con(const uint64_t *ids, const double *x, const double *f, size_t n, size_t nx, size_t nec, size_t nic, const double *tol,
                            int rate_is_frac, double rate, const uint64_t *mids, const double *mx, const double *mf, size_t nm,
                            uint64_t *ids_out, double *x_out, double *f_out);
int oracle_ring_connections(size_t n, size_t i, size_t *out, size_t *count);
int oracle_fully_connected_connections(size_t n, size_t i, size_t *out, size_t *count);
int oracle_population_init(const double *lb, const double *ub, size_t n, size_t nx, uint64_t seed, double *x, uint64_t *ids);

int oracle_sga_evolve(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t NP, size_t nx, unsigned gens,
                      double cr, double eta_c, double m, double param_m, unsigned param_s, unsigned crossover, unsigned mutation,
                      unsigned selection, uint64_t seed, uint32_t first_generation);

/* ---- Philox draws and NSGA-II generation operators (philox.h, restate_nsga2.c) ---- */
void oracle_philox_raw(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double oracle_philox_u01_at(uint64_t seed, uint32_t tag, uint32_t generation, uint32_t index, uint32_t slot);
int oracle_philox_perm(size_t n, uint64_t seed, uint32_t tag, uint32_t generation, size_t *perm);
/* integer alleles at the end of the chromosome (problem::get_nix()) for the NSGA-II operators below; per thread, default 0 */
void oracle_nsga2_set_nix(size_t nix);
int oracle_nsga2_variation(const double *x, const size_t *rank, const double *cd, size_t NP, size_t nx, const double *lb,
                           const double *ub, const size_t *sh1, const size_t *sh2, double cr, double eta_c, double m, double eta_m,
                           uint64_t seed, uint32_t generation, double *children);
int oracle_nsga2_rank_crowding(const double *f, size_t NP, size_t nobj, size_t *rank, double *cd);
int oracle_nsga2_evolve(int family, unsigned prob_id, size_t nx, size_t nobj, unsigned alpha, const double *lb, const double *ub,
                        double *x, double *f, size_t NP, unsigned gens, double cr, double eta_c, double m, double eta_m, uint64_t seed,
                        uint32_t first_generation);


/* ---- the same restatements on the REFERENCE's draw stream (sequential std::mt19937 + libstdc++ distributions, mt19937.h):
 * these are what tests/test_oracle_pin.py compares bit for bit with the compiled reference (ref_capi.h: ref_*_from) ---- */
/* 1: index sorts in libstdc++ std::sort order (ties as the compiled reference leaves them), 0 (default): stable */
void oracle_set_sort_mode(int libstdcxx);
int oracle_std_argsort(const double *keys, size_t n, int desc, size_t *out);
int oracle_mt_sequence(uint32_t seed, int kind, uint64_t a, uint64_t b, size_t n, double *out_real, uint64_t *out_int);
int oracle_mt_shuffles(uint32_t seed, size_t n, size_t rounds, size_t *perm);
int oracle_genetic_operators_mt(const double *p1, const double *p2, size_t nx, const double *lb, const double *ub, double p_cr, double eta_c,
                                double p_m, double eta_m, const size_t *rank, const double *cd, size_t n_pairs, uint32_t seed, double *c1,
                                double *c2, size_t *winners);
int oracle_population_init_mt(const double *lb, const double *ub, size_t n, size_t nx, uint32_t seed, double *x, uint64_t *ids);
int oracle_sga_evolve_mt(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t NP, size_t nx,
                         unsigned gens, double cr, double eta_c, double m, double param_m, unsigned param_s, unsigned crossover,
                         unsigned mutation, unsigned selection, uint32_t seed);
int oracle_mt_binomial_sequence(uint32_t seed, uint64_t t, double p, size_t n, uint64_t *out);
int oracle_de_evolve_sequential(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t NP, size_t dim,
                                unsigned gens, unsigned algo, unsigned variant, unsigned variant_adptv, double F, double CR,
                                const unsigned *allowed, unsigned n_allowed, double ftol, double xtol, uint64_t seed,
                                uint32_t first_generation, unsigned *gens_done, double *F_state, double *CR_state, unsigned *variant_state);
int oracle_de_evolve_mt(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t NP, size_t dim,
                        unsigned gens, unsigned algo, unsigned variant, unsigned variant_adptv, double F, double CR,
                        const unsigned *allowed, unsigned n_allowed, double ftol, double xtol, uint32_t seed, unsigned *gens_done);
int oracle_pso_evolve_mt(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t n, size_t dim,
                         unsigned gens, double omega, double eta1, double eta2, double max_vel, unsigned variant, unsigned neighb_type,
                         unsigned neighb_param, uint32_t seed);
int oracle_nsga2_evolve_mt(int family, unsigned prob_id, size_t nx, size_t nobj, unsigned alpha, const double *lb, const double *ub,
                           double *x, double *f, size_t NP, unsigned gens, double cr, double eta_c, double m, double eta_m, uint32_t seed);

#ifdef __cplusplus
}
#endif
/* nspso::evolve (src/algorithms/nspso.cpp:84-411); diversity: 0 crowding distance, 1 niche count, 2 max min.  vel / best_x / best_f:
 * the algorithm's memory (in/out), NULL = memory-less start (velocities drawn, archive = population) */
int oracle_nspso_evolve(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t n, size_t dim, size_t m,
                        unsigned gens, double omega, double c1, double c2, double chi, double v_coeff, unsigned leader_selection_range,
                        unsigned diversity, uint64_t seed, uint32_t first_generation, double *vel, double *best_x, double *best_f);
int oracle_nspso_evolve_mt(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t n, size_t dim, size_t m,
                           unsigned gens, double omega, double c1, double c2, double chi, double v_coeff, unsigned leader_selection_range,
                           unsigned diversity, uint32_t seed);

/* gaco::evolve (src/algorithms/gaco.cpp:104-445) on an unconstrained single-objective population, memory = false; the last nix
 * variables are integers.  The algorithm's scalar members that survive between evolve() calls travel in oracle_gaco_state. */
typedef struct {
    double oracle, q;
    unsigned n_evalstop, n_impstop, gen_mark;
    unsigned long long fevals;
    unsigned counter;  /* m_counter: evolve() calls so far, counted only with memory */
    int memory;        /* constructed with memory = true: the archive below survives between calls and is never written back */
    double *archive;   /* [ker x (1 + nx + 1)], caller-owned; used with memory only */
    int has_champion;  /* the POPULATION's champion (population.cpp:209-246: the best individual it has ever held) travels with the population */
    double champion;   /* from call to call; 0: taken as the best of the population handed in */
} oracle_gaco_state;
void oracle_gaco_state_init(oracle_gaco_state *s, double q, double oracle_par);
int oracle_gaco_evolve(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t n, size_t nx, size_t nix,
                       unsigned gens, unsigned ker, double acc, unsigned threshold, unsigned n_gen_mark, unsigned impstop, unsigned evalstop,
                       double focus, uint64_t seed, uint32_t first_generation, oracle_gaco_state *st, unsigned *gens_done);
int oracle_gaco_evolve_mt(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t n, size_t nx, size_t nix,
                          unsigned gens, unsigned ker, double q, double oracle_par, double acc, unsigned threshold, unsigned n_gen_mark,
                          unsigned impstop, unsigned evalstop, double focus, uint32_t seed, int memory, unsigned calls);

/* maco::evolve (src/algorithms/maco.cpp:88-533), memory = false; the algorithm object's m_q, m_n_evalstop, m_gen_mark travel in the state */
typedef struct {
    double q;
    unsigned n_evalstop, gen_mark;
} oracle_maco_state;
void oracle_maco_state_init(oracle_maco_state *s, double q);
int oracle_maco_evolve(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t n, size_t nx, size_t nix,
                       size_t m, unsigned gens, unsigned ker, unsigned threshold, unsigned n_gen_mark, unsigned evalstop, double focus,
                       uint64_t seed, uint32_t first_generation, oracle_maco_state *st, unsigned *gens_done);
int oracle_maco_evolve_mt(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t n, size_t nx, size_t nix,
                          size_t m, unsigned gens, unsigned ker, double q, unsigned threshold, unsigned n_gen_mark, unsigned evalstop,
                          double focus, uint32_t seed);

/* moead_gen::evolve (src/algorithms/moead_gen.cpp:128-345) with the weight vectors [NP x m] and their neighbourhoods [NP x T] given;
 * decomposition: 0 weighted, 1 tchebycheff, 2 bi */
int oracle_moead_gen_evolve(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t NP, size_t dim, size_t m,
                            unsigned gens, const double *weights, const size_t *neigh, size_t T, int decomposition, double CR, double F,
                            double eta_m, double realb, unsigned limit, int preserve_diversity, uint64_t seed, uint32_t first_generation,
                            size_t burn_draws);
int oracle_moead_gen_evolve_mt(const oracle_problem *prob, const double *lb, const double *ub, double *x, double *f, size_t NP, size_t dim,
                               size_t m, unsigned gens, const double *weights, const size_t *neigh, size_t T, int decomposition, double CR, double F,
                               double eta_m, double realb, unsigned limit, int preserve_diversity, uint32_t seed, size_t burn_draws);

#endif
