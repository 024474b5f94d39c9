/* oracle/ref_capi.h - plain-C handle API over the UNMODIFIED reference sources.  TEST INFRASTRUCTURE ONLY.
 *
 * oracle/_ref/libpagmo_ref.so is compiled by oracle/Makefile from /root/reference/src/... (where the
 * files lie, never copied) against oracle/shim/, plus this wrapper.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  All functions return 0 on success,
 * non-zero on a C++ exception (message via ref_last_error()).
 */
#ifndef ORACLE_REF_CAPI_H
#define ORACLE_REF_CAPI_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ref_problem ref_problem;

const char *ref_last_error(void);

/* family: "rastrigin","ackley","griewank","schwefel","rosenbrock" (p0=dim),
 *         "cec2014","cec2013" (p0=prob_id,p1=dim), "zdt" (p0=prob_id,p1=param),
 *         "dtlz" (p0=prob_id,p1=dim,p2=fdim,p3=alpha), "wfg" (p0=prob_id,p1=dim_dvs,p2=dim_obj,p3=dim_k),
 *         "lennard_jones" (p0=atoms).  Wraps `pagmo::problem{udp}` (problem.cpp:154-242). */
int ref_problem_create(const char *family, unsigned p0, unsigned p1, unsigned p2, unsigned p3, ref_problem **out);
/* pagmo::problem{pagmo::translate{inner, t}} (translate.hpp) and pagmo::problem{pagmo::decompose{inner, w, z, method, adapt_ideal}}
 * (decompose.hpp); the inner problem is copied, as the reference's meta-problems do. */
int ref_problem_translate(const ref_problem *inner, const double *t, size_t len, ref_problem **out);
int ref_problem_decompose(const ref_problem *inner, const double *w, const double *z, size_t len, const char *method, int adapt_ideal,
                          ref_problem **out);
/* pagmo::decompose_objectives (utils/multi_objective.cpp:582-638) */
int ref_decompose_objectives(const double *f, const double *w, const double *z, size_t m, const char *method, double *out);
void ref_problem_destroy(ref_problem *p);
size_t ref_problem_nx(const ref_problem *p);
size_t ref_problem_nf(const ref_problem *p);
size_t ref_problem_nobj(const ref_problem *p);
size_t ref_problem_nec(const ref_problem *p);
size_t ref_problem_nic(const ref_problem *p);
/* pagmo::problem{pagmo::unconstrain{inner, method, weights}} (unconstrain.hpp; methods "death penalty", "kuri", "weighted",
 * "ignore_c", "ignore_o") and problem::set_c_tol (problem.cpp:620-644) */
int ref_problem_unconstrain(const ref_problem *inner, const char *method, const double *weights, size_t len, ref_problem **out);
int ref_problem_set_c_tol(ref_problem *p, const double *tol, size_t len);
unsigned long long ref_problem_fevals(const ref_problem *p);
int ref_problem_bounds(const ref_problem *p, double *lb, double *ub);
int ref_problem_name(const ref_problem *p, char *buf, size_t buflen);
/* problem::fitness (problem.cpp:353-380) */
int ref_problem_fitness(const ref_problem *p, const double *x, double *f);
/* n sequential problem::fitness calls on one thread (what de/de1220/sade/pso/sga do) */
int ref_problem_fitness_loop(const ref_problem *p, const double *dvs, size_t n, double *fvs);
/* pagmo::bfe{thread_bfe{}}(prob, dvs) (bfe.cpp:91-110 -> thread_bfe.cpp:64-144); nthreads<=0: all cores */
int ref_thread_bfe(ref_problem *p, const double *dvs, size_t n, double *fvs, int nthreads);
/* pagmo::bfe{} == default_bfe (default_bfe.cpp:53-68) */
int ref_default_bfe(ref_problem *p, const double *dvs, size_t n, double *fvs, int nthreads);

/* the tables the reference constructors saw (synthetic, see cec_synth.h); sizes: Mr 10*dim*dim,
 * Os 10*100 (uncompacted lines), S 10*dim */
int ref_cec2014_tables(unsigned func, unsigned dim, double *Mr, double *Os, int *S);
int ref_cec2013_tables(unsigned dim, double *Mr, double *Os);
/* cec2014::get_origin_shift() (cec2014.hpp:104): the compacted shift vector, returns its length in *n */
int ref_cec2014_origin_shift(const ref_problem *p, double *out, size_t cap, size_t *n);

/* ---- multi-objective utilities (src/utils/multi_objective.cpp) on flat row-major [n x m] ---- */
int ref_pareto_dominance(const double *a, const double *b, size_t m, int *out);
/* fast_non_dominated_sorting :200-257.  rank[n], dom_count[n]; fronts flattened into front_idx[n] with
 * front_off[nfronts+1]; dom_list flattened into dl_idx[*dl_total<=cap] with dl_off[n+1] (pass NULL to skip). */
int ref_fnds(const double *f, size_t n, size_t m, size_t *rank, size_t *dom_count, size_t *front_idx,
             size_t *front_off, size_t *nfronts, size_t *dl_idx, size_t *dl_off, size_t dl_cap);
int ref_crowding_distance(const double *f, size_t n, size_t m, double *out);
int ref_sort_population_mo(const double *f, size_t n, size_t m, size_t *out);
int ref_select_best_N_mo(const double *f, size_t n, size_t m, size_t N, size_t *out, size_t *nout);
int ref_ideal(const double *f, size_t n, size_t m, double *out);
int ref_nadir(const double *f, size_t n, size_t m, double *out);

/* ---- unmodified reference algorithms (ref_algos.cpp): evolve a fresh population(prob, pop_size, pop_seed) for `gens`
 * generations with reference default parameters; returns the wall time of evolve() and the final population.
 * algo: "nsga2" (nsga2.cpp:91-307; thread_bfe when use_bfe), "de", "de1220", "sade", "pso", "pso_gen". */
int ref_evolve(ref_problem *p, const char *algo, unsigned pop_size, unsigned gens, unsigned pop_seed, unsigned algo_seed,
               int use_bfe, double *seconds, double *x_out, double *f_out, unsigned long long *fevals);

/* ---- unmodified reference migration policies (ref_policies.cpp) on flat row-major groups (ids[n], x[n x nx], f[n x nobj]);
 * unconstrained problems (nec = nic = 0).  fair_replace::replace fair_replace.cpp:63-221, select_best::select select_best.cpp:63-171.
 * rate: absolute count, or a fraction of n when rate_is_frac. */
int ref_fair_replace(const unsigned long long *ids, const double *x, const double *f, size_t n, size_t nx, size_t nobj, int rate_is_frac,
                     double rate, const unsigned long long *mids, const double *mx, const double *mf, size_t nm, unsigned long long *ids_out,
                     double *x_out, double *f_out);
int ref_select_best(const unsigned long long *ids, const double *x, const double *f, size_t n, size_t nx, size_t nobj, int rate_is_frac,
                    double rate, unsigned long long *ids_out, double *x_out, double *f_out, size_t *n_out);

/* ---- hypervolume::compute / contributions (hypervolume.cpp:196-330) with the reference's own algorithm choice ---- */
int ref_fair_replace_con(const unsigned long long *ids, const double *x, const double *f, size_t n, size_t nx, size_t nec, size_t nic,
                         const double *tol, int rate_is_frac, double rate, const unsigned long long *mids, const double *mx, const double *mf,
                         size_t nm, unsigned long long *ids_out, double *x_out, double *f_out);
int ref_select_best_con(const unsigned long long *ids, const double *x, const double *f, size_t n, size_t nx, size_t nec, size_t nic,
                        const double *tol, int rate_is_frac, double rate, unsigned long long *ids_out, double *x_out, double *f_out, size_t *n_out);
int ref_hv_fpras(const double *f, size_t n, size_t m, const double *r, double eps, double delta, unsigned seed, double *out);
int ref_hv_approx_extreme(const double *f, size_t n, size_t m, const double *r, int greatest, int use_exact, double eps, double delta,
                          unsigned seed, size_t *out);
int ref_hv_compute(const double *f, size_t n, size_t m, const double *r, double *out);
int ref_hv_contributions(const double *f, size_t n, size_t m, const double *r, double *out);

/* ---- pinning the restatements to the reference's own random stream (ref_pin.cpp) ----
 * The real std::mt19937 + libstdc++ distributions (what oracle/mt19937.h restates): kind 0 raw words, 1 uniform_real(0,1),
 * 2 uniform_int<size_t>(a,b), 3 one normal_distribution(0,1) object, 4 uniform_real(-a,b). */
int ref_std_sequence(unsigned seed, int kind, unsigned long long a, unsigned long long b, size_t n, double *out_real,
                     unsigned long long *out_int);
int ref_std_shuffles(unsigned seed, size_t n, size_t rounds, size_t *perm);
int ref_std_argsort(const double *keys, size_t n, int desc, size_t *out);
int ref_std_binomial(unsigned seed, unsigned long long t, double p, size_t n, unsigned long long *out);
/* detail::sbx_crossover_impl + polynomial_mutation_impl (both children) + n_pairs mo_tournament_selection_impl((2i,2i+1)) on one
 * engine seeded with `seed` (genetic_operators.cpp:71-211) */
int ref_genetic_operators(const double *p1, const double *p2, size_t nx, const double *lb, const double *ub, double p_cr, double eta_c,
                          double p_m, double eta_m, const size_t *rank, const double *cd, size_t n_pairs, unsigned seed, double *c1,
                          double *c2, size_t *winners);
/* an unmodified reference UDA ("nsga2", "pso_gen", "de", "sade", "de1220", "sga") evolving the population with decision vectors
 * x0 [n x nx] for `gens` generations with algorithm seed `seed`; par[] = the constructor's arguments between gen and seed */
int ref_evolve_from(ref_problem *p, const char *algo, const double *par, size_t npar, const char *strategies, const double *x0, size_t n,
                    unsigned gens, unsigned seed, double *x_out, double *f_out);
/* population(prob, n, seed): its decision vectors and ids (population.cpp:62-80) */
int ref_population_init(ref_problem *p, size_t n, unsigned seed, double *x_out, unsigned long long *ids_out);

#ifdef __cplusplus
}
#endif
#endif
