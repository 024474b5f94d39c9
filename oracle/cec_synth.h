/* oracle/cec_synth.h - seeded synthetic CEC2013/CEC2014 data tables.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference's real tables (src/problems/cec2014_data.cpp, cec2013_data.cpp) are missing from the
 * checkout (/root/reference/.MISSING_LARGE_BLOBS), so the oracle, the compiled reference (oracle/_ref)
 * and the CUDA engine are all fed the SAME synthetic tables produced here.  Shapes follow what the
 * reference constructors expect (cec2014.cpp:66-94, cec2013.cpp:65-68):
 *   cec2014 rotation_data[func][dim] : 10 row-major dim x dim orthogonal matrices, back to back
 *   cec2014 shift_data[func]         : 10 lines of 100 values in [-80,80)
 *   cec2014 shuffle_data[func][dim]  : 10 permutations of 1..dim (1-based), back to back
 *   cec2013 MD[dim]                  : 10 row-major dim x dim orthogonal matrices
 *   cec2013 shift_data               : 10 lines of 100 values in [-80,80)
 * Only + - * / sqrt and integer arithmetic are used (compiled with -ffp-contract=off), so the tables are
 * bit-reproducible on any IEEE-754 host.
 */
#ifndef ORACLE_CEC_SYNTH_H
#define ORACLE_CEC_SYNTH_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define CEC_SYNTH_NCOMP 10 /* matrices / shift lines / permutations per function */

/* building blocks */
void cec_synth_uniform(uint64_t seed, double lo, double hi, double *out, size_t n);
void cec_synth_rotation(uint64_t seed, unsigned dim, double *out /* dim*dim */);
void cec_synth_perm(uint64_t seed, unsigned dim, int *out /* dim, 1-based */);

/* suite tables (seed derived from suite, func, dim, component index) */
void cec2014_synth_rotation(unsigned func, unsigned dim, double *out /* 10*dim*dim */);
void cec2014_synth_shift(unsigned func, double *out /* 10*100 */);
void cec2014_synth_shuffle(unsigned func, unsigned dim, int *out /* 10*dim */);
void cec2013_synth_md(unsigned dim, double *out /* 10*dim*dim */);
void cec2013_synth_shift(double *out /* 10*100 */);

#ifdef __cplusplus
}
#endif
#endif
