/* Minimal stand-in for <gsl/gsl_rng.h> (GSL is not vendored by the reference and not
 * installed here; Makefile:54 links an unpinned -lgsl).  TEST INFRASTRUCTURE ONLY.
 * Implements gsl_rng_mt19937 = the standard MT19937 (Matsumoto & Nishimura 2002,
 * init_genrand seeding; seed 0 -> 4357 as GSL documents) with
 * gsl_rng_uniform = genrand_int32 / 2^32.  See gsl_stub.c.
 * Adds a draw log so tests can replay the exact stream the reference consumed.
 */
#ifndef ORACLE_STUB_GSL_RNG_H
#define ORACLE_STUB_GSL_RNG_H
typedef struct { int dummy; } gsl_rng_type;
typedef struct gsl_rng_s gsl_rng;
extern const gsl_rng_type *gsl_rng_mt19937;
extern const gsl_rng_type *gsl_rng_default;
const gsl_rng_type *gsl_rng_env_setup(void);
gsl_rng *gsl_rng_alloc(const gsl_rng_type *t);
void gsl_rng_set(gsl_rng *r, unsigned long seed);
double gsl_rng_uniform(gsl_rng *r);
void gsl_rng_free(gsl_rng *r);

/* oracle-side hooks (not part of GSL) */
void   oracle_rng_log_begin(double *buf, long cap);   /* record every uniform drawn */
long   oracle_rng_log_count(void);
void   oracle_rng_log_end(void);
long   oracle_rng_total_draws(void);
#endif
