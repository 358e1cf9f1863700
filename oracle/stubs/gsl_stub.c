/* MT19937 behind the tiny gsl_rng surface sidm_rand.c uses.  TEST INFRASTRUCTURE ONLY.
 * Algorithm: Matsumoto & Nishimura, "Mersenne Twister" (1998), 2002 initialisation. */
#include <stdlib.h>
#include "gsl/gsl_rng.h"

#define MT_N 624
#define MT_M 397

struct gsl_rng_s { unsigned int mt[MT_N]; int idx; };

static const gsl_rng_type mt_type = {0};
const gsl_rng_type *gsl_rng_mt19937 = &mt_type;
const gsl_rng_type *gsl_rng_default = &mt_type;

static double *log_buf = 0;
static long log_cap = 0, log_n = 0, total_draws = 0;

const gsl_rng_type *gsl_rng_env_setup(void) { return &mt_type; }

gsl_rng *gsl_rng_alloc(const gsl_rng_type *t)
{
  (void)t;
  gsl_rng *r = (gsl_rng *)malloc(sizeof(gsl_rng));
  gsl_rng_set(r, 0);
  return r;
}

void gsl_rng_free(gsl_rng *r) { free(r); }

void gsl_rng_set(gsl_rng *r, unsigned long seed)
{
  if (seed == 0) seed = 4357;
  r->mt[0] = (unsigned int)(seed & 0xffffffffUL);
  for (int i = 1; i < MT_N; i++)
    r->mt[i] = 1812433253U * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (unsigned int)i;
  r->idx = MT_N;
}

static unsigned int mt_next(gsl_rng *r)
{
  if (r->idx >= MT_N) {
    unsigned int *mt = r->mt;
    for (int k = 0; k < MT_N; k++) {
      unsigned int y = (mt[k] & 0x80000000U) | (mt[(k + 1) % MT_N] & 0x7fffffffU);
      mt[k] = mt[(k + MT_M) % MT_N] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
    }
    r->idx = 0;
  }
  unsigned int y = r->mt[r->idx++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680U;
  y ^= (y << 15) & 0xefc60000U;
  y ^= (y >> 18);
  return y;
}

double gsl_rng_uniform(gsl_rng *r)
{
  double u = mt_next(r) / 4294967296.0;
  total_draws++;
  if (log_buf && log_n < log_cap) log_buf[log_n] = u;
  if (log_buf) log_n++;
  return u;
}

void oracle_rng_log_begin(double *buf, long cap) { log_buf = buf; log_cap = cap; log_n = 0; }
long oracle_rng_log_count(void) { return log_n; }
void oracle_rng_log_end(void) { log_buf = 0; log_cap = 0; }
long oracle_rng_total_draws(void) { return total_draws; }
