// Declaration stub for the five GSL symbols the reference's tree.cpp uses (root Dirichlet noise only,
// outside the hot path). Definitions live in oracle/ref_link_stubs.cpp.
#pragma once
#include <cstddef>
struct gsl_rng_type;
struct gsl_rng;
extern const gsl_rng_type *gsl_rng_mt19937;
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_set(gsl_rng *r, unsigned long seed);
void gsl_rng_free(gsl_rng *r);
