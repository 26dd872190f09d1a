#pragma once
#include <cstddef>
#include "gsl_rng.h"
void gsl_ran_dirichlet(const gsl_rng *r, size_t K, const double alpha[], double theta[]);
