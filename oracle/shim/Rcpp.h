/* Rcpp.h -- header shim (test infrastructure): see RcppArmadillo.h in this directory. */
#include "RcppArmadillo.h"
