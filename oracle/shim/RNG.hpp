// oracle/shim/RNG.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Declaration shim for the reference's missing `class RNG` (jwindle/RNG, absent
// from /root/reference, see oracle/l0.h).  It exists so that the reference's own
// sampler sources -- PolyaGamma.cpp, PolyaGammaAlt.cpp, PolyaGammaSP.cpp,
// InvertY.cpp -- compile UNMODIFIED and IN PLACE (oracle/Makefile, target
// _ref/libpg_ref.so).  Only the members those four files call are declared:
//   PolyaGamma.cpp:61,74-75,89,94-96,106,110,147,170-171,176
//   PolyaGammaAlt.cpp:12-15,29-30,56,66,73,84,92,103,149,151,161
//   PolyaGammaSP.cpp:63-64,71,218,222,244,250,257
// Every member forwards to the plain-C L0 layer in oracle/l0.c, which draws
// from a variate tape or from the documented Philox stream.
#ifndef PGO_SHIM_RNG_HPP
#define PGO_SHIM_RNG_HPP

#include <cmath>
#include <cstdio>

#include "l0.h"

class RNG {
 public:
  pgo_src src;

  RNG() { pgo_src_philox(&src, 0, 0, 0); }

  double unif() { return pgo_unif(&src); }
  double expon_rate(double rate) { return pgo_expon(&src) / rate; }
  double norm(double sd) { return sd * pgo_norm(&src); }
  double norm(double mean, double sd) { return mean + sd * pgo_norm(&src); }
  double gamma_scale(double shape, double scale) { return scale * pgo_gamma(&src, shape); }
  double igauss(double mu, double lambda) { return pgo_igauss(&src, mu, lambda); }
  double ltgamma(double shape, double rate, double trunc) {
    return pgo_ltgamma(&src, shape, rate, trunc);
  }
  double rtinvchi2(double scale, double trunc) { return pgo_rtinvchi2(&src, scale, trunc); }

  static double p_norm(double x, int use_log = 0) { return pgo_p_norm(x, use_log); }
  static double p_gamma_rate(double x, double shape, double rate) {
    return pgo_p_gamma_rate(x, shape, rate);
  }
  static double p_igauss(double x, double mu, double lambda) {
    return pgo_p_igauss(x, mu, lambda);
  }
  static double Gamma(double x, int use_log = 0) { return pgo_Gamma(x, use_log); }
};

#endif
