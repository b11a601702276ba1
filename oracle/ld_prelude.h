// TEST INFRASTRUCTURE ONLY.  Force-included by the extended-precision build (liboracle_ld.so): the restatement
// calls the C math functions unqualified; without these using-declarations a long double argument would be
// silently narrowed to ::log(double) etc. and the build would compute in fp64 after all.
#pragma once
#include <cmath>
using std::acos;
using std::atan;
using std::cos;
using std::exp;
using std::fabs;
using std::log;
using std::pow;
using std::sin;
using std::sqrt;
using std::tanh;
using std::fmax;
using std::fmin;
using std::floor;
