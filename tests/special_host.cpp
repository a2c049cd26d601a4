// Host build of examodels.jl_b200/csrc/exb_special.h for tests/test_special_functions.py (g++, no CUDA): the hand-written
// special-function algorithms of the device path, callable through ctypes so they can be pinned against scipy without a GPU.
#include "../examodels.jl_b200/csrc/exb_special.h"
extern "C" {
double sf_digamma(double x) { return exb_digamma(x); }
double sf_trigamma(double x) { return exb_trigamma(x); }
double sf_polygamma2(double x) { return exb_polygamma2(x); }
double sf_polygamma3(double x) { return exb_polygamma3(x); }
double sf_invdigamma(double x) { return exb_invdigamma(x); }
double sf_dawson(double x) { return exb_dawson(x); }
double sf_erfi(double x) { return exb_erfi(x); }
double sf_airy(double x, int which) { return exb_airy(x, which); }
}
