// Fast build of the device code: algebraically identical rewrites that drop transcendental calls,
// compiled with FMA contraction (-fmad=true).  Held to the same parity tests as the strict build.
#define SENV_NS senv_fast
#define SENV_FAST_MATH 1
#include "shipenv_kernels.cuh"
