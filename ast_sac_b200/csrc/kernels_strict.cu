// Strict build of the device code: the reference's formulas statement by statement, compiled with
// -fmad=false (one IEEE rounding per operation, like CPython / NumPy scalar arithmetic).
#define SENV_NS senv_strict
#define SENV_FAST_MATH 0
#include "shipenv_kernels.cuh"
