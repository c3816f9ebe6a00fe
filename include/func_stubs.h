/*
 * func_stubs.h — compatibility shim.
 *
 * In SynchroPMU this header is the compile-time plug layer: the float_p / complex_p /
 * uint_p / bool_p type macros (reference src/func_stubs.h:33-47) and the overridable
 * pmue_fft_r / pmue_sin / pmue_cexp ... math stubs (src/func_stubs.h:60-143).  In the
 * B200 library the math runs inside the CUDA kernel, so only the type vocabulary is
 * left; it lives in pmu_estimator.h and this file exists so that
 * `#include "func_stubs.h"` in caller code keeps compiling.
 */
#ifndef FUNC_STUBS_H
#define FUNC_STUBS_H
#include "pmu_estimator.h"
#endif
