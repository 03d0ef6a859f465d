// Scalar types of the host layer.  Same compile-time switches as the reference
// (ref/types.h:61-81): -DPRECISION=1|2 selects MMD_float, -DPAD4 the AoS stride.
#pragma once

#ifndef PRECISION
#define PRECISION 2
#endif
#if PRECISION == 1
typedef float MMD_float;
#elif PRECISION == 2
typedef double MMD_float;
#else
#error "PRECISION must be 1 (float) or 2 (double)"
#endif
typedef int MMD_int;

#ifdef PAD4
#define PAD 4
#else
#define PAD 3
#endif

enum ForceStyle { FORCELJ, FORCEEAM };
enum { LJ = 0, METAL = 1 };

#define VARIANT_STRING "miniMD-B200 2.0 (sm_100a CUDA hot path, reference-compatible host layer)"
