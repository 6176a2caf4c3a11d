// Host stand-in for csrc/kernels/psad_common.cuh for the GENERIC kernels.  TEST INFRASTRUCTURE: lets g++ compile an
// emitted generic kernel unchanged and run it on the CPU, one call per CUDA thread (the kernel has no inter-thread
// communication: grid-stride loops over cells, scalar loads and stores).
#ifndef PSAD_COMMON_CUH
#define PSAD_COMMON_CUH

#include <cmath>
#include <cstring>

#include "psad_args.h"

typedef unsigned int psad_u32;
typedef unsigned long long psad_u64;

#define PSAD_DEV static inline
#define __device__
#define __global__
#define __launch_bounds__(...)
#define __grid_constant__

struct PsadEmuDim3 { unsigned x, y, z; };
static PsadEmuDim3 blockIdx, threadIdx, gridDim, blockDim;

static inline float psad_rsqrt(float x) { return 1.0f / std::sqrt(x); }
static inline double psad_rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
template <int N, typename T> static inline T psad_ipow(T x) {
  T r = x;
  for (int i = 1; i < N; ++i) r *= x;
  return r;
}

#endif
