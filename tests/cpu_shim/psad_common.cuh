// Host stand-in for csrc/kernels/psad_common.cuh.  TEST INFRASTRUCTURE: lets g++ compile an emitted march kernel's
// per-step body (psad_item_begin / psad_step) and replay it on the CPU, one OS thread per lane of a warp running in
// lock step at the warp-level operations (shuffles, __syncwarp).  Shared-memory loads read a host copy of the TMA
// box, stores go to host arrays, mbarrier arrivals are no-ops (the replay stages planes synchronously).
#ifndef PSAD_COMMON_CUH
#define PSAD_COMMON_CUH

#include <barrier>
#include <cmath>
#include <cstring>

#include "psad_args.h"

typedef unsigned int psad_u32;
typedef unsigned long long psad_u64;

#define PSAD_DEV static inline
#define __device__

struct PsadEmuWarp {
  std::barrier<> bar{32};
  double slots[32];
};
extern thread_local PsadEmuWarp* psad_emu_warp;
extern thread_local int psad_emu_lane;

extern thread_local std::barrier<>* psad_emu_cta;     // all consumer threads of the replayed CTA (exchange kernels)

static inline void __syncwarp() { psad_emu_warp->bar.arrive_and_wait(); }
static inline void psad_consumer_barrier(int) { psad_emu_cta->arrive_and_wait(); }
static inline void psad_mbar_arrive(psad_u32) {}

template <typename T> static inline T psad_emu_shift(T v, int delta) {
  PsadEmuWarp* w = psad_emu_warp;
  std::memcpy(&w->slots[psad_emu_lane], &v, sizeof(T));
  w->bar.arrive_and_wait();
  const int src = psad_emu_lane + delta;
  T r = v;   // lanes without a source keep their own value, like __shfl_up_sync / __shfl_down_sync
  if (src >= 0 && src < 32) std::memcpy(&r, &w->slots[src], sizeof(T));
  w->bar.arrive_and_wait();
  return r;
}
template <typename T> static inline T psad_from_left(T v) { return psad_emu_shift(v, -1); }
template <typename T> static inline T psad_from_right(T v) { return psad_emu_shift(v, +1); }

template <typename T> static inline void psad_lds_vec(const T* p, T* e) { std::memcpy(e, p, 16); }
// the device version orders its two 16-byte loads by lane to avoid bank conflicts; the values are the same 32 bytes
template <typename T> static inline void psad_lds_pair(const T* p, int hi, T* e) { (void)hi; std::memcpy(e, p, 32); }
template <typename T> static inline void psad_stg_vec(T* p, const T* e) { std::memcpy(p, e, 16); }
template <typename T> static inline void psad_sts_vec(T* p, const T* e) { std::memcpy(p, e, 16); }

static inline float psad_rsqrt(float x) { return 1.0f / std::sqrt(x); }
static inline double psad_rsqrt(double x) { return 1.0 / std::sqrt(x); }
template <int N, typename T> static inline T psad_ipow(T x) {
  T r = x;
  for (int i = 1; i < N; ++i) r *= x;
  return r;
}

#endif
