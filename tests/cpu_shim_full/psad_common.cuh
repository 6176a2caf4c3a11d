// Host stand-in for csrc/kernels/psad_common.cuh that lets g++ compile the REAL march template
// (csrc/kernels/psad_march.cuh: producer warp, TMA ring, full / empty mbarriers, consumer loop) together with an emitted
// kernel and run it on the CPU, one OS thread per CUDA thread of a CTA.  TEST INFRASTRUCTURE.
//   * shared memory = one host array; "shared-window addresses" (psad_u32) are offsets into it;
//   * mbarriers: pending-arrival count + transaction bytes + phase bit, completed exactly as the hardware defines it
//     (phase flips when both reach zero); waits block on a condition variable;
//   * TMA loads copy the box (zero fill outside the array) and then complete_tx on the barrier;
//   * warp shuffles / __syncwarp / __syncthreads / the consumers' named barrier are real barriers between the threads.
// What it cannot show: memory-ordering and proxy-fence subtleties of the hardware — only the protocol logic.
#ifndef PSAD_COMMON_CUH
#define PSAD_COMMON_CUH

#include <barrier>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "psad_args.h"

typedef unsigned int psad_u32;
typedef unsigned long long psad_u64;

#define PSAD_DEV static inline
#define __device__
#define __noinline__
#define __global__
#define __grid_constant__
#define __shared__
#define __align__(n)
#define __launch_bounds__(...)

struct PsadEmuDim3 { unsigned x, y, z; };
extern thread_local PsadEmuDim3 threadIdx, blockIdx, gridDim;

constexpr int PSAD_EMU_SMEM = 232 * 1024;
extern unsigned char psad_smem[];

struct PsadEmuWarp {
  std::barrier<> bar{32};
  double slots[32];
};
struct PsadEmuField {      // what a tensor map describes: one staged input field
  const void* ptr;
  long long stride[3];     // element strides (z, y, x)
  long long shape[3];
  int esize, boxw, boxh;
};
struct PsadEmuMbar { int count = 0, pending = 0; long long tx = 0; unsigned phase = 0; };

extern thread_local PsadEmuWarp* psad_emu_warp;
extern thread_local int psad_emu_lane;
extern std::barrier<>* psad_emu_cta_all;        // __syncthreads(): every thread of the CTA
extern std::barrier<>* psad_emu_cta_consumers;  // named barrier 1: the consumer warps
extern std::mutex psad_emu_mutex;
extern std::condition_variable psad_emu_cv;
extern PsadEmuMbar psad_emu_mbar[PSAD_EMU_SMEM / 8];
extern long long psad_emu_waits, psad_emu_tma_loads;

PSAD_DEV psad_u32 psad_smem_u32(const void* p) { return (psad_u32)((const unsigned char*)p - psad_smem); }
static inline void __syncthreads() { psad_emu_cta_all->arrive_and_wait(); }
static inline void __syncwarp() { psad_emu_warp->bar.arrive_and_wait(); }
static inline void psad_consumer_barrier(int) { psad_emu_cta_consumers->arrive_and_wait(); }

static inline void psad_emu_check(PsadEmuMbar& m) {   // caller holds the mutex
  if (m.pending == 0 && m.tx == 0) {
    m.phase ^= 1u;
    m.pending = m.count;
    psad_emu_cv.notify_all();
  }
}
PSAD_DEV void psad_mbar_init(psad_u32 bar, psad_u32 count) {
  std::lock_guard<std::mutex> g(psad_emu_mutex);
  psad_emu_mbar[bar / 8] = PsadEmuMbar{(int)count, (int)count, 0, 0u};
}
PSAD_DEV void psad_fence_barrier_init() {}
PSAD_DEV void psad_fence_proxy_async() {}
PSAD_DEV void psad_mbar_arrive_expect_tx(psad_u32 bar, psad_u32 bytes) {
  std::lock_guard<std::mutex> g(psad_emu_mutex);
  PsadEmuMbar& m = psad_emu_mbar[bar / 8];
  m.tx += bytes;
  --m.pending;
  psad_emu_check(m);
}
PSAD_DEV void psad_mbar_arrive(psad_u32 bar) {
  std::lock_guard<std::mutex> g(psad_emu_mutex);
  PsadEmuMbar& m = psad_emu_mbar[bar / 8];
  --m.pending;
  psad_emu_check(m);
}
PSAD_DEV void psad_mbar_wait(psad_u32 bar, psad_u32 parity) {
  std::unique_lock<std::mutex> g(psad_emu_mutex);
  PsadEmuMbar& m = psad_emu_mbar[bar / 8];
  ++psad_emu_waits;
  psad_emu_cv.wait(g, [&] { return (m.phase & 1u) != parity; });   // the phase with this parity has completed
}

static inline void psad_emu_tma(psad_u32 dst, const PsadTensorMap* tmap, psad_u32 bar, int x0, int y0, int z) {
  const PsadEmuField* F;
  std::memcpy(&F, &tmap->opaque[0], sizeof(F));
  for (int by = 0; by < F->boxh; ++by)
    for (int bx = 0; bx < F->boxw; ++bx) {
      const long long gx = x0 + bx, gy = y0 + by;
      unsigned char* d = psad_smem + dst + ((size_t)by * F->boxw + bx) * F->esize;
      if (gx < 0 || gx >= F->shape[2] || gy < 0 || gy >= F->shape[1] || z < 0 || z >= F->shape[0])
        std::memset(d, 0, F->esize);
      else
        std::memcpy(d, (const unsigned char*)F->ptr + (z * F->stride[0] + gy * F->stride[1] + gx * F->stride[2]) * F->esize, F->esize);
    }
  std::lock_guard<std::mutex> g(psad_emu_mutex);
  PsadEmuMbar& m = psad_emu_mbar[bar / 8];
  m.tx -= (long long)F->boxw * F->boxh * F->esize;
  ++psad_emu_tma_loads;
  psad_emu_check(m);
}
PSAD_DEV void psad_tma_load_2d(psad_u32 dst, const PsadTensorMap* tmap, psad_u32 bar, int c0, int c1) { psad_emu_tma(dst, tmap, bar, c0, c1, 0); }
PSAD_DEV void psad_tma_load_3d(psad_u32 dst, const PsadTensorMap* tmap, psad_u32 bar, int c0, int c1, int c2) { psad_emu_tma(dst, tmap, bar, c0, c1, c2); }
PSAD_DEV void psad_tma_prefetch_desc(const PsadTensorMap*) {}

// peer halos: the neighbours' counters are plain host words here; a counter that has not reached `expect` is what the
// device version would spin on — recorded as an error instead (the test sets the counters before the launch)
long long psad_emu_peer_waits = 0;
PSAD_DEV void psad_wait_peer(const unsigned* flag, unsigned expect, unsigned* error) {
  std::lock_guard<std::mutex> g(psad_emu_mutex);
  ++psad_emu_peer_waits;
  if ((int)(*flag - expect) < 0 && error) *error = 1u;
}

template <typename T> static inline T psad_emu_shift(T v, int delta) {
  PsadEmuWarp* w = psad_emu_warp;
  std::memcpy(&w->slots[psad_emu_lane], &v, sizeof(T));
  w->bar.arrive_and_wait();
  const int src = psad_emu_lane + delta;
  T r = v;
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
