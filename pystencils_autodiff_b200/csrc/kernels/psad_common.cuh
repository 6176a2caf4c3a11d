// psad_common.cuh — hand-written sm_100a device primitives used by every specialised stencil kernel:
// mbarrier handshakes, TMA (cp.async.bulk.tensor) tile loads, vector shared/global accessors, warp halo shuffles.
#ifndef PSAD_COMMON_CUH
#define PSAD_COMMON_CUH

#include "psad_args.h"

typedef unsigned int psad_u32;
typedef unsigned long long psad_u64;

#define PSAD_DEV __device__ __forceinline__

PSAD_DEV psad_u32 psad_smem_u32(const void* p) { return (psad_u32)__cvta_generic_to_shared(p); }

// ---- mbarrier (all addresses are 32-bit shared-window addresses: no generic->shared conversion in the hot loop) ----
PSAD_DEV void psad_mbar_init(psad_u32 bar, psad_u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
PSAD_DEV void psad_fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
PSAD_DEV void psad_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
PSAD_DEV void psad_mbar_arrive_expect_tx(psad_u32 bar, psad_u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
PSAD_DEV void psad_mbar_arrive(psad_u32 bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
PSAD_DEV void psad_mbar_wait(psad_u32 bar, psad_u32 parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "PSAD_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra PSAD_DONE;\n"
      "bra PSAD_WAIT;\n"
      "PSAD_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// ---- cross-GPU handshake for peer halos: wait until a counter in a NEIGHBOURING GPU's memory (mapped through CUDA IPC,
// read over NVLink) has reached `expect`.  Acquire at system scope orders the peer data written before the counter; the
// proxy fence orders the TMA (async proxy) loads that follow behind this generic-proxy load.  Bounded: after ~2 s the CTA
// raises *error and carries on (wrong halo values, but no hung GPU).
PSAD_DEV void psad_wait_peer(const unsigned* flag, unsigned expect, unsigned* error) {
  unsigned v;
  unsigned long long t0 = 0, t1;
  for (unsigned spins = 0;; ++spins) {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if ((int)(v - expect) >= 0) break;
    if ((spins & 1023u) == 1023u) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t0 == 0) t0 = t1;
      else if (t1 - t0 > 2000000000ull) { if (error) atomicExch(error, 1u); break; }
    }
  }
  asm volatile("fence.proxy.async.global;" ::: "memory");
}

// ---- TMA tile loads (global -> shared, completion on an mbarrier; out-of-bounds elements are zero-filled) ---
PSAD_DEV void psad_tma_load_2d(psad_u32 smem_dst, const PsadTensorMap* tmap, psad_u32 bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"((psad_u64)tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
PSAD_DEV void psad_tma_load_3d(psad_u32 smem_dst, const PsadTensorMap* tmap, psad_u32 bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"((psad_u64)tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
PSAD_DEV void psad_tma_prefetch_desc(const PsadTensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((psad_u64)tmap) : "memory");
}

// ---- 16-byte vectors of the field element type -----------------------------------------------------------------
template <typename T> struct PsadVec;
template <> struct PsadVec<float> {
  typedef float4 type;
  static constexpr int N = 4;
  PSAD_DEV static void unpack(const float4& v, float* e) { e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w; }
  PSAD_DEV static float4 pack(const float* e) { return make_float4(e[0], e[1], e[2], e[3]); }
};
template <> struct PsadVec<double> {
  typedef double2 type;
  static constexpr int N = 2;
  PSAD_DEV static void unpack(const double2& v, double* e) { e[0] = v.x; e[1] = v.y; }
  PSAD_DEV static double2 pack(const double* e) { return make_double2(e[0], e[1]); }
};

// shared -> registers, one aligned 16-byte vector (LDS.128)
template <typename T> PSAD_DEV void psad_lds_vec(const T* p, T* e) {
  typename PsadVec<T>::type v = *reinterpret_cast<const typename PsadVec<T>::type*>(p);
  PsadVec<T>::unpack(v, e);
}
// shared -> registers, the two aligned 16-byte vectors of a 32-byte strip (8-byte elements, 4 cells per lane), free of bank
// conflicts.  A warp-wide LDS.128 is served one quarter warp (8 lanes x 16 bytes = one 128-byte wavefront) at a time; with a
// lane pitch of 32 bytes the eight lanes of a quarter warp touch only the even 16-byte columns, twice each: every LDS.128 is
// a 2-way bank conflict (ncu, 27-point fp64 kernel of round 1: 36 % of all shared wavefronts).  Here lanes 0-3 of every
// quarter warp fetch their first vector while lanes 4-7 fetch their second one, then the other way round: each
// instruction covers all 32 banks exactly once.  The price is one select per 32-bit register to put the halves back in order.
// `hi` = (lane >> 2) & 1.
template <typename T> PSAD_DEV void psad_lds_pair(const T* p, int hi, T* e) {
  typedef typename PsadVec<T>::type V;
  constexpr int N = PsadVec<T>::N;
  const V a = *reinterpret_cast<const V*>(p + (hi ? N : 0));
  const V b = *reinterpret_cast<const V*>(p + (hi ? 0 : N));
  T ea[N], eb[N];
  PsadVec<T>::unpack(a, ea);
  PsadVec<T>::unpack(b, eb);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    e[i] = hi ? eb[i] : ea[i];
    e[N + i] = hi ? ea[i] : eb[i];
  }
}
// registers -> shared, one aligned 16-byte vector (STS.128)
template <typename T> PSAD_DEV void psad_sts_vec(T* p, const T* e) {
  *reinterpret_cast<typename PsadVec<T>::type*>(p) = PsadVec<T>::pack(e);
}
// Barrier over the consumer warps only (named barrier 1; the producer warp never joins it).  Orders the shared-memory
// writes of the participating threads before the reads that follow it.
PSAD_DEV void psad_consumer_barrier(int n_threads) { asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory"); }
// registers -> global, one aligned 16-byte streaming store (STG.128, evict-first: outputs are not re-read)
#ifndef PSAD_STORE_MODE
#define PSAD_STORE_MODE 1   // 0: default caching, 1: streaming / evict-first (.cs), 2: write-through (.wt)
#endif
template <typename T> PSAD_DEV void psad_stg_vec(T* p, const T* e) {
  typedef typename PsadVec<T>::type V;
#if PSAD_STORE_MODE == 0
  *reinterpret_cast<V*>(p) = PsadVec<T>::pack(e);
#elif PSAD_STORE_MODE == 2
  __stwt(reinterpret_cast<V*>(p), PsadVec<T>::pack(e));
#else
  __stcs(reinterpret_cast<V*>(p), PsadVec<T>::pack(e));
#endif
}

// ---- reciprocal square root.  float: the hardware approximation (rsqrt.approx.ftz.f32, max relative error 2^-22.4 — the
// accuracy class of CUDA's rsqrtf(), without its subnormal pre-scaling branches; subnormal arguments are treated as zero).
// -DPSAD_RSQRT_NEWTON=1 adds one Newton-Raphson step, y <- y + y/2 * (1 - x y^2) (two FMAs, result within 1 ulp).  Measured
// on B200 for the TV-denoising gradient (scripts/c5_accuracy.py, profiles/r2_c5_accuracy.md): the error against the fp64
// oracle does not change (adjoint 2.0e-7 .. 4.2e-7 norm-wise either way: it is set by the other ~60 fp32 operations per
// cell), while the issue-bound adjoint kernel gets 15 % slower (0.72 -> 0.83 ms) — so it is not the default.
// double: CUDA's rsqrt() (1 ulp).
PSAD_DEV float psad_rsqrt(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
#ifdef PSAD_RSQRT_NEWTON
  const float e = __fmaf_rn(-x * r, r, 1.0f);          // 1 - x y^2
  r = __fmaf_rn(0.5f * r, e, r);
#endif
  return r;
}
PSAD_DEV double psad_rsqrt(double x) { return rsqrt(x); }

// ---- small integer powers by repeated multiplication (fixed association: ((x*x)*x)*...)
template <int N, typename T> PSAD_DEV T psad_ipow(T x) {
  T r = x;
#pragma unroll
  for (int i = 1; i < N; ++i) r *= x;
  return r;
}

// ---- warp halo exchange: element of the lane to the left / right -----------------------------------------------
template <typename T> PSAD_DEV T psad_from_left(T v) { return __shfl_up_sync(0xffffffffu, v, 1); }
template <typename T> PSAD_DEV T psad_from_right(T v) { return __shfl_down_sync(0xffffffffu, v, 1); }

#endif
