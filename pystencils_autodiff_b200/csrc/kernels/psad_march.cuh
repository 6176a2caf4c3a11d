// psad_march.cuh — hand-written sm_100a pipeline for stencil kernels ("march" template).
//
// One persistent CTA per resident slot walks a static round-robin list of work items.  A work item is an (y, x)
// tile of TY x TX cells marched along z over a chunk of planes (3-D), or a column of TX cells marched over row
// tiles of TY rows (2-D).  For every step of the march one haloed box per input field is staged by TMA
// (cp.async.bulk.tensor; zero fill outside the array = the reference's 'zeros' boundary handling) into a ring of
// STAGES shared-memory slots.
//
// Warp specialisation: the last warp of the CTA is the producer — one elected lane streams the boxes of all the
// CTA's work items, back to back across item boundaries, and blocks only on the per-slot `empty` mbarrier.  The
// other warps are consumers: each waits on the slot's `full` mbarrier (TMA complete_tx), runs the per-step body
// psad_step() emitted for the stencil (register window along z, 128-bit shared loads, warp-shuffle x-halos,
// 128-bit streaming stores) and arrives on the `empty` barrier of the slot it will not read from shared memory
// again.  There is no CTA-wide barrier in the steady state, so consumer warps drift apart by up to the ring depth
// and the queue of outstanding HBM requests never drains.
//
// The including translation unit defines, before this header:
//   namespace cfg { NDIM, TX, TY, TXS, XORG, TYS, YORG, THREADS (consumer threads), MIN_CTAS, STAGES, HZL, HZH, JREL, NP, NTMA,
//                   STAGE_BYTES, TX_BYTES, F_OFF[], F_ORGX[], F_ORGY[] }
//   struct PsadCarry;  psad_item_begin(...);  psad_step(...);  PSAD_KERNEL_NAME
#ifndef PSAD_MARCH_CUH
#define PSAD_MARCH_CUH

#ifndef PSAD_PEER
#define PSAD_PEER 0     // 1: ghost planes along z come from the neighbouring GPUs' arrays (peer memory over NVLink)
#endif

#ifndef PSAD_PEER_PRODUCER
#define PSAD_PEER_PRODUCER 1        // 0: source chosen per plane inside one loop; 1: three plane loops per item (lower
#endif                              //    neighbour's planes, own planes, upper neighbour's planes), wait once per item
// Not inlined: measured on the 27-point fp64 kernel, the peer producer inlined into the kernel changes ptxas' register
// allocation / schedule of the CONSUMER loop (1.13 -> 1.25 ms with no neighbour at all); as a separate function the
// consumers' code is the plain kernel's again (1.137 ms).  For the same reason the kernel does not announce its own
// completion (a fence + count + flag store at its end cost 0.03-0.05 ms there, inlined or not): the caller writes the
// launch counter behind the kernel in stream order (psad_stream_write_u32).  profiles/r2_peer_halo.md.
#ifndef PSAD_PEER_PRODUCER_NOINLINE
#define PSAD_PEER_PRODUCER_NOINLINE 1
#endif
#ifndef PSAD_TMAPS_SETS
#define PSAD_TMAPS_SETS (PSAD_PEER ? 3 : 1)
#endif

struct PsadTmaps {
  PsadTensorMap m[cfg::NTMA * PSAD_TMAPS_SETS];   // [this GPU's arrays | lower neighbour's | upper neighbour's]
};

#include "psad_item.cuh"

#if PSAD_PEER && PSAD_PEER_PRODUCER
// The producer lane of a peer-halo kernel (3-D).  Per item: the planes below peer_lo_end come from the lower neighbour's
// array, the planes from peer_hi_begin on from the upper neighbour's, everything between from this GPU's — three loops
// with one TMA instruction each (constant descriptor address), the middle one being the loop of the plain kernel.  Before
// the first plane it takes from a neighbour the lane waits — once per kernel and side — for that neighbour's counter.
struct PsadRing { int slot; psad_u32 parity; };
PSAD_DEV void psad_stage_plane(const PsadTensorMap* maps, PsadRing& r, psad_u32 ring_s, psad_u32 full_s, psad_u32 empty_s,
                               int x0, int y0, int pz) {
  psad_mbar_wait(empty_s + 8 * r.slot, r.parity);
  psad_mbar_arrive_expect_tx(full_s + 8 * r.slot, cfg::TX_BYTES);
  const psad_u32 base = ring_s + r.slot * cfg::STAGE_BYTES;
#pragma unroll
  for (int f = 0; f < cfg::NTMA; ++f)
    psad_tma_load_3d(base + cfg::F_OFF[f], &maps[f], full_s + 8 * r.slot, x0 + cfg::F_ORGX[f], y0 + cfg::F_ORGY[f], pz);
  if (++r.slot == cfg::STAGES) {
    r.slot = 0;
    r.parity ^= 1;
  }
}

#if PSAD_PEER_PRODUCER_NOINLINE
__device__ __noinline__
#else
PSAD_DEV
#endif
void psad_produce_peer(const PsadArgs& A, const PsadTmaps& TM, psad_u32 ring_s, psad_u32 full_s, psad_u32 empty_s) {
#pragma unroll
  for (int f = 0; f < 3 * cfg::NTMA; ++f) psad_tma_prefetch_desc(&TM.m[f]);
  PsadRing r{0, 1u};   // the first pass over the ring does not wait (barrier phase -1 counts as complete)
  bool lo_ready = false, hi_ready = false;
  const int lo_end = A.peer_flag_lo != nullptr ? A.peer_lo_end : -(1 << 30);       // no neighbour: no plane qualifies
  const int hi_begin = A.peer_flag_hi != nullptr ? A.peer_hi_begin : (1 << 30);
#pragma unroll 1
  for (long long item = blockIdx.x; item < A.n_items; item += gridDim.x) {
    const PsadItem it = psad_decode_item(A, item);
    int p = it.p_first;
    if (p < lo_end) {
      if (!lo_ready) { psad_wait_peer(A.peer_flag_lo, A.peer_expect, A.peer_error); lo_ready = true; }
#pragma unroll 1
      for (; p < lo_end && p <= it.p_last; ++p)
        psad_stage_plane(&TM.m[cfg::NTMA], r, ring_s, full_s, empty_s, it.x0, it.y0, p + A.peer_lo_shift);
    }
    const int own_last = it.p_last < hi_begin ? it.p_last : hi_begin - 1;
#pragma unroll 1
    for (; p <= own_last; ++p) psad_stage_plane(&TM.m[0], r, ring_s, full_s, empty_s, it.x0, it.y0, p);
    if (p <= it.p_last) {
      if (!hi_ready) { psad_wait_peer(A.peer_flag_hi, A.peer_expect, A.peer_error); hi_ready = true; }
#pragma unroll 1
      for (; p <= it.p_last; ++p)
        psad_stage_plane(&TM.m[2 * cfg::NTMA], r, ring_s, full_s, empty_s, it.x0, it.y0, p - A.peer_hi_shift);
    }
  }
}
#endif

extern "C" __global__ void __launch_bounds__(cfg::THREADS + 32, cfg::MIN_CTAS)
PSAD_KERNEL_NAME(const __grid_constant__ PsadArgs A, const __grid_constant__ PsadTmaps TM) {
  extern __shared__ __align__(1024) unsigned char psad_smem[];
  unsigned char* ring = psad_smem;
  const psad_u32 ring_s = psad_smem_u32(psad_smem);                        // shared-window address of the ring
  const psad_u32 full_s = ring_s + cfg::STAGES * cfg::STAGE_BYTES;         // STAGES "full" barriers (8 bytes each)
  const psad_u32 empty_s = full_s + 8 * cfg::STAGES;                       // STAGES "empty" barriers

  constexpr int D = cfg::HZL + cfg::HZH;
  constexpr int NWARPS = cfg::THREADS / 32;        // consumer warps
  constexpr int REL_BACK = D - cfg::JREL;          // the slot released at a step holds the plane staged REL_BACK steps ago
  static_assert(cfg::STAGES >= REL_BACK + 2, "ring too small");

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < cfg::STAGES; ++s) {
      psad_mbar_init(full_s + 8 * s, 1);
      psad_mbar_init(empty_s + 8 * s, NWARPS);
    }
    psad_fence_barrier_init();
  }
  __syncthreads();

  if (warp == NWARPS) {
    // ================= producer warp =================
#if PSAD_PEER && PSAD_PEER_PRODUCER
    if (lane == 0) psad_produce_peer(A, TM, ring_s, full_s, empty_s);
    return;
#endif
    if (lane == 0) {
#pragma unroll
      for (int f = 0; f < cfg::NTMA * (PSAD_PEER ? 3 : 1); ++f) psad_tma_prefetch_desc(&TM.m[f]);
      int slot = 0;
      psad_u32 parity = 1;  // the first pass over the ring does not wait (barrier phase -1 counts as complete)
#if PSAD_PEER
      bool lo_ready = (A.peer_flag_lo == nullptr), hi_ready = (A.peer_flag_hi == nullptr);
#endif
#pragma unroll 1
      for (long long item = blockIdx.x; item < A.n_items; item += gridDim.x) {
        const PsadItem it = psad_decode_item(A, item);
#pragma unroll 1
        for (int p = it.p_first; p <= it.p_last; ++p) {
          int tm0 = 0, pz = p;   // which set of tensor maps, and the plane in that array
#if PSAD_PEER
          if (p < A.peer_lo_end && A.peer_flag_lo != nullptr) {
            if (!lo_ready) { psad_wait_peer(A.peer_flag_lo, A.peer_expect, A.peer_error); lo_ready = true; }
            tm0 = cfg::NTMA;
            pz = p + A.peer_lo_shift;
          } else if (p >= A.peer_hi_begin && A.peer_flag_hi != nullptr) {
            if (!hi_ready) { psad_wait_peer(A.peer_flag_hi, A.peer_expect, A.peer_error); hi_ready = true; }
            tm0 = 2 * cfg::NTMA;
            pz = p - A.peer_hi_shift;
          }
#endif
          psad_mbar_wait(empty_s + 8 * slot, parity);
          psad_mbar_arrive_expect_tx(full_s + 8 * slot, cfg::TX_BYTES);
          const psad_u32 base = ring_s + slot * cfg::STAGE_BYTES;
#pragma unroll
          for (int f = 0; f < cfg::NTMA; ++f) {
            if (cfg::NDIM == 3) {
              psad_tma_load_3d(base + cfg::F_OFF[f], &TM.m[tm0 + f], full_s + 8 * slot, it.x0 + cfg::F_ORGX[f],
                               it.y0 + cfg::F_ORGY[f], pz);
            } else {
              psad_tma_load_2d(base + cfg::F_OFF[f], &TM.m[f], full_s + 8 * slot, it.x0 + cfg::F_ORGX[f],
                               p * cfg::TY + cfg::F_ORGY[f]);
            }
          }
          if (++slot == cfg::STAGES) {
            slot = 0;
            parity ^= 1;
          }
        }
      }
    }
    return;
  }

  // ================= consumer warps =================
  int slot = 0;          // ring slot of the newest plane of this step
  psad_u32 parity = 0;   // phase parity of that slot's full barrier
  int warm = REL_BACK;   // steps until the first slot release (stream start only)
  int ph = 0;            // step index mod cfg::NP: which register-window slot receives the newest plane
  PsadCarry R;
#pragma unroll 1
  for (long long item = blockIdx.x; item < A.n_items; item += gridDim.x) {
    const PsadItem it = psad_decode_item(A, item);
    if (cfg::NDIM == 3) psad_item_begin(A, R, lane, warp, it.y0, it.x0);
#pragma unroll 1
    for (int p = it.p_first; p <= it.p_last; ++p) {
      psad_mbar_wait(full_s + 8 * slot, parity);
      const int zo = p - cfg::HZH;  // output plane (3-D) / row tile (2-D) of this step
      // the slot whose plane is read from shared memory for the last time in this step
      int rel = slot - REL_BACK;
      if (rel < 0) rel += cfg::STAGES;
      const psad_u32 rel_bar = (warm == 0) ? empty_s + 8 * rel : 0u;   // 0 = nothing to release yet
      if (warm > 0) --warm;
      if (cfg::NDIM == 3) {
        psad_step(A, ring, slot, R, lane, warp, zo >= it.z0, zo, it.y0, it.x0, rel_bar, ph);
      } else {
        psad_item_begin(A, R, lane, warp, zo * cfg::TY, it.x0);
        psad_step(A, ring, slot, R, lane, warp, true, 0, zo * cfg::TY, it.x0, rel_bar, ph);
      }
      if (++slot == cfg::STAGES) {
        slot = 0;
        parity ^= 1;
      }
      if (++ph == cfg::NP) ph = 0;
    }
  }
}

#endif
