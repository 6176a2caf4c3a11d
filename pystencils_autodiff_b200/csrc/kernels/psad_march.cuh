// psad_march.cuh — hand-written sm_100a pipeline for stencil kernels ("march" template).
//
// One persistent CTA per resident slot walks a static round-robin list of work items.  A work item is an (y, x)
// tile of TY x TX cells marched along z over a chunk of planes (3-D), or a column of TX cells marched over row
// tiles of TY rows (2-D).  For every step of the march one haloed box per input field is staged by TMA
// (cp.async.bulk.tensor, zero fill outside the array = the reference's 'zeros' boundary handling) into a ring of
// STAGES shared-memory slots; completion is signalled on one mbarrier per slot.  One elected thread issues the
// loads LA = STAGES - (HZL + HZH) - 1 steps ahead of the consumers, across work-item boundaries, so the queue of
// outstanding HBM requests never drains.  The per-step arithmetic — register windows along z, 128-bit shared
// loads, warp-shuffle x-halos, 128-bit streaming stores — is emitted per stencil by emit.py as psad_step().
//
// The including translation unit defines, before this header:
//   namespace cfg { NDIM, TX, TY, RY, SX, THREADS, MIN_CTAS, STAGES, HZL, HZH, NTMA, STAGE_BYTES,
//                   F_OFF[], F_BYTES[], F_ORGX[], F_ORGY[], TX_BYTES }
//   struct PsadCarry;  psad_step(...);  PSAD_KERNEL_NAME
#ifndef PSAD_MARCH_CUH
#define PSAD_MARCH_CUH

struct PsadTmaps {
  PsadTensorMap m[cfg::NTMA];
};

struct PsadItem {
  long long x0, y0;        // tile origin (y0 only meaningful for NDIM == 3)
  long long p_first, p_last;  // first / last plane (3-D) or row-tile (2-D) to stage
  long long z0;            // first output plane of the item
};

PSAD_DEV PsadItem psad_decode_item(const PsadArgs& A, long long item) {
  PsadItem it;
  const long long tx = item % A.tiles_x;
  long long rest = item / A.tiles_x;
  it.x0 = tx * cfg::TX;
  if (cfg::NDIM == 3) {
    const long long ty = rest % A.tiles_y;
    const long long c = rest / A.tiles_y;
    it.y0 = ty * cfg::TY;
    it.z0 = A.wr_lo[0] + c * A.chunk;
    long long z1 = it.z0 + A.chunk;
    if (z1 > A.wr_hi[0]) z1 = A.wr_hi[0];
    it.p_first = it.z0 - cfg::HZL;
    it.p_last = z1 - 1 + cfg::HZH;
  } else {
    const long long c = rest;
    it.y0 = 0;
    it.z0 = c * A.chunk;
    long long k1 = it.z0 + A.chunk;
    if (k1 > A.tiles_y) k1 = A.tiles_y;
    it.p_first = it.z0;
    it.p_last = k1 - 1;
  }
  return it;
}

// Producer side: stage plane / row-tile `p` of item `it` into ring slot `slot`.
PSAD_DEV void psad_issue(const PsadTmaps& TM, unsigned char* ring, psad_u64* full, int slot, const PsadItem& it,
                         long long p) {
  psad_mbar_arrive_expect_tx(&full[slot], cfg::TX_BYTES);
  unsigned char* base = ring + (long long)slot * cfg::STAGE_BYTES;
#pragma unroll
  for (int f = 0; f < cfg::NTMA; ++f) {
    if (cfg::NDIM == 3) {
      psad_tma_load_3d(base + cfg::F_OFF[f], &TM.m[f], &full[slot], (int)it.x0 + cfg::F_ORGX[f],
                       (int)it.y0 + cfg::F_ORGY[f], (int)p);
    } else {
      psad_tma_load_2d(base + cfg::F_OFF[f], &TM.m[f], &full[slot], (int)it.x0 + cfg::F_ORGX[f],
                       (int)(p * cfg::TY) + cfg::F_ORGY[f]);
    }
  }
}

extern "C" __global__ void __launch_bounds__(cfg::THREADS, cfg::MIN_CTAS)
PSAD_KERNEL_NAME(const __grid_constant__ PsadArgs A, const __grid_constant__ PsadTmaps TM) {
  extern __shared__ __align__(1024) unsigned char psad_smem[];
  unsigned char* ring = psad_smem;
  psad_u64* full = reinterpret_cast<psad_u64*>(psad_smem + (long long)cfg::STAGES * cfg::STAGE_BYTES);

  constexpr int D = cfg::HZL + cfg::HZH;
  constexpr int LA = cfg::STAGES - D - 1;
  static_assert(LA >= 1, "ring too small");

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int wy = tid >> 5;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < cfg::STAGES; ++s) psad_mbar_init(&full[s], 1);
    psad_fence_barrier_init();
#pragma unroll
    for (int f = 0; f < cfg::NTMA; ++f) psad_tma_prefetch_desc(&TM.m[f]);
  }
  __syncthreads();

  // ---- producer cursor (thread 0 only): next (item, plane) to stage and the ring slot it goes to
  long long pc_item = blockIdx.x;
  PsadItem pc_it;
  long long pc_p = 0;
  int pc_slot = 0;
  bool pc_valid = pc_item < A.n_items;
  if (tid == 0 && pc_valid) {
    pc_it = psad_decode_item(A, pc_item);
    pc_p = pc_it.p_first;
#pragma unroll 1
    for (int i = 0; i < LA && pc_valid; ++i) {
      psad_issue(TM, ring, full, pc_slot, pc_it, pc_p);
      pc_slot = (pc_slot + 1 == cfg::STAGES) ? 0 : pc_slot + 1;
      if (++pc_p > pc_it.p_last) {
        pc_item += gridDim.x;
        pc_valid = pc_item < A.n_items;
        if (pc_valid) {
          pc_it = psad_decode_item(A, pc_item);
          pc_p = pc_it.p_first;
        }
      }
    }
  }

  // ---- consumers
  int slot = 0;          // ring slot of the newest plane of this step
  psad_u32 parity = 0;   // phase parity of that slot
  PsadCarry R;
#pragma unroll 1
  for (long long item = blockIdx.x; item < A.n_items; item += gridDim.x) {
    const PsadItem it = psad_decode_item(A, item);
#pragma unroll 1
    for (long long p = it.p_first; p <= it.p_last; ++p) {
      __syncthreads();  // every thread is done with the previous step: its oldest slot may be refilled
      if (tid == 0 && pc_valid) {
        psad_issue(TM, ring, full, pc_slot, pc_it, pc_p);
        pc_slot = (pc_slot + 1 == cfg::STAGES) ? 0 : pc_slot + 1;
        if (++pc_p > pc_it.p_last) {
          pc_item += gridDim.x;
          pc_valid = pc_item < A.n_items;
          if (pc_valid) {
            pc_it = psad_decode_item(A, pc_item);
            pc_p = pc_it.p_first;
          }
        }
      }
      psad_mbar_wait(&full[slot], parity);
      const long long zo = p - cfg::HZH;  // output plane (3-D) / row tile (2-D) of this step
      if (cfg::NDIM == 3) {
        psad_step(A, ring, slot, R, lane, wy, zo >= it.z0, zo, it.y0, it.x0);
      } else {
        psad_step(A, ring, slot, R, lane, wy, true, 0, zo * cfg::TY, it.x0);
      }
      if (++slot == cfg::STAGES) {
        slot = 0;
        parity ^= 1;
      }
    }
  }
}

#endif
