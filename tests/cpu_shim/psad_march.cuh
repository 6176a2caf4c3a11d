// Host stand-in for csrc/kernels/psad_march.cuh: the consumer loop of the march template, replayed on the CPU.
// TEST INFRASTRUCTURE.  Work items are decoded by the product's own psad_item.cuh; the parameter block comes from
// the product's psad_plan_launch().  Each warp of the CTA is replayed by 32 lock-stepped OS threads; lane 0 plays
// the TMA producer (box copy with zero fill outside the array) right before the step that consumes the plane.
#ifndef PSAD_MARCH_CUH
#define PSAD_MARCH_CUH

#include <thread>
#include <vector>

#include "psad_item.cuh"

thread_local PsadEmuWarp* psad_emu_warp = nullptr;
thread_local int psad_emu_lane = 0;
thread_local std::barrier<>* psad_emu_cta = nullptr;

struct PsadEmuField {      // one staged (TMA) input field, plan order
  const void* ptr;
  long long stride[3];     // element strides (z, y, x)
  int esize, boxw, boxh;
};

static void psad_emu_stage(unsigned char* dst, const PsadArgs& A, const PsadEmuField& F, int x0, int y0, int z) {
  for (int by = 0; by < F.boxh; ++by)
    for (int bx = 0; bx < F.boxw; ++bx) {
      const long long gx = x0 + bx, gy = y0 + by;
      unsigned char* d = dst + ((size_t)by * F.boxw + bx) * F.esize;
      if (gx < 0 || gx >= A.shape[2] || gy < 0 || gy >= A.shape[1] || z < 0 || z >= A.shape[0])
        std::memset(d, 0, F.esize);
      else
        std::memcpy(d, (const unsigned char*)F.ptr + (z * F.stride[0] + gy * F.stride[1] + gx * F.stride[2]) * F.esize, F.esize);
    }
}

#ifndef PSAD_CTA_EXCHANGE
// ---- warps are independent: replay them one after the other, each with a private copy of the ring ----------------
extern "C" int psad_emulate(const PsadArgs* Ap, const PsadEmuField* tf, int n_tf, int n_ctas) {
  const PsadArgs& A = *Ap;
  if (n_tf != cfg::NTMA) return 1;
  constexpr int D = cfg::HZL + cfg::HZH;
  constexpr int NWARPS = cfg::THREADS / 32;
  constexpr int REL_BACK = D - cfg::JREL;
  static_assert(cfg::STAGES >= REL_BACK + 2, "ring too small");
  for (int cta = 0; cta < n_ctas; ++cta)
    for (int warp = 0; warp < NWARPS; ++warp) {
      std::vector<unsigned char> ring((size_t)cfg::STAGES * cfg::STAGE_BYTES, 0xff);   // 0xff..: NaN until staged
      PsadEmuWarp W;
      auto lane_main = [&](int lane) {
        psad_emu_warp = &W;
        psad_emu_lane = lane;
        int slot = 0, warm = REL_BACK, ph = 0;
        PsadCarry R;
        std::memset(&R, 0xff, sizeof(R));
        for (long long item = cta; item < A.n_items; item += n_ctas) {
          const PsadItem it = psad_decode_item(A, item);
          if (cfg::NDIM == 3) psad_item_begin(A, R, lane, warp, it.y0, it.x0);
          for (int p = it.p_first; p <= it.p_last; ++p) {
            if (lane == 0)
              for (int f = 0; f < cfg::NTMA; ++f) {
                unsigned char* dst = ring.data() + (size_t)slot * cfg::STAGE_BYTES + cfg::F_OFF[f];
                if (cfg::NDIM == 3) psad_emu_stage(dst, A, tf[f], it.x0 + cfg::F_ORGX[f], it.y0 + cfg::F_ORGY[f], p);
                else psad_emu_stage(dst, A, tf[f], it.x0 + cfg::F_ORGX[f], p * cfg::TY + cfg::F_ORGY[f], 0);
              }
            W.bar.arrive_and_wait();
            const int zo = p - cfg::HZH;
            const psad_u32 rel_bar = (warm == 0) ? 1u : 0u;
            if (warm > 0) --warm;
            if (cfg::NDIM == 3) {
              psad_step(A, ring.data(), slot, R, lane, warp, zo >= it.z0, zo, it.y0, it.x0, rel_bar, ph);
            } else {
              psad_item_begin(A, R, lane, warp, zo * cfg::TY, it.x0);
              psad_step(A, ring.data(), slot, R, lane, warp, true, 0, zo * cfg::TY, it.x0, rel_bar, ph);
            }
            W.bar.arrive_and_wait();
            if (lane == 0 && REL_BACK + 1 < cfg::STAGES) {
              // planes older than the released one are dead: poison them so a stale read shows up as NaN
              int dead = slot - REL_BACK - 1;
              if (dead < 0) dead += cfg::STAGES;
              if (dead != slot) std::memset(ring.data() + (size_t)dead * cfg::STAGE_BYTES, 0xff, cfg::STAGE_BYTES);
            }
            if (++slot == cfg::STAGES) slot = 0;
            if (++ph == cfg::NP) ph = 0;
          }
        }
      };
      std::vector<std::thread> lanes;
      for (int lane = 0; lane < 32; ++lane) lanes.emplace_back(lane_main, lane);
      for (auto& t : lanes) t.join();
    }
  return 0;
}
#else
// ---- kernels whose consumer warps talk to each other through shared memory (psad_consumer_barrier): all warps of the
// CTA run concurrently on one shared-memory image [ring | barriers | exchange buffers] of cfg::SMEM_BYTES; the plane is
// staged once per step by thread 0 between two CTA-wide barriers.  Inside a step the warps are ordered only by the
// kernel's own barrier, so a missing or misplaced one shows up as a race.
extern "C" int psad_emulate(const PsadArgs* Ap, const PsadEmuField* tf, int n_tf, int n_ctas) {
  const PsadArgs& A = *Ap;
  if (n_tf != cfg::NTMA) return 1;
  constexpr int D = cfg::HZL + cfg::HZH;
  constexpr int NWARPS = cfg::THREADS / 32;
  static_assert(D - cfg::JREL == 0 && cfg::NDIM == 3, "exchange kernels read only the newest plane of a 3-D march");
  for (int cta = 0; cta < n_ctas; ++cta) {
    std::vector<unsigned char> smem((size_t)cfg::SMEM_BYTES, 0xff);
    std::vector<PsadEmuWarp> W(NWARPS);
    std::barrier<> cta_bar(NWARPS * 32);
    auto thread_main = [&](int warp, int lane) {
      psad_emu_warp = &W[warp];
      psad_emu_lane = lane;
      psad_emu_cta = &cta_bar;
      int slot = 0, ph = 0;
      PsadCarry R;
      std::memset(&R, 0xff, sizeof(R));
      for (long long item = cta; item < A.n_items; item += n_ctas) {
        const PsadItem it = psad_decode_item(A, item);
        psad_item_begin(A, R, lane, warp, it.y0, it.x0);
        for (int p = it.p_first; p <= it.p_last; ++p) {
          if (warp == 0 && lane == 0)
            for (int f = 0; f < cfg::NTMA; ++f)
              psad_emu_stage(smem.data() + (size_t)slot * cfg::STAGE_BYTES + cfg::F_OFF[f], A, tf[f], it.x0 + cfg::F_ORGX[f],
                             it.y0 + cfg::F_ORGY[f], p);
          cta_bar.arrive_and_wait();
          const int zo = p - cfg::HZH;
          psad_step(A, smem.data(), slot, R, lane, warp, zo >= it.z0, zo, it.y0, it.x0, 1u, ph);
          cta_bar.arrive_and_wait();
          if (++slot == cfg::STAGES) slot = 0;
          if (++ph == cfg::NP) ph = 0;
        }
      }
    };
    std::vector<std::thread> threads;
    for (int warp = 0; warp < NWARPS; ++warp)
      for (int lane = 0; lane < 32; ++lane) threads.emplace_back(thread_main, warp, lane);
    for (auto& t : threads) t.join();
  }
  return 0;
}
#endif

#endif
