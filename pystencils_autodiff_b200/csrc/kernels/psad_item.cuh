// psad_item.cuh — decoding of march work items (shared by the device template psad_march.cuh and by the host replay
// of the per-step bodies in tests/cpu_shim/).  Needs namespace cfg { NDIM, TX, TY, TXS, XORG, TYS, YORG, HZL, HZH } and
// PsadArgs.
//
// Items are numbered x-tile fastest, then y-tile, then z-chunk, so CTAs that run concurrently (static round robin
// over consecutive items) work on neighbouring tiles and share their halo columns / rows in L2.  Tiles are TX cells
// wide and start every TXS cells at XORG + k * TXS: single-step kernels have TXS = TX, XORG = 0; kernels that fuse
// several steps recompute the columns they cannot complete, so their tiles overlap (TXS < TX, XORG < 0; likewise TYS,
// YORG along y for the kernels that exchange intermediate rows between warps instead of recomputing them).
#ifndef PSAD_ITEM_CUH
#define PSAD_ITEM_CUH

struct PsadItem {
  int x0, y0;            // tile origin (y0 only meaningful for NDIM == 3)
  int p_first, p_last;   // first / last plane (3-D) or row-tile (2-D) to stage
  int z0;                // first output plane of the item
};

PSAD_DEV PsadItem psad_decode_item(const PsadArgs& A, long long item) {
  PsadItem it;
  const int tx = (int)(item % A.tiles_x);
  const long long rest = item / A.tiles_x;
  it.x0 = tx * cfg::TXS + cfg::XORG;
  if (cfg::NDIM == 3) {
    const int ty = (int)(rest % A.tiles_y);
    int c = (int)(rest / A.tiles_y) + A.chunk_rot;
    if (c >= A.n_chunks) c -= A.n_chunks;
    it.y0 = ty * cfg::TYS + cfg::YORG;
    it.z0 = (int)A.wr_lo[0] + c * A.chunk;
    int z1 = it.z0 + A.chunk;
    if (z1 > (int)A.wr_hi[0]) z1 = (int)A.wr_hi[0];
    it.p_first = it.z0 - cfg::HZL;
    it.p_last = z1 - 1 + cfg::HZH;
  } else {
    const int c = (int)rest;
    it.y0 = 0;
    it.z0 = c * A.chunk;
    int k1 = it.z0 + A.chunk;
    if (k1 > A.tiles_y) k1 = A.tiles_y;
    it.p_first = it.z0;
    it.p_last = k1 - 1;
  }
  return it;
}

#endif
