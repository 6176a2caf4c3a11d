// psad_args.h — kernel parameter block shared by the host runtime (psad_runtime.cpp) and the NVRTC-compiled
// kernels.  Plain C types only: this header is compiled by g++ and by NVRTC (no system includes there).
//
// All kernels see fields as 3-D (z, y, x) with x the contiguous axis; 1-D / 2-D fields get leading extents of 1.
#ifndef PSAD_ARGS_H
#define PSAD_ARGS_H

#define PSAD_MAX_FIELDS 12
#define PSAD_MAX_SCALARS 16

struct PsadArgs {
  void* ptr[PSAD_MAX_FIELDS];              // field base pointers, plan order (outputs first, then inputs)
  long long stride[PSAD_MAX_FIELDS][4];    // element strides (z, y, x, index)
  long long shape[3];                      // (Z, Y, X)
  long long it_lo[3], it_hi[3];            // cells evaluated with the stencil expression
  long long wr_lo[3], wr_hi[3];            // cells written (0 outside the iteration range)
  double scalar[PSAD_MAX_SCALARS];         // free scalar symbols, sorted by name
  long long n_items;                       // march: number of work items
  int tiles_x, tiles_y, n_chunks, chunk;   // march: work decomposition
  // Peer halos (kernels built with PSAD_PEER, 3-D): the ghost planes [0, peer_lo_end) and [peer_hi_begin, Z) are not read
  // from this array but from the NEIGHBOURING GPU's array through a second / third tensor map — plane p of the lower
  // ghost block is the lower neighbour's plane p + peer_lo_shift, of the upper block the upper neighbour's p - peer_hi_shift.
  // Before its first such load a CTA waits until the neighbour's "launches completed" counter has reached peer_expect.
  const unsigned* peer_flag_lo;            // NULL: no lower neighbour (global border: the local ghost planes, zeros)
  const unsigned* peer_flag_hi;
  unsigned* peer_error;                    // set to 1 by a CTA that gave up waiting (timeout), never cleared by kernels
  unsigned peer_expect;
  int peer_lo_end, peer_hi_begin, peer_lo_shift, peer_hi_shift;
  // Order of the z-chunks (3-D march): chunk c of the item numbering works on planes of chunk (c + chunk_rot) % n_chunks.
  // Peer-halo launches start in the middle of the slab (chunk_rot = n_chunks / 2): the chunks that touch ghost planes — the
  // ones that wait for the neighbour's previous launch and load over NVLink — are then neither the first work of every
  // CTA (a wait at the start of every kernel) nor its tail.
  int chunk_rot;
};

// 128-byte opaque CUtensorMap image (cuTensorMapEncodeTiled output), 64-byte aligned as the driver requires.
struct __attribute__((aligned(64))) PsadTensorMap {
  unsigned long long opaque[16];
};

#endif
