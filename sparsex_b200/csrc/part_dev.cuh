// Per-partition kernel argument and small helpers shared by the stream kernel and the gather kernel.
// Included by engine.cu (nvcc, sm_100a).  tests/emul/ compiles the same text for the host with the warp
// intrinsics replaced by a lock-step fibre emulation (CSXB_EMUL), so that the kernels' logic is also checked
// by the CPU test suite; that build is test infrastructure and is never part of the product library.
#pragma once
#include "gpu_layout.hpp"

using namespace spxb;

// ------------------------------------------------------------ device side --
struct BtDev {                 // one block table (gpu_layout.hpp: BlockTable)
  const uint32_t *ptr;
  const uint2 *ent;            // {first value of the sub-block (device wide), first column / row}
  long long j0;
  int G, nloop, sf, sl, image;
  uint32_t magic;              // ceil(2^32 / G): g / G == umulhi(g, magic) or one more
};
struct PartDev {
  const uint8_t *ctl;          // this partition's ctl bytes (16-byte aligned, CTL_PAD readable bytes behind)
  const double *values;        // device-wide values array
  const uint32_t *tile_xoff;
  const uint4 *xdesc;
  const KindEntry *ktab;
  const double *dvalues;       // CSX-Sym: diagonal of the owned rows
  long long nrows, row_start;  // owned rows
  uint32_t val_base;
  int full_colind;
  int rpt;                     // rows per thread: a tile has CTA_THREADS * rpt rows
  uint32_t tile0;              // first tile of this launch (a launch may cover a sub-range)
  BtDev bt[BT_MAX];            // block tables (gpu_layout.hpp: BlockTable)
  int nbt;
  // stream kernel (stream_kernel.cuh; non-symmetric partitions)
  const uint4 *sk_chunks;      // 32-byte chunk entries (SkEntry)
  const uint16_t *sk_uoffs;    // unit head offsets inside the chunks
  double *sk_scratch;          // sums a chunk contributes to rows of other chunks
  const int32_t *sk_fix_rows;  // rows with such contributions, their scratch slots sk_fix_idx[ptr[i] .. ptr[i+1])
  const uint32_t *sk_fix_ptr, *sk_fix_idx;
  const long long *sk_gaps;    // pairs [lo, hi) of partition-relative rows no chunk window covers
  uint32_t sk_c0, sk_c1;       // chunks of this launch
  uint32_t sk_f0, sk_f1;       // fix rows of this launch
  uint32_t sk_g0, sk_g1;       // gaps of this launch
  const uint8_t *ctl_end;      // end of the device-wide ctl / values arrays (bounds of the L2 prefetches)
  const double *values_end;
  IdEntry idtab[64];
};

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}


