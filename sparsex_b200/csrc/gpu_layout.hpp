// GPU-side layout derived from the CSX byte stream (product code).
//
// The ctl stream and the values array go to HBM verbatim.  What the reference
// gets from sequential execution on one core per partition
// (src/templates/csx_spmv_tmpl.c:66-101) the GPU gets from two side tables
// built once at tune time by decoding ctl on the host:
//
//   * segment table — one entry per 32 consecutive rows ("warp segment"):
//     byte offset of the first unit that starts in the segment, index of its
//     first value, and the row it belongs to.  A warp enters the ctl stream
//     there and decodes unit heads (flags, size, varints) and delta bodies in
//     registers; it executes the units that live in one row (delta8/16/32/64,
//     horizontal) with one lane per element.
//
//   * cross-row unit table (XDT) — vertical, diagonal, anti-diagonal and block
//     units update several rows.  Each such unit gets a 16-byte descriptor
//     (value offset, start row, start column, kind/size) that is listed under
//     every row tile (256 or 1024 rows) it touches, including tiles after the one it starts
//     in ("carry-in").  The thread that owns a row gathers its contributions
//     from the descriptors of its tile: conflict free, no atomics, y written
//     once.  For CSX-Sym the transposed image of every cross-row unit is
//     listed under the tiles of its *columns*, which makes the symmetric
//     update of those units a gather as well.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "csx_host.hpp"

namespace spxb {

constexpr int SEG_ROWS = 32;     // rows per warp segment
constexpr int CTA_THREADS = 256; // threads per CTA; a tile has CTA_THREADS * rpt rows
constexpr int CTL_PAD = 32;      // readable bytes past the end of ctl

// unit kinds as the kernels see them
enum Kind : uint32_t {
  K_DELTA8 = 0, K_DELTA16 = 1, K_DELTA32 = 2, K_DELTA64 = 3, K_HORIZ = 4,   // row-local
  K_VERT = 5, K_DIAG = 6, K_ADIAG = 7, K_BROW = 8, K_BCOL = 9                // cross-row
};
inline bool kind_row_local(uint32_t k) { return k <= K_HORIZ; }

// 16-byte cross-row unit descriptor (device layout: uint4)
struct XDesc {
  uint32_t voff;   // index of the unit's first value in the device-wide values array
  int32_t row;     // global 0-based row of the unit's first element
  int32_t col;     // global 0-based column of the unit's first element
  uint32_t meta;   // [0:16) kind-table index  [16:24) size  [24:28) kind  [28] transposed
                   // [29:32) linear kinds: bit29 = (delta == 1); block kinds: align - 1
};
constexpr uint32_t XD_TRANSPOSED = 1u << 28;
constexpr uint32_t XD_DELTA1 = 1u << 29;

struct KindEntry { uint32_t kind_align; uint32_t delta; };  // kind | align << 8 ; stride or free block dim

struct PartLayout {
  int64_t nrows = 0, row_start = 0, nnz = 0, ctl_size = 0;
  uint64_t val_base = 0, ctl_base = 0;   // offsets into the device-wide arrays
  bool has_row_local = false, has_cross = false;
  bool xd_diag1_only = true;             // every descriptor of this partition's table is a direct diagonal unit of stride 1
  int64_t nseg = 0, ntiles = 0;
  int rpt = 1;                           // rows per thread (1 or 4): tile_rows = CTA_THREADS * rpt
  int64_t tile_rows() const { return (int64_t)CTA_THREADS * rpt; }
  KindEntry idtab[64];                   // ctl unit id -> kind
  std::vector<uint64_t> seg_ctl;         // nseg + 1 ; [63:56] row within segment, [55:0] ctl offset (partition relative)
  std::vector<uint32_t> seg_val;         // nseg + 1 ; value index (partition relative)
  std::vector<uint32_t> tile_xoff;       // ntiles + 1 ; bit 31 of entry t: tile t has row-local units
  std::vector<XDesc> xdesc;
};

struct DeviceLayout {
  std::vector<KindEntry> ktab;           // device-wide kind table referenced by XDesc.meta
  std::vector<PartLayout> parts;
  uint64_t total_values = 0, total_ctl = 0;
  bool symmetric = false, full_colind = false;
};

// Decodes every local partition's ctl stream and fills the tables.
std::string build_layout(const CsxMatrix &m, DeviceLayout &out);

}  // namespace spxb
