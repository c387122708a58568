// Device code of the chunk kernel (kernel 2 of the SpMV) and the structures shared by both kernels.
// Included by engine.cu (nvcc, sm_100a).  tests/emul/ compiles the same text for the host with the warp
// intrinsics replaced by a lock-step fibre emulation (CSXB_EMUL), so that the decode logic is also checked
// by the CPU test suite; that build is test infrastructure and is never part of the product library.
#pragma once
#include "gpu_layout.hpp"

using namespace spxb;

// ------------------------------------------------------------ device side --
struct PartDev {
  const uint8_t *ctl;          // this partition's ctl bytes (16-byte aligned, CTL_PAD readable bytes behind)
  const double *values;        // device-wide values array
  const ChunkEntry *chunks;    // chunk kernel entry points
  const uint16_t *uoffs;       // unit head offsets inside the chunks
  const uint32_t *tile_xoff;
  const uint4 *xdesc;
  const KindEntry *ktab;
  const double *dvalues;       // CSX-Sym: diagonal of the owned rows
  long long nrows, row_start;  // owned rows
  uint32_t val_base;
  uint32_t nchunks;
  int full_colind;
  int rpt;                     // rows per thread: a tile has CTA_THREADS * rpt rows
  uint32_t tile0, chunk0;      // first tile / chunk of this launch (a launch may cover a sub-range)
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


// ---- kernel 2: chunk kernel ------------------------------------------------------------------------
// One warp = one chunk of the ctl stream (gpu_layout.hpp).  Handles every unit that is not in the table:
// delta8/16/32/64 and horizontal units (delta_tmpl.c, horiz_tmpl.c), block-row / block-column units
// (block_row_tmpl.c, block_col_tmpl.c) and short vertical / diagonal / anti-diagonal units, plus their
// CSX-Sym transposed updates (*_sym_tmpl.c).  Runs after kernel 1 on the same stream and adds into y.
//
// Work distribution inside the warp: units are cut into slices of at most `slice` elements and every lane
// walks one slice per round (several elements per lane, so the warp-wide scans and reductions are paid once
// per slice instead of once per element).  ctl bytes and values are staged in shared memory with
// asynchronous copies; lanes read their slice's values from there, so the global loads stay coalesced.
constexpr int CHUNK_WARPS = 4;
constexpr int CHUNK_BATCH = 4;   // elements a lane has in flight per loop iteration
constexpr int CHUNK_VALS = CHUNK_MAX_ELEMS + CHUNK_MAX_ELEMS / 16 + 2;   // one pad double per 16: conflict-free strided reads
struct __align__(16) ChunkSmem {
  uint4 raw[(CHUNK_MAX_BYTES + 64) / 16];   // staged ctl bytes (16-byte aligned copy window + read-ahead slack)
  double vals[CHUNK_VALS];                  // staged values, element i at i + (i >> 4)
  uint4 units[CHUNK_MAX_UNITS];             // parsed unit heads
  uint16_t upos[CHUNK_MAX_UNITS];           // byte offset of every unit head
  uint16_t sstart[CHUNK_MAX_UNITS];         // first slice of every unit
  uint8_t smap[CHUNK_MAX_SLICES];           // slice -> unit
};

__device__ __forceinline__ uint64_t smem_varint(const uint8_t *c, uint32_t &pos) {  // CtlUtil.hpp:110-133
  uint64_t v = 0;
  unsigned shift = 0;
  for (;;) {
    uint32_t b = c[pos++];
    v |= (uint64_t)(b & 0x7f) << shift;
    if (!(b & 0x80)) break;
    shift += 7;
  }
  return v;
}
#ifndef CSXB_EMUL
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#else   // host emulation (tests/emul): synchronous copies
inline void cp_async16(void *smem, const void *gmem) { memcpy(smem, gmem, 16); }
inline void cp_async8(void *smem, const void *gmem) { memcpy(smem, gmem, 8); }
inline void cp_async_commit() {}
template <int N>
inline void cp_async_wait() {}
#endif

// little-endian fixed-width delta (1, 2, 4 or 8 bytes; the low 32 bits suffice) at byte position bp of the staged window
__device__ __forceinline__ uint32_t read_delta(const uint32_t *cw, uint32_t bp, uint32_t width) {
  const uint32_t w = __funnelshift_r(cw[bp >> 2], cw[(bp >> 2) + 1], (bp & 3) * 8);
  return width >= 4 ? w : (w & ((1u << (width * 8)) - 1u));
}

// sum of n consecutive deltas of `width` bytes starting at byte position bp (pass 1 of a delta slice)
__device__ __forceinline__ uint32_t sum_deltas(const uint32_t *cw, uint32_t bp, uint32_t n, uint32_t width) {
  uint32_t sum = 0;
  if (width == 8) {   // delta64 (columns below 2^32: the low words suffice); rare
    for (uint32_t t = 0; t < n; t++) sum += read_delta(cw, bp + t * 8, 8);
    return sum;
  }
  const uint32_t sh = (bp & 3) * 8;
  uint32_t wi = bp >> 2, prev = cw[wi];
  for (int rem = (int)(n * width); rem > 0; rem -= 4) {
    const uint32_t nxt = cw[++wi];
    uint32_t w = __funnelshift_r(prev, nxt, sh);
    prev = nxt;
    if (rem < 4) w &= (1u << (rem * 8)) - 1u;
    sum += width == 1 ? __dp4a(w, 0x01010101u, 0u) : (width == 2 ? (w & 0xffffu) + (w >> 16) : w);
  }
  return sum;
}
// four consecutive deltas starting at byte position bp
__device__ __forceinline__ void read_deltas4(const uint32_t *cw, uint32_t bp, uint32_t width, uint32_t R[4]) {
  const uint32_t sh = (bp & 3) * 8, wi = bp >> 2;
  if (width == 2) {
    const uint32_t a = cw[wi], b = cw[wi + 1], c = cw[wi + 2];
    const uint32_t w0 = __funnelshift_r(a, b, sh), w1 = __funnelshift_r(b, c, sh);
    R[0] = w0 & 0xffffu; R[1] = w0 >> 16; R[2] = w1 & 0xffffu; R[3] = w1 >> 16;
  } else if (width == 1) {
    const uint32_t w0 = __funnelshift_r(cw[wi], cw[wi + 1], sh);
    R[0] = w0 & 0xffu; R[1] = (w0 >> 8) & 0xffu; R[2] = (w0 >> 16) & 0xffu; R[3] = w0 >> 24;
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) R[i] = read_delta(cw, bp + i * width, width);
  }
}

// F::elem(slot, valid, value index within the chunk, partition-relative row, column) issues the loads of one element,
// F::use(slot, valid, row, column) returns its product (and performs per-element side effects),
// F::line(row, sum) adds a finished block / cross-row line, F::tail(pending, row, sum) combines the row-local
// slices of a round across the warp (warp-uniform call).
template <class F>
__device__ __forceinline__ void process_chunk(const PartDev &P, const uint32_t ch, ChunkSmem &S, int lane, F &f) {
  const uint4 *q = reinterpret_cast<const uint4 *>(P.chunks + ch);   // 32-byte entries
  const uint4 qa = __ldg(q), qb = __ldg(q + 1);
  const uint64_t ctl_off = (uint64_t)qa.x | ((uint64_t)qa.y << 32);
  const uint32_t val_off = qa.z, cursor0 = qa.w;
  const int row0 = (int)qb.x;
  const uint32_t nbytes = qb.y & 0xfff, ne = (qb.y >> 12) & 0x3ff, nu = qb.y >> 22;

  // 1. stage ctl bytes (16-byte copies of the aligned window that contains them), unit offsets and values
  const uint8_t *g = P.ctl + ctl_off;
  const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(g) & 15);
  const uint4 *src = reinterpret_cast<const uint4 *>(g - mis);
  for (uint32_t i = lane; i * 16 < mis + nbytes; i += 32) cp_async16(&S.raw[i], src + i);
  cp_async_commit();
  {
    const double *vsrc = P.values + P.val_base + val_off;
    for (uint32_t i = lane; i < ne; i += 32) cp_async8(&S.vals[i + (i >> 4)], vsrc + i);
    cp_async_commit();
  }
  for (uint32_t u = lane; u < nu; u += 32) S.upos[u] = __ldg(P.uoffs + qb.z + u);
  f.begin(val_off);
  cp_async_wait<1>();
  __syncwarp();
  const uint8_t *c = reinterpret_cast<const uint8_t *>(S.raw) + mis;
  const uint32_t *cw = reinterpret_cast<const uint32_t *>(S.raw);   // word view for unaligned delta reads

  // 2. unit records, one unit per lane: head decode, rows / element offsets / slice offsets by warp prefix sums
  uint32_t ne_seen = 0, nslices = 0;
  int row_base = row0;
  for (uint32_t u0 = 0; u0 < nu; u0 += 32) {
    const uint32_t u = u0 + lane;
    uint32_t size = 0, rowinc = 0, rec_x = 0, rec_y = 0, inc0 = 0, nsl = 0;
    if (u < nu) {
      uint32_t p = S.upos[u];
      const uint32_t flags = c[p];
      size = c[p + 1];
      p += 2;
      const bool nr = (flags & 0x80) != 0;
      if (nr) {  // csx_spmv_tmpl.c:86-91; the entry unit's row comes from the table
        uint32_t jmp = 1;
        if (flags & 0x40) jmp = (uint32_t)smem_varint(c, p);
        if (u != 0) rowinc = jmp;
      }
      uint32_t ucol;
      if (P.full_colind) { ucol = c[p] | (c[p + 1] << 8) | (c[p + 2] << 16) | ((uint32_t)c[p + 3] << 24); p += 4; }
      else ucol = (uint32_t)smem_varint(c, p);   // modulo 2^32 == modulo 2^64 truncated (negative ucol)
      const IdEntry ie = P.idtab[flags & 0x3f];
      const uint32_t kind = ie.kind_align & 0xff, align = (ie.kind_align >> 8) & 0xff;
      const bool reset = u == 0 || nr || P.full_colind;    // column cursor restarts at this unit
      inc0 = (u == 0 && !P.full_colind) ? cursor0 + ucol : ucol;
      nsl = unit_slices(kind, size, ie.delta, ie);
      rec_x = (size << 10) | (kind << 18) | ((uint32_t)reset << 22) | (align << 23);
      rec_y = p | (ie.delta << 12) | (ie.sl << 26);
    }
    // inclusive scans over the 32 units: element offsets, row numbers, slice offsets
    uint32_t es = size, rs = rowinc, ss = nsl;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t a = __shfl_up_sync(FULL, es, o), b = __shfl_up_sync(FULL, rs, o), d = __shfl_up_sync(FULL, ss, o);
      if (lane >= o) { es += a; rs += b; ss += d; }
    }
    if (u < nu) {
      const uint32_t estart = ne_seen + es - size, sstart = nslices + ss - nsl;
      S.units[u] = make_uint4(rec_x | estart, rec_y, inc0, (uint32_t)(row_base + (int)rs));
      S.sstart[u] = (uint16_t)sstart;
      for (uint32_t k = 0; k < nsl; k++) S.smap[sstart + k] = (uint8_t)u;
    }
    ne_seen += __shfl_sync(FULL, es, 31);
    nslices += __shfl_sync(FULL, ss, 31);
    row_base += (int)__shfl_sync(FULL, rs, 31);
  }
  cp_async_wait<0>();
  __syncwarp();

  // 3. rounds of 32 slices, one slice per lane
  uint32_t carry = 0;   // column cursor after the last slice of the previous round
  for (uint32_t s0 = 0; s0 < nslices; s0 += 32) {
    const uint32_t s = s0 + lane;
    // slice geometry: `total` elements in `nlines` lines of `llen`; element (l, e) has value index
    // vi0 + l*vsl + e*vse, row rbase + l*rl, column cbase + l*cl + e (substructures) or the running delta sum
    uint32_t total = 0, llen = 1, vi = 0, vse = 1, vwrap = 0, T = 0, body = 0, dw = 0, j = 0, inc0 = 0, kind = 0, cdelta = 0;
    int row = -1, rl = 0, coff = 0, cwrap = 0;
    bool reset = false;
    if (s < nslices) {
      const uint32_t u = S.smap[s];
      const uint4 rec = S.units[u];
      const uint32_t k = s - S.sstart[u];
      const uint32_t estart = rec.x & 0x3ff, size = (rec.x >> 10) & 0xff, align = (rec.x >> 23) & 0xf;
      const uint32_t delta = (rec.y >> 12) & 0x3fff, sl = rec.y >> 26;
      kind = (rec.x >> 18) & 0xf;
      body = rec.y & 0xfff;
      row = (int)rec.w;
      inc0 = k == 0 ? rec.z : 0u;
      reset = k == 0 && ((rec.x >> 22) & 1);
      T = inc0;
      const uint32_t j0 = k * sl;
      if (kind <= K_HORIZ) {          // row-local: one line; columns advance by the deltas (delta_tmpl.c, horiz_tmpl.c)
        total = min(sl, size - j0); llen = total; vi = estart + j0; j = j0;
        if (kind == K_HORIZ) { cdelta = delta; T += (total - (k == 0 ? 1u : 0u)) * delta; }
        else {
          dw = delta;
          const uint32_t first = k == 0 ? 1u : 0u;
          T += sum_deltas(cw, mis + body + (j0 + first - 1) * dw, total - first, dw);
        }
      } else if (kind <= K_ADIAG) {   // short vertical / diagonal / anti-diagonal run: one element per line
        total = min(sl, size - j0); llen = 1; vi = estart + j0;
        row += (int)(j0 * delta); rl = (int)delta;
        const int cstep = kind == K_DIAG ? (int)delta : (kind == K_ADIAG ? -(int)delta : 0);
        coff = (int)j0 * cstep; cwrap = cstep - 1;
      } else if (kind == K_BROW) {    // align rows x delta columns, values column-major (block_row_tmpl.c); slice = column range
        const uint32_t w = min(sl, delta - j0);
        total = w * align; llen = w; vi = estart + j0 * align; vse = align; vwrap = 1 - w * align; rl = 1;
        coff = (int)j0; cwrap = -(int)w;
      } else {                        // K_BCOL: delta rows x align columns, values row-major (block_col_tmpl.c); slice = row range
        const uint32_t h = min(sl, delta - j0);
        total = h * align; llen = align; vi = estart + j0 * align; rl = 1;
        row += (int)j0; cwrap = -(int)align;
      }
    }
    // column cursor before every slice: inclusive scan of the slice totals, restarted where the cursor restarts
    uint32_t incl = T;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t a = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += a;
    }
    const uint32_t rmask = __ballot_sync(FULL, reset) & (0xffffffffu >> (31 - lane));   // restarts at or before this lane
    const int seg = 31 - __clz(rmask);                                                    // lane of the last restart (-1: none)
    const uint32_t before = __shfl_sync(FULL, incl, seg > 0 ? seg - 1 : 0);
    const uint32_t base = seg < 0 ? 0u - carry : (seg > 0 ? before : 0u);                // incl - base = cursor after this slice
    uint32_t col = incl - base - T;                                                       // cursor before this slice
    carry = __shfl_sync(FULL, incl - base, 31);
    const bool rowlocal = kind <= K_HORIZ;
    // `col` is the column before the slice's first element: the cursor for row-local slices (their elements add
    // their deltas), first column - 1 for the others (their elements add 1 and wrap at the end of a line)
    if (!rowlocal) col += inc0 + (uint32_t)coff - 1u;

    // walk the slices, CHUNK_BATCH elements per lane in flight
    const uint32_t tmax = __reduce_max_sync(FULL, total);
    uint32_t e = 0;
    double acc = 0.0;
    const uint32_t bp0 = mis + body;
    // Rounds that hold only row-local slices (delta / horizontal units: all of an R-MAT stream, most of a mixed one)
    // take a loop without the line bookkeeping of block and cross-row slices.
    const bool all_rowlocal = __all_sync(FULL, rowlocal || total == 0);
    for (uint32_t t0 = 0; all_rowlocal && t0 < tmax; t0 += CHUNK_BATCH) {
      if (t0 >= total) continue;   // this lane's slice is done (no warp collective inside the loop)
      uint32_t D[CHUNK_BATCH];
      if (kind == K_HORIZ) {
#pragma unroll
        for (int i = 0; i < CHUNK_BATCH; i++) D[i] = cdelta;
      } else {   // element j of a delta unit adds body[j - 1]; the unit's first element adds ucol instead
        const uint32_t first = (j + t0 == 0) ? 1u : 0u;
        uint32_t R[CHUNK_BATCH];
        read_deltas4(cw, bp0 + (j + t0 + first - 1) * dw, dw, R);
        D[0] = first ? 0u : R[0];
#pragma unroll
        for (int i = 1; i < CHUNK_BATCH; i++) D[i] = first ? R[i - 1] : R[i];
      }
      if (j + t0 == 0) D[0] = inc0;
      const uint32_t left = total - t0;
      uint32_t cl[CHUNK_BATCH];
#pragma unroll
      for (int i = 0; i < CHUNK_BATCH; i++) {
        col += D[i];
        cl[i] = col;
        f.elem(i, (uint32_t)i < left, vi + i, row, col);
      }
      vi += CHUNK_BATCH;
#pragma unroll
      for (int i = 0; i < CHUNK_BATCH; i++) acc += f.use(i, (uint32_t)i < left, row, cl[i]);
    }
    for (uint32_t t0 = 0; !all_rowlocal && t0 < tmax; t0 += CHUNK_BATCH) {
      uint32_t D[CHUNK_BATCH];
#pragma unroll
      for (int i = 0; i < CHUNK_BATCH; i++) D[i] = 1;
      if (rowlocal && t0 < total) {
        if (kind == K_HORIZ) {
#pragma unroll
          for (int i = 0; i < CHUNK_BATCH; i++) D[i] = cdelta;
        } else {   // element j of a delta unit adds body[j - 1]; the unit's first element adds ucol instead
          const uint32_t first = (j + t0 == 0) ? 1u : 0u;
          uint32_t R[CHUNK_BATCH];
          read_deltas4(cw, bp0 + (j + t0 + first - 1) * dw, dw, R);
          D[0] = first ? 0u : R[0];
#pragma unroll
          for (int i = 1; i < CHUNK_BATCH; i++) D[i] = first ? R[i - 1] : R[i];
        }
        if (j + t0 == 0) D[0] = inc0;
      }
      bool ok[CHUNK_BATCH], eol[CHUNK_BATCH];
      int rw[CHUNK_BATCH];
      uint32_t cl[CHUNK_BATCH];
#pragma unroll
      for (int i = 0; i < CHUNK_BATCH; i++) {
        ok[i] = t0 + i < total;
        eol[i] = false;
        if (ok[i]) {
          col += D[i];
          rw[i] = row; cl[i] = col;
          f.elem(i, true, vi, row, col);
          vi += vse; e++;
          if (e == llen) { e = 0; eol[i] = !rowlocal; vi += vwrap; col += (uint32_t)cwrap; row += rl; }
        } else { rw[i] = row; cl[i] = col; f.elem(i, false, 0, 0, 0); }
      }
#pragma unroll
      for (int i = 0; i < CHUNK_BATCH; i++) {
        acc += f.use(i, ok[i], rw[i], cl[i]);
        if (eol[i]) { f.line(rw[i], acc); acc = 0.0; }
      }
    }
    f.tail(rowlocal && total > 0, row, acc);
  }
}

template <bool SYM>
struct SpmvChunkOp {
  const double *__restrict__ x;
  double *__restrict__ y;
  const double *vals;   // staged values of the chunk (shared memory)
  long long row_start;
  double alpha;
  int lane;
  double v[CHUNK_BATCH], xc[CHUNK_BATCH], xr[CHUNK_BATCH];
  __device__ __forceinline__ void begin(uint32_t) {}
  __device__ __forceinline__ void elem(int i, bool valid, uint32_t vi, int row, uint32_t col) {
    v[i] = 0.0; xc[i] = 0.0; xr[i] = 0.0;
    if (valid) {
      v[i] = vals[vi + (vi >> 4)];
      xc[i] = __ldg(x + col);
      if (SYM) xr[i] = __ldg(x + row_start + row);
    }
  }
  __device__ __forceinline__ double use(int i, bool valid, int, uint32_t col) {
    if (SYM && valid) atomicAdd(y + col, alpha * v[i] * xr[i]);   // transposed update (*_sym_tmpl.c)
    return v[i] * xc[i];
  }
  __device__ __forceinline__ void line(int row, double sum) { atomicAdd(y + row_start + row, alpha * sum); }
  __device__ __forceinline__ void tail(bool pending, int row, double p) {
    // row-local slices: combine the lanes that hold pieces of the same row, one red operation per row
    if (!__any_sync(FULL, pending)) return;   // a round of block / cross-row slices only
    const int key = pending ? row : -1 - lane;
    const int key0 = __shfl_sync(FULL, key, 0);
    if (__all_sync(FULL, key == key0)) {   // the whole round inside one row (long rows)
      const double sum = warp_sum(p);
      if (lane == 0) atomicAdd(y + row_start + row, alpha * sum);
      return;
    }
    const int prev = __shfl_up_sync(FULL, key, 1), next = __shfl_down_sync(FULL, key, 1);
    const uint32_t heads = __ballot_sync(FULL, lane == 0 || prev != key) & (0xffffffffu >> (31 - lane));
    const int seg = 31 - __clz(heads);     // first lane of this lane's run
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double tv = __shfl_up_sync(FULL, p, o);
      if (lane - o >= seg) p += tv;
    }
    if (pending && (lane == 31 || next != key)) atomicAdd(y + row_start + row, alpha * p);
  }
};

// Parity aid: the same traversals, storing the decoded coordinates per value (csxb_decode_coords).
struct DecodeChunkOp {
  int *rows, *cols;   // partition base applied
  long long row_start;
  uint32_t val_off;
  __device__ __forceinline__ void begin(uint32_t vo) { val_off = vo; }
  __device__ __forceinline__ void elem(int, bool valid, uint32_t vi, int row, uint32_t col) {
    if (valid) { rows[val_off + vi] = (int)(row_start + row); cols[val_off + vi] = (int)col; }
  }
  __device__ __forceinline__ double use(int, bool, int, uint32_t) { return 0.0; }
  __device__ __forceinline__ void line(int, double) {}
  __device__ __forceinline__ void tail(bool, int, double) {}
};
