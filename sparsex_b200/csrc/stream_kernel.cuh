// Device code of the stream kernel (kernel 2 of non-symmetric partitions): decodes the CSX ctl stream in registers
// and multiplies every unit that is not a table unit — delta8/16/32/64 and horizontal units (delta_tmpl.c,
// horiz_tmpl.c), block-row and block-column units (block_row_tmpl.c, block_col_tmpl.c).  gpu_layout.hpp describes
// the chunk table it walks.  Included by engine.cu (nvcc, sm_100a); tests/emul/ compiles the same text for the host on
// the fibre emulation of the warp intrinsics (CSXB_EMUL).
//
// One warp = one chunk.  (1) One unit head per lane is parsed straight from global memory (three aligned words and
// funnel shifts give the 12 bytes behind the head; varints are decoded branch-free for up to four bytes); warp
// prefix sums give every unit its row, its first value and its first task.  (2) Units are cut into lane tasks: up to
// SK_RL_E consecutive elements of a delta / horizontal unit, a column range of a block-row unit or a row range of a
// block-column unit.  32 tasks are walked at a time: a lane finds its unit with a ballot / population count, reads
// the unit record with three shuffles, loads its deltas once (aligned words + funnel shifts) and sums them; a
// segmented warp prefix sum over the task totals gives every task its column cursor; the lane then gathers x and
// multiplies — a block task loads x once per column and keeps one register accumulator per block row.  (3) Tasks
// that start in the same row sit in neighbouring lanes: a segmented shuffle reduction combines them and the last
// lane of a run adds the sums into the warp's shared-memory window of y rows.  (4) The chunk's own rows leave the
// window with coalesced plain stores, y = alpha*sum + beta*y; contributions to rows of other chunks go to the
// scratch array (csx_stream_fixup_kernel adds them in chunk order).  No atomics: the result is bit-reproducible.
//
// The kernel is instantiated per pattern set (R = register accumulators = rows of a block task, KM = mask of unit
// kinds): the counterpart of the reference's per-partition JIT (CsxJit.hpp:359-732), which emits one loop per unit
// kind of the partition's id_map.
#pragma once
#include "part_dev.cuh"

constexpr uint32_t SKM_DELTA = (1u << K_DELTA8) | (1u << K_DELTA16) | (1u << K_DELTA32);
constexpr uint32_t SKM_ROWLOCAL = SKM_DELTA | (1u << K_DELTA64) | (1u << K_HORIZ);
constexpr uint32_t SKM_BROW = 1u << K_BROW, SKM_BCOL = 1u << K_BCOL;
constexpr int SK_WARPS = 8;                         // warps (chunks) per CTA
constexpr int SK_WIN = SK_WROWS + 8;                // window doubles per warp (slack for idle accumulators)

// Source / target vectors of one launch.  `step` != nullptr: vectors by step parity of the multi-GPU exchange
// (gather_kernel.cuh: XchgDev), read from the device-resident step counter so that a captured graph can be replayed.
constexpr int SK_MAX_PEERS = 15;
struct SkIO {
  const double *x;
  double *y;
  const unsigned long long *step;
  double *vec[2];
  // multi-GPU exchange fused into the kernels that write the final rows of a partition without a gather pass: a row
  // that rank p reads (global rows [push_lo[p], push_hi[p])) is also stored into p's vector over NVLink as it is written
  int npush;
  long long push_lo[SK_MAX_PEERS], push_hi[SK_MAX_PEERS];
  double *push_vec[SK_MAX_PEERS][2];
};
// stores v to row g of the target vector of every peer that reads the row
__device__ __forceinline__ void sk_push(const SkIO &io, int par, long long g, double v) {
  for (int p = 0; p < io.npush; p++)
    if (g >= io.push_lo[p] && g < io.push_hi[p]) io.push_vec[p][par][g] = v;
}

// Bulk prefetch of [p, p + bytes) into L2 (cp.async.bulk.prefetch: one instruction, no registers, no LSU wavefronts),
// clipped to `end`.  The stream kernel's loads form a dependent chain (chunk entry -> unit offsets -> heads -> deltas ->
// x); with the stream already in L2 every link costs an L2 hit instead of a DRAM access.
__device__ __forceinline__ void sk_prefetch_l2(const void *p, uint32_t bytes, const void *end) {
#ifndef CSXB_EMUL
  const uintptr_t a = reinterpret_cast<uintptr_t>(p) & ~uintptr_t(15), e = reinterpret_cast<uintptr_t>(end) & ~uintptr_t(15);
  if (a + bytes > e) bytes = a < e ? (uint32_t)(e - a) : 0u;
  if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(bytes) : "memory");
#endif
}
constexpr uint32_t SK_PF_CTL = 2048, SK_PF_VAL = 4096;   // bytes prefetched for the first round of a far chunk
constexpr uint32_t SK_PF_CTL_NEXT = 1024, SK_PF_VAL_NEXT = 2048;   // ... and behind the current round
constexpr uint32_t SK_PF_DIST = 148 * 32;                 // chunks in flight on the device: prefetch that far ahead

// 32 bits at any byte address (ctl is readable CTL_PAD bytes past its end)
__device__ __forceinline__ uint32_t sk_ld32(const uint8_t *p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
  return __funnelshift_r(__ldg(w), __ldg(w + 1), (uint32_t)(a & 3) * 8);
}
// variable-length integer (CtlUtil.hpp:110-133) at the low end of `win` (`mem` = its address), modulo 2^32;
// `len` = bytes consumed
__device__ __forceinline__ uint32_t sk_varint(uint64_t win, const uint8_t *mem, uint32_t &len) {
  const uint32_t lo = (uint32_t)win;
  const uint32_t stop = ~lo & 0x80808080u;
  if (stop) {   // at most four bytes
    len = (uint32_t)__ffs((int)stop) >> 3;
    const uint32_t v = (lo & 0x7fu) | ((lo >> 1) & 0x3f80u) | ((lo >> 2) & 0x1fc000u) | ((lo >> 3) & 0xfe00000u);
    return v & ((1u << (7 * len)) - 1u);
  }
  // five to ten bytes (columns from 2^28, "negative" column jumps: SURVEY App. A): walk the bytes
  uint32_t v = 0, shift = 0;
  len = 0;
  for (;;) {
    const uint32_t b = mem[len++];
    if (shift < 32) v |= (b & 0x7fu) << shift;
    shift += 7;
    if (!(b & 0x80u)) break;
  }
  return v;
}

// Four consecutive deltas of `w` bytes starting at byte address p (low 32 bits of delta64).
template <uint32_t KM>
__device__ __forceinline__ void sk_deltas4(const uint8_t *p, uint32_t kind, uint32_t D[4]) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
  const uint32_t sh = (uint32_t)(a & 3) * 8;
  if ((KM >> K_DELTA8 & 1) && kind == K_DELTA8) {
    const uint32_t q = __funnelshift_r(__ldg(w), __ldg(w + 1), sh);
    D[0] = q & 0xffu; D[1] = (q >> 8) & 0xffu; D[2] = (q >> 16) & 0xffu; D[3] = q >> 24;
  } else if ((KM >> K_DELTA16 & 1) && kind == K_DELTA16) {
    const uint32_t a0 = __ldg(w), a1 = __ldg(w + 1), a2 = __ldg(w + 2);
    const uint32_t q0 = __funnelshift_r(a0, a1, sh), q1 = __funnelshift_r(a1, a2, sh);
    D[0] = q0 & 0xffffu; D[1] = q0 >> 16; D[2] = q1 & 0xffffu; D[3] = q1 >> 16;
  } else if ((KM >> K_DELTA32 & 1) && kind == K_DELTA32) {
    const uint32_t a0 = __ldg(w), a1 = __ldg(w + 1), a2 = __ldg(w + 2), a3 = __ldg(w + 3), a4 = __ldg(w + 4);
    D[0] = __funnelshift_r(a0, a1, sh); D[1] = __funnelshift_r(a1, a2, sh);
    D[2] = __funnelshift_r(a2, a3, sh); D[3] = __funnelshift_r(a3, a4, sh);
  } else if (KM >> K_DELTA64 & 1) {   // K_DELTA64: columns below 2^32, the low words suffice
#pragma unroll
    for (int i = 0; i < 4; i++) D[i] = sk_ld32(p + 8 * i);
  }
}

// Instantiations of the kernel, first match wins: X(R, KM, BC, BRC).  The shaped ones (BC / BRC > 0) need the
// partition's block tasks to have exactly that shape (PartLayout::sk_bc / sk_brc); the generic ones take any
// partition whose kinds are in KM and whose block tasks have at most R rows.
#define SK_INSTANCES(X)                         \
  X(2, SKM_ROWLOCAL | SKM_BCOL, 2, 0)           \
  X(3, SKM_ROWLOCAL | SKM_BCOL, 3, 0)           \
  X(4, SKM_ROWLOCAL | SKM_BCOL, 2, 0)           \
  X(2, SKM_ROWLOCAL | SKM_BROW, 0, 4)           \
  X(3, SKM_ROWLOCAL | SKM_BROW, 0, 4)           \
  X(1, SKM_DELTA, 0, 0)                         \
  X(1, SKM_ROWLOCAL, 0, 0)                      \
  X(2, SKM_ROWLOCAL | SKM_BCOL, 0, 0)           \
  X(4, SKM_ROWLOCAL | SKM_BCOL, 0, 0)           \
  X(4, SKM_ROWLOCAL | SKM_BROW | SKM_BCOL, 0, 0) \
  X(8, SKM_ROWLOCAL | SKM_BROW | SKM_BCOL, 0, 0)
// does instance (RR, KK, BCC, BRR) serve a partition with kinds km, r rows per block task and shapes bc / brc?
inline bool sk_instance_serves(int RR, uint32_t KK, int BCC, int BRR, uint32_t km, int r, int bc, int brc) {
  if (km & ~KK) return false;
  if (BCC == 0 && BRR == 0) return r <= RR;
  return r == RR && (!(KK & SKM_BCOL) || bc == BCC) && (!(KK & SKM_BROW) || brc == BRR);
}

// Per-chunk constants of a warp.
struct SkCtx {
  const uint8_t *cbase;      // ctl bytes of the chunk
  const double *values;      // values of the current round of units
  const double *x;
  double *sacc;              // the warp's window of y rows
  const uint4 *sid;          // unit id -> kind | align << 8, delta, lines per task, its reciprocal (shared memory)
  int lane;
  uint32_t lemask;           // lanes up to and including this one
  bool multib;               // a block-column unit of the chunk has several tasks
  long long grow0;           // global row of the window base
  int *drows, *dcols;        // DECODE: coordinates per value of the current round
};

// 32 lane tasks.  A, B, C = record of the task's unit (see sk_chunk), ie = its kind entry, k = index of the task inside
// the unit.  SINGLE: every unit of the round is one task (lane = unit, k = 0).
// BC > 0: every block-column task of the partition is exactly R rows x BC columns; BRC > 0: every block-row task is
// exactly R rows x BRC columns (compile-time shapes: no predicates, all loads issued together).
template <int R, uint32_t KM, int BC, int BRC, bool DECODE, bool SINGLE>
__device__ __forceinline__ void sk_window(const SkCtx &c, const uint32_t A, const uint32_t B, const uint32_t C, const uint4 ie,
                                          const uint32_t k, const uint32_t nvalid, uint32_t &carry) {
  const int lane = c.lane;
  const bool valid = (uint32_t)lane < nvalid;   // the first nvalid lanes hold tasks
  const uint32_t ues = (A >> 8) & 0x3ffu, usize = (A >> 18) & 0xffu;
  const uint32_t kind = ie.x & 0xffu, align = (ie.x >> 8) & 0xffu, delta = ie.y, tpar = ie.z;
  const uint32_t rowrel = B & 0xffu;
  const bool first = SINGLE || k == 0;
  // geometry of the task: n elements; row-local: column steps D[]; blocks: nl lines of the free dimension from l0
  uint32_t n = 0, adv = 0, key = 0x1000u + lane, vi = ues, l0 = 0, nl = 0;
  uint32_t D[SK_RL_E] = {0, 0, 0, 0};
  double vrl[SK_RL_E] = {0.0, 0.0, 0.0, 0.0};   // values of a row-local task
  if (valid) {
    key = rowrel;
    if (kind <= K_HORIZ) {
      const uint32_t j0 = SINGLE ? 0u : k * SK_RL_E;
      n = SINGLE ? usize : min((uint32_t)SK_RL_E, usize - j0);
      vi = ues + j0;
      if ((KM >> K_HORIZ & 1) && kind == K_HORIZ) {
#pragma unroll
        for (int i = 0; i < SK_RL_E; i++) D[i] = delta;
        if (first) D[0] = C;
      } else {   // element j of a delta unit adds body[j - 1]; the unit's first element adds ucol instead (delta_tmpl.c)
        const uint8_t *bp = c.cbase + ((B >> 8) & 0x1fffu);
        if (first) {
          uint32_t Rw[SK_RL_E] = {0, 0, 0, 0};
          if (usize > 1) sk_deltas4<KM>(bp, kind, Rw);
          D[0] = C; D[1] = Rw[0]; D[2] = Rw[1]; D[3] = Rw[2];
        } else {
          sk_deltas4<KM>(bp + (j0 - 1) * delta, kind, D);   // delta = bytes per step
        }
      }
#pragma unroll
      for (int i = 0; i < SK_RL_E; i++) adv += (uint32_t)i < n ? D[i] : 0u;
      if (!DECODE) {   // the values do not wait for the column cursor
#pragma unroll
        for (int i = 0; i < SK_RL_E; i++)
          if ((uint32_t)i < n) vrl[i] = __ldg(c.values + vi + i);
      }
    } else if ((KM & SKM_BROW) && kind == K_BROW) {   // align rows x delta columns, values column-major; task = column range
      if (BRC > 0) { l0 = k * BRC; nl = BRC; n = BRC * R; vi = ues + l0 * R; }
      else { l0 = k * tpar; nl = min(tpar, delta - l0); n = nl * align; vi = ues + l0 * align; }
      if (first) adv = C;
    } else if ((KM & SKM_BCOL) && kind == K_BCOL) {   // delta rows x align columns, values row-major; task = row range
      if (BC > 0) { l0 = k * R; nl = R; n = R * BC; vi = ues + l0 * BC; }
      else { l0 = k * tpar; nl = min(tpar, delta - l0); n = nl * align; vi = ues + l0 * align; }
      key = rowrel + l0;
      if (first) adv = C;
    } else {   // table unit: moves the cursor to its start column
      adv = C;
    }
  }
  const bool reset = valid && first && ((B >> 21) & 1u);
  // column cursor before every task: inclusive scan of the advances, restarted where the cursor restarts; as many
  // steps as the longest run of tasks behind a restart needs
  uint32_t incl = adv;
  {
    const uint32_t rmask = __ballot_sync(FULL, reset) & c.lemask;   // restarts at or before this lane
    const int rseg = 31 - __clz(rmask);                              // lane of the last restart (-1: none)
    const uint32_t maxd = __reduce_max_sync(FULL, valid ? (uint32_t)(lane - max(rseg, 0)) : 0u);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      if ((uint32_t)o > maxd) break;
      const uint32_t a = __shfl_up_sync(FULL, incl, o);
      if (lane - o >= rseg && lane >= o) incl += a;
    }
    if (rseg < 0) incl += carry;
    carry = __shfl_sync(FULL, incl, (int)nvalid - 1);   // idle lanes behind the last task did not take part
  }
  uint32_t col = incl - adv;   // cursor before this task (0 at a restart)

  // ---- elements ----
  double acc[R];
#pragma unroll
  for (int a = 0; a < R; a++) acc[a] = 0.0;
  if (kind <= K_HORIZ) {
    if (n) {
      uint32_t cl[SK_RL_E];
#pragma unroll
      for (int i = 0; i < SK_RL_E; i++) { col += D[i]; cl[i] = col; }
      if (DECODE) {
#pragma unroll
        for (int i = 0; i < SK_RL_E; i++)
          if ((uint32_t)i < n) { c.drows[vi + i] = (int)(c.grow0 + rowrel); c.dcols[vi + i] = (int)cl[i]; }
      } else {
        double xv[SK_RL_E];
#pragma unroll
        for (int i = 0; i < SK_RL_E; i++) {
          xv[i] = 0.0;
          if ((uint32_t)i < n) xv[i] = __ldg(c.x + cl[i]);
        }
#pragma unroll
        for (int i = 0; i < SK_RL_E; i++) acc[0] += vrl[i] * xv[i];
      }
    }
  } else if ((KM & SKM_BROW) && kind == K_BROW) {
    const uint32_t c0 = col + (first ? C : 0u) + l0;   // first column of the task
    if (DECODE) {
      for (uint32_t e = 0; e < n; e++) {
        c.drows[vi + e] = (int)(c.grow0 + rowrel + e % align);
        c.dcols[vi + e] = (int)(c0 + e / align);
      }
    } else if (BRC > 0) {
      if (valid) {
        double xv[BRC > 0 ? BRC : 1], v[BRC > 0 ? BRC * R : 1];
#pragma unroll
        for (int j = 0; j < BRC; j++) xv[j] = __ldg(c.x + c0 + j);   // x once per column (block_row_tmpl.c)
#pragma unroll
        for (int e = 0; e < BRC * R; e++) v[e] = __ldg(c.values + vi + e);
#pragma unroll
        for (int j = 0; j < BRC; j++)
#pragma unroll
          for (int a = 0; a < R; a++) acc[a] += v[j * R + a] * xv[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < SK_BLK_LINES; j++) {
        if ((uint32_t)j < nl) {
          const double xv = __ldg(c.x + c0 + j);
#pragma unroll
          for (int a = 0; a < R; a++)
            if ((uint32_t)a < align) acc[a] += __ldg(c.values + vi + j * align + a) * xv;
        }
      }
    }
  } else if ((KM & SKM_BCOL) && kind == K_BCOL) {
    const uint32_t c0 = col + (first ? C : 0u);
    if (DECODE) {
      for (uint32_t e = 0; e < n; e++) {
        c.drows[vi + e] = (int)(c.grow0 + key + e / align);
        c.dcols[vi + e] = (int)(c0 + e % align);
      }
    } else if (BC > 0) {
      if (valid) {
        double xv[BC > 0 ? BC : 1], v[BC > 0 ? BC * R : 1];
#pragma unroll
        for (int j = 0; j < BC; j++) xv[j] = __ldg(c.x + c0 + j);   // x once per column (block_col_tmpl.c)
#pragma unroll
        for (int e = 0; e < BC * R; e++) v[e] = __ldg(c.values + vi + e);
#pragma unroll
        for (int a = 0; a < R; a++)
#pragma unroll
          for (int j = 0; j < BC; j++) acc[a] += v[a * BC + j] * xv[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        if ((uint32_t)j < align) {
          const double xv = __ldg(c.x + c0 + j);
#pragma unroll
          for (int a = 0; a < R && a < SK_BLK_LINES; a++)
            if ((uint32_t)a < nl) acc[a] += __ldg(c.values + vi + a * align + j) * xv;
        }
      }
    }
  }
  if (DECODE) return;

  // ---- row sums: tasks that start in the same row are neighbours; the last lane of a run adds into the window ----
  const uint32_t kprev = __shfl_up_sync(FULL, key, 1);
  const uint32_t heads = __ballot_sync(FULL, lane == 0 || kprev != key);
  const int seg = 31 - __clz(heads & c.lemask);
  const bool last = valid && (lane == 31 || ((heads >> (lane + 1)) & 1u));
  const uint32_t maxrun = __reduce_max_sync(FULL, valid ? (uint32_t)(lane - seg) : 0u);
#pragma unroll
  for (int a = 0; a < R; a++) {
    double s = acc[a];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      if ((uint32_t)o > maxrun) break;
      const double tv = __shfl_up_sync(FULL, s, o);
      if (lane - o >= seg) s += tv;
    }
    bool add = last;
    if (R > 1 && c.multib) {
      // runs of different block-column tasks can end in the same row: the first lane of such a group adds the others
      const uint32_t target = last ? key + a : 0x2000u + lane;
      const uint32_t peers = __match_any_sync(FULL, target);
      uint32_t others = peers & ~(1u << lane);
      add = last && (peers & ((1u << lane) - 1u)) == 0;
      const double s0 = s;
      while (__any_sync(FULL, others != 0)) {
        const int src = others ? __ffs((int)others) - 1 : lane;
        const double ov = __shfl_sync(FULL, s0, src);
        if (others) { s += ov; others &= others - 1; }
      }
    }
    if (add) c.sacc[key + a] += s;
    __syncwarp();
  }
}

// One chunk = up to SK_MAX_ROUNDS rounds of 32 units.  DECODE: parity aid — store the decoded (row, column) of every
// value instead of multiplying.
template <int R, uint32_t KM, int BC, int BRC, bool DECODE>
__device__ __forceinline__ void sk_chunk(const PartDev &P, const uint32_t ch, double *sacc, const uint4 *sid, const int lane,
                                         const double *__restrict__ x, double *__restrict__ y, const double alpha, const double beta,
                                         const int overwrite, int *drows, int *dcols, const SkIO *io = nullptr, const int ypar = 0) {
  const uint4 *q = P.sk_chunks + 2 * (size_t)ch;
  const uint4 qa = __ldg(q), qb = __ldg(q + 1);
  const uint32_t cursor0 = qa.z;
  const int wrow = (int)qa.w;
  const uint32_t c6 = qb.z, c7 = qb.w;
  const uint32_t nunits = (c6 & 0xffu) + 1u, row0rel = (c6 >> 8) & 0xffu, f_lo = (c6 >> 16) & 0xffu;
  const bool headf = (c6 >> 30) & 1u;
  const uint32_t f_hi = c7 & 0x1ffu, t_hi = (c7 >> 9) & 0x1ffu;
  SkCtx c;
  c.cbase = P.ctl + ((uint64_t)qa.x | ((uint64_t)((c6 >> 24) & 0x3fu) << 32));
  c.values = P.values + P.val_base + qa.y;
  c.x = x; c.sacc = sacc; c.sid = sid; c.lane = lane;
  c.lemask = 0xffffffffu >> (31 - lane);
  c.multib = (c6 >> 31) & 1u;
  c.grow0 = P.row_start + wrow;
  c.drows = DECODE ? drows + qa.y : nullptr; c.dcols = DECODE ? dcols + qa.y : nullptr;
  if (!DECODE) {   // the window rows this chunk adds into start from zero
    const uint32_t zn = max(f_hi, t_hi);
    if (lane < (int)zn) sacc[lane] = 0.0;
    for (uint32_t i = lane + 32; i < zn; i += 32) sacc[i] = 0.0;
    __syncwarp();
  }

  // the chunk a later CTA starts with: its entry is loaded now and its first round is prefetched below
  uint4 pq = make_uint4(0, 0, 0, 0), pq2 = make_uint4(0, 0, 0, 0);
  const bool pf_far = !DECODE && lane == 0 && ch + SK_PF_DIST < P.sk_c1;
  if (pf_far) { pq = __ldg(q + 2 * (size_t)SK_PF_DIST); pq2 = __ldg(q + 2 * (size_t)SK_PF_DIST + 1); }
  uint32_t carry = 0;         // column cursor behind the last task walked so far
  uint32_t rowbase = row0rel; // window row of the last unit of the previous round
  const uint16_t *uo = P.sk_uoffs + qb.x;
  for (uint32_t u0 = 0, nu; u0 < nunits; u0 += nu) {
    // ---- 1. unit heads, one per lane; the round ends at the first unit whose offset carries the end mark ----------
    const uint32_t oraw = u0 + lane < nunits ? __ldg(uo + u0 + lane) : 0x8000u;
    nu = (uint32_t)__ffs((int)__ballot_sync(FULL, (oraw & 0x8000u) != 0));
    uint32_t size = 0, nt = 0, rowinc = 0, ucol = 0, body = 0, id = 0;
    bool ureset = false, rjmp = false;
    uint4 ie = make_uint4(0, 0, 1, 65536);
    if ((uint32_t)lane < nu) {
      const uint32_t off = oraw & 0x7fffu;
      const uint8_t *hp = c.cbase + off;
      const uintptr_t a = reinterpret_cast<uintptr_t>(hp);
      const uint32_t *wp = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
      const uint32_t sh = (uint32_t)(a & 3) * 8;
      const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3);
      const uint32_t b0 = __funnelshift_r(w0, w1, sh), b1 = __funnelshift_r(w1, w2, sh), b2 = __funnelshift_r(w2, w3, sh);
      const uint32_t flags = b0 & 0xffu;
      size = (b0 >> 8) & 0xffu;
      id = flags & 0x3fu;
      uint64_t win = (uint64_t)__funnelshift_r(b0, b1, 16) | ((uint64_t)__funnelshift_r(b1, b2, 16) << 32);   // bytes 2..9
      uint32_t hl = 2;
      const bool nr = (flags & 0x80u) != 0;
      const bool chunk_first = (u0 | (uint32_t)lane) == 0;
      if (nr) {   // csx_spmv_tmpl.c:86-91; the row of the chunk's first unit comes from the chunk entry
        uint32_t jmp = 1;
        if (flags & 0x40u) {
          uint32_t len;
          jmp = sk_varint(win, hp + hl, len);
          hl += len;
          win = len >= 8 ? 0 : win >> (8 * len);
          rjmp = true;
        }
        if (!chunk_first) rowinc = jmp;
      }
      if (P.full_colind) { ucol = hl == 2 ? (uint32_t)win : sk_ld32(hp + hl); hl += 4; }
      else {
        uint32_t len;
        ucol = hl <= 6 ? sk_varint(win, hp + hl, len) : sk_varint(sk_ld32(hp + hl) | ((uint64_t)sk_ld32(hp + hl + 4) << 32), hp + hl, len);
        hl += len;
      }
      ureset = chunk_first || nr || P.full_colind;   // the column cursor restarts at this unit
      if (chunk_first && !P.full_colind) ucol += cursor0;
      body = off + hl;
      ie = sid[id];
      nt = sk_unit_tasks(ie.x & 0xffu, size, ie.y, ie.z, ie.w);
    }
    // inclusive scans over the units: elements | tasks << 16, rows
    uint32_t et = size | (nt << 16), rs = rowinc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t a = __shfl_up_sync(FULL, et, o);
      if (lane >= o) et += a;
    }
    if (__any_sync(FULL, rjmp)) {
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t b = __shfl_up_sync(FULL, rs, o);
        if (lane >= o) rs += b;
      }
    } else {
      rs = (uint32_t)__popc(__ballot_sync(FULL, rowinc != 0) & c.lemask);
    }
    const uint32_t tot = __shfl_sync(FULL, et, 31);
    const uint32_t ntasks = tot >> 16;
    const uint32_t ts = (et >> 16) - nt, es = (et & 0xffffu) - size;
    // unit record: A = first task | first element << 8 | size << 18 | id << 26, B = row | body offset << 8 | restart << 21, C = ucol
    const uint32_t recA = ts | (es << 8) | (size << 18) | (id << 26);
    const uint32_t recB = (rowbase + rs) | (body << 8) | ((uint32_t)ureset << 21);
    rowbase += __shfl_sync(FULL, rs, 31);
    if (!DECODE && (uint32_t)lane == nu - 1) {   // what the next round reads: ctl and values behind this round
      sk_prefetch_l2(c.cbase + body + 64, SK_PF_CTL_NEXT, P.ctl_end);
      sk_prefetch_l2(c.values + es + size, SK_PF_VAL_NEXT, P.values_end);
    }
    if (pf_far && u0 == 0) {
      sk_prefetch_l2(P.ctl + ((uint64_t)pq.x | ((uint64_t)((pq2.z >> 24) & 0x3fu) << 32)), SK_PF_CTL, P.ctl_end);
      sk_prefetch_l2(P.values + P.val_base + pq.y, SK_PF_VAL, P.values_end);
      sk_prefetch_l2(P.sk_uoffs + pq2.x, 256, P.sk_uoffs + pq2.x + 128);
    }

    // ---- 2. tasks, 32 at a time -----------------------------------------------------------------------------------
    if (ntasks == nu) {   // every unit is one task: lane = unit
      sk_window<R, KM, BC, BRC, DECODE, true>(c, recA, recB, ucol, ie, 0u, nu, carry);
    } else {
      for (uint32_t t0 = 0; t0 < ntasks; t0 += 32) {
        const uint32_t t = t0 + lane;
        const uint32_t rel = ts - t0;
        const uint32_t bit = ((uint32_t)lane < nu && rel < 32u) ? 1u << rel : 0u;
        const uint32_t M = __reduce_or_sync(FULL, bit);
        const uint32_t nb = (uint32_t)__popc(__ballot_sync(FULL, (uint32_t)lane < nu && ts < t0));
        const int U = (int)(nb + (uint32_t)__popc(M & c.lemask)) - 1;
        const uint32_t A = __shfl_sync(FULL, recA, U), B = __shfl_sync(FULL, recB, U), C = __shfl_sync(FULL, ucol, U);
        sk_window<R, KM, BC, BRC, DECODE, false>(c, A, B, C, sid[A >> 26], t - (A & 0xffu), min(32u, ntasks - t0), carry);
      }
    }
    c.values += tot & 0xffffu;
    if (DECODE) { c.drows += tot & 0xffffu; c.dcols += tot & 0xffffu; }
  }
  if (DECODE) return;

  // ---- 3. the chunk's rows leave the window ---------------------------------------------------------------------
  {
    const bool push = io != nullptr && io->npush > 0;
    for (uint32_t i = f_lo + lane; i < f_hi; i += 32) {
      double *yp = y + (c.grow0 + i);
      double v = alpha * sacc[i];
      if (!overwrite) v += beta * *yp;
      *yp = v;
      if (push) sk_push(*io, ypar, c.grow0 + i, v);
    }
  }
  if (headf | (t_hi > f_hi)) {   // rows of other chunks: to the scratch array
    double *sc = P.sk_scratch + qb.y;
    if (headf && lane == 0) sc[0] = sacc[row0rel];
    for (uint32_t i = f_hi + lane; i < t_hi; i += 32) sc[(headf ? 1u : 0u) + (i - f_hi)] = sacc[i];
  }
  __syncwarp();
}
