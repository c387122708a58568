// CUDA kernels (sm_100a) and the csxb_* C-ABI of the B200 CSX SpMV engine.
//
// Execution model (see gpu_layout.hpp for the tables), two kernels per SpMV on one stream:
//   1. csx_spmv_kernel   one CTA per tile of 256 / 1024 rows; every thread owns rows and gathers the
//                        contributions of the long cross-row units (vertical, diagonal, anti-diagonal)
//                        listed for its tile, then writes y = alpha*acc + beta*y once per row.
//                        No atomics, deterministic, streams values with coalesced loads.
//   2. csx_chunk_kernel  one warp per chunk of the ctl stream: unit heads and varints are parsed from a
//                        shared-memory copy of the chunk, delta bodies are turned into columns with a
//                        segmented warp prefix sum, block / short substructure elements get their
//                        coordinates from the unit geometry; rows are reduced with segmented shuffle
//                        reductions and added to y with fp64 red operations.
// SpMV is HBM-bound fp64 work: no tensor cores; every value and ctl byte is read once.
//
// Reference semantics reproduced: src/templates/csx_spmv_tmpl.c:66-101 and the
// nine unit templates (delta/horiz/vert/diag/rdiag/block_row/block_col), the
// symmetric variants (csx_sym_spmv_tmpl.c:60-106, *_sym_tmpl.c) and the
// y handling of CsxKernels.cpp:35-129 / CsxSpmv.cpp:28-86.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/csx_b200.h"
#include "csx_host.hpp"
#include "gpu_layout.hpp"
#include "chunk_kernel.cuh"

// Table units are vertical / diagonal / anti-diagonal runs (gpu_layout.hpp: goes_to_xdt).
// Op::add(device-wide value index, x index) for every element of descriptor d that contributes to `myrow`.
template <bool SYM, class Op>
__device__ __forceinline__ void gather_desc(const uint4 d, const KindEntry *__restrict__ ktab, int myrow, Op &op) {
  const uint32_t meta = d.w, kind = (meta >> 24) & 0xf, size = (meta >> 16) & 0xff;
  const int r = (int)d.y, c = (int)d.z;
  const uint32_t voff = d.x;
  const uint32_t delta = (meta & XD_DELTA1) ? 1u : __ldg(&ktab[meta & 0xffff].delta);
  if (!SYM || !(meta & XD_TRANSPOSED)) {  // vert_tmpl.c, diag_tmpl.c, rdiag_tmpl.c
    const int t = myrow - r;
    if (t < 0) return;
    const uint32_t k = (uint32_t)t / delta;
    if (k * delta != (uint32_t)t || k >= size) return;
    op.add(voff + k, kind == K_VERT ? c : (kind == K_DIAG ? c + t : c - t));
  } else if (kind == K_VERT) {
    // transposed image (CSX-Sym): element (r+a, c+b, v) adds v * x[r+a] to y[c+b];
    // vert_sym_tmpl.c: cur[x_indx] += sum v_k x[y_indx + k*delta]
    if (myrow != c) return;
    for (uint32_t k = 0; k < size; k++) op.add(voff + k, r + (int)(k * delta));
  } else {  // diag_sym_tmpl.c, rdiag_sym_tmpl.c
    const int u = kind == K_DIAG ? myrow - c : c - myrow;
    if (u < 0) return;
    const uint32_t k = (uint32_t)u / delta;
    if (k * delta != (uint32_t)u || k >= size) return;
    op.add(voff + k, r + u);
  }
}

// rows a descriptor can contribute to: [lo, hi] (global rows; columns for a transposed image)
template <bool SYM>
__device__ __forceinline__ bool desc_touches(const uint4 d, const KindEntry *__restrict__ ktab, int row_lo, int row_hi) {
  const uint32_t meta = d.w, kind = (meta >> 24) & 0xf, size = (meta >> 16) & 0xff;
  const int r = (int)d.y, c = (int)d.z;
  const uint32_t delta = (meta & XD_DELTA1) ? 1u : __ldg(&ktab[meta & 0xffff].delta);
  const int span = (int)((size - 1) * delta);
  int lo, hi;
  if (!SYM || !(meta & XD_TRANSPOSED)) { lo = r; hi = r + span; }
  else if (kind == K_VERT) { lo = hi = c; }
  else if (kind == K_DIAG) { lo = c; hi = c + span; }
  else { lo = c - span; hi = c; }
  return lo <= row_hi && hi >= row_lo;
}

// Linear kinds contribute at most one element per row: returns its value index and x index.
template <bool SYM>
__device__ __forceinline__ bool linear_probe(const uint4 d, const KindEntry *__restrict__ ktab, int myrow, uint32_t &vi, int &xi) {
  const uint32_t meta = d.w, kind = (meta >> 24) & 0xf, size = (meta >> 16) & 0xff;
  const int r = (int)d.y, c = (int)d.z;
  const bool tr = SYM && (meta & XD_TRANSPOSED);
  int t = !tr ? myrow - r : (kind == K_ADIAG ? c - myrow : myrow - c);
  if (t < 0) return false;
  uint32_t k = (uint32_t)t;
  if (!(meta & XD_DELTA1)) {
    uint32_t delta = __ldg(&ktab[meta & 0xffff].delta);
    k = (uint32_t)t / delta;
    if (k * delta != (uint32_t)t) return false;
  }
  if (k >= size) return false;
  vi = d.x + k;
  xi = !tr ? (kind == K_VERT ? c : (kind == K_DIAG ? c + t : c - t)) : r + t;
  return true;
}

struct SpmvGatherOp {
  const double *__restrict__ values;  // device-wide
  const double *__restrict__ x;
  double acc;
  __device__ __forceinline__ void add(uint32_t vi, int xi) { acc += __ldg(values + vi) * __ldg(x + xi); }
};

// ---- kernel 1: gather over the cross-row unit table + y initialisation --------------------------
// One CTA = one tile of CTA_THREADS * RPT rows; warp w owns RPT consecutive 32-row groups and lane L owns
// rows  tile0 + (w*RPT + k)*32 + L,  k < RPT  (RPT independent accumulators per thread).  Warps never
// synchronise with each other.  Every owned row of y is written exactly once here
// (y = alpha*acc + beta*y); the chunk kernel adds the remaining units afterwards.
// KSET specialises the gather for the set of unit kinds the partition's table holds:
//   KSET_ANY    vertical / diagonal / anti-diagonal units of any stride, CSX-Sym images included
//   KSET_DIAG1  only diagonal units of stride 1 (what the stencil matrices of the baseline configs encode to)
// The kernel is bandwidth-bound and latency-sensitive: compiled for 8 resident CTAs per SM (32 registers).
// VAR = 1 (4-rows-per-thread diagonal instantiation) issues the eight loads of a unit as one inline-PTX block so
// that all of them are in flight before the first FMA; ptxas otherwise interleaves loads and FMAs at 32 registers.
enum { KSET_ANY = 0, KSET_DIAG1 = 1 };
template <bool XD, bool SYM, int RPT, int KSET, int MINB = 8, int VAR = 0>
__global__ void __launch_bounds__(CTA_THREADS, MINB) csx_spmv_kernel(const __grid_constant__ PartDev P,
                                                                  const double *__restrict__ x,
                                                                  double *__restrict__ y, double alpha, double beta,
                                                                  int overwrite) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long tile = blockIdx.x;
  const long long lrow0 = ((tile * (CTA_THREADS / 32) + warp) * RPT) * 32;   // first row of this warp (partition relative)
  if (lrow0 >= P.nrows) return;
  double acc[RPT];
#pragma unroll
  for (int k = 0; k < RPT; k++) acc[k] = 0.0;

  if (XD) {
    const uint32_t b = __ldg(P.tile_xoff + tile), e = __ldg(P.tile_xoff + tile + 1);
    const int grow0 = (int)(P.row_start + lrow0);       // global rows of this warp: [grow0, grow0 + 32*RPT)
    const double *__restrict__ values = P.values;
    // each lane inspects one descriptor of the tile; the ones that reach this warp's rows are
    // then visited by the whole warp (warp-uniform loop over the ballot mask)
    for (uint32_t base = b; base < e; base += 32) {
      const uint32_t j = base + lane;
      bool hit = false;
      if (j < e) {
        const uint4 d = __ldg(P.xdesc + j);
        if (KSET == KSET_DIAG1) hit = (int)d.y <= grow0 + 32 * RPT - 1 && (int)d.y + (int)((d.w >> 16) & 0xff) > grow0;
        else hit = desc_touches<SYM>(d, P.ktab, grow0, grow0 + 32 * RPT - 1);
      }
      uint32_t mask = __ballot_sync(FULL, hit);
      while (mask) {
        const uint4 d = __ldg(P.xdesc + base + __ffs(mask) - 1);
        mask &= mask - 1;
        if (KSET == KSET_DIAG1) {  // diag_tmpl.c with delta 1: y[r+k] += x[c+k] * v[k]
          const int t0 = grow0 + lane - (int)d.y;
          const uint32_t size = (d.w >> 16) & 0xff;
          // one base pointer per stream; the RPT rows of this lane sit at fixed 256-byte strides from it
          const double *__restrict__ vp = values + ((long long)d.x + t0);
          const double *__restrict__ xp = x + ((long long)(int)d.z + t0);
          if (VAR == 1 && RPT == 4) {
            // the eight loads of a unit as one block, so that all of them are in flight before the first FMA
            double v0, v1, v2, v3, x0, x1, x2, x3;
            asm volatile(
                "{\n\t.reg .pred p0, p1, p2, p3;\n\t"
                "setp.lt.u32 p0, %8, %12;\n\tsetp.lt.u32 p1, %9, %12;\n\tsetp.lt.u32 p2, %10, %12;\n\tsetp.lt.u32 p3, %11, %12;\n\t"
                "mov.f64 %0, 0d0000000000000000;\n\tmov.f64 %1, 0d0000000000000000;\n\t"
                "mov.f64 %2, 0d0000000000000000;\n\tmov.f64 %3, 0d0000000000000000;\n\t"
                "mov.f64 %4, 0d0000000000000000;\n\tmov.f64 %5, 0d0000000000000000;\n\t"
                "mov.f64 %6, 0d0000000000000000;\n\tmov.f64 %7, 0d0000000000000000;\n\t"
                "@p0 ld.global.nc.f64 %0, [%13];\n\t@p0 ld.global.nc.f64 %4, [%14];\n\t"
                "@p1 ld.global.nc.f64 %1, [%13+256];\n\t@p1 ld.global.nc.f64 %5, [%14+256];\n\t"
                "@p2 ld.global.nc.f64 %2, [%13+512];\n\t@p2 ld.global.nc.f64 %6, [%14+512];\n\t"
                "@p3 ld.global.nc.f64 %3, [%13+768];\n\t@p3 ld.global.nc.f64 %7, [%14+768];\n\t}"
                : "=d"(v0), "=d"(v1), "=d"(v2), "=d"(v3), "=d"(x0), "=d"(x1), "=d"(x2), "=d"(x3)
                : "r"((uint32_t)t0), "r"((uint32_t)(t0 + 32)), "r"((uint32_t)(t0 + 64)), "r"((uint32_t)(t0 + 96)), "r"(size),
                  "l"(vp), "l"(xp));
            acc[0] += v0 * x0; acc[1 % RPT] += v1 * x1; acc[2 % RPT] += v2 * x2; acc[3 % RPT] += v3 * x3;
            continue;
          }
          double v[RPT], xv[RPT];
#pragma unroll
          for (int k = 0; k < RPT; k++) {
            v[k] = 0.0; xv[k] = 0.0;
            if ((uint32_t)(t0 + k * 32) < size) { v[k] = __ldg(vp + k * 32); xv[k] = __ldg(xp + k * 32); }
          }
#pragma unroll
          for (int k = 0; k < RPT; k++) acc[k] += v[k] * xv[k];
          continue;
        }
        // one element per row, except the transposed image of a vertical unit, which folds the whole
        // unit into the single row of its column
        if (!(SYM && ((d.w >> 24) & 0xf) == K_VERT && (d.w & XD_TRANSPOSED))) {
          double v[RPT], xv[RPT];  // issue all RPT value / x loads of this unit before using them
#pragma unroll
          for (int k = 0; k < RPT; k++) {
            uint32_t vi; int xi;
            v[k] = 0.0; xv[k] = 0.0;
            if (linear_probe<SYM>(d, P.ktab, grow0 + k * 32 + lane, vi, xi)) { v[k] = __ldg(values + vi); xv[k] = __ldg(x + xi); }
          }
#pragma unroll
          for (int k = 0; k < RPT; k++) acc[k] += v[k] * xv[k];
        } else {
#pragma unroll
          for (int k = 0; k < RPT; k++) {
            SpmvGatherOp op{values, x, 0.0};
            gather_desc<SYM>(d, P.ktab, grow0 + k * 32 + lane, op);
            acc[k] += op.acc;
          }
        }
      }
    }
  }

#pragma unroll
  for (int k = 0; k < RPT; k++) {
    const long long lrow = lrow0 + k * 32 + lane;
    if (lrow < P.nrows) {
      const long long g = P.row_start + lrow;
      double a = acc[k];
      if (SYM) a += __ldg(P.dvalues + lrow) * __ldg(x + g);   // diagonal (CsxJit.hpp:373-394 new-row hook)
      y[g] = overwrite ? alpha * a : alpha * a + beta * y[g];
    }
  }
}

template <bool SYM>
__global__ void __launch_bounds__(CHUNK_WARPS * 32, 9) csx_chunk_kernel(const __grid_constant__ PartDev P,
                                                                     const double *__restrict__ x,
                                                                     double *__restrict__ y, double alpha) {
  __shared__ ChunkSmem smem[CHUNK_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t ch = blockIdx.x * CHUNK_WARPS + warp;
  if (ch >= P.nchunks) return;
  SpmvChunkOp<SYM> op;
  op.x = x; op.y = y; op.vals = smem[warp].vals; op.row_start = P.row_start; op.alpha = alpha; op.lane = lane;
  process_chunk(P, ch, smem[warp], lane, op);
}

__global__ void __launch_bounds__(CHUNK_WARPS * 32) csx_decode_chunk_kernel(const __grid_constant__ PartDev P, int *rows,
                                                                            int *cols) {
  __shared__ ChunkSmem smem[CHUNK_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t ch = blockIdx.x * CHUNK_WARPS + warp;
  if (ch >= P.nchunks) return;
  DecodeChunkOp op{rows + P.val_base, cols + P.val_base, P.row_start, 0};
  process_chunk(P, ch, smem[warp], lane, op);
}
struct DecodeGatherOp {
  int *rows, *cols;   // device-wide
  int myrow;
  __device__ __forceinline__ void add(uint32_t vi, int col) { rows[vi] = myrow; cols[vi] = col; }
};
__global__ void __launch_bounds__(CTA_THREADS) csx_decode_gather_kernel(const __grid_constant__ PartDev P, int *rows, int *cols) {
  const long long tile = blockIdx.x;
  const uint32_t b = __ldg(P.tile_xoff + tile), e = __ldg(P.tile_xoff + tile + 1);
  for (int k = 0; k < P.rpt; k++) {
    const long long lrow = (tile * P.rpt + k) * CTA_THREADS + threadIdx.x;
    if (lrow >= P.nrows) return;
    DecodeGatherOp op{rows, cols, (int)(P.row_start + lrow)};
    for (uint32_t j = b; j < e; j++) {
      const uint4 d = __ldg(P.xdesc + j);
      if (d.w & XD_TRANSPOSED) continue;
      gather_desc<false>(d, P.ktab, op.myrow, op);
    }
  }
}

// --------------------------------------------------------------- host side --
static thread_local std::string g_last_error;
static int fail(const std::string &m) { g_last_error = m; return -1; }
#define CUDA_TRY(call)                                                                         \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_));                         \
  } while (0)

struct csxb_matrix {
  CsxMatrix host;
  DeviceLayout layout;
  bool uploaded = false;
  int device = -1;
  std::vector<void *> allocs;
  std::vector<PartDev> pdev;
  double *d_values = nullptr;
  double *d_x = nullptr, *d_y = nullptr;   // staging for csxb_spmv_host
  int64_t covered_rows_end = 0;
  int64_t sym_halo_lo = 0, sym_halo_hi = 0;   // CSX-Sym, partial device: rows of other devices this one adds into
  int64_t bytes[7] = {0, 0, 0, 0, 0, 0, 0};
  std::vector<std::string> logs;
  ~csxb_matrix() {
    if (!allocs.empty() || d_x || d_y) {
      int cur = -1;
      cudaGetDevice(&cur);
      if (device >= 0) cudaSetDevice(device);
      for (void *p : allocs) cudaFree(p);
      if (d_x) cudaFree(d_x);
      if (d_y) cudaFree(d_y);
      if (cur >= 0) cudaSetDevice(cur);
    }
  }
};

static std::string parse_options(const char *options, TuneOptions &o) {
  if (!options) return "";
  std::stringstream ss(options);
  std::string kv;
  while (std::getline(ss, kv, ';')) {
    if (kv.empty()) continue;
    size_t eq = kv.find('=');
    if (eq == std::string::npos) return "malformed option \"" + kv + "\"";
    std::string e = o.set(kv.substr(0, eq), kv.substr(eq + 1));
    if (!e.empty()) return e;
  }
  return "";
}
static void put_err(char *err, size_t n, const std::string &m) {
  g_last_error = m;
  if (err && n) { strncpy(err, m.c_str(), n - 1); err[n - 1] = 0; }
}

// C++ entry used by api.cpp for inputs already parsed into memory (spx_input_load_mmf)
csxb_matrix_t *csxb_tune_coo_internal(const CooHost &coo, const char *options, char *err, size_t errlen) {
  TuneOptions o;
  std::string e = parse_options(options, o);
  if (!e.empty()) { put_err(err, errlen, e); return nullptr; }
  csxb_matrix *m = new csxb_matrix;
  e = tune_coo(coo, o, 0, o.nr_threads, m->host);
  if (!e.empty()) { put_err(err, errlen, e); delete m; return nullptr; }
  return m;
}

extern "C" {

csxb_matrix_t *csxb_tune_csr(const int32_t *rowptr, const int32_t *colind, const double *values, int64_t nrows,
                             int64_t ncols, const char *options, int part_lo, int part_hi, char *err, size_t errlen) {
  TuneOptions o;
  std::string e = parse_options(options, o);
  if (e.empty() && (!rowptr || !colind || !values || nrows < 0 || ncols < 0)) e = "invalid CSR arguments";
  if (e.empty() && rowptr[0] != 0) e = "CSR arrays must be zero-based";
  if (!e.empty()) { put_err(err, errlen, e); return nullptr; }
  if (part_hi < 0) { part_lo = 0; part_hi = o.nr_threads; }
  csxb_matrix *m = new csxb_matrix;
  CsrView v{rowptr, colind, values, nrows, ncols};
  e = tune_csr(v, o, part_lo, part_hi, m->host);
  if (!e.empty()) { put_err(err, errlen, e); delete m; return nullptr; }
  return m;
}

csxb_matrix_t *csxb_tune_mmf(const char *path, const char *options, int part_lo, int part_hi, char *err, size_t errlen) {
  TuneOptions o;
  std::string e = parse_options(options, o);
  CooHost coo;
  if (e.empty()) e = path ? read_mmf(path, coo) : "invalid file name";
  if (!e.empty()) { put_err(err, errlen, e); return nullptr; }
  if (part_hi < 0) { part_lo = 0; part_hi = o.nr_threads; }
  csxb_matrix *m = new csxb_matrix;
  e = tune_coo(coo, o, part_lo, part_hi, m->host);
  if (!e.empty()) { put_err(err, errlen, e); delete m; return nullptr; }
  return m;
}

void csxb_destroy(csxb_matrix_t *m) { delete m; }

int64_t csxb_info(const csxb_matrix_t *m, int what) {
  switch (what) {
    case CSXB_NROWS: return m->host.nrows;
    case CSXB_NCOLS: return m->host.ncols;
    case CSXB_NNZ: return m->host.nnz;
    case CSXB_SYMMETRIC: return m->host.symmetric;
    case CSXB_NPARTS: return (int64_t)m->host.parts.size();
    case CSXB_NPARTS_TOTAL: return m->host.nparts_total;
    case CSXB_PART_LO: return m->host.part_lo;
    case CSXB_FULL_COLIND: return m->host.full_colind;
    case CSXB_SYM_HALO_LO: return m->sym_halo_lo;
    case CSXB_SYM_HALO_HI: return m->sym_halo_hi;
  }
  return -1;
}

int64_t csxb_part_info(const csxb_matrix_t *m, int part, int what) {
  if (part < 0 || (size_t)part >= m->host.parts.size()) return -1;
  const CsxPartition &p = m->host.parts[part];
  switch (what) {
    case CSXB_P_NNZ: return p.nnz;
    case CSXB_P_NROWS: return p.nrows;
    case CSXB_P_NCOLS: return p.ncols;
    case CSXB_P_ROW_START: return p.row_start;
    case CSXB_P_CTL_SIZE: return (int64_t)p.ctl.size();
    case CSXB_P_ROW_JUMPS: return p.row_jumps;
    case CSXB_P_ID_MAP_LEN: return (int64_t)p.id_map.size();
    case CSXB_P_MAP_LEN: return (int64_t)p.map_cpus.size();
    case CSXB_P_DVALUES_LEN: return (int64_t)p.dvalues.size();
    case CSXB_P_ROWS_INFO_LEN: return (int64_t)p.rows_info.size();
    case CSXB_P_SAMPLING_UNDEFINED: return p.sampling_undefined;
    case CSXB_P_COL_MIN: return p.col_min;
    case CSXB_P_COL_MAX: return p.col_max;
  }
  return -1;
}

int csxb_part_copy(const csxb_matrix_t *m, int part, int what, void *dst) {
  if (part < 0 || (size_t)part >= m->host.parts.size() || !dst) return fail("invalid argument");
  const CsxPartition &p = m->host.parts[part];
  switch (what) {
    case CSXB_A_VALUES:
      if (p.values.size() != (size_t)p.nnz) return fail("host values were released at upload");
      memcpy(dst, p.values.data(), p.values.size() * 8); break;
    case CSXB_A_CTL: memcpy(dst, p.ctl.data(), p.ctl.size()); break;
    case CSXB_A_ID_MAP: { int64_t *d = (int64_t *)dst; for (size_t i = 0; i < p.id_map.size(); i++) d[i] = p.id_map[i]; break; }
    case CSXB_A_ROWS_INFO: {
      struct R { int64_t rowptr, valptr; int32_t span, pad; } *d = (R *)dst;
      for (size_t i = 0; i < p.rows_info.size(); i++) d[i] = R{p.rows_info[i].rowptr, p.rows_info[i].valptr, p.rows_info[i].span, 0};
      break;
    }
    case CSXB_A_DVALUES: memcpy(dst, p.dvalues.data(), p.dvalues.size() * 8); break;
    case CSXB_A_MAP_CPUS: memcpy(dst, p.map_cpus.data(), p.map_cpus.size() * 4); break;
    case CSXB_A_MAP_POS: memcpy(dst, p.map_pos.data(), p.map_pos.size() * 4); break;
    default: return fail("unknown array");
  }
  return 0;
}

const char *csxb_part_log(const csxb_matrix_t *m, int part) {
  if (part < 0 || (size_t)part >= m->host.parts.size()) return "";
  return m->host.parts[part].encoding_log.c_str();
}

const char *csxb_last_error(void) { return g_last_error.c_str(); }

}  // extern "C"

template <class T>
static int dev_copy(csxb_matrix *m, const T *src, size_t n, T **dst, size_t extra_zero = 0) {
  void *p = nullptr;
  size_t bytes = (n + extra_zero) * sizeof(T);
  CUDA_TRY(cudaMalloc(&p, bytes ? bytes : 16));
  m->allocs.push_back(p);
  if (n) CUDA_TRY(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
  if (extra_zero) CUDA_TRY(cudaMemset((char *)p + n * sizeof(T), 0, extra_zero * sizeof(T)));
  *dst = (T *)p;
  return 0;
}

extern "C" {

int csxb_upload(csxb_matrix_t *m, int device, int free_host) {
  if (m->uploaded) return fail("matrix already uploaded");
  // CSX-Sym: a partition owns dvalues.size() rows (SparsePartitionSym::GetNrRows, SparsePartition.hpp:420-423)
  CsxMatrix &H = m->host;
  std::vector<int64_t> saved_nrows;
  for (auto &p : H.parts) { saved_nrows.push_back(p.nrows); if (H.symmetric) p.nrows = (int64_t)p.dvalues.size(); }
  std::string e = build_layout(H, m->layout);
  for (size_t i = 0; i < H.parts.size(); i++) H.parts[i].nrows = saved_nrows[i];
  if (!e.empty()) return fail("layout: " + e);
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail("invalid device");
  CUDA_TRY(cudaSetDevice(device));
  m->device = device;
  DeviceLayout &L = m->layout;
  // device-wide arrays
  void *dv = nullptr, *dc = nullptr;
  CUDA_TRY(cudaMalloc(&dv, std::max<uint64_t>(L.total_values, 1) * 8));
  m->allocs.push_back(dv);
  CUDA_TRY(cudaMalloc(&dc, std::max<uint64_t>(L.total_ctl, 16)));
  m->allocs.push_back(dc);
  CUDA_TRY(cudaMemset(dc, 0, std::max<uint64_t>(L.total_ctl, 16)));
  m->d_values = (double *)dv;
  KindEntry *d_ktab = nullptr;
  if (dev_copy(m, L.ktab.data(), L.ktab.size(), &d_ktab)) return -1;
  int64_t tables = (int64_t)L.ktab.size() * 8, nnz_stored = 0, ctl_bytes = 0, rows_owned = 0, launches = 0;
  m->pdev.resize(L.parts.size());
  for (size_t i = 0; i < L.parts.size(); i++) {
    PartLayout &pl = L.parts[i];
    CsxPartition &hp = H.parts[i];
    PartDev &P = m->pdev[i];
    memset(&P, 0, sizeof(P));
    if (hp.nnz) CUDA_TRY(cudaMemcpy(m->d_values + pl.val_base, hp.values.data(), (size_t)hp.nnz * 8, cudaMemcpyHostToDevice));
    if (!hp.ctl.empty()) CUDA_TRY(cudaMemcpy((uint8_t *)dc + pl.ctl_base, hp.ctl.data(), hp.ctl.size(), cudaMemcpyHostToDevice));
    P.ctl = (const uint8_t *)dc + pl.ctl_base;
    P.values = m->d_values;
    uint32_t *tx = nullptr; XDesc *xd = nullptr; ChunkEntry *ch = nullptr; uint16_t *uo = nullptr;
    if (dev_copy(m, pl.chunks.data(), pl.chunks.size(), &ch)) return -1;
    if (dev_copy(m, pl.uoffs.data(), pl.uoffs.size(), &uo)) return -1;
    P.uoffs = uo;
    if (dev_copy(m, pl.tile_xoff.data(), pl.tile_xoff.size(), &tx)) return -1;
    if (dev_copy(m, pl.xdesc.data(), pl.xdesc.size(), &xd)) return -1;
    P.chunks = ch; P.nchunks = (uint32_t)pl.chunks.size();
    P.tile_xoff = tx; P.xdesc = (const uint4 *)xd; P.ktab = d_ktab;
    if (H.symmetric) {
      double *dd = nullptr;
      if (dev_copy(m, hp.dvalues.data(), hp.dvalues.size(), &dd)) return -1;
      P.dvalues = dd;
      tables += (int64_t)hp.dvalues.size() * 8;
    }
    P.nrows = pl.nrows; P.row_start = pl.row_start; P.val_base = (uint32_t)pl.val_base;
    P.full_colind = L.full_colind;
    P.rpt = pl.rpt;
    memcpy(P.idtab, pl.idtab, sizeof(P.idtab));
    nnz_stored += hp.nnz; ctl_bytes += (int64_t)hp.ctl.size(); rows_owned += pl.nrows;
    tables += (int64_t)pl.tile_xoff.size() * 4 + (int64_t)pl.xdesc.size() * 16;
    tables += (int64_t)pl.chunks.size() * (int64_t)sizeof(ChunkEntry) + (int64_t)pl.uoffs.size() * 2;
    if (pl.nrows) launches += 1 + (pl.chunks.empty() ? 0 : 1);
    m->covered_rows_end = std::max<int64_t>(m->covered_rows_end, pl.row_start + pl.nrows);
    if (free_host) std::vector<double>().swap(hp.values);
  }
  if (H.symmetric && (int)H.parts.size() != H.nparts_total && !H.parts.empty()) {
    // transposed updates of the local lower triangle reach rows [col_min, first local row) of lower ranks
    int64_t first_row = H.parts.front().row_start, lo = first_row;
    for (auto &p : H.parts) if (p.col_max >= p.col_min) lo = std::min(lo, p.col_min);
    m->sym_halo_lo = lo; m->sym_halo_hi = first_row;
  }
  m->bytes[CSXB_B_VALUES] = nnz_stored * 8;
  m->bytes[CSXB_B_CTL] = ctl_bytes;
  m->bytes[CSXB_B_TABLES] = tables;
  // x: the column window the local partitions read (the whole vector when every partition is local)
  int64_t cmin = H.ncols, cmax = -1;
  for (auto &p : H.parts) if (p.col_max >= p.col_min) { cmin = std::min(cmin, p.col_min); cmax = std::max(cmax, p.col_max); }
  int64_t xcols = (int)H.parts.size() == H.nparts_total ? H.ncols : std::max<int64_t>(0, cmax - cmin + 1);
  if (H.symmetric) xcols = H.ncols;
  m->bytes[CSXB_B_X] = xcols * 8;
  m->bytes[CSXB_B_Y] = rows_owned * 8;
  m->bytes[CSXB_B_TOTAL] = nnz_stored * 8 + ctl_bytes + tables + xcols * 8 + rows_owned * 8;
  m->bytes[CSXB_B_LAUNCHES] = launches;
  m->uploaded = true;
  return 0;
}

int64_t csxb_traffic(const csxb_matrix_t *m, int what) {
  if (what < 0 || what > CSXB_B_LAUNCHES) return -1;
  return m->bytes[what];
}

}  // extern "C"

template <bool SYM, int RPT, int KSET>
static void launch_gather_k(const PartDev &P, const PartLayout &pl, const double *x, double *y, double alpha, double beta,
                            int overwrite, cudaStream_t s) {
  dim3 grid((unsigned)pl.ntiles), block(CTA_THREADS);
  // descriptors can also come from other partitions (transposed images under CSX-Sym)
  if (!pl.xdesc.empty()) csx_spmv_kernel<true, SYM, RPT, KSET><<<grid, block, 0, s>>>(P, x, y, alpha, beta, overwrite);
  else csx_spmv_kernel<false, SYM, RPT, KSET_ANY><<<grid, block, 0, s>>>(P, x, y, alpha, beta, overwrite);
}
template <bool SYM>
static void launch_gather(const PartDev &P, const PartLayout &pl, const double *x, double *y, double alpha, double beta,
                          int overwrite, cudaStream_t s) {
  // kernels are pre-compiled per (tile shape, unit-kind set) — the counterpart of the per-partition JIT (CsxJit.hpp)
  const bool diag1 = !SYM && pl.xd_diag1_only && !pl.xdesc.empty();
  if (pl.rpt == 4) {
    if (diag1) {  // the instantiation the stencil configs run: loads of a unit issued as one PTX block
      dim3 grid((unsigned)pl.ntiles), block(CTA_THREADS);
      csx_spmv_kernel<true, false, 4, KSET_DIAG1, 8, 1><<<grid, block, 0, s>>>(P, x, y, alpha, beta, overwrite);
    }
    else launch_gather_k<SYM, 4, KSET_ANY>(P, pl, x, y, alpha, beta, overwrite, s);
  } else {
    if (diag1) launch_gather_k<SYM, 1, SYM ? KSET_ANY : KSET_DIAG1>(P, pl, x, y, alpha, beta, overwrite, s);
    else launch_gather_k<SYM, 1, KSET_ANY>(P, pl, x, y, alpha, beta, overwrite, s);
  }
}

extern "C" {

int csxb_spmv(csxb_matrix_t *m, double alpha, const double *d_x, double beta, double *d_y, int overwrite, void *stream) {
  if (!m->uploaded) return fail("matrix not uploaded (csxb_upload)");
  cudaStream_t s = (cudaStream_t)stream;
  const bool sym = m->host.symmetric;
  if (m->sym_halo_hi > m->sym_halo_lo)   // halo rows owned by other devices: start from zero, the caller reduces them
    CUDA_TRY(cudaMemsetAsync(d_y + m->sym_halo_lo, 0, (size_t)(m->sym_halo_hi - m->sym_halo_lo) * 8, s));
  // kernel 1 of every partition first: it initialises y, and under CSX-Sym the chunk kernel of one
  // partition adds into rows that another partition owns
  for (size_t i = 0; i < m->pdev.size(); i++) {
    const PartLayout &pl = m->layout.parts[i];
    if (!pl.ntiles) continue;
    if (sym) launch_gather<true>(m->pdev[i], pl, d_x, d_y, alpha, beta, overwrite, s);
    else launch_gather<false>(m->pdev[i], pl, d_x, d_y, alpha, beta, overwrite, s);
  }
  for (size_t i = 0; i < m->pdev.size(); i++) {
    const PartLayout &pl = m->layout.parts[i];
    if (pl.chunks.empty()) continue;
    const unsigned grid = (unsigned)((pl.chunks.size() + CHUNK_WARPS - 1) / CHUNK_WARPS);
    if (sym) csx_chunk_kernel<true><<<grid, CHUNK_WARPS * 32, 0, s>>>(m->pdev[i], d_x, d_y, alpha);
    else csx_chunk_kernel<false><<<grid, CHUNK_WARPS * 32, 0, s>>>(m->pdev[i], d_x, d_y, alpha);
  }
  // rows after the last partition's last non-empty row belong to nobody; VecInit(y,0) clears them (CsxKernels.cpp:93)
  if (overwrite && m->host.part_lo + (int)m->host.parts.size() == m->host.nparts_total && m->covered_rows_end < m->host.nrows)
    CUDA_TRY(cudaMemsetAsync(d_y + m->covered_rows_end, 0, (size_t)(m->host.nrows - m->covered_rows_end) * 8, s));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int csxb_spmv_host(csxb_matrix_t *m, double alpha, const double *h_x, double beta, double *h_y, int overwrite) {
  if (!m->uploaded) return fail("matrix not uploaded (csxb_upload)");
  int cur = -1;
  CUDA_TRY(cudaGetDevice(&cur));
  if (cur != m->device) CUDA_TRY(cudaSetDevice(m->device));
  size_t nx = (size_t)std::max<int64_t>(m->host.ncols, 1), ny = (size_t)std::max<int64_t>(m->host.nrows, 1);
  if (!m->d_x) CUDA_TRY(cudaMalloc((void **)&m->d_x, nx * 8));
  if (!m->d_y) CUDA_TRY(cudaMalloc((void **)&m->d_y, ny * 8));
  CUDA_TRY(cudaMemcpyAsync(m->d_x, h_x, (size_t)m->host.ncols * 8, cudaMemcpyHostToDevice, 0));
  if (!overwrite) CUDA_TRY(cudaMemcpyAsync(m->d_y, h_y, (size_t)m->host.nrows * 8, cudaMemcpyHostToDevice, 0));
  if (csxb_spmv(m, alpha, m->d_x, beta, m->d_y, overwrite, nullptr)) return -1;
  // only the rows this handle computes travel back (one process per GPU owns one row range)
  int64_t lo = m->layout.parts.empty() ? 0 : m->layout.parts.front().row_start;
  int64_t hi = (m->host.part_lo + (int)m->host.parts.size() == m->host.nparts_total) ? m->host.nrows : m->covered_rows_end;
  if (!overwrite) { lo = 0; hi = m->host.nrows; }
  if (hi > lo) CUDA_TRY(cudaMemcpyAsync(h_y + lo, m->d_y + lo, (size_t)(hi - lo) * 8, cudaMemcpyDeviceToHost, 0));
  CUDA_TRY(cudaStreamSynchronize(0));
  if (cur != m->device && cur >= 0) CUDA_TRY(cudaSetDevice(cur));
  return 0;
}

int csxb_decode_coords(const csxb_matrix_t *mc, int part, int32_t *rows, int32_t *cols) {
  csxb_matrix *m = const_cast<csxb_matrix *>(mc);
  if (!m->uploaded) return fail("matrix not uploaded (csxb_upload)");
  if (part < 0 || (size_t)part >= m->pdev.size()) return fail("invalid partition");
  CUDA_TRY(cudaSetDevice(m->device));
  const PartLayout &pl = m->layout.parts[part];
  size_t n = (size_t)m->layout.total_values;
  int *dr = nullptr, *dcl = nullptr;
  CUDA_TRY(cudaMalloc((void **)&dr, std::max<size_t>(n, 1) * 4));
  CUDA_TRY(cudaMalloc((void **)&dcl, std::max<size_t>(n, 1) * 4));
  CUDA_TRY(cudaMemset(dr, 0xff, std::max<size_t>(n, 1) * 4));
  CUDA_TRY(cudaMemset(dcl, 0xff, std::max<size_t>(n, 1) * 4));
  if (pl.ntiles && !pl.xdesc.empty()) csx_decode_gather_kernel<<<(unsigned)pl.ntiles, CTA_THREADS>>>(m->pdev[part], dr, dcl);
  if (!pl.chunks.empty())
    csx_decode_chunk_kernel<<<(unsigned)((pl.chunks.size() + CHUNK_WARPS - 1) / CHUNK_WARPS), CHUNK_WARPS * 32>>>(m->pdev[part], dr, dcl);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(rows, dr + pl.val_base, (size_t)pl.nnz * 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(cols, dcl + pl.val_base, (size_t)pl.nnz * 4, cudaMemcpyDeviceToHost));
  cudaFree(dr); cudaFree(dcl);
  return 0;
}

}  // extern "C"
