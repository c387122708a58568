// Host emulation of both SpMV kernels (test infrastructure): tunes a CSR matrix with the product encoder, builds
// the product's GPU layout and then executes sparsex_b200/csrc/gather_kernel.cuh (kernel 1, generic instantiations)
// and part_dev.cuh (kernel 2) — the same text nvcc compiles — warp by warp on the fibre emulation of
// warp_emul.hpp.  Lets the CPU test suite check the traversal logic of the kernels against the CSR input without
// a GPU.
#include "warp_emul.hpp"

#include <cmath>
#include <sstream>
#include <string>
#include <vector>

#define CSXB_EMUL 1
#include "../../sparsex_b200/csrc/csx_host.hpp"
#include <type_traits>
#include "../../sparsex_b200/csrc/stream_kernel.cuh"
#include "../../sparsex_b200/csrc/gather_kernel.cuh"

namespace {
void put_err(char *err, size_t n, const std::string &m) {
  if (err && n) { strncpy(err, m.c_str(), n - 1); err[n - 1] = 0; }
}
}  // namespace

extern "C" int emul_spmv(const int32_t *rowptr, const int32_t *colind, const double *values, int64_t nrows, int64_t ncols,
                         const char *options, double alpha, const double *x, double *y, int32_t *dec_rows,
                         int32_t *dec_cols, int64_t *stats, char *err, size_t errlen) {
  TuneOptions o;
  if (options) {
    std::stringstream ss(options);
    std::string kv;
    while (std::getline(ss, kv, ';')) {
      if (kv.empty()) continue;
      size_t eq = kv.find('=');
      std::string e = eq == std::string::npos ? "malformed option" : o.set(kv.substr(0, eq), kv.substr(eq + 1));
      if (!e.empty()) { put_err(err, errlen, e); return -1; }
    }
  }
  CsxMatrix M;
  CsrView v{rowptr, colind, values, nrows, ncols};
  std::string e = tune_csr(v, o, 0, o.nr_threads, M);
  if (!e.empty()) { put_err(err, errlen, e); return -1; }
  std::vector<int64_t> saved;
  for (auto &p : M.parts) { saved.push_back(p.nrows); if (M.symmetric) p.nrows = (int64_t)p.dvalues.size(); }
  DeviceLayout L;
  e = build_layout(M, L);
  if (!e.empty()) { put_err(err, errlen, "layout: " + e); return -1; }
  // device-wide arrays as the upload would lay them out
  std::vector<double> vals((size_t)L.total_values + 2, 0.0);
  std::vector<uint8_t> ctl((size_t)L.total_ctl + 64, 0);
  for (size_t i = 0; i < L.parts.size(); i++) {
    memcpy(vals.data() + L.parts[i].val_base, M.parts[i].values.data(), M.parts[i].values.size() * 8);
    memcpy(ctl.data() + L.parts[i].ctl_base, M.parts[i].ctl.data(), M.parts[i].ctl.size());
  }
  // stale content of y must not leak: kernel 1 writes every owned row; rows behind the last partition are cleared by
  // the host side of csxb_spmv (VecInit(y, 0), CsxKernels.cpp:93)
  int64_t covered = 0;
  for (size_t i = 0; i < L.parts.size(); i++) covered = std::max<int64_t>(covered, L.parts[i].row_start + L.parts[i].nrows);
  for (int64_t i = 0; i < nrows; i++) y[i] = i < covered ? std::nan("") : 0.0;
  for (int64_t i = 0; i < (int64_t)L.total_values; i++) { dec_rows[i] = -1; dec_cols[i] = -1; }
  int64_t nchunks = 0, nunits = 0, nxd = 0;
  std::vector<PartDev> pdev;
  std::vector<std::vector<double>> scratch(L.parts.size());
  for (size_t i = 0; i < L.parts.size(); i++) {
    const PartLayout &pl = L.parts[i];
    const CsxPartition &hp = M.parts[i];
    // decoded coordinates of the table units, each unit once (under the tile of its first row)
    for (int64_t t = 0; t < pl.ntiles; t++)
      for (uint32_t j = pl.tile_xoff[t]; j < pl.tile_xoff[t + 1]; j++) {
        const XDesc &d = pl.xdesc[j];
        if (d.meta & XD_TRANSPOSED) continue;
        if (((int64_t)d.row - pl.row_start) / pl.tile_rows() != t) continue;
        nxd++;
        const uint32_t kind = (d.meta >> 24) & 0xf, size = (d.meta >> 16) & 0xff;
        const uint32_t delta = (d.meta & XD_DELTA1) ? 1u : L.ktab[d.meta & 0xffff].delta;
        for (uint32_t k = 0; k < size; k++) {
          const int64_t r = d.row + (int64_t)k * delta;
          const int64_t c = kind == K_VERT ? d.col : (kind == K_DIAG ? d.col + (int64_t)k * delta : d.col - (int64_t)k * delta);
          dec_rows[d.voff + k] = (int32_t)r; dec_cols[d.voff + k] = (int32_t)c;
        }
      }
    PartDev P;
    memset(&P, 0, sizeof(P));
    P.ctl = ctl.data() + pl.ctl_base;
    P.values = vals.data();
    P.tile_xoff = pl.tile_xoff.data();
    P.xdesc = reinterpret_cast<const uint4 *>(pl.xdesc.data());
    P.ktab = L.ktab.data();
    P.dvalues = hp.dvalues.data();
    P.nrows = pl.nrows; P.row_start = pl.row_start; P.val_base = (uint32_t)pl.val_base;
    P.full_colind = L.full_colind;
    P.rpt = pl.rpt;
    memcpy(P.idtab, pl.idtab, sizeof(P.idtab));
    for (size_t t = 0; t < pl.bt.size(); t++) {
      const BlockTable &T = pl.bt[t];
      P.bt[t] = BtDev{T.ptr.data(), reinterpret_cast<const uint2 *>(T.ent.data()), (long long)T.j0, T.G, T.nloop, T.sf, T.sl, T.image, (uint32_t)((0x100000000ull + T.G - 1) / T.G)};
      if (stats) stats[12] += (int64_t)T.ent.size();
      if (T.image) continue;   // decoded coordinates of the block-table units (csx_decode_bt_kernel)
      for (int64_t lr = 0; lr < pl.nrows; lr++) {
        const int64_t g = pl.row_start + lr, J = g / T.G;
        const int f = (int)(g - J * T.G) * T.sf;
        for (uint32_t e = T.ptr[J - T.j0]; e < T.ptr[J - T.j0 + 1]; e++)
          for (int l = 0; l < T.nloop && !(T.ent[e].other & BT_IMAGE); l++) {
            dec_rows[T.ent[e].voff + f + l * T.sl] = (int32_t)g; dec_cols[T.ent[e].voff + f + l * T.sl] = (int32_t)T.ent[e].other + l;
          }
      }
    }
    P.nbt = (int)pl.bt.size();
    P.sk_chunks = reinterpret_cast<const uint4 *>(pl.sk_chunks.data());
    P.sk_uoffs = pl.sk_uoffs.data();
    scratch[i].assign(pl.sk_scratch + 2, std::nan(""));
    P.sk_scratch = scratch[i].data();
    P.sk_fix_rows = pl.sk_fix_rows.data(); P.sk_fix_ptr = pl.sk_fix_ptr.data(); P.sk_fix_idx = pl.sk_fix_idx.data();
    P.sk_gaps = reinterpret_cast<const long long *>(pl.sk_gaps.data());
    P.sk_c0 = 0; P.sk_c1 = (uint32_t)pl.sk_chunks.size();
    nchunks += (int64_t)pl.sk_chunks.size();
    nunits += (int64_t)pl.sk_uoffs.size();
    if (stats) for (int k = 0; k < 4; k++) stats[8 + k] += pl.sk_stat[k];
    if (stats) { stats[4] += (int64_t)pl.sk_fix_idx.size(); stats[5] += (int64_t)pl.sk_gaps.size(); stats[6] = pl.sk_rows; stats[7] = pl.sk_kmask; }
    pdev.push_back(P);
  }
  if (getenv("CSXB_EMUL_LAYOUT_ONLY")) { if (stats) { stats[0] = nchunks; stats[1] = nunits; } return 0; }
  static double sacc[SK_WIN];
  static uint4 sid[64];
  std::vector<bool> streamed(L.parts.size(), false);
  int inst = 0;
  for (size_t i = 0; i < L.parts.size(); i++) {
    const PartLayout &pl = L.parts[i];
    const PartDev &P = pdev[i];
    if (pl.sk_chunks.empty()) continue;
    streamed[i] = true;
    for (int k = 0; k < 64; k++) sid[k] = make_uint4(P.idtab[k].kind_align, P.idtab[k].delta, P.idtab[k].sl, P.idtab[k].recip);
    for (uint32_t ch = 0; ch < P.sk_c1; ch++) {
      for (int k = 0; k < SK_WIN; k++) sacc[k] = std::nan("");   // the kernel must clear what it uses
      warp_emul::run_warp([&](int lane) {
        sk_chunk<8, SKM_ROWLOCAL | SKM_BROW | SKM_BCOL, 0, 0, true>(P, ch, sacc, sid, lane, nullptr, nullptr, 0.0, 0.0, 1, dec_rows + P.val_base, dec_cols + P.val_base);
      });
      bool done = false;
#define SK_TRY(RR, KK, BCC, BRR)                                                                                    \
  if (!done && sk_instance_serves(RR, (KK), BCC, BRR, pl.sk_kmask, pl.sk_rows, pl.sk_bc, pl.sk_brc)) {              \
    done = true;                                                                                                    \
    inst = RR * 1000 + BCC * 100 + BRR * 10;                                                                        \
    warp_emul::run_warp([&](int lane) { sk_chunk<RR, (KK), BCC, BRR, false>(P, ch, sacc, sid, lane, x, y, alpha, 0.0, 1, nullptr, nullptr); }); \
  }
      SK_INSTANCES(SK_TRY)
#undef SK_TRY
      if (!done) { put_err(err, errlen, "no stream kernel instantiation"); return -1; }
    }
    // csx_stream_fixup_kernel
    for (size_t f = 0; f < pl.sk_fix_rows.size(); f++) {
      double sum = 0.0;
      for (uint32_t j = pl.sk_fix_ptr[f]; j < pl.sk_fix_ptr[f + 1]; j++) sum += scratch[i][pl.sk_fix_idx[j]];
      y[pl.row_start + pl.sk_fix_rows[f]] += alpha * sum;
    }
    for (const SkGap &g : pl.sk_gaps)
      for (int64_t rr = g.lo; rr < g.hi; rr++) y[pl.row_start + rr] = 0.0;
  }
  // kernel 1 of every partition first (it initialises y; under CSX-Sym chunks add into rows of other partitions),
  // with the instantiation launch_gather would pick (the PTX variant of the diagonal kernel is device-only)
  for (size_t i = 0; i < L.parts.size(); i++) {
    const PartLayout &pl = L.parts[i];
    const PartDev &P = pdev[i];
    const bool xd = !pl.xdesc.empty(), diag1 = !M.symmetric && pl.xd_diag1_only && xd;
    const bool bt = !pl.bt.empty();
    if (streamed[i] && !xd && !bt && !M.symmetric) continue;
    const double beta = streamed[i] ? 1.0 : 0.0;
    const int ow = streamed[i] ? 0 : 1;
    for (int64_t t = 0; t < pl.ntiles; t++)
      for (int w = 0; w < CTA_THREADS / 32; w++) {
        warp_emul::warp_in_cta() = w;
        warp_emul::run_warp([&](int) {
          const NoXchg nx;
          if (M.symmetric) {
            if (pl.rpt == 4) { if (xd) (bt ? spmv_tile<true, true, 4, KSET_ANY, 0, false, NoXchg, true> : spmv_tile<true, true, 4, KSET_ANY, 0, false, NoXchg, false>)(P, x, y, alpha, beta, ow, t, t, nx, 0);
                               else (bt ? spmv_tile<false, true, 4, KSET_ANY, 0, false, NoXchg, true> : spmv_tile<false, true, 4, KSET_ANY, 0, false, NoXchg, false>)(P, x, y, alpha, beta, ow, t, t, nx, 0); }
            else { if (xd) (bt ? spmv_tile<true, true, 1, KSET_ANY, 0, false, NoXchg, true> : spmv_tile<true, true, 1, KSET_ANY, 0, false, NoXchg, false>)(P, x, y, alpha, beta, ow, t, t, nx, 0);
                   else (bt ? spmv_tile<false, true, 1, KSET_ANY, 0, false, NoXchg, true> : spmv_tile<false, true, 1, KSET_ANY, 0, false, NoXchg, false>)(P, x, y, alpha, beta, ow, t, t, nx, 0); }
          } else if (pl.rpt == 4) {
            if (diag1) (bt ? spmv_tile<true, false, 4, KSET_DIAG1, 0, false, NoXchg, true> : spmv_tile<true, false, 4, KSET_DIAG1, 0, false, NoXchg, false>)(P, x, y, alpha, beta, ow, t, t, nx, 0);
            else if (xd) (bt ? spmv_tile<true, false, 4, KSET_ANY, 0, false, NoXchg, true> : spmv_tile<true, false, 4, KSET_ANY, 0, false, NoXchg, false>)(P, x, y, alpha, beta, ow, t, t, nx, 0);
            else (bt ? spmv_tile<false, false, 4, KSET_ANY, 0, false, NoXchg, true> : spmv_tile<false, false, 4, KSET_ANY, 0, false, NoXchg, false>)(P, x, y, alpha, beta, ow, t, t, nx, 0);
          } else {
            if (diag1) (bt ? spmv_tile<true, false, 1, KSET_DIAG1, 0, false, NoXchg, true> : spmv_tile<true, false, 1, KSET_DIAG1, 0, false, NoXchg, false>)(P, x, y, alpha, beta, ow, t, t, nx, 0);
            else if (xd) (bt ? spmv_tile<true, false, 1, KSET_ANY, 0, false, NoXchg, true> : spmv_tile<true, false, 1, KSET_ANY, 0, false, NoXchg, false>)(P, x, y, alpha, beta, ow, t, t, nx, 0);
            else (bt ? spmv_tile<false, false, 1, KSET_ANY, 0, false, NoXchg, true> : spmv_tile<false, false, 1, KSET_ANY, 0, false, NoXchg, false>)(P, x, y, alpha, beta, ow, t, t, nx, 0);
          }
        });
      }
  }
  warp_emul::warp_in_cta() = 0;
  for (size_t i = 0; i < M.parts.size(); i++) M.parts[i].nrows = saved[i];
  if (stats) { stats[0] = nchunks; stats[1] = nunits; stats[2] = nxd; stats[3] = inst; }
  return 0;
}
