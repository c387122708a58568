// Host-side helpers on the tuned CSX arrays (product code): a Boost-free container for tuned matrices
// (spx_mat_save / spx_mat_restore; reference: CsxSaveRestore.hpp:77-370 with boost::archive) and the element
// search behind spx_mat_get_entry / spx_mat_set_entry (reference: CsxGetSet.hpp:84-537, SearchValue :194-313).
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "gpu_layout.hpp"

namespace spxb {
namespace {

inline uint64_t get_varint(const uint8_t *ctl, uint64_t &p) {
  uint64_t v = 0;
  unsigned shift = 0;
  for (;;) {
    uint8_t b = ctl[p++];
    v |= (uint64_t)(b & 0x7f) << shift;
    if (!(b & 0x80)) break;
    shift += 7;
  }
  return v;
}

struct Unit {   // one decoded unit head
  uint32_t kind, align, delta, size;
  int64_t col;        // column of the first element
  uint64_t body;      // offset of the delta body (delta kinds)
  bool new_row;
  int64_t row_jump;
};

// decodes the unit head at p (advances p past head and body); `col` is the running column cursor of the row
bool next_unit(const CsxPartition &cp, bool full_colind, const KindEntry *kinds, size_t nkinds, uint64_t &p, int64_t &col, Unit &u) {
  const uint8_t *ctl = cp.ctl.data();
  const uint64_t end = cp.ctl.size();
  if (p + 2 > end) return false;
  const uint8_t flags = ctl[p++];
  u.size = ctl[p++];
  u.new_row = (flags & 0x80) != 0;
  u.row_jump = 0;
  if (u.new_row) { u.row_jump = (flags & 0x40) ? (int64_t)get_varint(ctl, p) : 1; col = 0; }
  if (full_colind) { uint32_t c; memcpy(&c, ctl + p, 4); p += 4; col = c; }
  else col = (int64_t)((uint64_t)col + get_varint(ctl, p));
  const uint32_t id = flags & 0x3f;
  if (id >= nkinds || u.size == 0) return false;
  u.kind = kinds[id].kind_align & 0xff; u.align = (kinds[id].kind_align >> 8) & 0xff; u.delta = kinds[id].delta;
  u.col = col;
  u.body = p;
  if (u.kind <= K_DELTA64) {
    for (uint32_t k = 1; k < u.size; k++) { uint64_t d = 0; memcpy(&d, ctl + p, u.delta); p += u.delta; col += (int64_t)d; }
  } else if (u.kind == K_HORIZ) col += (int64_t)(u.size - 1) * u.delta;
  return p <= end;
}

inline int64_t unit_span(const Unit &u) {
  if (u.kind >= K_VERT && u.kind <= K_ADIAG) return (int64_t)(u.size - 1) * u.delta;
  if (u.kind == K_BROW) return (int64_t)u.align - 1;
  if (u.kind == K_BCOL) return (int64_t)u.delta - 1;
  return 0;
}

bool kinds_of(const CsxPartition &cp, std::vector<KindEntry> &kinds) {
  kinds.clear();
  for (size_t i = 0; i < cp.id_map.size() && cp.id_map[i] != -1; i++) {
    KindEntry ke;
    if (!classify(cp.id_map[i], ke)) return false;
    kinds.push_back(ke);
  }
  return true;
}

}  // namespace

std::string build_row_index(const CsxPartition &cp, bool full_colind, RowIndex &ri) {
  std::vector<KindEntry> kinds;
  if (!kinds_of(cp, kinds)) return "unsupported pattern id";
  const size_t nrows = (size_t)std::max<int64_t>(cp.nrows, (int64_t)cp.dvalues.size());
  ri.ctl_off.assign(nrows, UINT64_MAX);
  ri.val_off.assign(nrows, 0);
  ri.span.assign(nrows, 0);
  ri.max_span = 0;
  uint64_t p = 0;
  int64_t row = 0, col = 0, v = 0;
  while (p < cp.ctl.size()) {
    const uint64_t at = p;
    Unit u;
    if (!next_unit(cp, full_colind, kinds.data(), kinds.size(), p, col, u)) return "malformed ctl stream";
    if (u.new_row) row += u.row_jump;   // the stream starts in row 0; a leading empty row shows as a row bit on the first unit
    if (row < 0 || (size_t)row >= nrows) return "ctl stream leaves the partition";
    if (ri.ctl_off[row] == UINT64_MAX) { ri.ctl_off[row] = at; ri.val_off[row] = (uint64_t)v; }
    const int64_t sp = unit_span(u);
    ri.span[row] = std::max<int32_t>(ri.span[row], (int32_t)sp);
    ri.max_span = std::max<int32_t>(ri.max_span, (int32_t)sp);
    v += u.size;
  }
  if (v != cp.nnz) return "ctl stream does not cover the values";
  return "";
}

int64_t find_entry(const CsxPartition &cp, bool full_colind, const RowIndex &ri, int64_t prow, int64_t col) {
  std::vector<KindEntry> kinds;
  if (!kinds_of(cp, kinds)) return -1;
  const uint8_t *ctl = cp.ctl.data();
  for (int64_t r = prow; r >= 0 && prow - r <= ri.max_span; r--) {
    if ((size_t)r >= ri.ctl_off.size() || ri.ctl_off[r] == UINT64_MAX || ri.span[r] < prow - r) continue;
    uint64_t p = ri.ctl_off[r];
    int64_t cur = 0, v = (int64_t)ri.val_off[r];
    bool first = true;
    while (p < cp.ctl.size()) {
      const uint64_t save = p;
      Unit u;
      int64_t c2 = cur;
      if (!next_unit(cp, full_colind, kinds.data(), kinds.size(), p, c2, u)) return -1;
      if (u.new_row && !first) { p = save; break; }   // next row begins
      first = false;
      cur = c2;
      const int64_t dr = prow - r, c0 = u.col;
      if (u.kind <= K_DELTA64) {
        if (dr == 0) {
          int64_t cc = c0;
          for (uint32_t k = 0; k < u.size; k++) {
            if (k) { uint64_t d = 0; memcpy(&d, ctl + u.body + (uint64_t)(k - 1) * u.delta, u.delta); cc += (int64_t)d; }
            if (cc == col) return v + k;
            if (cc > col) break;
          }
        }
      } else if (u.kind == K_HORIZ) {
        if (dr == 0 && col >= c0 && (col - c0) % u.delta == 0 && (col - c0) / u.delta < u.size) return v + (col - c0) / u.delta;
      } else if (u.kind == K_VERT || u.kind == K_DIAG || u.kind == K_ADIAG) {
        if (dr % u.delta == 0 && dr / u.delta < u.size) {
          const int64_t k = dr / u.delta;
          const int64_t cc = u.kind == K_VERT ? c0 : (u.kind == K_DIAG ? c0 + k * u.delta : c0 - k * u.delta);
          if (cc == col) return v + k;
        }
      } else if (u.kind == K_BROW) {   // align rows x delta columns, column-major
        if (dr < (int64_t)u.align && col >= c0 && col - c0 < (int64_t)u.delta) return v + (col - c0) * u.align + dr;
      } else {                          // K_BCOL: delta rows x align columns, row-major
        if (dr < (int64_t)u.delta && col >= c0 && col - c0 < (int64_t)u.align) return v + dr * u.align + (col - c0);
      }
      v += u.size;
    }
  }
  return -1;
}

// ---- container --------------------------------------------------------------------------------------------
namespace {
const char MAGIC[8] = {'C', 'S', 'X', 'B', '2', '0', '0', 1};
struct Writer {
  FILE *f; bool ok = true;
  void raw(const void *p, size_t n) { if (ok && n && fwrite(p, 1, n, f) != n) ok = false; }
  void i64(int64_t v) { raw(&v, 8); }
  template <class T> void vec(const std::vector<T> &v) { i64((int64_t)v.size()); raw(v.data(), v.size() * sizeof(T)); }
  void str(const std::string &s) { i64((int64_t)s.size()); raw(s.data(), s.size()); }
};
struct Reader {
  FILE *f; bool ok = true;
  int64_t left = 0;   // bytes of the file not read yet: no array can be longer than that
  void raw(void *p, size_t n) {
    if (ok && n && ((int64_t)n > left || fread(p, 1, n, f) != n)) ok = false;
    if (ok) left -= (int64_t)n;
  }
  int64_t i64() { int64_t v = 0; raw(&v, 8); return v; }
  template <class T> void vec(std::vector<T> &v) {
    int64_t n = i64();
    if (!ok || n < 0 || n > left / (int64_t)sizeof(T)) { ok = false; return; }
    v.resize((size_t)n);
    raw(v.data(), (size_t)n * sizeof(T));
  }
  void str(std::string &s) { int64_t n = i64(); if (!ok || n < 0 || n > left || n > (1 << 24)) { ok = false; return; } s.resize((size_t)n); raw(&s[0], (size_t)n); }
};
}  // namespace

std::string save_matrix(const CsxMatrix &m, const char *path) {
  FILE *f = fopen(path, "wb");
  if (!f) return std::string("cannot open ") + path + " for writing";
  Writer w{f};
  w.raw(MAGIC, 8);
  w.i64(m.nrows); w.i64(m.ncols); w.i64(m.nnz); w.i64(m.symmetric); w.i64(m.full_colind);
  w.i64(m.rows_per_thread); w.i64(m.slice_elems); w.i64(m.slab_rows); w.i64(m.nparts_total); w.i64(m.part_lo);
  w.i64((int64_t)m.parts.size());
  for (const CsxPartition &p : m.parts) {
    if ((int64_t)p.values.size() != p.nnz) { fclose(f); return "partition values are not on the host"; }
    w.i64(p.nnz); w.i64(p.nrows); w.i64(p.ncols); w.i64(p.row_start); w.i64(p.row_jumps); w.i64(p.col_min); w.i64(p.col_max);
    w.i64(p.sampling_undefined);
    w.vec(p.values); w.vec(p.ctl);
    std::vector<int64_t> ids(p.id_map.begin(), p.id_map.end());
    w.vec(ids);
    w.vec(p.rows_info); w.vec(p.dvalues); w.vec(p.map_cpus); w.vec(p.map_pos);
    w.str(p.encoding_log);
  }
  w.vec(m.permutation);   // SaveTuned stores the permutation with the matrix (matvec.c:422, CsxSaveRestore.hpp)
  const bool ok = w.ok && fclose(f) == 0;
  return ok ? "" : std::string("write error on ") + path;
}

std::string load_matrix(const char *path, CsxMatrix &m) {
  FILE *f = fopen(path, "rb");
  if (!f) return std::string("cannot open ") + path;
  Reader r{f};
  if (fseek(f, 0, SEEK_END) == 0) { r.left = (int64_t)ftell(f); rewind(f); }
  char magic[8];
  r.raw(magic, 8);
  if (!r.ok || memcmp(magic, MAGIC, 8) != 0) { fclose(f); return std::string(path) + " is not a CSX container of this engine (or of another version)"; }
  m = CsxMatrix();
  m.nrows = r.i64(); m.ncols = r.i64(); m.nnz = r.i64(); m.symmetric = r.i64() != 0; m.full_colind = r.i64() != 0;
  m.rows_per_thread = (int)r.i64(); m.slice_elems = (int)r.i64(); m.slab_rows = r.i64(); m.nparts_total = (int)r.i64();
  m.part_lo = (int)r.i64();
  const int64_t np = r.i64();
  if (!r.ok || np < 0 || np > 4096) { fclose(f); return "corrupt container header"; }
  m.parts.resize((size_t)np);
  for (CsxPartition &p : m.parts) {
    p.nnz = r.i64(); p.nrows = r.i64(); p.ncols = r.i64(); p.row_start = r.i64(); p.row_jumps = r.i64() != 0;
    p.col_min = r.i64(); p.col_max = r.i64(); p.sampling_undefined = r.i64() != 0;
    r.vec(p.values); r.vec(p.ctl);
    std::vector<int64_t> ids;
    r.vec(ids);
    p.id_map.assign(ids.begin(), ids.end());
    r.vec(p.rows_info); r.vec(p.dvalues); r.vec(p.map_cpus); r.vec(p.map_pos);
    r.str(p.encoding_log);
    if (!r.ok || (int64_t)p.values.size() != p.nnz) { fclose(f); return "corrupt container (partition arrays)"; }
    if (p.nrows < 0 || p.ncols != m.ncols || p.row_start < 0 || p.row_start + p.nrows > m.nrows || p.id_map.size() > 64 ||
        (m.symmetric && (int64_t)p.dvalues.size() + p.row_start > m.nrows)) { fclose(f); return "corrupt container (partition header)"; }
  }
  if (r.ok && r.left > 0) {   // containers written before the permutation was stored end here
    r.vec(m.permutation);
    if (r.ok && !m.permutation.empty()) {
      bool good = (int64_t)m.permutation.size() == m.nrows;
      std::vector<bool> seen(good ? (size_t)m.nrows : 0, false);
      for (size_t i = 0; good && i < m.permutation.size(); i++) {
        const int32_t v = m.permutation[i];
        good = v >= 0 && v < m.nrows && !seen[(size_t)v];
        if (good) seen[(size_t)v] = true;
      }
      if (!good) { fclose(f); return "corrupt container (permutation)"; }
    }
  }
  fclose(f);
  if (!r.ok) return "truncated container";
  // the same range checks the options go through at tune time (TuneOptions::set); the ctl stream itself is validated
  // when the GPU tables are built (build_layout walks it with bounds checks)
  if (m.nrows < 0 || m.ncols < 0 || (m.rows_per_thread != 0 && m.rows_per_thread != 1 && m.rows_per_thread != 4) || m.slab_rows < 0 ||
      m.nparts_total < (int)np || m.part_lo < 0 || m.part_lo + (int)np > m.nparts_total)
    return "corrupt container (options)";
  return "";
}

}  // namespace spxb
