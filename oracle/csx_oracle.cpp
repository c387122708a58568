// ORACLE — TEST INFRASTRUCTURE ONLY (see csx_oracle.hpp).
// CPU restatement of the SparseX CSX / CSX-Sym encoder.  It follows the
// reference function by function and keeps its container choices (std::map /
// std::set iteration order, std::sort on (row, col)) so that order-dependent
// behaviour is inherited rather than re-derived.  Non-NUMA branches only
// (SPX_USE_NUMA == 0).  Citations are file:line into /root/reference/.
// PINNED TO THE REFERENCE: oracle/build_refenc.py compiles the reference's own encoder (its headers and
// Encodings/Runtime/Statistics/CtlBuilder/CsxUtil sources, g++ against the stand-in headers of oracle/refshim);
// this restatement reproduces its ctl / values / id_map / dvalues bit for bit on the 1002 seeded cases of
// tests/refpin_cases.py (tests/golden/ref_encodings.json, tests/test_cpu_refpin.py).
#include "csx_oracle.hpp"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace csxo {

struct OracleError : std::runtime_error {
  explicit OracleError(const std::string &m) : std::runtime_error(m) {}
};

// ---------------------------------------------------------------- options --
static bool parse_bool(const std::string &v) { return v == "true" || v == "1"; }

// Runtime.cpp:65-95 (mnemonics)
std::string Options::set(const std::string &k, const std::string &v) {
  try {
    if (k == "spx.rt.nr_threads") nr_threads = std::stoi(v);
    else if (k == "spx.rt.cpu_affinity") {}
    else if (k == "spx.preproc.heuristic") { if (v != "ratio") return "only the ratio heuristic is restated"; }
    else if (k == "spx.preproc.xform") xform = v;
    else if (k == "spx.preproc.sampling") sampling = v;
    else if (k == "spx.preproc.sampling.nr_samples") nr_samples = std::stoul(v);
    else if (k == "spx.preproc.sampling.portion") portion = std::stod(v);
    else if (k == "spx.preproc.sampling.window_size") window_size = std::stoul(v);
    else if (k == "spx.matrix.symmetric") symmetric = parse_bool(v);
    else if (k == "spx.matrix.split_blocks") split_blocks = parse_bool(v);
    else if (k == "spx.matrix.full_colind") full_colind = parse_bool(v);
    else if (k == "spx.matrix.min_unit_size") min_unit_size = std::stoul(v);
    else if (k == "spx.matrix.max_unit_size") max_unit_size = std::stoul(v);
    else if (k == "spx.matrix.min_coverage") min_coverage = std::stod(v);
    else if (k == "oracle.onedim_blocks") onedim_blocks = parse_bool(v);
    else if (k == "oracle.undefined_sampling") undefined_sampling = v;
    else return "unknown option " + k;
  } catch (std::exception &) { return "bad value for " + k; }
  return "";
}

// ------------------------------------------------- xform string (Encodings) --
// Encodings.cpp:32-57 short names.
static int type_from_name(const std::string &s) {
  static const char *names[] = {"none", "h", "v", "d", "ad", "br1", "br2", "br3", "br4", "br5", "br6",
                                "br7", "br8", "bc1", "bc2", "bc3", "bc4", "bc5", "bc6", "bc7", "bc8"};
  for (int i = 0; i < TypeMax; i++) if (s == names[i]) return i;
  if (s == "br") return 100;
  if (s == "bc") return 101;
  if (s == "all") return 102;
  return -1;
}
// Encodings.cpp:78-98
static void group_types(int t, std::vector<int> &out) {
  if (t == 100) for (int i = BlockRow1; i <= BlockRow8; i++) out.push_back(i);
  else if (t == 101) for (int i = BlockCol1; i <= BlockCol8; i++) out.push_back(i);
  else if (t == 102) for (int i = None; i < TypeMax; i++) out.push_back(i);
  else out.push_back(t);
}

struct EncSeq {  // Encodings.cpp:108-138, regex ([a-z]+([0-9]*))(\{([0-9]+(,[0-9]+)*)\})?
  std::vector<std::pair<int, std::vector<size_t>>> seq;
  bool is_explicit = false;
  explicit EncSeq(const std::string &s) {
    size_t i = 0, n = s.size();
    while (i < n) {
      if (!(s[i] >= 'a' && s[i] <= 'z')) { i++; continue; }  // regex_search skips non-matching text
      size_t j = i;
      while (j < n && s[j] >= 'a' && s[j] <= 'z') j++;
      while (j < n && s[j] >= '0' && s[j] <= '9') j++;
      std::string name = s.substr(i, j - i);
      int t = type_from_name(name);
      if (t < 0) throw OracleError("invalid value \"" + name + "\" while setting property \"spx.preproc.xform\"");
      std::vector<size_t> deltas;
      if (j < n && s[j] == '{') {
        size_t k = j + 1, start = k;
        bool ok = false;
        std::vector<size_t> tmp;
        while (k < n) {
          if (s[k] >= '0' && s[k] <= '9') { k++; continue; }
          if ((s[k] == ',' || s[k] == '}') && k > start) {
            tmp.push_back(std::stoul(s.substr(start, k - start)));
            if (s[k] == '}') { ok = true; k++; break; }
            k++; start = k; continue;
          }
          break;
        }
        if (ok) { deltas = tmp; j = k; }
      }
      if (!deltas.empty()) is_explicit = true;
      seq.push_back(std::make_pair(t, deltas));
      i = j;
    }
  }
};

// ------------------------------------------------------------- Xform.hpp --
typedef std::pair<int, int> RC;
static RC xform_from_horiz(int to, RC p, int R, int C) {  // Xform.hpp:37-248, 258-329
  int r = p.first, c = p.second;
  switch (to) {
    case Horizontal: return RC(r, c);
    case Vertical: return RC(c, r);
    case Diagonal: return RC(R + c - r, (c < r) ? c : r);                    // :106-112
    case AntiDiagonal: { int nr = r + c - 1; return RC(nr, (nr <= C) ? r : C - c + 1); }  // :142-148
    default: break;
  }
  if (is_block_row(to)) { int k = (int)block_align(to); return RC((r - 1) / k + 1, (r - 1) % k + k * (c - 1) + 1); }  // :181-187
  if (is_block_col(to)) { int k = (int)block_align(to); return RC((c - 1) / k + 1, (c - 1) % k + k * (r - 1) + 1); }  // :216-222
  throw OracleError("bad xform target");
}
static RC xform_to_horiz(int from, RC p, int R, int C) {  // Xform.hpp:331-418
  int r = p.first, c = p.second;
  switch (from) {
    case Horizontal: return RC(r, c);
    case Vertical: return RC(c, r);
    case Diagonal: return (r < R) ? RC(R + c - r, c) : RC(c, r + c - R);    // :123-131
    case AntiDiagonal: return (r <= C) ? RC(c, r - c + 1) : RC(r + c - C, C - c + 1);  // :160-168
    default: break;
  }
  if (is_block_row(from)) { int k = (int)block_align(from); return RC(k * (r - 1) + (c - 1) % k + 1, (c - 1) / k + 1); }  // :198-204
  if (is_block_col(from)) { int k = (int)block_align(from); RC t(k * (r - 1) + (c - 1) % k + 1, (c - 1) / k + 1); return RC(t.second, t.first); }  // :233-239
  throw OracleError("bad xform source");
}
static RC xform(int from, int to, RC p, int R, int C) {  // Xform.hpp:420-443
  if (from == to) return p;
  if (from == Horizontal) return xform_from_horiz(to, p, R, C);
  if (to == Horizontal) return xform_to_horiz(from, p, R, C);
  return xform_from_horiz(to, xform_to_horiz(from, p, R, C), R, C);
}

// ----------------------------------------------------- SparsePartition --
struct Partition {
  size_t nr_rows = 0, nr_cols = 0, nr_nzeros = 0;
  int type = None;
  std::vector<Elem> elems;     // elems_[0 .. elems_size_)
  std::vector<int> rowptr;     // rowptr_size_ == rowptr.size()
  int row_start = 0;

  size_t rowptr_size() const { return rowptr.size(); }

  // SparsePartition.hpp:543-563 + Builder :852-891
  void set_rowptr() {
    rowptr.clear();
    rowptr.push_back(0);
    int row_prev = 1;
    size_t cnt = 0;
    for (const Elem &e : elems) {
      if (e.row != row_prev) {
        if (e.row < row_prev)
          throw OracleError("set_rowptr: rows not ascending (row " + std::to_string(e.row) + " after " +
                            std::to_string(row_prev) + ", type " + std::to_string(type) + ", nr_rows " +
                            std::to_string(nr_rows) + ")");
        for (int k = 0; k < e.row - row_prev; k++) rowptr.push_back((int)cnt);
        row_prev = e.row;
      }
      cnt++;
    }
    if ((size_t)rowptr.back() != cnt) rowptr.push_back((int)cnt);
  }

  // SparsePartition.hpp:680-744 (the h<->br / v<->bc partial-sort shortcut is a
  // full lexicographic sort on disjoint, ordered groups; restated as one sort)
  void transform(int t) {
    if (type == t) return;
    for (Elem &e : elems) {
      RC n = xform(type, t, RC(e.row, e.col), (int)nr_rows, (int)nr_cols);
      e.row = n.first; e.col = n.second;
    }
    std::sort(elems.begin(), elems.end(), [](const Elem &a, const Elem &b) {
      return a.row < b.row || (a.row == b.row && a.col < b.col);  // Element.hpp:486-489
    });
    if (!elems.empty()) set_rowptr();
    type = t;
  }
};

// ---------------------------------------------------------- Statistics --
struct StatsData {  // Statistics.hpp:54-174
  size_t enc = 0, pat = 0, del = 0;
  StatsData() {}
  StatsData(size_t e, size_t p, size_t d = 0) : enc(e), pat(p), del(d) {}
  StatsData &operator+=(const StatsData &o) { enc += o.enc; pat += o.pat; del += o.del; return *this; }
  void scale(double f) { enc = (size_t)(enc * f); pat = (size_t)(pat * f); del = (size_t)(del * f); }  // :135-144
  bool zero() const { return enc == 0 && pat == 0 && del == 0; }
};
typedef std::map<size_t, StatsData> InstStats;
struct TypeNode { InstStats stats; StatsData data; };
struct StatsCollection {  // Statistics.hpp:228-524
  std::map<int, TypeNode> types;
  void append(const Inst &inst, const StatsData &d) {  // :311-314, 466-475, 415-423
    auto it = types.find(inst.first);
    if (it == types.end()) { TypeNode n; n.stats[inst.second] = d; n.data = d; types[inst.first] = n; }
    else {
      auto ii = it->second.stats.find(inst.second);
      if (ii == it->second.stats.end()) it->second.stats[inst.second] = d; else ii->second += d;
      it->second.data += d;
    }
  }
};

// Manipulators — Statistics.hpp:596-626 walks types, instantiations, then the type subtree.
struct Manipulator {
  virtual bool inst(const Inst &, StatsData &) = 0;
  virtual bool type(int, InstStats &) = 0;
  virtual ~Manipulator() {}
};
static void manipulate(StatsCollection &sc, Manipulator &m) {
  std::vector<int> types_erase;
  std::vector<Inst> inst_erase;
  for (auto &ti : sc.types) {
    int recalc = 0;
    for (auto &ii : ti.second.stats) {
      recalc += m.inst(Inst(ti.first, ii.first), ii.second);
      if (ii.second.zero()) inst_erase.push_back(Inst(ti.first, ii.first));
    }
    recalc += m.type(ti.first, ti.second.stats);
    if (recalc) {  // InstStatsNode::RecalculateStats :440-447
      StatsData d;
      for (auto &ii : ti.second.stats) d += ii.second;
      ti.second.data = d;
    }
    if (ti.second.data.zero()) types_erase.push_back(ti.first);
  }
  for (auto &i : inst_erase) sc.types[i.first].stats.erase(i.second);
  for (int t : types_erase) sc.types.erase(t);
}
struct Scaler : Manipulator {  // Statistics.hpp:651-689
  double f;
  explicit Scaler(double f_) : f(f_) {}
  bool inst(const Inst &, StatsData &d) override { d.scale(f); return true; }
  bool type(int, InstStats &) override { return false; }
};
struct CoverageFilter : Manipulator {  // Statistics.hpp:697-756
  size_t nnz; double minc; std::set<Inst> &unf;
  CoverageFilter(size_t n, double m, std::set<Inst> &u) : nnz(n), minc(m), unf(u) {}
  bool inst(const Inst &i, StatsData &d) override {
    double cov = d.enc / (double)nnz;
    if (cov < minc) { d = StatsData(); return true; }
    unf.insert(i);
    return false;
  }
  bool type(int, InstStats &) override { return false; }
};
// Statistics.cpp:28-41
static void split_block_data(size_t fixed_dim, size_t var_dim, size_t max_var_dim, const StatsData &data,
                             InstStats &stats) {
  size_t nr_chunks = var_dim / max_var_dim;
  size_t rem_dim = var_dim % max_var_dim;
  size_t max_block_size = max_var_dim * fixed_dim;
  size_t nr_max_blocks = nr_chunks * data.pat;
  size_t rem_nr_nzeros = data.enc - nr_max_blocks * max_block_size;
  stats[max_var_dim] += StatsData(nr_max_blocks * max_block_size, nr_max_blocks, 0);
  if (rem_dim >= 2) stats[rem_dim] += StatsData(rem_nr_nzeros, data.pat, 0);
}
struct BlockSplitter : Manipulator {  // Statistics.hpp:778-822, Statistics.cpp:51-87
  size_t max_patt, nnz; double minc;
  BlockSplitter(size_t mp, size_t n, double m) : max_patt(mp), nnz(n), minc(m) {}
  bool inst(const Inst &, StatsData &) override { return false; }
  bool type(int t, InstStats &stats) override {
    if (!is_block(t)) return false;
    size_t fixed_dim = block_align(t);
    size_t max_block_dim = max_patt / fixed_dim;
    int ret = 0;
    std::vector<size_t> erase;
    InstStats::reverse_iterator i = stats.rbegin();
    for (; i != stats.rend() && i->first * fixed_dim > max_patt; ++i) {
      StatsData d = i->second;  // value copy: map insertion below never invalidates, but keep the read stable
      split_block_data(fixed_dim, i->first, max_block_dim, d, stats);
      erase.push_back(i->first);
      ++ret;
    }
    for (size_t d : erase) stats.erase(d);
    erase.clear();
    InstStats::reverse_iterator j = stats.rbegin();
    for (i = stats.rbegin(); i != stats.rend(); ++i) {
      if ((i->second.enc / (double)nnz) < minc) continue;
      for (; j != stats.rend() && j->first >= i->first && j->second.enc / (double)nnz < minc; ++j) {
        StatsData d = j->second;
        split_block_data(fixed_dim, j->first, i->first, d, stats);
        erase.push_back(j->first);
        ++ret;
      }
    }
    for (size_t d : erase) stats.erase(d);
    return ret != 0;
  }
};

// ------------------------------------------------------ EncodingManager --
struct RLE { size_t freq; int val; };
static void delta_encode(std::vector<int> &d) {  // EncodingManager.hpp:457-466
  for (size_t i = d.size() - 1; i > 0; --i) d[i] -= d[i - 1];
}
static void rl_encode(const std::vector<int> &in, std::vector<RLE> &out) {  // :475-502
  RLE r; r.freq = 1; r.val = in[0];
  for (size_t k = 1; k < in.size(); k++) {
    if (r.val != in[k]) { out.push_back(r); r.freq = 1; r.val = in[k]; } else ++r.freq;
  }
  out.push_back(r);
}

struct EncodingManager {
  Partition *spm;
  const Options &opt;
  size_t min_limit, max_limit;
  double min_perc;
  bool sampling_enabled = false;
  size_t sort_window_size = 0, samples_max = 0;
  bool split_blocks, onedim_blocks;
  std::vector<size_t> sort_splits;
  std::vector<int> sort_splits_nzeros;
  std::vector<size_t> selected_splits;
  std::set<Inst> encoded_inst;
  bool ignore[TypeMax];
  std::vector<int> cols_buff;
  std::vector<double> vals_buff;
  std::string *log;
  bool undefined_hit = false;

  // EncodingManager.hpp:560-619
  EncodingManager(Partition *p, const Options &o, std::string *lg)
      : spm(p), opt(o), min_limit(o.min_unit_size), max_limit(o.max_unit_size), min_perc(o.min_coverage),
        split_blocks(o.split_blocks), onedim_blocks(o.onedim_blocks), log(lg) {
    for (int i = 0; i < TypeMax; i++) ignore[i] = true;  // IgnoreAll
    sort_window_size = o.window_size;
    samples_max = o.nr_samples;
    if (o.sampling == "none") {
      sampling_enabled = false;
    } else if (o.sampling == "portion" || o.sampling == "window") {
      sampling_enabled = true;
      samples_max = (size_t)std::ceil((float)samples_max / o.nr_threads);  // :595
      if (samples_max == 0) throw OracleError("invalid samples number");
      if (o.sampling == "portion") {
        if (!(o.portion > 0 && o.portion <= 1)) throw OracleError("invalid sampling portion");
        sort_window_size = (size_t)(o.portion * spm->nr_nzeros / samples_max);  // :600-601
      } else if (sort_window_size == 0) {
        throw OracleError("invalid window size");
      }
      compute_sort_splits_by_nnz();
      if (samples_max > sort_splits.size()) samples_max = sort_splits.size();  // :608-609
      select_splits();
    } else {
      throw OracleError("invalid value \"" + o.sampling + "\" while setting property \"spx.preproc.sampling\"");
    }
  }

  // :1568-1599
  void compute_sort_splits_by_nnz() {
    size_t nzeros_cnt = 0;
    size_t nr_rows = spm->rowptr_size() - 1;
    sort_splits.push_back(0);
    for (size_t i = 0; i < nr_rows; ++i) {
      size_t new_cnt = nzeros_cnt + spm->rowptr[i + 1] - spm->rowptr[i];
      if (new_cnt < sort_window_size) {
        nzeros_cnt = new_cnt;
      } else {
        sort_splits.push_back(i + 1);
        sort_splits_nzeros.push_back((int)new_cnt);
        nzeros_cnt = 0;
      }
    }
    if (nzeros_cnt) {
      if (sort_splits_nzeros.empty())
        // reference writes through back() of an empty vector here (UB); only
        // reachable when the whole partition is smaller than one window.
        throw OracleError("undefined: sampling window larger than the partition (EncodingManager.hpp:1589-1591)");
      sort_splits_nzeros.back() += (int)nzeros_cnt;
      if (nzeros_cnt > sort_window_size / 2) {
        sort_splits.push_back(nr_rows);
      } else {
        sort_splits.pop_back();
        sort_splits.push_back(nr_rows);
      }
    }
  }

  // :1489-1516.  selected_valid[i] == false marks entries the reference leaves
  // uninitialised (nr_samples > nr_splits/2); reading one is undefined there
  // and is flagged here at the moment gen_all_stats would touch it.
  std::vector<char> selected_valid;
  void select_splits() {
    size_t nr_splits = sort_splits.size();
    size_t nr_samples = samples_max;
    selected_splits.assign(nr_samples, 0);
    selected_valid.assign(nr_samples, 0);
    if (nr_samples == nr_splits) {
      for (size_t i = 0; i < nr_splits; ++i) { selected_splits[i] = i; selected_valid[i] = 1; }
      return;
    }
    if (nr_samples > nr_splits / 2) {
      for (size_t i = 0; i < nr_splits / 2; ++i) { selected_splits[i] = i; selected_valid[i] = 1; }
      nr_samples -= nr_splits / 2;
      nr_splits -= nr_splits / 2;
    }
    size_t skip = nr_splits / (nr_samples + 1);
    for (size_t i = 0; i < nr_samples; ++i) { selected_splits[i] = (i + 1) * skip; selected_valid[i] = 1; }
  }

  void remove_ignore(int t) {  // :144-152
    if (!onedim_blocks && (t == BlockRow1 || t == BlockCol1)) return;
    ignore[t] = false;
  }
  void remove_ignore_group(int g) {
    std::vector<int> ts; group_types(g, ts);
    for (int t : ts) remove_ignore(t);
  }

  // :1321-1408 (non-NUMA branch).  Marker bookkeeping lands on copies in the
  // reference (Element.hpp:275-281) and never influences control flow.
  void update_stats(Partition *sp, std::vector<int> &xs, StatsCollection &stats) {
    size_t ba = block_align(sp->type);
    if (ba) { update_stats_block(sp->type, xs, ba, stats); return; }
    if (xs.empty()) return;
    std::vector<RLE> rles;
    delta_encode(xs);
    rl_encode(xs, rles);
    int col = 0;
    bool last_rle_patt = false;
    for (const RLE &rle : rles) {
      size_t real_limit = (col && !last_rle_patt) ? min_limit - 1 : min_limit;
      if (rle.freq > 1 && rle.freq >= real_limit) {
        size_t real_nnz = (col && !last_rle_patt) ? rle.freq + 1 : rle.freq;
        size_t rem_nnz = real_nnz % max_limit;
        size_t patt_nnz = real_nnz;
        size_t patt_npatterns = real_nnz / max_limit + (rem_nnz != 0);
        if (rem_nnz && rem_nnz < min_limit) { --patt_npatterns; patt_nnz -= rem_nnz; }
        stats.append(Inst(sp->type, (size_t)rle.val), StatsData(patt_nnz, patt_npatterns));
        last_rle_patt = true;
      } else {
        last_rle_patt = false;
      }
      col += rle.val;
    }
    xs.clear();
  }

  // :1410-1487
  void update_stats_block(int type, std::vector<int> &xs, size_t ba, StatsCollection &stats) {
    if (xs.empty()) return;
    std::vector<RLE> rles;
    delta_encode(xs);
    rl_encode(xs, rles);
    int unit_start = 0;
    for (const RLE &rle : rles) {
      unit_start += rle.val;
      if (rle.val == 1) {
        size_t nr_elem, skip_front;
        if (unit_start == 1) { skip_front = 0; nr_elem = rle.freq; }
        else {
          skip_front = (unit_start - 2) % ba;
          if (skip_front != 0) skip_front = ba - skip_front;
          nr_elem = rle.freq + 1;
        }
        if (nr_elem > skip_front) nr_elem -= skip_front; else nr_elem = 0;
        size_t other_dim = nr_elem / ba;
        if (other_dim >= 2) stats.append(Inst(type, other_dim), StatsData(other_dim * ba, 1));
      }
      unit_start += rle.val * ((int)rle.freq - 1);
    }
    xs.clear();
  }

  // :621-645 — every element (pattern or not) is a point; one buffer per row.
  void generate_stats(Partition *sp, size_t rs, size_t re, StatsCollection &stats) {
    for (size_t i = rs; i < re; ++i) {
      for (int k = sp->rowptr[i]; k < sp->rowptr[i + 1]; ++k) cols_buff.push_back(sp->elems[k].col);
      update_stats(sp, cols_buff, stats);
    }
  }

  // SparsePartition.hpp:775-839 GetWindow / PutWindow; returns false for an empty window
  bool get_window(size_t rs, size_t length, Partition &w, int &es_out) {
    if (rs + length > spm->rowptr_size() - 1) length = spm->rowptr_size() - rs - 1;
    int es = spm->rowptr[rs], ee = spm->rowptr[rs + length];
    if (es == ee) return false;
    w.elems.assign(spm->elems.begin() + es, spm->elems.begin() + ee);
    for (Elem &e : w.elems) e.row -= (int)rs;
    w.set_rowptr();
    w.nr_rows = length; w.nr_cols = spm->nr_cols; w.nr_nzeros = w.elems.size();
    w.row_start = spm->row_start + (int)rs; w.type = spm->type;
    es_out = es;
    return true;
  }
  void put_window(Partition &w, size_t rs, int es) {
    if (spm->type == Horizontal) for (Elem &e : w.elems) e.row += (int)rs;
    std::copy(w.elems.begin(), w.elems.end(), spm->elems.begin() + es);
  }

  // :707-813
  void gen_all_stats(StatsCollection &stats) {
    encoded_inst.clear();
    bool use_sampling = sampling_enabled && spm->rowptr_size() - 1 > samples_max;
    if (use_sampling) {
      size_t samples_nnz = 0;
      spm->transform(Horizontal);
      for (size_t i = 0; i < samples_max; ++i) {
        // Reads the reference performs on uninitialised / out-of-bounds data
        // (selected_splits_[i], sort_splits_[..+1], sort_splits_nzeros_[..]).
        bool undefined = !selected_valid[i] || selected_splits[i] + 1 >= sort_splits.size();
        if (!undefined) {
          size_t ws = sort_splits[selected_splits[i]], we = sort_splits[selected_splits[i] + 1];
          if (!(ws >= we - 1) && selected_splits[i] >= sort_splits_nzeros.size()) undefined = true;
        }
        if (undefined) {
          if (opt.undefined_sampling == "break") { undefined_hit = true; break; }
          throw OracleError("undefined: reference reads uninitialised sampling-window data for this "
                            "(rows, nnz, nr_samples) regime (EncodingManager.hpp:716-732,1489-1516)");
        }
        size_t window_start = sort_splits[selected_splits[i]];
        size_t window_end = sort_splits[selected_splits[i] + 1];
        size_t window_size = window_end - window_start;
        if (window_start >= window_end - 1) break;  // "quick fix for windows of size 0"
        if (window_start > spm->rowptr_size() - 1) {
          // after an encoding round the trailing rows can be empty and the rebuilt rowptr shorter than
          // the split table: GetWindow then reads rowptr_[rs] out of bounds (SparsePartition.hpp:780-786)
          if (opt.undefined_sampling == "break") { undefined_hit = true; break; }
          throw OracleError("undefined: sampling window starts past the rebuilt rowptr "
                            "(SparsePartition.hpp:780-786)");
        }
        Partition w; int es = 0;
        if (!get_window(window_start, window_size, w, es)) break;
        samples_nnz += sort_splits_nzeros[selected_splits[i]];
        for (int t = Horizontal; t < TypeMax; ++t) {
          if (ignore[t]) continue;
          w.transform(t);
          generate_stats(&w, 0, w.rowptr_size() - 1, stats);
        }
        w.transform(Horizontal);
        put_window(w, window_start, es);
      }
      if (samples_nnz) { Scaler sc(spm->nr_nzeros / (double)samples_nnz); manipulate(stats, sc); }
      if (split_blocks) { BlockSplitter bs(max_limit, spm->nr_nzeros, min_perc); manipulate(stats, bs); }
      CoverageFilter cf(spm->nr_nzeros, min_perc, encoded_inst);
      manipulate(stats, cf);
    } else {
      for (int t = Horizontal; t < TypeMax; ++t) {
        if (ignore[t]) continue;
        spm->transform(t);
        generate_stats(spm, 0, spm->rowptr_size() - 1, stats);
        if (block_align(t) && split_blocks) { BlockSplitter bs(max_limit, spm->nr_nzeros, min_perc); manipulate(stats, bs); }
        CoverageFilter cf(spm->nr_nzeros, min_perc, encoded_inst);
        manipulate(stats, cf);
      }
    }
  }

  // :815-861 (ratio heuristic)
  int choose_type(const StatsCollection &stats) {
    int ret = None;
    unsigned long max_score = 0;
    for (auto &ti : stats.types) {
      unsigned long score = ti.second.data.enc - ti.second.data.pat;
      if (score == 0) ignore[ti.first] = true;
      else if (score > max_score) { max_score = score; ret = ti.first; }
    }
    return ret;
  }

  static Elem make_pattern(int row, int col, const double *v, size_t size, int type, size_t delta) {
    Elem e; e.row = row; e.col = col;
    if (size == 1) { e.val = v[0]; return e; }  // Element.hpp:234-236
    e.type = type; e.delta = (uint32_t)delta; e.size = (uint32_t)size; e.vals.assign(v, v + size);
    return e;
  }
  static Elem make_single(int row, int col, double v) { Elem e; e.row = row; e.col = col; e.val = v; return e; }

  // :1003-1082
  void do_encode(int row_no, std::vector<int> &xs, std::vector<double> &vs, std::vector<Elem> &encoded) {
    int type = spm->type;
    if (is_block(type)) {
      if (!split_blocks) do_encode_block(row_no, xs, vs, encoded);
      else do_encode_block_alt(row_no, xs, vs, encoded);
      return;
    }
    size_t vi = 0;
    std::vector<RLE> rles;
    delta_encode(xs);
    rl_encode(xs, rles);
    int col = 0;
    for (const RLE &rle : rles) {
      size_t rle_freq = rle.freq, rle_start;
      if (rle_freq != 1 && encoded_inst.count(Inst(type, (size_t)rle.val))) {
        col += rle.val;
        if (col != rle.val) {
          rle_start = col;
          rle_freq = rle.freq;
          if (!encoded.back().is_pattern()) {  // include the previous element, too
            rle_start -= rle.val;
            rle_freq++;
            encoded.pop_back();
            --vi;
          }
        } else {
          rle_start = col;
          rle_freq = rle.freq;
        }
        while (rle_freq >= min_limit) {
          size_t curr_freq = std::min(max_limit, rle_freq);
          encoded.push_back(make_pattern(row_no, (int)rle_start, &vs[vi], curr_freq, type, (size_t)rle.val));
          vi += curr_freq;
          rle_start += rle.val * curr_freq;
          rle_freq -= curr_freq;
        }
        col = (int)(rle_start - rle.val);
      }
      for (size_t i = 0; i < rle_freq; ++i) {
        col += rle.val;
        encoded.push_back(make_single(row_no, col, vs[vi++]));
      }
    }
    if (vi != vs.size()) throw OracleError("do_encode: not all elements processed");
    xs.clear(); vs.clear();
  }

  // :1085-1192
  void do_encode_block(int row_no, std::vector<int> &xs, std::vector<double> &vs, std::vector<Elem> &encoded) {
    size_t vi = 0;
    std::vector<RLE> rles;
    delta_encode(xs);
    rl_encode(xs, rles);
    int type = spm->type;
    int ba = (int)block_align(type);
    int col = 0;
    for (const RLE &rle : rles) {
      size_t skip_front, skip_back, nr_elem;
      col += rle.val;
      if (col == 1) { skip_front = 0; nr_elem = rle.freq; }
      else {
        skip_front = (col - 2) % ba;
        if (skip_front != 0) skip_front = ba - skip_front;
        nr_elem = rle.freq + 1;
      }
      if (nr_elem > skip_front) nr_elem -= skip_front; else nr_elem = 0;
      skip_back = nr_elem % ba;
      if (nr_elem > skip_back) nr_elem -= skip_back; else nr_elem = 0;
      if (rle.val == 1 && encoded_inst.count(Inst(type, nr_elem / ba)) && nr_elem >= (size_t)2 * ba) {
        size_t rle_start;
        if (col != 1) { rle_start = col - 1; encoded.pop_back(); --vi; }
        else rle_start = col;
        for (size_t i = 0; i < skip_front; ++i) encoded.push_back(make_single(row_no, (int)(rle_start + i), vs[vi++]));
        size_t max_limit_a = max_limit / ba * ba;
        size_t nr_blocks = nr_elem / max_limit_a;
        size_t nr_elem_block = std::min(max_limit_a, nr_elem);
        if (nr_blocks == 0) nr_blocks = 1;
        else skip_back += nr_elem - nr_elem_block * nr_blocks;
        for (size_t i = 0; i < nr_blocks; ++i) {
          encoded.push_back(make_pattern(row_no, (int)(rle_start + skip_front + i * nr_elem_block), &vs[vi],
                                         nr_elem_block, type, nr_elem_block / ba));
          vi += nr_elem_block;
        }
        for (size_t i = 0; i < skip_back; ++i)
          encoded.push_back(make_single(row_no, (int)(rle_start + skip_front + nr_elem_block * nr_blocks + i), vs[vi++]));
      } else {
        for (size_t i = 0; i < rle.freq; ++i) encoded.push_back(make_single(row_no, col + (int)i * rle.val, vs[vi++]));
      }
      col += rle.val * ((int)rle.freq - 1);
    }
    if (vi != vs.size()) throw OracleError("do_encode_block: out of bounds");
    xs.clear(); vs.clear();
  }

  // :1194-1290
  void do_encode_block_alt(int row_no, std::vector<int> &xs, std::vector<double> &vs, std::vector<Elem> &encoded) {
    size_t vi = 0;
    std::vector<RLE> rles;
    delta_encode(xs);
    rl_encode(xs, rles);
    int type = spm->type;
    size_t ba = block_align(type);
    int col = 0;
    for (const RLE &rle : rles) {
      size_t skip_front, skip_back, nr_elem;
      col += rle.val;
      if (col == 1) { skip_front = 0; nr_elem = rle.freq; }
      else {
        skip_front = (col - 2) % ba;
        if (skip_front != 0) skip_front = ba - skip_front;
        nr_elem = rle.freq + 1;
      }
      if (nr_elem > skip_front) nr_elem -= skip_front; else nr_elem = 0;
      skip_back = nr_elem % ba;
      nr_elem -= skip_back;
      if (rle.val == 1 && nr_elem >= 2 * ba) {
        size_t rle_start;
        if (col != 1) { rle_start = col - 1; encoded.pop_back(); --vi; }
        else rle_start = col;
        for (size_t i = 0; i < skip_front; ++i) encoded.push_back(make_single(row_no, (int)rle_start++, vs[vi++]));
        size_t other_dim = nr_elem / ba;
        for (auto it = encoded_inst.rbegin(); it != encoded_inst.rend(); ++it) {
          if (it->first != type) continue;
          while (other_dim >= it->second) {
            size_t nr_elem_block = ba * it->second;
            encoded.push_back(make_pattern(row_no, (int)rle_start, &vs[vi], nr_elem_block, type, it->second));
            rle_start += nr_elem_block;
            vi += nr_elem_block;
            nr_elem -= nr_elem_block;
            other_dim -= it->second;
          }
        }
        skip_back += nr_elem;
        for (size_t i = 0; i < skip_back; ++i) encoded.push_back(make_single(row_no, (int)rle_start++, vs[vi++]));
      } else {
        for (size_t i = 0; i < rle.freq; ++i) encoded.push_back(make_single(row_no, col + (int)i * rle.val, vs[vi++]));
      }
      col += rle.val * ((int)rle.freq - 1);
    }
    if (vi != vs.size()) throw OracleError("do_encode_block_alt: out of bounds");
    xs.clear(); vs.clear();
  }

  // :1292-1319
  void encode_row(size_t rb, size_t re, std::vector<Elem> &newrow) {
    if (rb == re) return;
    int row_no = spm->elems[rb].row;
    for (size_t k = rb; k < re; ++k) {
      const Elem &e = spm->elems[k];
      if (!e.is_pattern()) { cols_buff.push_back(e.col); vals_buff.push_back(e.val); continue; }
      if (!cols_buff.empty()) do_encode(row_no, cols_buff, vals_buff, newrow);
      newrow.push_back(e);
    }
    if (!cols_buff.empty()) do_encode(row_no, cols_buff, vals_buff, newrow);
  }

  // :863-903
  void encode(int type) {
    if (type == None) return;
    spm->transform(type);
    std::vector<Elem> out, new_row;
    out.reserve(spm->elems.size());
    for (size_t i = 0; i + 1 < spm->rowptr_size(); ++i) {
      encode_row(spm->rowptr[i], spm->rowptr[i + 1], new_row);
      for (Elem &e : new_row) out.push_back(std::move(e));
      new_row.clear();
    }
    spm->elems.swap(out);
    spm->set_rowptr();
    ignore[type] = true;
  }

  static const char *tname(int t) {
    static const char *n[] = {"none", "h", "v", "d", "ad", "br1", "br2", "br3", "br4", "br5", "br6",
                              "br7", "br8", "bc1", "bc2", "bc3", "bc4", "bc5", "bc6", "bc7", "bc8"};
    return n[t];
  }

  // :905-960
  void encode_all() {
    if (!spm->nr_nzeros) return;
    for (;;) {
      StatsCollection st;
      gen_all_stats(st);
      int type = choose_type(st);
      if (type == None) break;
      if (log) {
        *log += std::string(tname(type)) + "{";
        bool first = true;
        for (auto &i : encoded_inst) if (i.first == type) { *log += (first ? "" : ",") + std::to_string(i.second); first = false; }
        *log += "} ";
      }
      encode(type);
    }
    spm->transform(Horizontal);
  }

  // :962-986
  void encode_serial(const EncSeq &seq) {
    if (!spm->nr_nzeros) return;
    for (int i = 0; i < TypeMax; i++) ignore[i] = true;
    for (auto &s : seq.seq) {
      if (s.first >= 100) throw OracleError("explicit xform sequence needs concrete types");
      remove_ignore(s.first);
      for (size_t d : s.second) encoded_inst.insert(Inst(s.first, d));
      encode(s.first);
      ignore[s.first] = true;
    }
    spm->transform(Horizontal);
  }
};

// ------------------------------------------------------------ CtlBuilder --
struct CtlBuilder {  // CtlBuilder.cpp:32-81
  std::vector<uint8_t> ctl;
  void var_int(unsigned long val) {
    for (;;) {
      uint8_t byte = val & 0x7f;
      if (val < 0x80) { ctl.push_back(byte); break; }
      ctl.push_back(byte | 0x80);
      val >>= 7;
    }
  }
  void fixed_int(unsigned long val, size_t nr_bytes) {
    for (size_t i = 0; i < nr_bytes; ++i) ctl.push_back((uint8_t)(val >> (8 * i)));
  }
  void head(bool nr, size_t rowjmp, uint8_t id, uint8_t size, size_t ucol, size_t ucol_size, bool full_colind) {
    uint8_t flag = id;
    if (nr) flag |= 1 << 7;      // CTL_NR_BIT   (CtlUtil.hpp:46-66)
    if (rowjmp) flag |= 1 << 6;  // CTL_RJMP_BIT
    ctl.push_back(flag);
    ctl.push_back(size);
    if (rowjmp) var_int(rowjmp);
    if (full_colind) fixed_int(ucol, ucol_size); else var_int(ucol);
  }
};

static size_t delta_size(size_t val) {  // Delta.hpp:35-48
  if ((uint8_t)val == val) return 1;
  if ((uint16_t)val == val) return 2;
  if ((uint32_t)val == val) return 4;
  return 8;
}

// ------------------------------------------------------------ CsxManager --
struct CsxManager {
  Partition *spm;
  bool full_colind;
  std::map<long, uint8_t> patterns;  // pattern id -> flag (CsxManager.hpp:60-71)
  uint8_t flag_avail = 0;
  bool row_jmps = false, new_row = false;
  uint64_t empty_rows = 0;
  int last_col = 0;
  size_t span = 0;
  CtlBuilder bld;
  std::vector<double> values;

  CsxManager(Partition *p, bool fc) : spm(p), full_colind(fc) {}

  uint8_t get_flag(unsigned long pid) {  // :237-258
    auto it = patterns.find((long)pid);
    if (it != patterns.end()) return it->second;
    uint8_t ret = flag_avail++;
    if (ret > 63) throw OracleError("too many patterns (CTL_PATTERNS_MAX)");
    patterns[(long)pid] = ret;
    return ret;
  }
  std::pair<bool, size_t> update_new_row() {  // :615-633
    bool nr = false; size_t rowjmp = 0;
    if (new_row) {
      nr = true; new_row = false;
      if (empty_rows != 0) { rowjmp = empty_rows + 1; empty_rows = 0; row_jmps = true; }
    }
    return std::make_pair(nr, rowjmp);
  }
  void add_cols(std::vector<int> &cols) {  // :635-682
    size_t n = cols.size();
    int last = cols[n - 1], col_start = cols[0];
    int prev = last_col;
    for (size_t i = 0; i < n; i++) { int tmp = cols[i]; cols[i] -= prev; prev = tmp; }  // DeltaEncode :212-224
    last_col = last;
    int mx = 0;
    if (n > 1) mx = *std::max_element(cols.begin() + 1, cols.end());
    size_t dbytes = delta_size((size_t)mx);
    unsigned long pid = (dbytes << 3);  // CsxUtil.cpp:27-33
    auto nri = update_new_row();
    int ucol = full_colind ? col_start - 1 : cols[0];
    bld.head(nri.first, nri.second, get_flag(pid), (uint8_t)n, (size_t)ucol, sizeof(int), full_colind);
    for (size_t i = 1; i < n; ++i) bld.fixed_int((unsigned long)cols[i], dbytes);
    cols.clear();
  }
  void add_pattern(const Elem &e) {  // :684-706, CsxUtil.hpp:58-74
    unsigned long pid = is_block(e.type) ? e.type * 10000UL + e.size / block_align(e.type)
                                         : e.type * 10000UL + e.delta;
    auto nri = update_new_row();
    int ucol = full_colind ? e.col - 1 : e.col - last_col;
    bld.head(nri.first, nri.second, get_flag(pid), (uint8_t)e.size, (size_t)ucol, sizeof(int), full_colind);
    // GetLastCol (Element.hpp:657-666); spm type is Horizontal here
    last_col = e.col;
    if (e.type == spm->type) last_col += (int)((e.size - 1) * e.delta);
  }
  void update_row_span(const Elem &e) {  // :452-496
    size_t s;
    if (e.type == Vertical || e.type == Diagonal || e.type == AntiDiagonal) s = (e.size - 1) * e.delta;
    else if (is_block_row(e.type)) s = e.type - BlockRow1;
    else if (is_block_col(e.type)) s = e.size / block_align(e.type) - 1;
    else s = 0;
    if (s > span) span = s;
  }
  void do_elems(size_t &k, size_t re, bool stop_at_boundary, std::vector<int> &cols) {  // body of DoRow/DoSymRow
    for (; k < re; ++k) {
      const Elem &e = spm->elems[k];
      if (stop_at_boundary && !(e.col < spm->row_start + 1)) break;
      if (e.is_pattern()) {
        update_row_span(e);
        if (!cols.empty()) add_cols(cols);
        add_pattern(e);
        values.insert(values.end(), e.vals.begin(), e.vals.end());
        continue;
      }
      if (cols.size() == 255) add_cols(cols);  // CTL_SIZE_MAX
      cols.push_back(e.col);
      values.push_back(e.val);
    }
    if (!cols.empty()) add_cols(cols);
  }
  void do_row(size_t rb, size_t re) {  // :504-541
    std::vector<int> cols;
    span = 0; last_col = 1;
    size_t k = rb;
    do_elems(k, re, false, cols);
  }
  void do_sym_row(size_t rb, size_t re) {  // :549-613
    std::vector<int> cols;
    span = 0; last_col = 1;
    size_t k = rb;
    do_elems(k, re, true, cols);
    do_elems(k, re, false, cols);
  }

  // :300-437
  void make_csx(bool symmetric, CsxPart &csx) {
    size_t nrows = spm->nr_rows;
    csx.rows_info.assign(nrows, RowInfo{0, 0, 0});
    csx.nnz = (long)spm->nr_nzeros; csx.nrows = (long)nrows; csx.ncols = (long)spm->nr_cols;
    csx.row_start = spm->row_start;
    new_row = false;
    size_t nrp = spm->rowptr_size() - 1;
    for (size_t i = 0; i < nrp; ++i) {
      size_t rb = spm->rowptr[i], re = spm->rowptr[i + 1];
      RowInfo &ri = csx.rows_info[i];
      if (rb == re) {
        if (!new_row) { ri.rowptr = 0; new_row = true; }
        else { empty_rows++; ri.rowptr = csx.rows_info[i - 1].rowptr; }
        ri.valptr = 0; ri.span = 0;
        continue;
      }
      ri.rowptr = (i > 0) ? (int)bld.ctl.size() : 0;
      ri.valptr = (int)values.size();
      if (symmetric) do_sym_row(rb, re); else do_row(rb, re);
      ri.span = (int)span;
      new_row = true;
    }
    for (size_t i = nrp; i < nrows; i++) {
      csx.rows_info[i].valptr = 0;
      csx.rows_info[i].rowptr = i ? csx.rows_info[i - 1].rowptr : 0;
      csx.rows_info[i].span = 0;
    }
    csx.row_jumps = row_jmps;
    csx.ctl = bld.ctl;
    if (values.size() != spm->nr_nzeros) throw OracleError("make_csx: value count mismatch");
    csx.values = values;
    csx.id_map.assign(patterns.size() + 1, -1);  // :439-450
    for (auto &p : patterns) csx.id_map[p.second] = p.first;
  }
};

// ----------------------------------------------------- partition building --
struct InputIter {
  const Coo &in; size_t pos = 0;
  explicit InputIter(const Coo &c) : in(c) {}
  bool end() const { return pos >= in.row.size(); }
};

// SparsePartition.hpp:508-541
static size_t set_elems(Partition &p, InputIter &it, int row_start1, size_t limit) {
  int row_prev = 1;
  size_t cnt = 0;
  p.rowptr.clear(); p.rowptr.push_back(0);
  for (; !it.end(); ++it.pos) {
    int row = it.in.row[it.pos] - row_start1 + 1;
    if (row != row_prev) {
      if (row < row_prev) throw OracleError("input not sorted by row");
      if (limit && cnt >= limit) break;
      for (int k = 0; k < row - row_prev; k++) p.rowptr.push_back((int)cnt);
      row_prev = row;
    }
    Elem e; e.row = row; e.col = it.in.col[it.pos]; e.val = it.in.val[it.pos];
    p.elems.push_back(e);
    cnt++;
  }
  if ((size_t)p.rowptr.back() != cnt) p.rowptr.push_back((int)cnt);
  return cnt;
}

struct PartitionSym {  // SparsePartition.hpp:358-497
  Partition lower, m1, m2;
  std::vector<double> diagonal;
};

// SparsePartition.hpp:1087-1129
static size_t set_elems_sym(PartitionSym &ps, InputIter &it, int row_start1, size_t limit) {
  Partition &p = ps.lower;
  int row_prev = 1;
  size_t cnt = 0, diag = 0;
  p.rowptr.clear(); p.rowptr.push_back(0);
  for (; !it.end(); ++it.pos) {
    int row = it.in.row[it.pos] - row_start1 + 1;
    int col = it.in.col[it.pos];
    if (row_start1 + row - 1 > col) {
      if (row != row_prev) {
        if (row < row_prev) throw OracleError("input not sorted by row");
        if (limit && diag + cnt >= limit && row_prev == row - 1) break;
        for (int k = 0; k < row - row_prev; k++) p.rowptr.push_back((int)cnt);
        row_prev = row;
      }
      Elem e; e.row = row; e.col = col; e.val = it.in.val[it.pos];
      p.elems.push_back(e);
      cnt++;
    } else if (row_start1 + row - 1 == col) {
      ps.diagonal.push_back(it.in.val[it.pos]);
      diag++;
    }
  }
  if ((size_t)p.rowptr.back() != cnt) p.rowptr.push_back((int)cnt);
  return cnt + diag;
}

// SparsePartition.hpp:965-1024
static void divide_matrix(PartitionSym &ps) {
  Partition &m = ps.lower, &m1 = ps.m1, &m2 = ps.m2;
  int row_start = m.row_start;
  int nr_rows = (int)m.rowptr_size() - 1;
  for (Partition *q : {&m1, &m2}) {
    q->type = Horizontal; q->row_start = row_start; q->nr_cols = m.nr_cols; q->nr_nzeros = 0;
    q->rowptr.clear(); q->rowptr.push_back(0);
  }
  int rows1 = 0, rows2 = 0;
  for (int i = 0; i < nr_rows; i++) {
    for (int j = m.rowptr[i]; j < m.rowptr[i + 1]; j++) {
      const Elem &e = m.elems[j];
      if (e.col < row_start + 1) {
        if (rows1 < i) { for (int k = 0; k < i - rows1; k++) m1.rowptr.push_back((int)m1.elems.size()); rows1 = i; }
        m1.nr_nzeros++; m1.elems.push_back(e);
      } else {
        if (rows2 < i) { for (int k = 0; k < i - rows2; k++) m2.rowptr.push_back((int)m2.elems.size()); rows2 = i; }
        m2.nr_nzeros++; m2.elems.push_back(e);
      }
    }
  }
  for (Partition *q : {&m1, &m2}) {
    if ((size_t)q->rowptr.back() != q->elems.size()) q->rowptr.push_back((int)q->elems.size());
    q->nr_rows = q->rowptr_size() - 1;
  }
}

// SparsePartition.hpp:1026-1074
static void merge_matrix(PartitionSym &ps) {
  Partition &m = ps.lower, &m1 = ps.m1, &m2 = ps.m2;
  Partition tmp;
  size_t nr_rows = m.rowptr_size() - 1;
  tmp.type = Horizontal; tmp.row_start = m.row_start; tmp.nr_cols = m.nr_cols; tmp.nr_nzeros = m.nr_nzeros;
  tmp.rowptr.push_back(0);
  for (size_t i = 0; i < nr_rows; i++) {
    if (m1.rowptr_size() - 1 > i) for (int j = m1.rowptr[i]; j < m1.rowptr[i + 1]; j++) tmp.elems.push_back(m1.elems[j]);
    if (m2.rowptr_size() - 1 > i) for (int j = m2.rowptr[i]; j < m2.rowptr[i + 1]; j++) tmp.elems.push_back(m2.elems[j]);
    tmp.rowptr.push_back((int)tmp.elems.size());
  }
  tmp.nr_rows = tmp.rowptr_size() - 1;
  m = tmp;
}

// CsxBuild.hpp:400-581
static void make_map(std::vector<PartitionSym> &ps, std::vector<CsxPart> &out) {
  unsigned ncpus = (unsigned)ps.size();
  unsigned n = (unsigned)ps[0].lower.nr_cols;
  std::vector<unsigned> count(n + 1, 0);
  std::vector<std::vector<char>> imap(ncpus, std::vector<char>(n + 1, 0));
  for (unsigned i = 0; i < ncpus; i++) {
    Partition &spm = ps[i].lower;
    unsigned start = spm.row_start;
    for (const Elem &e : spm.elems) {
      unsigned col = e.col;
      if (col < start + 1 && !imap[i][col]) { imap[i][col] = 1; count[col]++; }
    }
  }
  unsigned total = 0;
  for (unsigned i = 0; i < n; i++) total += count[i];
  unsigned end = 0, start;
  for (unsigned i = 0; i + 1 < ncpus; i++) {
    start = end;
    unsigned limit = total / (ncpus - i), temp = 0;
    while (temp < limit) temp += count[end++];
    total -= temp;
    for (unsigned j = start; j < end; j++)
      for (unsigned k = 0; k < ncpus; k++)
        if (imap[k][j]) { out[i].map_cpus.push_back(k); out[i].map_pos.push_back(j - 1); }
  }
  start = end; end = n;
  for (unsigned j = start; j < end; j++)
    for (unsigned k = 0; k < ncpus; k++)
      if (imap[k][j]) { out[ncpus - 1].map_cpus.push_back(k); out[ncpus - 1].map_pos.push_back(j - 1); }
}

// CsxBuild.hpp:134-166 / 204-242
static void run_encoder(Partition &p, const Options &opt, const EncSeq &seq, bool do_all, std::string *log) {
  EncodingManager mg(&p, opt, log);
  if (seq.is_explicit) mg.encode_serial(seq);
  else {
    for (auto &s : seq.seq) mg.remove_ignore_group(s.first);
    if (do_all) mg.encode_all();
  }
}

std::string tune(const Coo &in, const Options &opt, Tuned &out) {
  try {
    if (opt.nr_threads < 1) throw OracleError("invalid nr_threads");
    EncSeq seq(opt.xform);
    size_t nr = opt.nr_threads;
    out = Tuned();
    out.nrows = in.nrows; out.ncols = in.ncols; out.nnz = (long)in.row.size();
    out.symmetric = opt.symmetric; out.full_colind = opt.full_colind;
    out.parts.resize(nr);
    InputIter it(in);
    if (!opt.symmetric) {
      // SparseInternal.hpp:119-152
      std::vector<Partition> parts(nr);
      size_t nnz_total = in.row.size(), cnt = 0;
      int row_start = 0;
      for (size_t i = 0; i < nr; ++i) {
        Partition &p = parts[i];
        size_t limit = (nnz_total - cnt) / (nr - i);
        size_t nnz = set_elems(p, it, row_start + 1, limit);
        p.nr_nzeros = nnz; p.nr_rows = p.rowptr_size() - 1; p.nr_cols = in.ncols;
        p.row_start = row_start; p.type = Horizontal;
        row_start += (int)p.nr_rows;
        cnt += nnz;
      }
      if (cnt != nnz_total) throw OracleError("error in input matrix (matrix has less elements than claimed)");
      for (size_t i = 0; i < nr; ++i) {
        out.log += "p" + std::to_string(i) + ": ";
        run_encoder(parts[i], opt, seq, true, &out.log);
        out.log += "; ";
        CsxManager mg(&parts[i], opt.full_colind);
        mg.make_csx(false, out.parts[i]);
      }
    } else {
      if (in.nrows != in.ncols) throw OracleError("symmetric requires a square matrix");
      // SparseInternal.hpp:83-95
      std::vector<PartitionSym> ps(nr);
      size_t nnz_total = (in.row.size() + in.ncols) / 2, cnt = 0;
      int row_start = 0;
      for (size_t i = 0; i < nr; ++i) {
        PartitionSym &s = ps[i];
        size_t limit = (nnz_total - cnt) / (nr - i);
        size_t nnz = set_elems_sym(s, it, row_start + 1, limit);
        s.lower.nr_nzeros = nnz - s.diagonal.size();
        s.lower.nr_rows = s.lower.rowptr_size() - 1;
        s.lower.nr_cols = in.ncols;
        s.lower.row_start = row_start;
        s.lower.type = Horizontal;
        row_start += (int)s.diagonal.size();  // SparsePartitionSym::GetNrRows == diagonal_size_
        cnt += nnz;
      }
      if (cnt != nnz_total) throw OracleError("error in input matrix (matrix has less elements than claimed)");
      make_map(ps, out.parts);  // before preprocessing (CsxBuild.hpp:393-395)
      for (size_t i = 0; i < nr; ++i) {
        PartitionSym &s = ps[i];
        divide_matrix(s);
        out.log += "p" + std::to_string(i) + ": m1 ";
        run_encoder(s.m1, opt, seq, i != 0, &out.log);  // thread 0 skips m1 (CsxBuild.hpp:240)
        out.log += "m2 ";
        run_encoder(s.m2, opt, seq, true, &out.log);
        out.log += "; ";
        merge_matrix(s);
        CsxManager mg(&s.lower, opt.full_colind);
        mg.make_csx(true, out.parts[i]);
        out.parts[i].dvalues = s.diagonal;  // CsxManager.hpp:260-298
      }
    }
  } catch (std::exception &e) {
    return e.what();
  }
  return "";
}

// ---------------------------------------------------------------- MMF ---
static bool read_tokens(std::istream &in, std::vector<std::string> &args) {  // Mmf.cpp:27-55 (DoRead)
  std::string line;
  args.clear();
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    std::string t;
    while (ss >> t) args.push_back(t);
    if (!args.empty()) return true;
  }
  return false;
}

std::string load_mmf(const char *path, Coo &out) {
  std::ifstream in(path);
  if (!in.is_open()) return "MMF file error";
  std::vector<std::string> a;
  if (!read_tokens(in, a)) return "MMF file error";
  bool symmetric = false, col_wise = true, zero_based = false;
  int file_mode = 0;
  // Mmf.hpp:364-421
  if (a[0] != "%%MatrixMarket") {
    if (a[0].size() > 2 && a[0][0] == '%' && a[0][1] == '%') return "invalid header line in MMF file";
    file_mode = 1; col_wise = false;
  } else {
    if (a.size() < 5) return "less arguments in header line of MMF file";
    for (auto &t : a) for (auto &c : t) c = (char)tolower(c);
    if (a[1] != "matrix") return "unsupported object in header line of MMF file";
    if (a[2] != "coordinate") return "unsupported matrix format in header line of MMF file";
    if (a[4] == "general") symmetric = false;
    else if (a[4] == "symmetric") symmetric = true;
    else return "unsupported symmetry in header line of MMF file";
    for (size_t i = 5; i < a.size(); i++) {
      if (a[i] == "0-base") zero_based = true;
      else if (a[i] == "1-base") zero_based = false;
      else if (a[i] == "column") col_wise = true;
      else if (a[i] == "row") col_wise = false;
    }
  }
  // Mmf.hpp:423-443
  bool ignore_comments = file_mode && a[0][0] == '%';
  if (!file_mode || ignore_comments) {
    while (in.peek() == '%') { std::string skip; std::getline(in, skip); }
    if (!read_tokens(in, a)) return "size line error in MMF file";
  }
  if (a.size() != 3) return "bad input, less arguments in line of MMF file";
  long nr = atol(a[0].c_str()), nc = atol(a[1].c_str()), nnz = atol(a[2].c_str());
  out = Coo(); out.nrows = nr; out.ncols = nc;
  struct E { int r, c; double v; };
  std::vector<E> es;
  int rp = 0, cp = 0;
  for (long i = 0; i < nnz; i++) {
    if (!read_tokens(in, a)) return "Requesting dereference, but mmf ended.";
    if (a.size() != 3) return "bad input, less arguments in line of MMF file";
    E e{atoi(a[0].c_str()), atoi(a[1].c_str()), strtod(a[2].c_str(), nullptr)};
    if (zero_based) { e.r++; e.c++; }
    if (symmetric || col_wise) {  // Mmf.hpp:445-478
      es.push_back(e);
      if (symmetric && e.r != e.c) es.push_back(E{e.c, e.r, e.v});
    } else {  // streaming iterator enforces sortedness, Mmf.hpp:256-272
      if (e.r < rp || (e.r == rp && e.c < cp)) return "indices are not sorted in MMF file";
      cp = (e.r == rp) ? e.c : 1; rp = e.r;
      es.push_back(e);
    }
  }
  if (symmetric || col_wise)
    std::sort(es.begin(), es.end(), [](const E &x, const E &y) { return x.r < y.r || (x.r == y.r && x.c < y.c); });
  for (auto &e : es) { out.row.push_back(e.r); out.col.push_back(e.c); out.val.push_back(e.v); }
  return "";
}

// ------------------------------------------------- SpMV (unit templates) --
static inline uint64_t ul_get(const uint8_t *&ctl) {  // CtlUtil.hpp:110-133
  uint64_t ret = 0; unsigned shift = 0;
  for (;;) {
    uint8_t b = *ctl++;
    ret |= (uint64_t)(b & 0x7f) << shift;
    if (!(b & 0x80)) break;
    shift += 7;
  }
  return ret;
}
static inline uint64_t fixed_get(const uint8_t *&ctl, int bytes) {
  uint64_t v = 0;
  for (int i = 0; i < bytes; i++) v |= (uint64_t)ctl[i] << (8 * i);
  ctl += bytes;
  return v;
}

// csx_spmv_tmpl.c:66-101 + unit templates (delta/horiz/vert/diag/rdiag/block_row/block_col *_tmpl.c)
static void part_multiply(const CsxPart &csx, bool full_colind, const double *x, double *y, double scale_f) {
  if (csx.ctl.empty()) return;  // reference would read one unit regardless
  const double *v = csx.values.data();
  const uint8_t *ctl = csx.ctl.data(), *ctl_end = ctl + csx.ctl.size();
  int64_t x_curr = 0;
  double *y_curr = y + csx.row_start;
  double yr = 0;
  do {
    uint8_t flags = *ctl++, size = *ctl++;
    if (flags & 0x80) {
      *y_curr += yr; yr = 0;
      if (flags & 0x40) y_curr += ul_get(ctl); else y_curr++;
      x_curr = 0;
    }
    if (full_colind) x_curr = (int64_t)fixed_get(ctl, 4); else x_curr += (int64_t)ul_get(ctl);
    long pid = csx.id_map[flags & 0x3f];
    int type = (int)(pid / 10000); long d = pid % 10000;
    const double *xc = x + x_curr;
    if (type == None) {
      int w = (int)d / 8;
      double s = xc[0] * *v++;
      for (uint8_t i = 1; i < size; i++) { x_curr += (int64_t)fixed_get(ctl, w); s += x[x_curr] * *v++; }
      yr += s * scale_f;
    } else if (type == Horizontal) {
      double s = 0;
      for (long i = 0; i < d * size; i += d) s += xc[i] * *v++;
      x_curr += d * size - d;
      yr += s * scale_f;
    } else if (type == Vertical) {
      double xr = xc[0];
      for (long i = 0; i < d * size; i += d) y_curr[i] += xr * *v++ * scale_f;
    } else if (type == Diagonal) {
      for (long i = 0; i < d * size; i += d) y_curr[i] += xc[i] * *v++ * scale_f;
    } else if (type == AntiDiagonal) {
      for (long i = 0; i < d * size; i += d) y_curr[i] += xc[-i] * *v++ * scale_f;
    } else if (is_block_row(type)) {
      long r = (long)block_align(type), c = d;
      if (r == 1) { double s = 0; for (long i = 0; i < c; i++) s += xc[i] * *v++; yr += s * scale_f; }
      else for (long i = 0; i < c; i++) { double xr = xc[i]; for (long j = 0; j < r; j++) y_curr[j] += xr * *v++ * scale_f; }
    } else {
      long r = d, c = (long)block_align(type);
      if (c == 1) { double xr = xc[0]; for (long i = 0; i < r; i++) y_curr[i] += xr * *v++ * scale_f; }
      else for (long i = 0; i < r; i++) { double s = 0; for (long j = 0; j < c; j++) s += xc[j] * *v++; y_curr[i] += s * scale_f; }
    }
  } while (ctl < ctl_end);
  *y_curr += yr;
}

// csx_sym_spmv_tmpl.c:60-106 + *_sym_tmpl.c
static void part_multiply_sym(const CsxPart &csx, bool full_colind, const double *x, double *y, double *tmp,
                              double scale_f) {
  const double *v = csx.values.data();
  const double *dv = csx.dvalues.data();
  long x_indx = 0, y_indx = csx.row_start, y_end = csx.row_start + csx.nrows;
  double yr = 0;
  double *cur = tmp;
  if (!csx.ctl.empty()) {
    const uint8_t *ctl = csx.ctl.data(), *ctl_end = ctl + csx.ctl.size();
    do {
      uint8_t flags = *ctl++, size = *ctl++;
      if (flags & 0x80) {
        y[y_indx] += yr;
        long jmp = (flags & 0x40) ? (long)ul_get(ctl) : 1;  // CsxJit.hpp:373-394
        for (long i = 0; i < jmp; i++) { y[y_indx] += x[y_indx] * *dv * scale_f; y_indx++; dv++; }
        yr = 0; x_indx = 0; cur = tmp;
      }
      if (full_colind) x_indx = (long)fixed_get(ctl, 4); else x_indx += (long)ul_get(ctl);
      if (cur != y && x_indx >= csx.row_start) cur = y;
      long pid = csx.id_map[flags & 0x3f];
      int type = (int)(pid / 10000); long d = pid % 10000;
      double rx = x[y_indx];
      if (type == None) {
        int w = (int)d / 8;
        double s = x[x_indx] * *v; cur[x_indx] += rx * *v * scale_f; v++;
        for (uint8_t i = 1; i < size; i++) {
          x_indx += (long)fixed_get(ctl, w);
          s += x[x_indx] * *v; cur[x_indx] += rx * *v * scale_f; v++;
        }
        yr += s * scale_f;
      } else if (type == Horizontal) {
        double s = 0;
        for (long i = 0; i < d * size; i += d) { s += x[x_indx + i] * *v; cur[x_indx + i] += rx * *v * scale_f; v++; }
        x_indx += d * size - d;
        yr += s * scale_f;
      } else if (type == Vertical) {
        double xv = x[x_indx], ry = 0;
        for (long i = 0; i < d * size; i += d) { y[y_indx + i] += xv * *v * scale_f; ry += x[y_indx + i] * *v; v++; }
        cur[x_indx] += ry * scale_f;
      } else if (type == Diagonal) {
        for (long i = 0; i < d * size; i += d) {
          y[y_indx + i] += x[x_indx + i] * *v * scale_f; cur[x_indx + i] += x[y_indx + i] * *v * scale_f; v++;
        }
      } else if (type == AntiDiagonal) {
        for (long i = 0; i < d * size; i += d) {
          y[y_indx + i] += x[x_indx - i] * *v * scale_f; cur[x_indx - i] += x[y_indx + i] * *v * scale_f; v++;
        }
      } else if (is_block_row(type)) {
        long r = (long)block_align(type), c = d;
        for (long i = 0; i < c; i++) {
          double cx = x[x_indx + i], cry = 0;
          for (long j = 0; j < r; j++) { y[y_indx + j] += cx * *v * scale_f; cry += x[y_indx + j] * *v; v++; }
          cur[x_indx + i] += cry * scale_f;
        }
      } else {
        long r = d, c = (long)block_align(type);
        for (long i = 0; i < r; i++) {
          double cy = 0, crx = x[y_indx + i];
          for (long j = 0; j < c; j++) { cy += x[x_indx + j] * *v; cur[x_indx + j] += crx * *v * scale_f; v++; }
          y[y_indx + i] += cy * scale_f;
        }
      }
    } while (ctl < ctl_end);
    y[y_indx] += yr;
  }
  for (long i = y_indx; i < y_end; i++) { y[i] += x[i] * *dv * scale_f; dv++; }
}

// CsxKernels.cpp:35-129, CsxSpmv.cpp:28-86 (sequential over partitions; the
// barrier protocol only orders phases)
void spmv(const Tuned &A, double alpha, const double *x, double beta, double *y, bool overwrite) {
  size_t nt = A.parts.size();
  if (overwrite) for (long i = 0; i < A.nrows; i++) y[i] = 0;           // VecInit(y,0), CsxKernels.cpp:93
  if (!A.symmetric) {
    for (size_t t = 0; t < nt; t++) {
      const CsxPart &p = A.parts[t];
      if (!overwrite && beta != 1) for (long i = p.row_start; i < p.row_start + p.nrows; i++) y[i] *= beta;  // VecScalePart
      part_multiply(p, A.full_colind, x, y, alpha);
    }
    return;
  }
  std::vector<std::vector<double>> local(nt);
  for (size_t t = 1; t < nt; t++) local[t].assign(A.nrows, 0.0);  // matvec.c:302-318
  std::vector<double *> temp(nt);
  temp[0] = y;
  for (size_t t = 1; t < nt; t++) temp[t] = local[t].data();
  for (size_t t = 0; t < nt; t++)  // VecInitFromMap
    for (size_t k = 0; k < A.parts[t].map_cpus.size(); k++) temp[A.parts[t].map_cpus[k]][A.parts[t].map_pos[k]] = 0;
  if (!overwrite && beta != 1)
    for (size_t t = 0; t < nt; t++) {
      const CsxPart &p = A.parts[t];
      for (long i = p.row_start; i < p.row_start + p.nrows; i++) y[i] *= beta;
    }
  for (size_t t = 0; t < nt; t++) part_multiply_sym(A.parts[t], A.full_colind, x, y, temp[t], alpha);
  for (size_t t = 0; t < nt; t++)  // VecAddFromMap
    for (size_t k = 0; k < A.parts[t].map_cpus.size(); k++) {
      unsigned pos = A.parts[t].map_pos[k];
      y[pos] = y[pos] + temp[A.parts[t].map_cpus[k]][pos];
    }
}

// Second, independent walk of the grammar that yields the
// coordinates each value belongs to — the definition of "decoded column indices".
void decode_coords(const Tuned &A, int part, std::vector<int> &rows, std::vector<int> &cols) {
  const CsxPart &csx = A.parts[part];
  rows.clear(); cols.clear();
  if (csx.ctl.empty()) return;
  const uint8_t *ctl = csx.ctl.data(), *ctl_end = ctl + csx.ctl.size();
  long row = csx.row_start; int64_t col = 0;
  do {
    uint8_t flags = *ctl++, size = *ctl++;
    if (flags & 0x80) { row += (flags & 0x40) ? (long)ul_get(ctl) : 1; col = 0; }
    if (A.full_colind) col = (int64_t)fixed_get(ctl, 4); else col += (int64_t)ul_get(ctl);
    long pid = csx.id_map[flags & 0x3f];
    int type = (int)(pid / 10000); long d = pid % 10000;
    auto emit = [&](long r, long c) { rows.push_back((int)r); cols.push_back((int)c); };
    if (type == None) {
      emit(row, col);
      for (uint8_t i = 1; i < size; i++) { col += (int64_t)fixed_get(ctl, (int)d / 8); emit(row, col); }
    } else if (type == Horizontal) { for (long i = 0; i < size; i++) emit(row, col + i * d); col += (size - 1) * d; }
    else if (type == Vertical) for (long i = 0; i < size; i++) emit(row + i * d, col);
    else if (type == Diagonal) for (long i = 0; i < size; i++) emit(row + i * d, col + i * d);
    else if (type == AntiDiagonal) for (long i = 0; i < size; i++) emit(row + i * d, col - i * d);
    else if (is_block_row(type)) { long r = (long)block_align(type); for (long i = 0; i < d; i++) for (long j = 0; j < r; j++) emit(row + j, col + i); }
    else { long c = (long)block_align(type); for (long i = 0; i < d; i++) for (long j = 0; j < c; j++) emit(row + i, col + j); }
  } while (ctl < ctl_end);
}

}  // namespace csxo

namespace csxo {
void part_multiply_public(const CsxPart &csx, bool full_colind, const double *x, double *y, double scale_f) {
  part_multiply(csx, full_colind, x, y, scale_f);
}
}  // namespace csxo
