// CSX / CSX-Sym tuning stage of the B200 engine (host side of spx_mat_tune).
//
// Produces, per row partition, the CSX byte stream that SparseX would produce
// for the same input and options: substructure mining on sampled windows,
// greedy type selection, run encoding, ctl emission.  The data model is this
// engine's own (24-byte POD records over a shared value pool, LSD radix sort
// on packed (row, col) keys, streaming run detection); the decisions follow
// the reference and are cited inline (file:line into the SparseX tree):
//   partition split      SparseInternal.hpp:119-152, SparsePartition.hpp:508-541, 1087-1129
//   iteration orders     Xform.hpp:37-248, SparsePartition.hpp:661-744
//   statistics           EncodingManager.hpp:621-645, 707-813, 1321-1487; Statistics.hpp/.cpp
//   selection / encoding EncodingManager.hpp:815-1319
//   sampling windows     EncodingManager.hpp:560-619, 1489-1599
//   ctl emission         CsxManager.hpp:237-706, CtlBuilder.cpp:32-81, Delta.hpp:35-48
//   CSX-Sym              SparsePartition.hpp:965-1074, CsxBuild.hpp:204-288, 400-581
// Non-NUMA semantics (SPX_USE_NUMA == 0) throughout.
#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <thread>

#include "csx_host.hpp"

namespace spxb {
namespace {

// SPXB_TUNE_TRACE=1 prints phase timings of spx_mat_tune to stderr
struct Trace {
  const char *what; std::chrono::steady_clock::time_point t0;
  explicit Trace(const char *w) : what(w), t0(std::chrono::steady_clock::now()) {}
  ~Trace() {
    static const bool on = getenv("SPXB_TUNE_TRACE") != nullptr;
    if (on && what) fprintf(stderr, "[tune] %-22s %.3f s\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  }
};

struct TuneError : std::runtime_error {
  explicit TuneError(const std::string &m) : std::runtime_error(m) {}
};

// ---------------------------------------------------------------- records --
// A generic CSX element (reference: Element.hpp:192-608): a single non-zero
// (type == 0) or a substructure instance whose `size` values sit contiguously
// in the value pool starting at vptr.
struct Rec {
  int32_t r, c;      // 1-based coordinates in the partition's current order
  uint32_t delta;    // instantiation parameter: stride, or the free block dimension
  uint8_t type;      // EncType, 0 for singles
  uint8_t size;      // 1..255
  uint16_t pad;
  uint64_t vptr;
};
static_assert(sizeof(Rec) == 24, "Rec layout");

// resize() without the serial zero fill: the big arrays are written in full by parallel loops right afterwards
// (which also spreads the first touch of their pages over the threads)
template <class T>
struct default_init_alloc : std::allocator<T> {
  template <class U> struct rebind { using other = default_init_alloc<U>; };
  default_init_alloc() = default;
  template <class U> default_init_alloc(const default_init_alloc<U> &) {}
  template <class U, class... A>
  void construct(U *ptr, A &&...a) {
    if constexpr (sizeof...(A) == 0) ::new ((void *)ptr) U;
    else ::new ((void *)ptr) U(std::forward<A>(a)...);
  }
  // big arrays: 2 MB alignment and a transparent-huge-page hint (the first touch of gigabytes in 4 KB pages
  // costs more than filling them)
  T *allocate(size_t n) {
    const size_t bytes = n * sizeof(T);
    if (bytes >= (size_t(64) << 20)) {
      void *ptr = nullptr;
      if (posix_memalign(&ptr, size_t(2) << 20, bytes) == 0 && ptr) {
        madvise(ptr, bytes, MADV_HUGEPAGE);
        return (T *)ptr;
      }
    }
    void *ptr = malloc(bytes ? bytes : 1);
    if (!ptr) throw std::bad_alloc();
    return (T *)ptr;
  }
  void deallocate(T *ptr, size_t) { free(ptr); }
};
using RecVec = std::vector<Rec, default_init_alloc<Rec>>;
using ValVec = std::vector<double, default_init_alloc<double>>;

inline bool is_pattern(const Rec &e) { return e.delta != 0; }  // Element.hpp:372-377

// --------------------------------------------------------------- transforms --
struct RC { int32_t r, c; };
inline RC from_horiz(int to, RC p, int32_t R, int32_t C) {
  switch (to) {
    case T_HORIZ: return p;
    case T_VERT: return RC{p.c, p.r};
    case T_DIAG: return RC{R + p.c - p.r, p.c < p.r ? p.c : p.r};
    case T_ADIAG: { int32_t n = p.r + p.c - 1; return RC{n, n <= C ? p.r : C - p.c + 1}; }
  }
  if (is_brow(to)) { int k = blk_align(to); return RC{(p.r - 1) / k + 1, (p.r - 1) % k + k * (p.c - 1) + 1}; }
  int k = blk_align(to);
  return RC{(p.c - 1) / k + 1, (p.c - 1) % k + k * (p.r - 1) + 1};
}
inline RC to_horiz(int from, RC p, int32_t R, int32_t C) {
  switch (from) {
    case T_HORIZ: return p;
    case T_VERT: return RC{p.c, p.r};
    case T_DIAG: return p.r < R ? RC{R + p.c - p.r, p.c} : RC{p.c, p.r + p.c - R};
    case T_ADIAG: return p.r <= C ? RC{p.c, p.r - p.c + 1} : RC{p.r + p.c - C, C - p.c + 1};
  }
  int k = blk_align(from);
  RC t{k * (p.r - 1) + (p.c - 1) % k + 1, (p.c - 1) / k + 1};
  return is_brow(from) ? t : RC{t.c, t.r};
}
inline RC retarget(int from, int to, RC p, int32_t R, int32_t C) {
  if (from == to) return p;
  if (from == T_HORIZ) return from_horiz(to, p, R, C);
  if (to == T_HORIZ) return to_horiz(from, p, R, C);
  return from_horiz(to, to_horiz(from, p, R, C), R, C);
}

// ------------------------------------------------------------- radix sort --
// Slices [0, n) over worker threads (the sorts dominate spx_mat_tune on large partitions).
int g_sort_threads = 1;
// Arrays shorter than this are processed by one thread.  SPXB_PAR_MIN lowers it so that the test suite can run the
// threaded paths on small inputs (tests/test_cpu_refpin.py).
inline size_t par_min() {
  static const size_t v = getenv("SPXB_PAR_MIN") ? (size_t)atoll(getenv("SPXB_PAR_MIN")) : (size_t(1) << 20);
  return v;
}
template <class Fn>
void parallel_slices(size_t n, Fn fn) {
  int T = (n < par_min()) ? 1 : g_sort_threads;
  if (T <= 1) { fn(0, size_t(0), n); return; }
  std::vector<std::thread> th;
  size_t per = (n + T - 1) / T;
  for (int t = 0; t < T; t++) {
    size_t b = std::min(n, per * t), e = std::min(n, per * (t + 1));
    th.emplace_back([=]() { fn(t, b, e); });
  }
  for (auto &x : th) x.join();
}

// Stable LSD radix sort of (key, index) pairs; only the populated bit ranges
// of the packed (row << 32 | col) key are visited.  Each pass: per-thread
// histograms, one exclusive scan over (digit, thread), per-thread stable scatter.
using KeyVec = std::vector<uint64_t, default_init_alloc<uint64_t>>;
using IdxVec = std::vector<uint32_t, default_init_alloc<uint32_t>>;
void radix_sort_pairs(KeyVec &key, IdxVec &idx, int bits_lo, int bits_hi) {
  size_t n = key.size();
  KeyVec key2(n);
  IdxVec idx2(n);
  const int RB = 11;
  const size_t NB = size_t(1) << RB;
  int T = (n < par_min()) ? 1 : g_sort_threads;
  std::vector<size_t> hist((size_t)T * NB);
  auto pass = [&](int shift, int nbits) {
    uint64_t mask = (uint64_t(1) << nbits) - 1;
    std::fill(hist.begin(), hist.end(), 0);
    parallel_slices(n, [&](int t, size_t b, size_t e) {
      size_t *h = hist.data() + (size_t)t * NB;
      for (size_t i = b; i < e; i++) h[(key[i] >> shift) & mask]++;
    });
    size_t sum = 0;
    for (size_t d = 0; d <= mask; d++)
      for (int t = 0; t < T; t++) { size_t &h = hist[(size_t)t * NB + d]; size_t c = h; h = sum; sum += c; }
    parallel_slices(n, [&](int t, size_t b, size_t e) {
      size_t *h = hist.data() + (size_t)t * NB;
      for (size_t i = b; i < e; i++) {
        size_t d = h[(key[i] >> shift) & mask]++;
        key2[d] = key[i]; idx2[d] = idx[i];
      }
    });
    key.swap(key2); idx.swap(idx2);
  };
  for (int s = 0; s < bits_lo; s += RB) pass(s, std::min(RB, bits_lo - s));
  for (int s = 0; s < bits_hi; s += RB) pass(32 + s, std::min(RB, bits_hi - s));
}
inline int bit_width32(uint32_t v) { int b = 0; while (v) { b++; v >>= 1; } return b; }

// --------------------------------------------------------------- partition --
struct Part {
  int64_t nr_rows = 0, nr_cols = 0, nr_nzeros = 0;
  int type = T_NONE;
  int64_t row_start = 0;
  RecVec e;
  std::vector<int64_t> rowptr;   // rowptr[j] = #records with row <= j ; size = last row + 1
  ValVec *pool = nullptr;

  // The reference rebuilds a full rowptr after every reordering (SparsePartition.hpp:543-563, 852-891) — up to
  // nr_rows + nr_cols entries in the diagonal orders.  Only its length is ever used outside the Horizontal
  // order, so the array is materialised for Horizontal only; other orders are walked by runs of equal row.
  size_t nrowptr_ = 1;
  size_t nrowptr() const { return nrowptr_; }

  void build_rowptr() {
    rowptr.clear();
    if (e.empty()) { rowptr.push_back(0); nrowptr_ = 1; return; }
    int32_t last = e.back().r;
    nrowptr_ = (size_t)last + 1;
    if (type != T_HORIZ) return;
    rowptr.assign((size_t)last + 1, 0);
    const size_t n = e.size();
    if (n >= par_min()) {
      // rows are non-decreasing in horizontal order: rowptr[j] = index behind the last element of a row <= j,
      // written where the row number changes (disjoint ranges, one per change; threads share nothing)
      std::vector<char> bad(64, 0);
      parallel_slices(n, [&](int t, size_t b, size_t en) {
        for (size_t i = b; i < en; i++) {
          const int32_t r0 = e[i].r, r1 = i + 1 < n ? e[i + 1].r : last + 1;
          if (r1 < r0) { bad[t] = 1; return; }
          for (int32_t j = r0; j < r1 && j <= last; j++) rowptr[j] = (int64_t)(i + 1);
        }
      });
      bool ok = true;
      for (char c : bad) if (c) ok = false;
      if (ok) return;
      std::fill(rowptr.begin(), rowptr.end(), 0);
    }
    for (const Rec &x : e) rowptr[x.r]++;
    int64_t s = 0;
    for (size_t j = 1; j < rowptr.size(); j++) { s += rowptr[j]; rowptr[j] = s; }
  }

  // SparsePartition.hpp:680-744: re-coordinate, sort lexicographically, rebuild rowptr.
  void transform(int t) {
    if (type == t) return;
    Trace tr(e.size() >= (size_t(1) << 20) ? "transform" : nullptr);
    const int t2type = t;
    size_t n = e.size();
    if (n) {
      KeyVec key(n);
      IdxVec idx(n);
      const int from = type;
      const int32_t R = (int32_t)nr_rows, C = (int32_t)nr_cols;
      std::vector<uint32_t> mr(64, 0), mc(64, 0);
      std::vector<char> unsorted(64, 0);
      parallel_slices(n, [&](int t, size_t b, size_t en) {
        uint32_t maxr = 0, maxc = 0;
        uint64_t prev = 0;
        bool uns = false;
        for (size_t i = b; i < en; i++) {
          RC p = retarget(from, t2type, RC{e[i].r, e[i].c}, R, C);
          e[i].r = p.r; e[i].c = p.c;
          uint64_t k = (uint64_t(uint32_t(p.r)) << 32) | uint32_t(p.c);
          key[i] = k; idx[i] = (uint32_t)i;
          maxr = std::max(maxr, uint32_t(p.r)); maxc = std::max(maxc, uint32_t(p.c));
          if (i > b && k < prev) uns = true;
          prev = k;
        }
        mr[t] = maxr; mc[t] = maxc; unsorted[t] = uns;
      });
      uint32_t maxr = *std::max_element(mr.begin(), mr.end()), maxc = *std::max_element(mc.begin(), mc.end());
      bool sorted = true;
      for (char u : unsorted) if (u) sorted = false;
      if (sorted) {  // slices are sorted inside; check the seams
        int T = (n < par_min()) ? 1 : g_sort_threads;
        size_t per = (n + T - 1) / T;
        for (int t = 1; t < T && sorted; t++) { size_t b = std::min(n, per * t); if (b > 0 && b < n && key[b - 1] > key[b]) sorted = false; }
      }
      if (!sorted) {
        if (n < std::min<size_t>(32768, par_min())) {
          std::sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
        } else {
          radix_sort_pairs(key, idx, bit_width32(maxc), bit_width32(maxr));
        }
        RecVec out(n);
        parallel_slices(n, [&](int, size_t b, size_t en) { for (size_t i = b; i < en; i++) out[i] = e[idx[i]]; });
        e.swap(out);
      }
      type = t;
      build_rowptr();  // reference rebuilds only when elems_size_ != 0 (:741-742)
    }
    type = t;
  }
};

// -------------------------------------------------------------- statistics --
struct Stat {
  size_t enc = 0, pat = 0;
  Stat() {}
  Stat(size_t e, size_t p) : enc(e), pat(p) {}
  void add(const Stat &o) { enc += o.enc; pat += o.pat; }
  bool zero() const { return enc == 0 && pat == 0; }
};
typedef std::map<size_t, Stat> DimStats;
struct TypeStats { DimStats inst; Stat total; };
typedef std::map<int, TypeStats> Stats;  // Statistics.hpp:228-524 (ordered containers on purpose)
typedef std::pair<int, size_t> Inst;

inline void stats_add(Stats &s, int type, size_t d, const Stat &v) {
  auto it = s.find(type);
  if (it == s.end()) { TypeStats n; n.inst[d] = v; n.total = v; s[type] = n; return; }
  auto ii = it->second.inst.find(d);
  if (ii == it->second.inst.end()) it->second.inst[d] = v; else ii->second.add(v);
  it->second.total.add(v);
}

// StatsCollection::ManipulateStats (Statistics.hpp:596-626): per type, visit
// every instantiation, then the type subtree; recompute the type total when
// anything changed; drop zeroed instantiations / types afterwards.
template <class InstFn, class TypeFn>
void stats_visit(Stats &s, InstFn inst_fn, TypeFn type_fn) {
  std::vector<Inst> dead_inst;
  std::vector<int> dead_types;
  for (auto &t : s) {
    int changed = 0;
    for (auto &i : t.second.inst) {
      changed += inst_fn(t.first, i.first, i.second) ? 1 : 0;
      if (i.second.zero()) dead_inst.push_back(Inst(t.first, i.first));
    }
    changed += type_fn(t.first, t.second.inst) ? 1 : 0;
    if (changed) {
      Stat sum;
      for (auto &i : t.second.inst) sum.add(i.second);
      t.second.total = sum;
    }
    if (t.second.total.zero()) dead_types.push_back(t.first);
  }
  for (auto &d : dead_inst) s[d.first].inst.erase(d.second);
  for (int t : dead_types) s.erase(t);
}

// Statistics.cpp:28-41
void split_block(size_t fixed, size_t dim, size_t max_dim, Stat d, DimStats &st) {
  size_t chunks = dim / max_dim, rem = dim % max_dim;
  size_t big = max_dim * fixed, nbig = chunks * d.pat;
  size_t rem_nnz = d.enc - nbig * big;
  st[max_dim].add(Stat(nbig * big, nbig));
  if (rem >= 2) st[rem].add(Stat(rem_nnz, d.pat));
}
// Statistics.cpp:51-87 — keeps the reference's loop shape over an ordered map
// that is mutated while being walked in reverse.
bool block_splitter(int type, DimStats &st, size_t max_patt, size_t nnz, double minc) {
  if (!is_blk(type)) return false;
  size_t fixed = blk_align(type), max_dim = max_patt / fixed;
  int ret = 0;
  std::vector<size_t> dead;
  DimStats::reverse_iterator i = st.rbegin();
  for (; i != st.rend() && i->first * fixed > max_patt; ++i) {
    split_block(fixed, i->first, max_dim, i->second, st);
    dead.push_back(i->first);
    ++ret;
  }
  for (size_t d : dead) st.erase(d);
  dead.clear();
  DimStats::reverse_iterator j = st.rbegin();
  for (i = st.rbegin(); i != st.rend(); ++i) {
    if (i->second.enc / (double)nnz < minc) continue;
    for (; j != st.rend() && j->first >= i->first && j->second.enc / (double)nnz < minc; ++j) {
      split_block(fixed, j->first, i->first, j->second, st);
      dead.push_back(j->first);
      ++ret;
    }
  }
  for (size_t d : dead) st.erase(d);
  return ret != 0;
}

// ----------------------------------------------------------- xform string --
struct XformSeq {  // Encodings.cpp:108-138
  std::vector<std::pair<int, std::vector<size_t>>> items;  // type or group code (100 br, 101 bc, 102 all)
  bool is_explicit = false;
};
int xform_name(const std::string &s) {
  static const char *names[] = {"none", "h", "v", "d", "ad", "br1", "br2", "br3", "br4", "br5", "br6",
                                "br7", "br8", "bc1", "bc2", "bc3", "bc4", "bc5", "bc6", "bc7", "bc8"};
  for (int i = 0; i < T_MAX; i++) if (s == names[i]) return i;
  if (s == "br") return 100;
  if (s == "bc") return 101;
  if (s == "all") return 102;
  return -1;
}
const char *xform_short(int t) {
  static const char *names[] = {"none", "h", "v", "d", "ad", "br1", "br2", "br3", "br4", "br5", "br6",
                                "br7", "br8", "bc1", "bc2", "bc3", "bc4", "bc5", "bc6", "bc7", "bc8"};
  return names[t];
}
XformSeq parse_xform(const std::string &s) {
  XformSeq q;
  size_t i = 0, n = s.size();
  auto lower = [](char ch) { return ch >= 'a' && ch <= 'z'; };
  auto digit = [](char ch) { return ch >= '0' && ch <= '9'; };
  while (i < n) {
    if (!lower(s[i])) { i++; continue; }
    size_t j = i;
    while (j < n && lower(s[j])) j++;
    while (j < n && digit(s[j])) j++;
    std::string name = s.substr(i, j - i);
    int t = xform_name(name);
    if (t < 0) throw TuneError("invalid value \"" + name + "\" while setting property \"spx.preproc.xform\"");
    std::vector<size_t> deltas;
    if (j < n && s[j] == '{') {  // optional {d1,d2,...}
      size_t k = j + 1, st = k;
      std::vector<size_t> tmp;
      bool closed = false;
      while (k < n) {
        if (digit(s[k])) { k++; continue; }
        if ((s[k] == ',' || s[k] == '}') && k > st) {
          tmp.push_back(std::stoul(s.substr(st, k - st)));
          if (s[k] == '}') { closed = true; k++; break; }
          st = ++k;
          continue;
        }
        break;
      }
      if (closed) { deltas.swap(tmp); j = k; }
    }
    if (!deltas.empty()) q.is_explicit = true;
    q.items.push_back(std::make_pair(t, deltas));
    i = j;
  }
  return q;
}
void expand_group(int g, std::vector<int> &out) {  // Encodings.cpp:78-98
  if (g == 100) for (int t = T_BROW1; t <= T_BROW8; t++) out.push_back(t);
  else if (g == 101) for (int t = T_BCOL1; t <= T_BCOL8; t++) out.push_back(t);
  else if (g == 102) for (int t = T_NONE; t < T_MAX; t++) out.push_back(t);
  else out.push_back(g);
}

// ------------------------------------------------------------ run scanner --
struct Run { size_t freq; int32_t val; };
// Delta + run-length view of the column sequence of records [b, e) (the first
// delta is the absolute column, EncodingManager.hpp:457-502).
inline void scan_runs(const Rec *recs, size_t b, size_t e, std::vector<Run> &runs) {
  runs.clear();
  if (b == e) return;
  int32_t prev = 0;
  Run cur{0, 0};
  for (size_t k = b; k < e; k++) {
    int32_t d = recs[k].c - prev;
    prev = recs[k].c;
    if (cur.freq && cur.val == d) cur.freq++;
    else { if (cur.freq) runs.push_back(cur); cur.freq = 1; cur.val = d; }
  }
  runs.push_back(cur);
}

// ---------------------------------------------------------------- the miner --
class Miner {
 public:
  Miner(Part *p, const TuneOptions &o, std::string *log, bool *undefined)
      : spm_(p), opt_(o), min_(o.min_unit_size), max_(o.max_unit_size), minc_(o.min_coverage),
        log_(log), undefined_(undefined) {
    for (bool &b : ignore_) b = true;
    samples_ = o.nr_samples;
    window_ = o.window_size;
    if (o.sampling == "none") {
      sampling_ = false;
    } else if (o.sampling == "portion" || o.sampling == "window") {
      sampling_ = true;
      samples_ = (size_t)std::ceil((float)samples_ / o.nr_threads);  // EncodingManager.hpp:595
      if (samples_ == 0) throw TuneError("invalid number of samples");
      if (o.sampling == "portion") {
        if (!(o.portion > 0 && o.portion <= 1)) throw TuneError("invalid sampling portion");
        window_ = (size_t)(o.portion * spm_->nr_nzeros / samples_);      // :600-601
      } else if (window_ == 0) {
        throw TuneError("invalid window size");
      }
      split_by_nnz();
      if (samples_ > splits_.size()) samples_ = splits_.size();           // :608-609
      pick_splits();
    } else {
      throw TuneError("invalid value \"" + o.sampling + "\" while setting property \"spx.preproc.sampling\"");
    }
  }

  void allow(int t) {  // RemoveIgnore, :144-152 (one-dimensional blocks stay off: no mnemonic exists)
    if (t == T_BROW1 || t == T_BCOL1) return;
    ignore_[t] = false;
  }

  // EncodeAll, :905-960
  void mine_all() {
    if (!spm_->nr_nzeros) return;
    for (;;) {
      Stats st;
      gather_stats(st);
      int t = choose(st);
      if (t == T_NONE) break;
      if (log_) {
        *log_ += std::string(xform_short(t)) + "{";
        bool first = true;
        for (auto &i : chosen_) if (i.first == t) { *log_ += (first ? "" : ",") + std::to_string(i.second); first = false; }
        *log_ += "} ";
      }
      encode(t);
    }
    spm_->transform(T_HORIZ);
  }

  // EncodeSerial, :962-986
  void mine_serial(const XformSeq &q) {
    if (!spm_->nr_nzeros) return;
    for (bool &b : ignore_) b = true;
    for (auto &it : q.items) {
      if (it.first >= 100) throw TuneError("explicit xform sequences need concrete types");
      allow(it.first);
      for (size_t d : it.second) chosen_.insert(Inst(it.first, d));
      encode(it.first);
      ignore_[it.first] = true;
    }
    spm_->transform(T_HORIZ);
  }

 private:
  // DoComputeSortSplitsByNNZ, :1568-1599
  void split_by_nnz() {
    size_t acc = 0, nr = spm_->nrowptr() - 1;
    splits_.push_back(0);
    for (size_t i = 0; i < nr; ++i) {
      size_t nxt = acc + (size_t)(spm_->rowptr[i + 1] - spm_->rowptr[i]);
      if (nxt < window_) acc = nxt;
      else { splits_.push_back(i + 1); split_nnz_.push_back(nxt); acc = 0; }
    }
    if (acc) {
      if (split_nnz_.empty())
        throw TuneError("sampling window larger than the partition (undefined in the reference, "
                        "EncodingManager.hpp:1589-1591)");
      split_nnz_.back() += acc;
      if (acc > window_ / 2) splits_.push_back(nr);
      else { splits_.pop_back(); splits_.push_back(nr); }
    }
  }

  // SelectSplits, :1489-1516.  picked_ok_[i] is false for the entries the
  // reference leaves uninitialised.
  void pick_splits() {
    size_t ns = splits_.size(), want = samples_;
    picked_.assign(want, 0);
    picked_ok_.assign(want, 0);
    if (want == ns) { for (size_t i = 0; i < ns; i++) { picked_[i] = i; picked_ok_[i] = 1; } return; }
    if (want > ns / 2) {
      for (size_t i = 0; i < ns / 2; i++) { picked_[i] = i; picked_ok_[i] = 1; }
      want -= ns / 2; ns -= ns / 2;
    }
    size_t skip = ns / (want + 1);
    for (size_t i = 0; i < want; i++) { picked_[i] = (i + 1) * skip; picked_ok_[i] = 1; }
  }

  // UpdateStats, :1321-1408 — linear types.  A run that is neither first in
  // the row nor preceded by a detected run absorbs the element before it.
  void stats_linear(int type, const std::vector<Run> &runs, Stats &st) {
    bool started = false, prev_hit = false;
    for (const Run &r : runs) {
      bool absorb = started && !prev_hit;
      size_t need = absorb ? min_ - 1 : min_;
      if (r.freq > 1 && r.freq >= need) {
        size_t nnz = absorb ? r.freq + 1 : r.freq;
        size_t rem = nnz % max_;
        size_t units = nnz / max_ + (rem != 0);
        size_t covered = nnz;
        if (rem && rem < min_) { --units; covered -= rem; }
        stats_add(st, type, (size_t)r.val, Stat(covered, units));
        prev_hit = true;
      } else {
        prev_hit = false;
      }
      if (r.val) started = true;  // `col += rle.val` turns non-zero (:1403)
    }
  }
  // UpdateStatsBlock, :1410-1487
  void stats_block(int type, const std::vector<Run> &runs, Stats &st) {
    size_t a = blk_align(type);
    int64_t pos = 0;
    for (const Run &r : runs) {
      pos += r.val;
      if (r.val == 1) {
        size_t cnt, skip;
        if (pos == 1) { skip = 0; cnt = r.freq; }
        else {
          skip = (size_t)(pos - 2) % a;
          if (skip) skip = a - skip;
          cnt = r.freq + 1;
        }
        cnt = cnt > skip ? cnt - skip : 0;
        size_t other = cnt / a;
        if (other >= 2) stats_add(st, type, other, Stat(other * a, 1));
      }
      pos += (int64_t)r.val * ((int64_t)r.freq - 1);
    }
  }
  // GenerateStats, :621-645: every record of a row (substructure or not) is a point.
  void stats_of(Part &p, Stats &st) {
    const Rec *recs = p.e.data();
    size_t n = p.e.size(), b = 0;
    bool blk = is_blk(p.type);
    while (b < n) {
      size_t e = b + 1;
      while (e < n && recs[e].r == recs[b].r) e++;
      scan_runs(recs, b, e, runs_);
      if (blk) stats_block(p.type, runs_, st); else stats_linear(p.type, runs_, st);
      b = e;
    }
  }

  void filter_coverage(Stats &st) {  // CoverageFilter, Statistics.hpp:697-756
    size_t nnz = spm_->nr_nzeros;
    stats_visit(st,
                [&](int t, size_t d, Stat &v) {
                  if (v.enc / (double)nnz < minc_) { v = Stat(); return true; }
                  chosen_.insert(Inst(t, d));
                  return false;
                },
                [](int, DimStats &) { return false; });
  }
  void split_blocks(Stats &st) {  // BlockSplitter
    size_t nnz = spm_->nr_nzeros;
    stats_visit(st, [](int, size_t, Stat &) { return false; },
                [&](int t, DimStats &d) { return block_splitter(t, d, max_, nnz, minc_); });
  }

  // GenAllStats, :707-813
  void gather_stats(Stats &st) {
    Trace tr("gather_stats");
    chosen_.clear();
    if (sampling_ && spm_->nrowptr() - 1 > samples_) {
      size_t sampled = 0;
      spm_->transform(T_HORIZ);
      for (size_t i = 0; i < samples_; i++) {
        // The reference indexes sort_splits_/sort_splits_nzeros_ with entries
        // it never initialised in some (rows, nnz, nr_samples) regimes.  That
        // is undefined there; this engine ends the sampling loop at that
        // point (what the reference does for an empty window) and records it.
        bool undef = !picked_ok_[i] || picked_[i] + 1 >= splits_.size();
        if (!undef) {
          size_t a = splits_[picked_[i]], b = splits_[picked_[i] + 1];
          if (!(a >= b - 1) && picked_[i] >= split_nnz_.size()) undef = true;
        }
        if (undef) { if (undefined_) *undefined_ = true; break; }
        size_t ws = splits_[picked_[i]], we = splits_[picked_[i] + 1];
        if (ws >= we - 1) break;                       // windows of one row end the sampling (:720-722)
        if (ws > spm_->nrowptr() - 1) {                // window starts past the rebuilt rowptr: the reference
          if (undefined_) *undefined_ = true;          // reads rowptr_[rs] out of bounds here (undefined)
          break;
        }
        Part w;
        if (!window(ws, we - ws, w)) break;            // empty window (:726-729)
        sampled += split_nnz_[picked_[i]];
        for (int t = T_HORIZ; t < T_MAX; t++) {
          if (ignore_[t]) continue;
          w.transform(t);
          stats_of(w, st);
        }
      }
      if (sampled) {  // StatsDataScaler, Statistics.hpp:135-144, 651-689
        double f = spm_->nr_nzeros / (double)sampled;
        stats_visit(st, [&](int, size_t, Stat &v) { v.enc = (size_t)(v.enc * f); v.pat = (size_t)(v.pat * f); return true; },
                    [](int, DimStats &) { return false; });
      }
      if (opt_.split_blocks) split_blocks(st);
      filter_coverage(st);
    } else {
      for (int t = T_HORIZ; t < T_MAX; t++) {
        if (ignore_[t]) continue;
        spm_->transform(t);
        stats_of(*spm_, st);
        if (is_blk(t) && opt_.split_blocks) split_blocks(st);
        filter_coverage(st);
      }
    }
  }

  // GetWindow, SparsePartition.hpp:775-816 (the window is a private copy here;
  // the reference moves the records out and back, which leaves spm_ unchanged)
  bool window(size_t rs, size_t len, Part &w) {
    if (rs + len > spm_->nrowptr() - 1) len = spm_->nrowptr() - rs - 1;
    int64_t es = spm_->rowptr[rs], ee = spm_->rowptr[rs + len];
    if (es == ee) return false;
    w.e.assign(spm_->e.begin() + es, spm_->e.begin() + ee);
    for (Rec &x : w.e) x.r -= (int32_t)rs;
    w.build_rowptr();
    w.nr_rows = (int64_t)len; w.nr_cols = spm_->nr_cols; w.nr_nzeros = (int64_t)w.e.size();
    w.type = spm_->type; w.pool = spm_->pool;
    return true;
  }

  // ChooseType / GetTypeScore (ratio heuristic), :815-861
  int choose(const Stats &st) {
    int best = T_NONE;
    unsigned long best_score = 0;
    for (auto &t : st) {
      unsigned long score = t.second.total.enc - t.second.total.pat;
      if (score == 0) ignore_[t.first] = true;
      else if (score > best_score) { best_score = score; best = t.first; }
    }
    return best;
  }

  // ---- encoding --------------------------------------------------------
  Rec single(int32_t row, int32_t col, uint64_t vptr) { return Rec{row, col, 0, 0, 1, 0, vptr}; }
  // Build a substructure record from buffer members [m0, m0+cnt): their values
  // are copied behind each other at the end of the pool.
  // `lp` (threaded encoding): the values go to the thread's own pool and the record is tagged; encode() moves them
  // behind the shared pool afterwards and rewrites the tagged pointers.
  static constexpr uint64_t LOCAL_VPTR = uint64_t(1) << 63;
  Rec pattern(int32_t row, int32_t col, const Rec *buf, size_t m0, size_t cnt, int type, size_t delta, ValVec *lp) {
    if (cnt == 1) return single(row, col, buf[m0].vptr);  // Element.hpp:234-236
    ValVec &pool = *spm_->pool;
    if (lp) {
      uint64_t at = lp->size();
      for (size_t k = 0; k < cnt; k++) lp->push_back(pool[buf[m0 + k].vptr]);
      return Rec{row, col, (uint32_t)delta, (uint8_t)type, (uint8_t)cnt, 0, at | LOCAL_VPTR};
    }
    uint64_t at = pool.size();
    for (size_t k = 0; k < cnt; k++) pool.push_back(pool[buf[m0 + k].vptr]);
    return Rec{row, col, (uint32_t)delta, (uint8_t)type, (uint8_t)cnt, 0, at};
  }

  // DoEncode, :1003-1082 — buf[0..n) are the consecutive singles of one row.
  void encode_linear(int32_t row, const Rec *buf, size_t n, RecVec &out, std::vector<Run> &runs, ValVec *lp) {
    int type = spm_->type;
    scan_runs(buf, 0, n, runs);
    size_t vi = 0;
    int64_t col = 0;
    for (const Run &r : runs) {
      size_t left = r.freq;
      if (left != 1 && chosen_.count(Inst(type, (size_t)r.val))) {
        col += r.val;
        int64_t start = col;
        if (col != r.val && !is_pattern(out.back())) {  // pull in the element before the run
          start -= r.val; left++; out.pop_back(); --vi;
        }
        while (left >= min_) {
          size_t take = std::min(max_, left);
          out.push_back(pattern(row, (int32_t)start, buf, vi, take, type, (size_t)r.val, lp));
          vi += take; start += (int64_t)r.val * (int64_t)take; left -= take;
        }
        col = start - r.val;
      }
      for (size_t k = 0; k < left; k++) { col += r.val; out.push_back(single(row, (int32_t)col, buf[vi++].vptr)); }
    }
    if (vi != n) throw TuneError("internal: encode_linear consumed " + std::to_string(vi) + " of " + std::to_string(n));
  }

  // DoEncodeBlock (:1085-1192, split_blocks=false) and DoEncodeBlockAlt (:1194-1290)
  void encode_block(int32_t row, const Rec *buf, size_t n, RecVec &out, std::vector<Run> &runs, ValVec *lp) {
    int type = spm_->type;
    size_t a = blk_align(type);
    scan_runs(buf, 0, n, runs);
    size_t vi = 0;
    int64_t col = 0;
    for (const Run &r : runs) {
      size_t skip_front, skip_back, cnt;
      col += r.val;
      if (col == 1) { skip_front = 0; cnt = r.freq; }
      else {
        skip_front = (size_t)(col - 2) % a;
        if (skip_front) skip_front = a - skip_front;
        cnt = r.freq + 1;
      }
      cnt = cnt > skip_front ? cnt - skip_front : 0;
      skip_back = cnt % a;
      bool hit;
      if (opt_.split_blocks) { cnt -= skip_back; hit = r.val == 1 && cnt >= 2 * a; }
      else {
        cnt = cnt > skip_back ? cnt - skip_back : 0;
        hit = r.val == 1 && chosen_.count(Inst(type, cnt / a)) && cnt >= 2 * a;
      }
      if (hit) {
        int64_t start = col;
        if (col != 1) { start = col - 1; out.pop_back(); --vi; }
        for (size_t k = 0; k < skip_front; k++) out.push_back(single(row, (int32_t)start++, buf[vi++].vptr));
        if (opt_.split_blocks) {
          size_t other = cnt / a;  // carve with the surviving dims of this type, largest first
          for (auto it = chosen_.rbegin(); it != chosen_.rend(); ++it) {
            if (it->first != type) continue;
            while (other >= it->second) {
              size_t take = a * it->second;
              out.push_back(pattern(row, (int32_t)start, buf, vi, take, type, it->second, lp));
              start += (int64_t)take; vi += take; cnt -= take; other -= it->second;
            }
          }
          skip_back += cnt;
        } else {
          size_t cap = max_ / a * a;
          size_t nblocks = cnt / cap, per = std::min(cap, cnt);
          if (nblocks == 0) nblocks = 1; else skip_back += cnt - per * nblocks;
          for (size_t b = 0; b < nblocks; b++) {
            out.push_back(pattern(row, (int32_t)start, buf, vi, per, type, per / a, lp));
            start += (int64_t)per; vi += per;
          }
        }
        for (size_t k = 0; k < skip_back; k++) out.push_back(single(row, (int32_t)start++, buf[vi++].vptr));
      } else {
        for (size_t k = 0; k < r.freq; k++) out.push_back(single(row, (int32_t)(col + (int64_t)k * r.val), buf[vi++].vptr));
      }
      col += (int64_t)r.val * ((int64_t)r.freq - 1);
    }
    if (vi != n) throw TuneError("internal: encode_block consumed " + std::to_string(vi) + " of " + std::to_string(n));
  }

  // EncodeRow, :1292-1319, for the records [b0, b1) (whole rows)
  void encode_rows(const Rec *recs, size_t b0, size_t b1, bool blk, RecVec &out, std::vector<Run> &runs, ValVec *lp) {
    size_t b = b0;
    while (b < b1) {
      size_t e = b + 1;
      while (e < b1 && recs[e].r == recs[b].r) e++;
      int32_t row = recs[b].r;
      size_t k = b;
      while (k < e) {
        if (is_pattern(recs[k])) { out.push_back(recs[k++]); continue; }
        size_t m = k;
        while (m < e && !is_pattern(recs[m])) m++;
        if (blk) encode_block(row, recs + k, m - k, out, runs, lp); else encode_linear(row, recs + k, m - k, out, runs, lp);
        k = m;
      }
      b = e;
    }
  }

  // Encode, :863-903.  Rows are independent: big partitions are cut at row boundaries and encoded by several threads,
  // each into its own record list and value pool, which are then put behind each other in row order (the final CSX
  // arrays depend on the order of the records only, not on where a substructure's values sit in the pool).
  void encode(int t) {
    if (t == T_NONE) return;
    spm_->transform(t);
    Trace tr("encode rows");
    const Rec *recs = spm_->e.data();
    const size_t n = spm_->e.size();
    bool blk = is_blk(t);
    const int T = (n < par_min()) ? 1 : g_sort_threads;
    if (T <= 1) {
      RecVec out;
      out.reserve(n);
      encode_rows(recs, 0, n, blk, out, runs_, nullptr);
      spm_->e.swap(out);
    } else {
      std::vector<size_t> cut((size_t)T + 1, n);
      cut[0] = 0;
      for (int k = 1; k < T; k++) {
        size_t b = std::max(cut[k - 1], n / T * k);
        while (b < n && b > 0 && recs[b].r == recs[b - 1].r) b++;   // to the next row start
        cut[k] = b;
      }
      std::vector<RecVec> outs((size_t)T);
      std::vector<ValVec> pools((size_t)T);
      std::vector<std::string> errs((size_t)T);
      std::vector<std::thread> th;
      for (int k = 0; k < T; k++)
        th.emplace_back([&, k]() {
          try {
            std::vector<Run> runs;
            outs[k].reserve(cut[k + 1] - cut[k]);
            encode_rows(recs, cut[k], cut[k + 1], blk, outs[k], runs, &pools[k]);
          } catch (std::exception &ex) { errs[k] = ex.what(); }
        });
      for (auto &x : th) x.join();
      for (auto &er : errs) if (!er.empty()) throw TuneError(er);
      ValVec &pool = *spm_->pool;
      std::vector<size_t> obase((size_t)T + 1, 0), pbase((size_t)T + 1, pool.size());
      for (int k = 0; k < T; k++) { obase[k + 1] = obase[k] + outs[k].size(); pbase[k + 1] = pbase[k] + pools[k].size(); }
      pool.resize(pbase[T]);
      RecVec out(obase[T]);
      th.clear();
      for (int k = 0; k < T; k++)
        th.emplace_back([&, k]() {
          std::copy(pools[k].begin(), pools[k].end(), pool.begin() + (std::ptrdiff_t)pbase[k]);
          Rec *dst = out.data() + obase[k];
          for (size_t i = 0; i < outs[k].size(); i++) {
            Rec r = outs[k][i];
            if (r.vptr & LOCAL_VPTR) r.vptr = (r.vptr & ~LOCAL_VPTR) + pbase[k];
            dst[i] = r;
          }
        });
      for (auto &x : th) x.join();
      spm_->e.swap(out);
    }
    spm_->build_rowptr();
    ignore_[t] = true;
  }

  Part *spm_;
  const TuneOptions &opt_;
  size_t min_, max_;
  double minc_;
  bool sampling_ = false;
  size_t samples_ = 0, window_ = 0;
  std::vector<size_t> splits_, split_nnz_, picked_;
  std::vector<char> picked_ok_;
  std::set<Inst> chosen_;   // encoded_inst_ (ordered: reverse walk in encode_block)
  bool ignore_[T_MAX];
  std::vector<Run> runs_;
  std::string *log_;
  bool *undefined_;
};

// ----------------------------------------------------------- ctl emission --
inline size_t delta_bytes(uint64_t v) {  // Delta.hpp:35-48
  return v <= 0xff ? 1 : (v <= 0xffff ? 2 : (v <= 0xffffffffull ? 4 : 8));
}

class CtlWriter {
 public:
  CtlWriter(Part *p, bool full_colind, bool want_rows_info)
      : spm_(p), full_(full_colind), want_ri_(want_rows_info) {}

  // MakeCsx, CsxManager.hpp:300-437
  void run(bool sym, CsxPartition &out) {
    Trace tr("ctl emission");
    size_t nrows = (size_t)spm_->nr_rows;
    out.nnz = spm_->nr_nzeros; out.nrows = spm_->nr_rows; out.ncols = spm_->nr_cols;
    out.row_start = spm_->row_start;
    out.values.reserve((size_t)spm_->nr_nzeros);
    if (want_ri_) out.rows_info.assign(nrows, RowInfo64{0, 0, 0});
    ctl_ = &out.ctl; vals_ = &out.values;
    size_t nrp = spm_->nrowptr() - 1;
    int64_t prev_rowptr = 0;
    for (size_t i = 0; i < nrp; i++) {
      size_t b = spm_->rowptr[i], e = spm_->rowptr[i + 1];
      int64_t rp;
      if (b == e) {
        if (!new_row_) { rp = 0; new_row_ = true; } else { empty_rows_++; rp = prev_rowptr; }
        if (want_ri_) out.rows_info[i] = RowInfo64{rp, 0, 0};
        prev_rowptr = rp;
        continue;
      }
      rp = i ? (int64_t)ctl_->size() : 0;
      int64_t vp = (int64_t)vals_->size();
      span_ = 0; last_col_ = 1;
      size_t k = b;
      if (sym) emit(k, e, true);   // DoSymRow: columns left of the partition first (:559-584)
      emit(k, e, false);
      if (want_ri_) out.rows_info[i] = RowInfo64{rp, vp, (int32_t)span_};
      prev_rowptr = rp;
      new_row_ = true;
    }
    if (want_ri_) for (size_t i = nrp; i < nrows; i++) out.rows_info[i] = RowInfo64{i ? out.rows_info[i - 1].rowptr : 0, 0, 0};
    if ((int64_t)vals_->size() != spm_->nr_nzeros) throw TuneError("internal: value count mismatch in ctl emission");
    out.row_jumps = row_jumps_;
    out.id_map.assign(ids_.size() + 1, -1);  // AddMappings, :439-450
    for (auto &p : ids_) out.id_map[p.second] = p.first;
  }

 private:
  void varint(uint64_t v) {  // CtlBuilder.cpp:32-48
    for (;;) {
      uint8_t b = v & 0x7f;
      if (v < 0x80) { ctl_->push_back(b); break; }
      ctl_->push_back(b | 0x80);
      v >>= 7;
    }
  }
  void fixed(uint64_t v, size_t nbytes) { for (size_t i = 0; i < nbytes; i++) ctl_->push_back((uint8_t)(v >> (8 * i))); }
  uint8_t unit_id(long pattern_id) {  // GetFlag, :237-258
    auto it = ids_.find(pattern_id);
    if (it != ids_.end()) return it->second;
    if (ids_.size() >= 64) throw TuneError("too many unit kinds in one partition (CTL_PATTERNS_MAX)");
    uint8_t id = (uint8_t)ids_.size();
    ids_[pattern_id] = id;
    return id;
  }
  void head(long pattern_id, uint8_t size, int32_t ucol) {  // UpdateNewRow + AppendCtlHead (:615-633, CtlBuilder.cpp:62-81)
    bool nr = false; uint64_t jmp = 0;
    if (new_row_) {
      nr = true; new_row_ = false;
      if (empty_rows_) { jmp = empty_rows_ + 1; empty_rows_ = 0; row_jumps_ = true; }
    }
    uint8_t flags = unit_id(pattern_id);
    if (nr) flags |= 0x80;
    if (jmp) flags |= 0x40;
    ctl_->push_back(flags);
    ctl_->push_back(size);
    if (jmp) varint(jmp);
    if (full_) fixed((uint64_t)(int64_t)ucol, 4); else varint((uint64_t)(int64_t)ucol);  // int -> size_t sign extension
  }
  void flush_cols() {  // AddCols, :635-682
    size_t n = cols_.size();
    if (!n) return;
    int32_t first = cols_[0], last = cols_[n - 1];
    int32_t prev = last_col_, mx = 0;
    for (size_t i = 0; i < n; i++) { int32_t t = cols_[i]; cols_[i] -= prev; prev = t; if (i && cols_[i] > mx) mx = cols_[i]; }
    last_col_ = last;
    size_t w = delta_bytes((uint64_t)(int64_t)mx);
    head((long)(w << 3), (uint8_t)n, full_ ? first - 1 : cols_[0]);
    for (size_t i = 1; i < n; i++) fixed((uint64_t)(int64_t)cols_[i], w);
    cols_.clear();
  }
  void emit(size_t &k, size_t e, bool left_only) {  // DoRow / DoSymRow, :504-613
    const ValVec &pool = *spm_->pool;
    for (; k < e; k++) {
      const Rec &x = spm_->e[k];
      if (left_only && !(x.c < spm_->row_start + 1)) break;
      if (is_pattern(x)) {
        size_t sp = 0;  // UpdateRowSpan, :452-496
        if (x.type == T_VERT || x.type == T_DIAG || x.type == T_ADIAG) sp = (size_t)(x.size - 1) * x.delta;
        else if (is_brow(x.type)) sp = x.type - T_BROW1;
        else if (is_bcol(x.type)) sp = x.size / blk_align(x.type) - 1;
        span_ = std::max(span_, sp);
        flush_cols();
        long pid = is_blk(x.type) ? x.type * PATTERN_ID_OFFSET + x.size / blk_align(x.type)
                                  : x.type * PATTERN_ID_OFFSET + (long)x.delta;
        head(pid, x.size, full_ ? x.c - 1 : x.c - last_col_);
        last_col_ = x.c;  // GetLastCol (Element.hpp:657-666) at Horizontal order
        if (x.type == T_HORIZ) last_col_ += (int32_t)((x.size - 1) * x.delta);
        vals_->insert(vals_->end(), pool.begin() + x.vptr, pool.begin() + x.vptr + x.size);
        continue;
      }
      if (cols_.size() == 255) flush_cols();
      cols_.push_back(x.c);
      vals_->push_back(pool[x.vptr]);
    }
    flush_cols();
  }

  Part *spm_;
  bool full_, want_ri_;
  std::vector<uint8_t> *ctl_ = nullptr;
  std::vector<double> *vals_ = nullptr;
  std::map<long, uint8_t> ids_;
  bool new_row_ = false, row_jumps_ = false;
  uint64_t empty_rows_ = 0;
  int32_t last_col_ = 1;
  size_t span_ = 0;
  std::vector<int32_t> cols_;
};

// ------------------------------------------------------------ input cursor --
struct Cursor {
  // uniform 1-based (row, col, val) stream over CSR or COO input
  const CsrView *csr = nullptr;
  const CooHost *coo = nullptr;
  int64_t pos = 0, n = 0, row = 0;  // csr: row = current 0-based row
  int64_t roff = 0;                 // csr slab input: global row of the slab's first row
  void init() {
    if (csr) { n = csr->nnz(); row = 0; while (row < csr->nrows && csr->rowptr[row + 1] <= 0) row++; }
    else n = (int64_t)coo->row.size();
  }
  bool end() const { return pos >= n; }
  int32_t r() const { return csr ? (int32_t)(row + roff) + 1 : coo->row[pos]; }
  int32_t c() const { return csr ? csr->colind[pos] + 1 : coo->col[pos]; }   // Csr.hpp:352-367
  double v() const { return csr ? csr->values[pos] : coo->val[pos]; }
  void next() {
    pos++;
    if (csr) while (row < csr->nrows && csr->rowptr[row + 1] <= pos) row++;
  }
};

// SparsePartition::SetElems (:508-541) / SparsePartitionSym::SetElems (:1087-1129).
// `keep == false` only advances the cursor (partition not owned by this process).
struct SplitResult { int64_t taken = 0; int64_t rows = 0; int64_t diag = 0; int32_t cmin = INT32_MAX, cmax = 0; };
SplitResult take_partition(Cursor &cur, int64_t row_start, size_t limit, bool sym, bool keep, Part &p,
                           ValVec &pool, std::vector<double> &diag) {
  Trace tr("take_partition");
  SplitResult res;
  if (cur.csr && !sym && keep && !cur.end()) {
    // CSR input, plain CSX: the loop below ends the partition at the first row start at or after `limit` elements
    // (:525-533), which is a lookup in rowptr; the records are then filled in parallel.
    const CsrView &A = *cur.csr;
    const int64_t pos0 = cur.pos, n = cur.n;
    int64_t pos_end = n;
    if (limit && pos0 + (int64_t)limit < n)
      pos_end = *std::lower_bound(A.rowptr, A.rowptr + A.nrows + 1, (int)(pos0 + (int64_t)limit));
    const size_t cnt = (size_t)(pos_end - pos0);
    const size_t pbase = pool.size();
    p.e.resize(cnt);
    pool.resize(pbase + cnt);
    int T = (cnt < par_min()) ? 1 : g_sort_threads;
    std::vector<int32_t> tmin((size_t)T, INT32_MAX), tmax((size_t)T, 0);
    parallel_slices(cnt, [&](int t, size_t b, size_t en) {
      if (b >= en) return;
      // row of element pos0 + b: last row whose rowptr is <= that position and that is not empty there
      int64_t row = std::upper_bound(A.rowptr, A.rowptr + A.nrows + 1, (int)(pos0 + (int64_t)b)) - A.rowptr - 1;
      int32_t lo = INT32_MAX, hi = 0;
      for (size_t k = b; k < en; k++) {
        const int64_t pos = pos0 + (int64_t)k;
        while (A.rowptr[row + 1] <= pos) row++;
        const int32_t col = A.colind[pos] + 1;
        p.e[k] = Rec{(int32_t)(row + cur.roff + 1 - row_start), col, 0, 0, 1, 0, (uint64_t)(pbase + k)};
        pool[pbase + k] = A.values[pos];
        lo = std::min(lo, col); hi = std::max(hi, col);
      }
      tmin[t] = lo; tmax[t] = hi;
    });
    for (int t = 0; t < T; t++) { res.cmin = std::min(res.cmin, tmin[t]); res.cmax = std::max(res.cmax, tmax[t]); }
    const int32_t last_row = cnt ? p.e[cnt - 1].r : 0;
    cur.pos = pos_end;
    while (cur.row < A.nrows && A.rowptr[cur.row + 1] <= cur.pos) cur.row++;
    res.taken = (int64_t)cnt;
    res.rows = (int64_t)last_row;
    res.diag = 0;
    p.nr_nzeros = (int64_t)cnt;
    p.nr_rows = last_row;
    p.row_start = row_start;
    p.type = T_HORIZ;
    p.pool = &pool;
    p.build_rowptr();
    return res;
  }
  int32_t row_prev = 1, last_row = 0;
  size_t cnt = 0, dcnt = 0;
  if (keep) {
    size_t guess = limit ? std::min<size_t>(limit + (limit >> 6) + 1024, (size_t)(cur.n - cur.pos)) : (size_t)(cur.n - cur.pos);
    if (sym) guess = guess / 2 + 1024;
    p.e.reserve(guess); pool.reserve(guess);
  }
  for (; !cur.end(); cur.next()) {
    int32_t row = cur.r() - (int32_t)row_start;  // 1-based, partition relative
    int32_t col = cur.c();
    if (sym) {
      int64_t grow = row_start + row;            // 1-based global row
      if (grow == col) { dcnt++; if (keep) diag.push_back(cur.v()); continue; }
      if (grow < col) continue;                  // upper triangle is implied
    }
    if (row != row_prev) {
      if (row < row_prev) throw TuneError("input matrix rows are not sorted");
      if (sym ? (limit && dcnt + cnt >= limit && row_prev == row - 1) : (limit && cnt >= limit)) break;
      row_prev = row;
    }
    if (keep) {
      p.e.push_back(Rec{row, col, 0, 0, 1, 0, (uint64_t)pool.size()});
      pool.push_back(cur.v());
    }
    res.cmin = std::min(res.cmin, col); res.cmax = std::max(res.cmax, col);
    last_row = row;
    cnt++;
  }
  res.taken = (int64_t)(cnt + dcnt);
  res.rows = sym ? (int64_t)dcnt : (int64_t)last_row;   // SparsePartitionSym::GetNrRows == diagonal_size_
  res.diag = (int64_t)dcnt;
  if (keep) {
    p.nr_nzeros = (int64_t)cnt;
    p.nr_rows = last_row;       // rowptr_size - 1
    p.row_start = row_start;
    p.type = T_HORIZ;
    p.pool = &pool;
    p.build_rowptr();
  }
  return res;
}

void check_sorted_cols(const Part &p) {
  Trace tr("check_sorted_cols");
  std::vector<char> bad(64, 0);
  parallel_slices(p.e.size(), [&](int t, size_t b, size_t en) {
    for (size_t i = std::max<size_t>(b, 1); i < en; i++)
      if (p.e[i].r == p.e[i - 1].r && p.e[i].c <= p.e[i - 1].c) { bad[t] = 1; return; }
  });
  for (char c : bad)
    if (c) throw TuneError("column indices must be strictly increasing within each row");
}

void mine(Part &p, const TuneOptions &opt, const XformSeq &q, bool run_all, std::string *log, bool *undef) {
  Miner m(&p, opt, log, undef);   // CsxBuild.hpp:134-166
  if (q.is_explicit) { m.mine_serial(q); return; }
  for (auto &it : q.items) {
    std::vector<int> ts;
    expand_group(it.first, ts);
    for (int t : ts) if (t != T_NONE) m.allow(t);
  }
  if (run_all) m.mine_all();
}

// DivideMatrix / MergeMatrix, SparsePartition.hpp:965-1074
void split_lower(const Part &lower, Part &m1, Part &m2) {
  for (Part *q : {&m1, &m2}) {
    q->type = T_HORIZ; q->row_start = lower.row_start; q->nr_cols = lower.nr_cols; q->pool = lower.pool;
  }
  for (const Rec &x : lower.e) (x.c < lower.row_start + 1 ? m1 : m2).e.push_back(x);
  for (Part *q : {&m1, &m2}) {
    q->nr_nzeros = (int64_t)q->e.size();
    q->build_rowptr();
    q->nr_rows = (int64_t)q->nrowptr() - 1;
  }
}
void merge_lower(Part &lower, const Part &m1, const Part &m2) {
  RecVec out;
  out.reserve(m1.e.size() + m2.e.size());
  size_t nr = (size_t)lower.nr_rows;
  std::vector<int64_t> rp(nr + 1, 0);
  for (size_t i = 0; i < nr; i++) {
    if (m1.nrowptr() - 1 > i) out.insert(out.end(), m1.e.begin() + m1.rowptr[i], m1.e.begin() + m1.rowptr[i + 1]);
    if (m2.nrowptr() - 1 > i) out.insert(out.end(), m2.e.begin() + m2.rowptr[i], m2.e.begin() + m2.rowptr[i + 1]);
    rp[i + 1] = (int64_t)out.size();
  }
  lower.e.swap(out);
  lower.rowptr.swap(rp);
  lower.nrowptr_ = lower.rowptr.size();
  lower.type = T_HORIZ;
}

// MakeMap, CsxBuild.hpp:400-581 (needs every partition's lower triangle)
void build_sym_map(const std::vector<Part> &lowers, int64_t ncols, std::vector<CsxPartition *> outs) {
  size_t np = lowers.size(), n = (size_t)ncols;
  std::vector<uint32_t> count(n + 1, 0);
  std::vector<std::vector<uint8_t>> seen(np, std::vector<uint8_t>(n + 1, 0));
  for (size_t i = 0; i < np; i++)
    for (const Rec &x : lowers[i].e)
      if ((int64_t)x.c < lowers[i].row_start + 1 && !seen[i][x.c]) { seen[i][x.c] = 1; count[x.c]++; }
  uint64_t total = 0;
  for (size_t i = 0; i < n; i++) total += count[i];
  size_t end = 0;
  for (size_t i = 0; i + 1 < np; i++) {
    size_t start = end;
    uint64_t limit = total / (np - i), got = 0;
    while (got < limit) got += count[end++];
    total -= got;
    if (outs[i])
      for (size_t j = start; j < end; j++)
        for (size_t k = 0; k < np; k++)
          if (seen[k][j]) { outs[i]->map_cpus.push_back((uint32_t)k); outs[i]->map_pos.push_back((uint32_t)(j - 1)); }
  }
  if (outs[np - 1])
    for (size_t j = end; j < n; j++)
      for (size_t k = 0; k < np; k++)
        if (seen[k][j]) { outs[np - 1]->map_cpus.push_back((uint32_t)k); outs[np - 1]->map_pos.push_back((uint32_t)(j - 1)); }
}

// slab_part >= 0: the input holds exactly the rows of partition `slab_part` of the nr_threads-way split (first row
// cur.roff); the partition is encoded from them alone.
std::string tune_impl(Cursor &cur, int64_t nrows, int64_t ncols, const TuneOptions &opt, int part_lo, int part_hi,
                      CsxMatrix &out, int slab_part = -1) {
  try {
    if (opt.nr_threads < 1) throw TuneError("invalid value for spx.rt.nr_threads");
    if (opt.heuristic != "ratio") throw TuneError("spx.preproc.heuristic=" + opt.heuristic + " is not supported (ratio only)");
    if (opt.min_unit_size < 2 || opt.max_unit_size > 255 || opt.max_unit_size < opt.min_unit_size)
      throw TuneError("unit size limits must satisfy 2 <= min <= max <= 255");
    int np = opt.nr_threads;
    if (part_lo < 0 || part_hi > np || part_lo >= part_hi) throw TuneError("invalid partition range");
    XformSeq q = parse_xform(opt.xform);
    cur.init();
    out = CsxMatrix();
    out.nrows = nrows; out.ncols = ncols; out.nnz = cur.n;
    out.symmetric = opt.symmetric; out.full_colind = opt.full_colind;
    out.nparts_total = np; out.part_lo = part_lo;
    out.rows_per_thread = opt.rows_per_thread;
    out.slice_elems = opt.slice_elems;
    out.slab_rows = opt.slab_rows;
    out.parts.resize(part_hi - part_lo);
    bool sym = opt.symmetric;
    if (sym && nrows != ncols) throw TuneError("spx.matrix.symmetric requires a square matrix");
    // BuildPartitions, SparseInternal.hpp:119-152 (sym: nnz := (nnz + ncols) / 2, :92)
    uint64_t total = sym ? (uint64_t)(cur.n + ncols) / 2 : (uint64_t)cur.n, done = 0;
    int64_t row_start = 0;
    std::vector<Part> parts(np);
    std::vector<ValVec> pools(np);
    std::vector<std::vector<double>> diags(np);
    if (slab_part >= 0) { row_start = cur.roff; total = sym ? (uint64_t)(cur.n + cur.csr->nrows) / 2 : (uint64_t)cur.n; }
    for (int i = 0; i < np; i++) {
      if (slab_part >= 0 && i != slab_part) continue;
      size_t limit = slab_part >= 0 ? 0 : (size_t)((total - done) / (uint64_t)(np - i));   // 0: every element of the slab
      // the sym reduction map needs every partition's lower triangle
      bool keep = (i >= part_lo && i < part_hi) || (sym && np > 1);
      parts[i].nr_cols = ncols;
      SplitResult r = take_partition(cur, row_start, limit, sym, keep, parts[i], pools[i], diags[i]);
      if (!keep) { parts[i].row_start = row_start; parts[i].nr_rows = r.rows; }
      if (i >= part_lo && i < part_hi && r.cmax >= r.cmin) {
        out.parts[i - part_lo].col_min = r.cmin - 1;
        out.parts[i - part_lo].col_max = r.cmax - 1;
      }
      row_start += r.rows;
      done += (uint64_t)r.taken;
    }
    if (slab_part < 0 && done != total) throw TuneError("error in input matrix (matrix has less elements than claimed)");
    if (sym && slab_part < 0) {   // (a slab does not see the other partitions: no reduction map, the GPU path needs none)
      std::vector<CsxPartition *> outs(np, nullptr);
      for (int i = part_lo; i < part_hi; i++) outs[i] = &out.parts[i - part_lo];
      build_sym_map(parts, ncols, outs);
    }
    auto work = [&](int i) {
      Part &p = parts[i];
      CsxPartition &o = out.parts[i - part_lo];
      check_sorted_cols(p);
      if (!sym) {
        mine(p, opt, q, true, &o.encoding_log, &o.sampling_undefined);
        CtlWriter(&p, opt.full_colind, opt.build_rows_info).run(false, o);
      } else {
        Part m1, m2;
        split_lower(p, m1, m2);
        o.encoding_log += "m1 ";
        mine(m1, opt, q, i != 0, &o.encoding_log, &o.sampling_undefined);  // partition 0 skips m1 (CsxBuild.hpp:240)
        o.encoding_log += "m2 ";
        mine(m2, opt, q, true, &o.encoding_log, &o.sampling_undefined);
        merge_lower(p, m1, m2);
        CtlWriter(&p, opt.full_colind, opt.build_rows_info).run(true, o);
        o.dvalues = diags[i];
      }
      RecVec().swap(p.e);
      ValVec().swap(pools[i]);
    };
    // one preprocessing thread per partition like CsxBuild.hpp:290-326, capped by host_threads
    int hw = opt.host_threads > 0 ? opt.host_threads : (int)std::thread::hardware_concurrency();
    int nthreads = std::max(1, std::min(hw, part_hi - part_lo));
    g_sort_threads = std::max(1, std::min(32, hw / nthreads));   // workers inside one partition's sorts
    std::vector<std::string> errs(part_hi - part_lo);
    if (nthreads == 1) {
      for (int i = part_lo; i < part_hi; i++) work(i);
    } else {
      std::vector<std::thread> th;
      for (int t = 0; t < nthreads; t++)
        th.emplace_back([&, t]() {
          for (int i = part_lo + t; i < part_hi; i += nthreads) {
            try { work(i); } catch (std::exception &e) { errs[i - part_lo] = e.what(); }
          }
        });
      for (auto &t : th) t.join();
      for (auto &e : errs) if (!e.empty()) throw TuneError(e);
    }
  } catch (std::exception &e) {
    return e.what();
  }
  return "";
}

}  // namespace

std::string TuneOptions::set(const std::string &k, const std::string &v) {
  auto as_bool = [&](bool &dst) -> std::string {
    if (v == "true" || v == "1") dst = true;
    else if (v == "false" || v == "0") dst = false;
    else return "invalid value \"" + v + "\" while setting property \"" + k + "\"";
    return "";
  };
  try {
    if (k == "spx.rt.nr_threads") nr_threads = std::stoi(v);
    else if (k == "spx.rt.cpu_affinity") {}  // no CPU worker threads to pin
    else if (k == "spx.preproc.heuristic") heuristic = v;
    else if (k == "spx.preproc.xform") xform = v;
    else if (k == "spx.preproc.sampling") sampling = v;
    else if (k == "spx.preproc.sampling.nr_samples") nr_samples = std::stoul(v);
    else if (k == "spx.preproc.sampling.portion") portion = std::stod(v);
    else if (k == "spx.preproc.sampling.window_size") window_size = std::stoul(v);
    else if (k == "spx.matrix.symmetric") return as_bool(symmetric);
    else if (k == "spx.matrix.split_blocks") return as_bool(split_blocks);
    else if (k == "spx.matrix.full_colind") return as_bool(full_colind);
    else if (k == "spx.matrix.min_unit_size") min_unit_size = std::stoul(v);
    else if (k == "spx.matrix.max_unit_size") max_unit_size = std::stoul(v);
    else if (k == "spx.matrix.min_coverage") min_coverage = std::stod(v);
    else if (k == "spx.b200.rows_info") return as_bool(build_rows_info);
    else if (k == "spx.b200.host_threads") host_threads = std::stoi(v);
    else if (k == "spx.b200.rows_per_thread") {
      rows_per_thread = std::stoi(v);
      if (rows_per_thread != 0 && rows_per_thread != 1 && rows_per_thread != 4) return "spx.b200.rows_per_thread must be 0, 1 or 4";
    }
    else if (k == "spx.b200.slab_rows") {
      slab_rows = std::stoll(v);
      if (slab_rows < 0) return "spx.b200.slab_rows must not be negative";
    }
    else if (k == "spx.b200.slice") {
      slice_elems = std::stoi(v);
      if (slice_elems != 0 && (slice_elems < 4 || slice_elems > 32)) return "spx.b200.slice must be 0 or 4..32";
    }
    else return "unknown option \"" + k + "\"";
  } catch (std::exception &) {
    return "invalid value \"" + v + "\" while setting property \"" + k + "\"";
  }
  return "";
}

std::string tune_csr(const CsrView &in, const TuneOptions &opt, int part_lo, int part_hi, CsxMatrix &out) {
  Cursor c; c.csr = &in;
  return tune_impl(c, in.nrows, in.ncols, opt, part_lo, part_hi, out);
}
std::string tune_csr_slab(const CsrView &slab, int64_t nrows_total, int64_t row_start, int part, const TuneOptions &opt, CsxMatrix &out) {
  if (part < 0 || part >= opt.nr_threads) return "invalid partition index";
  if (row_start < 0 || row_start + slab.nrows > nrows_total) return "slab rows outside the matrix";
  Cursor c; c.csr = &slab; c.roff = row_start;
  return tune_impl(c, nrows_total, slab.ncols, opt, part, part + 1, out, part);
}
std::string tune_coo(const CooHost &in, const TuneOptions &opt, int part_lo, int part_hi, CsxMatrix &out) {
  Cursor c; c.coo = &in;
  return tune_impl(c, in.nrows, in.ncols, opt, part_lo, part_hi, out);
}

}  // namespace spxb
