// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the SparseX (cslab-ntua/sparsex v1.1.0) CSX / CSX-Sym
// encoder and of its SpMV unit semantics, written to follow the reference
// source function by function.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load this code; the
// product library (sparsex_b200/csrc, libsparsex_b200.so) never links it.
//
// Parity pinning status:
//   * multiply half: pinned against the reference's own C kernel templates
//     compiled by gcc (oracle/_ref, built from /root/reference/src/templates).
//   * encoder half: pinned against the hand-derived known-answer vector of
//     SURVEY.md Appendix C and, functionally, by decoding every emitted stream
//     with the reference templates; the reference ships no golden ctl streams.
//
// All citations are file:line into /root/reference/.
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>

namespace csxo {

// include/sparsex/internals/Encodings.hpp:38-66 — numeric values are part of
// the format (pattern id = type*10000 + delta).
enum Type : int {
  None = 0, Horizontal, Vertical, Diagonal, AntiDiagonal,
  BlockRow1, BlockRow2, BlockRow3, BlockRow4, BlockRow5, BlockRow6, BlockRow7, BlockRow8,
  BlockCol1, BlockCol2, BlockCol3, BlockCol4, BlockCol5, BlockCol6, BlockCol7, BlockCol8,
  TypeMax  // == Encoding::Max == __EndOfTypes__
};

inline bool is_block_row(int t) { return t >= BlockRow1 && t <= BlockRow8; }
inline bool is_block_col(int t) { return t >= BlockCol1 && t <= BlockCol8; }
inline bool is_block(int t) { return is_block_row(t) || is_block_col(t); }
// Encodings.hpp:118-126
inline size_t block_align(int t) {
  if (is_block_row(t)) return t - BlockRow1 + 1;
  if (is_block_col(t)) return t - BlockCol1 + 1;
  return 0;
}

typedef std::pair<int, size_t> Inst;  // Encoding::Instantiation

// Runtime.cpp:37-63 (non-NUMA defaults)
struct Options {
  int nr_threads = 1;
  std::string xform = "all";
  std::string sampling = "portion";  // none | window | portion
  size_t nr_samples = 48;
  double portion = 0.01;
  size_t window_size = 0;
  bool symmetric = false;
  bool split_blocks = true;
  bool onedim_blocks = false;
  bool full_colind = false;
  size_t min_unit_size = 4;
  size_t max_unit_size = 255;
  double min_coverage = 0.1;
  // Not a reference option: what to do where EncodingManager::SelectSplits
  // (EncodingManager.hpp:1489-1516) reads uninitialised / out-of-bounds data.
  // "error": refuse (parity undefined); "break": end the sampling loop there,
  // as the reference does for an empty window.
  std::string undefined_sampling = "error";
  // returns "" or an error message
  std::string set(const std::string &key, const std::string &val);
};

// Element.hpp:192-608 — generic element: a single non-zero or a substructure.
struct Elem {
  int row = 0, col = 0;   // 1-based, in the partition's current iteration order
  int type = None;        // inst.first
  uint32_t delta = 0;     // inst.second (0 => not a pattern, Element.hpp:372-377)
  uint32_t size = 1;
  double val = 0;         // size == 1
  std::vector<double> vals;  // size > 1
  bool is_pattern() const { return delta != 0; }
};

// Csx.hpp:29-48
struct RowInfo { int rowptr, valptr, span; };

struct CsxPart {
  std::vector<double> values;
  std::vector<uint8_t> ctl;
  long nnz = 0, ncols = 0, nrows = 0, row_start = 0;
  int row_jumps = 0;
  std::vector<long> id_map;      // terminated by -1 (CsxManager.hpp:439-450)
  std::vector<RowInfo> rows_info;
  // CSX-Sym only
  std::vector<double> dvalues;   // Csx.hpp:50-53
  std::vector<unsigned> map_cpus, map_pos;  // Map.hpp:23-27
};

struct Tuned {
  long nrows = 0, ncols = 0, nnz = 0;
  bool symmetric = false;
  bool full_colind = false;
  std::vector<CsxPart> parts;
  std::string log;  // chosen encoding sequence per partition
};

struct Coo { long nrows = 0, ncols = 0; std::vector<int> row, col; std::vector<double> val; };

// Mmf.hpp:331-478; returns "" or error text.  Elements come back 1-based, in
// the order the reference iterator would yield them.
std::string load_mmf(const char *path, Coo &out);

// spx_mat_tune restatement.  Input elements 1-based, row-major sorted.
// Returns "" on success, else an error message (the reference would exit(1)).
std::string tune(const Coo &in, const Options &opt, Tuned &out);

// Reference SpMV semantics (templates + CsxKernels.cpp + CsxSpmv.cpp),
// y = alpha*A*x + beta*y ; mult (VecInit(y,0)) is beta == 0 with y overwritten.
void spmv(const Tuned &A, double alpha, const double *x, double beta, double *y, bool overwrite);

// Independent decoder: (row, col) 0-based global for every value, in values order.
void decode_coords(const Tuned &A, int part, std::vector<int> &rows, std::vector<int> &cols);

}  // namespace csxo
