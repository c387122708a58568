// Host-side CSX data model of the B200 engine (product code).
//
// The tuning stage (spx_mat_tune) produces, per row partition, the CSX byte
// stream of SparseX: `ctl` + `values` + `rows_info` + `id_map`
// (reference: include/sparsex/internals/Csx.hpp:29-53, grammar in
// CtlBuilder.cpp:62-81).  These arrays are bit-compatible with the reference;
// everything GPU specific (segment table, cross-row unit table) is derived
// from them in gpu_layout.cpp.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace spxb {

// Encoding::Type numeric values (Encodings.hpp:38-60) are part of the format:
// pattern id = type * 10000 + delta (CsxUtil.hpp:58-74, CsxUtil.cpp:27-33).
enum EncType : int {
  T_NONE = 0, T_HORIZ = 1, T_VERT = 2, T_DIAG = 3, T_ADIAG = 4,
  T_BROW1 = 5, T_BROW8 = 12, T_BCOL1 = 13, T_BCOL8 = 20, T_MAX = 21
};
inline bool is_brow(int t) { return t >= T_BROW1 && t <= T_BROW8; }
inline bool is_bcol(int t) { return t >= T_BCOL1 && t <= T_BCOL8; }
inline bool is_blk(int t) { return is_brow(t) || is_bcol(t); }
inline int blk_align(int t) { return is_brow(t) ? t - T_BROW1 + 1 : (is_bcol(t) ? t - T_BCOL1 + 1 : 0); }
constexpr long PATTERN_ID_OFFSET = 10000;

// RtConfig properties (src/internals/Runtime.cpp:37-95), non-NUMA defaults.
struct TuneOptions {
  int nr_threads = 1;              // spx.rt.nr_threads  (= number of row partitions)
  std::string xform = "all";       // spx.preproc.xform
  std::string sampling = "portion";// spx.preproc.sampling
  size_t nr_samples = 48;          // spx.preproc.sampling.nr_samples
  double portion = 0.01;           // spx.preproc.sampling.portion
  size_t window_size = 0;          // spx.preproc.sampling.window_size
  bool symmetric = false;          // spx.matrix.symmetric
  bool split_blocks = true;        // spx.matrix.split_blocks
  bool full_colind = false;        // spx.matrix.full_colind
  size_t min_unit_size = 4;        // spx.matrix.min_unit_size
  size_t max_unit_size = 255;      // spx.matrix.max_unit_size
  double min_coverage = 0.1;       // spx.matrix.min_coverage
  std::string heuristic = "ratio"; // spx.preproc.heuristic
  // engine additions (no reference counterpart)
  bool build_rows_info = true;     // spx.b200.rows_info : keep the per-row table of Csx.hpp:29-35
  int host_threads = 0;            // spx.b200.host_threads : 0 = hardware concurrency
  int rows_per_thread = 0;         // spx.b200.rows_per_thread : GPU tile shape, 0 = by partition size, else 1 or 4
  int slice_elems = 0;             // spx.b200.slice : elements one lane of the chunk kernel handles, 0 = from the unit mix
  long long slab_rows = 0;         // spx.b200.slab_rows : rows per slab of the pipelined host-buffer SpMV, 0 = sizes chosen by the engine
  // Returns "" or an error message.  Unknown mnemonics are an error
  // (Runtime.hpp:108-134 logs and ignores; the C API layer downgrades this to a warning).
  std::string set(const std::string &mnemonic, const std::string &value);
};

struct RowInfo64 { int64_t rowptr, valptr; int32_t span; };

// One row partition in CSX form == csx_matrix_t (+ csx_sym_matrix_t extras).
struct CsxPartition {
  std::vector<double> values;
  std::vector<uint8_t> ctl;
  int64_t nnz = 0, nrows = 0, ncols = 0, row_start = 0;
  bool row_jumps = false;
  int64_t col_min = 0, col_max = -1;   // zero-based column window the partition reads (empty: min > max)
  std::vector<long> id_map;            // unit id -> pattern id, terminated by -1
  std::vector<RowInfo64> rows_info;    // optional (build_rows_info)
  std::vector<double> dvalues;         // CSX-Sym: diagonal of the partition's rows
  std::vector<uint32_t> map_cpus, map_pos;  // CSX-Sym reduction map (Map.hpp:23-27)
  std::string encoding_log;            // chosen sequence, e.g. "d{1} h{1}"
  bool sampling_undefined = false;     // reference behaviour was undefined (see encoder.cpp)
};

struct CsxMatrix {
  int64_t nrows = 0, ncols = 0, nnz = 0;   // nnz = input non-zeros (full matrix also when symmetric)
  bool symmetric = false, full_colind = false;
  int rows_per_thread = 0;                 // requested GPU tile shape (0 = automatic)
  int slice_elems = 0;                     // requested chunk-kernel slice length (0 = automatic)
  long long slab_rows = 0;                 // rows per slab of the pipelined host-buffer SpMV (0 = chosen by the engine)
  int nparts_total = 0;                    // partitions the matrix was split into
  int part_lo = 0;                         // parts[k] is global partition part_lo + k
  std::vector<CsxPartition> parts;
  std::vector<int32_t> permutation;        // RCM (SPX_MAT_REORDER): old index -> new index; empty = given ordering
};

// Input views -----------------------------------------------------------
struct CsrView {   // spx_input_load_csr wraps user arrays without copying (Csr.hpp:56-68)
  const int *rowptr; const int *colind; const double *values;
  int64_t nrows, ncols;
  int64_t nnz() const { return rowptr[nrows]; }
};
struct CooHost {   // MMF reader output, 1-based, row-major sorted
  int64_t nrows = 0, ncols = 0;
  std::vector<int> row, col;
  std::vector<double> val;
  bool buffered = false;   // symmetric or column-wise file: the reference holds it in memory (Mmf.hpp:218), RCM sees its elements
};

// mmf.cpp — reference: include/sparsex/internals/Mmf.hpp:331-478
std::string read_mmf(const char *path, CooHost &out);

// rcm.cpp — reference: include/sparsex/internals/Rcm.hpp:116-340 (Boost Graph Library's cuthill_mckee_ordering restated)
bool rcm_find_perm(int64_t n, const std::vector<int32_t> &eu, const std::vector<int32_t> &ev, std::vector<int32_t> &perm,
                   std::vector<int32_t> &inv_perm, int64_t *bandwidth, int64_t start = -1);
void rcm_edges_csr(const int32_t *rowptr, const int32_t *colind, int64_t nrows, bool symmetric, std::vector<int32_t> &eu,
                   std::vector<int32_t> &ev);
void rcm_edges_coo(const CooHost &coo, std::vector<int32_t> &eu, std::vector<int32_t> &ev);
void rcm_apply_coo(CooHost &coo, const std::vector<int32_t> &perm);

// encoder.cpp — reference call stack SURVEY.md §3.2.  Encodes partitions
// [part_lo, part_hi) of the nr_threads-way split; the other partitions are
// only scanned to find their row boundaries.
std::string tune_csr(const CsrView &in, const TuneOptions &opt, int part_lo, int part_hi, CsxMatrix &out);
std::string tune_coo(const CooHost &in, const TuneOptions &opt, int part_lo, int part_hi, CsxMatrix &out);
// One process per GPU on matrices too large for one host: `slab` holds exactly the rows [row_start, row_start +
// slab.nrows) of partition `part` of the nr_threads-way split of a matrix with nrows_total rows (the caller applies the
// split rule to the row lengths); the partition is encoded from its own rows alone, as the reference's per-thread
// preprocessing does (CsxBuild.hpp:134-288).  CSX-Sym: no reduction map (it needs every partition).
std::string tune_csr_slab(const CsrView &slab, int64_t nrows_total, int64_t row_start, int part, const TuneOptions &opt, CsxMatrix &out);

}  // namespace spxb
