// ORACLE / reference pin — TEST INFRASTRUCTURE ONLY.
//
// Drives the reference's own encoder: this translation unit includes the SparseX headers where they lie under
// /root/reference/include (SparseInternal, SparsePartition, EncodingManager, Statistics, CsxManager, CtlBuilder,
// ...) and is linked with the reference's Encodings.cpp, Runtime.cpp, Statistics.cpp, CtlBuilder.cpp and
// CsxUtil.cpp compiled from /root/reference/src/internals.  Boost, the logger, numa.h and the generated
// Config.hpp are replaced by the stand-ins of oracle/refshim (no CSX logic there).  What is restated here is
// only the per-thread driver of CsxBuild.hpp:134-288 (PreprocessThread / PreprocessThreadSym) without the
// thread pool and the LLVM JIT, and ThreadCtx::SetData (Runtime.hpp:318-330): the JIT headers need LLVM/Clang,
// which this image does not have.  Built by oracle/build_refenc.py into oracle/_ref/libcsxref_enc.so; used to
// generate tests/golden/ref_encodings.npz and to check oracle/csx_oracle.cpp against the reference itself.
#include <sparsex/internals/Csr.hpp>
#include <sparsex/internals/CsxManager.hpp>
#include <sparsex/internals/EncodingManager.hpp>
#include <sparsex/internals/Runtime.hpp>
#include <sparsex/internals/SparseInternal.hpp>

#include <cstring>
#include <sstream>
#include <string>
#include <vector>

using namespace sparsex;
using namespace sparsex::csx;
using namespace sparsex::runtime;
using sparsex::io::CSR;

namespace {
typedef int idx_t;
typedef double val_t;

struct PartOut {
  std::vector<uint8_t> ctl;
  std::vector<double> values, dvalues;
  std::vector<long> id_map;
  long nnz = 0, nrows = 0, ncols = 0, row_start = 0, row_jumps = 0;
  std::string log;
};
struct Handle { std::vector<PartOut> parts; };

void copy_csx(CsxMatrix<idx_t, val_t> *csx, PartOut &o) {
  o.nnz = (long)csx->nnz; o.nrows = (long)csx->nrows; o.ncols = (long)csx->ncols; o.row_start = (long)csx->row_start;
  o.row_jumps = csx->row_jumps ? 1 : 0;
  o.ctl.assign(csx->ctl, csx->ctl + csx->ctl_size);
  o.values.assign(csx->values, csx->values + csx->nnz);
  for (int i = 0; i < 64; i++) { o.id_map.push_back(csx->id_map[i]); if (csx->id_map[i] == -1) break; }
}

void set_options(const char *opts) {
  RtConfig &cfg = RtConfig::GetInstance();
  // defaults of the non-NUMA build (Runtime.cpp:37-63) before applying the caller's options
  const char *defaults[][2] = {{"spx.rt.nr_threads", "1"}, {"spx.rt.cpu_affinity", "0"}, {"spx.preproc.heuristic", "ratio"},
                               {"spx.preproc.xform", "all"}, {"spx.preproc.sampling", "portion"},
                               {"spx.preproc.sampling.nr_samples", "48"}, {"spx.preproc.sampling.portion", "0.01"},
                               {"spx.preproc.sampling.window_size", "0"}, {"spx.matrix.symmetric", "false"},
                               {"spx.matrix.split_blocks", "true"}, {"spx.matrix.full_colind", "false"},
                               {"spx.matrix.min_unit_size", "4"}, {"spx.matrix.max_unit_size", "255"},
                               {"spx.matrix.min_coverage", "0.1"}};
  for (auto &d : defaults) cfg.SetProperty(cfg.GetPropertyByMnemonic(d[0]), d[1]);
  if (!opts) return;
  std::stringstream ss(opts);
  std::string kv;
  while (std::getline(ss, kv, ';')) {
    size_t eq = kv.find('=');
    if (kv.empty() || eq == std::string::npos) continue;
    cfg.SetProperty(cfg.GetPropertyByMnemonic(kv.substr(0, eq)), kv.substr(eq + 1));   // spx_option_set, matvec.c:753-756
  }
}
}  // namespace

extern "C" {

// Tunes a zero-based CSR matrix with the reference encoder; returns a handle or NULL.
void *refenc_tune(const int *rowptr, const int *colind, const double *values, int nrows, int ncols, const char *opts) {
  set_options(opts);
  RtConfig &cfg = RtConfig::GetInstance();
  const size_t nr = cfg.GetProperty<size_t>(RtConfig::RtNrThreads);
  const bool sym = cfg.GetProperty<bool>(RtConfig::MatrixSymmetric);
  const bool full_colind = cfg.GetProperty<bool>(RtConfig::MatrixFullColind);
  EncodingSequence encseq(cfg.GetProperty<string>(RtConfig::PreprocXform));
  CSR<idx_t, val_t> csr(const_cast<int *>(rowptr), const_cast<int *>(colind), const_cast<double *>(values), (idx_t)nrows,
                        (idx_t)ncols, true);   // zero-based input (spx_input_load_csr, matvec.c:163-215)
  Handle *h = new Handle;
  h->parts.resize(nr);
  std::ostringstream buf;
  if (!sym) {
    typedef SparsePartition<idx_t, val_t> Part;
    SparseInternal<Part> *spi = SparseInternal<Part>::DoLoadMatrix(csr, nr);
    for (size_t i = 0; i < nr; i++) {   // PreprocessThread, CsxBuild.hpp:134-197
      Part *spm = spi->GetPartition(i);
      CsxManager<idx_t, val_t> mgr(spm);
      mgr.SetFullColumnIndices(full_colind);
      EncodingManager<idx_t, val_t> *drle = new EncodingManager<idx_t, val_t>(spm, cfg);
      std::ostringstream log;
      if (encseq.IsExplicit()) drle->EncodeSerial(encseq);
      else { drle->RemoveIgnore(encseq); drle->EncodeAll(log); }
      CsxMatrix<idx_t, val_t> *csx = mgr.MakeCsx(false);
      copy_csx(csx, h->parts[i]);
      h->parts[i].log = log.str();
      delete drle;
    }
    delete spi;
  } else {
    typedef SparsePartitionSym<idx_t, val_t> PartSym;
    SparseInternal<PartSym> *spi = SparseInternal<PartSym>::DoLoadMatrixSym(csr, nr);
    for (size_t i = 0; i < nr; i++) {   // PreprocessThreadSym, CsxBuild.hpp:199-288
      PartSym *spm = spi->GetPartition(i);
      CsxManager<idx_t, val_t> mgr(spm);
      mgr.SetFullColumnIndices(full_colind);
      spm->DivideMatrix();
      EncodingManager<idx_t, val_t> *d1 = new EncodingManager<idx_t, val_t>(spm->GetFirstMatrix(), cfg);
      EncodingManager<idx_t, val_t> *d2 = new EncodingManager<idx_t, val_t>(spm->GetSecondMatrix(), cfg);
      std::ostringstream log;
      if (encseq.IsExplicit()) { d1->EncodeSerial(encseq); d2->EncodeSerial(encseq); }
      else {
        d1->RemoveIgnore(encseq); d2->RemoveIgnore(encseq);
        if (i) d1->EncodeAll(log);
        d2->EncodeAll(log);
      }
      spm->MergeMatrix();
      CsxSymMatrix<idx_t, val_t> *csx = mgr.MakeCsxSym();
      copy_csx(csx->lower_matrix, h->parts[i]);
      h->parts[i].dvalues.assign(csx->dvalues, csx->dvalues + spm->GetDiagonalSize());
      h->parts[i].log = log.str();
      delete d1;
      delete d2;
    }
    delete spi;
  }
  return h;
}

int refenc_nparts(void *hv) { return (int)((Handle *)hv)->parts.size(); }
// what: 0 nnz, 1 nrows, 2 ncols, 3 row_start, 4 ctl_size, 5 row_jumps, 6 id_map length, 7 dvalues length
long refenc_info(void *hv, int p, int what) {
  PartOut &o = ((Handle *)hv)->parts[p];
  switch (what) {
    case 0: return o.nnz; case 1: return o.nrows; case 2: return o.ncols; case 3: return o.row_start;
    case 4: return (long)o.ctl.size(); case 5: return o.row_jumps; case 6: return (long)o.id_map.size();
    case 7: return (long)o.dvalues.size();
  }
  return -1;
}
// what: 0 values f64, 1 ctl u8, 2 id_map i64, 3 dvalues f64
void refenc_copy(void *hv, int p, int what, void *dst) {
  PartOut &o = ((Handle *)hv)->parts[p];
  if (what == 0) memcpy(dst, o.values.data(), o.values.size() * 8);
  else if (what == 1) memcpy(dst, o.ctl.data(), o.ctl.size());
  else if (what == 2) memcpy(dst, o.id_map.data(), o.id_map.size() * sizeof(long));
  else if (what == 3) memcpy(dst, o.dvalues.data(), o.dvalues.size() * 8);
}
const char *refenc_log(void *hv, int p) { return ((Handle *)hv)->parts[p].log.c_str(); }
void refenc_free(void *hv) { delete (Handle *)hv; }

}  // extern "C"
