// ORACLE — TEST INFRASTRUCTURE ONLY.  Plain C entry points over csx_oracle.cpp
// so that tests/ and bench.py (cpu_baseline leg) can drive the restatement
// through ctypes.  Never linked into the product library.
#include <chrono>
#include <cstring>
#include <sstream>
#include <thread>

#include "csx_oracle.hpp"

using namespace csxo;

namespace {
struct Handle {
  Coo coo;
  Tuned tuned;
  bool have_tuned = false;
};

std::string parse_opts(const char *opts, Options &o) {
  if (!opts) return "";
  std::stringstream ss(opts);
  std::string kv;
  while (std::getline(ss, kv, ';')) {
    if (kv.empty()) continue;
    size_t eq = kv.find('=');
    if (eq == std::string::npos) return "bad option " + kv;
    std::string e = o.set(kv.substr(0, eq), kv.substr(eq + 1));
    if (!e.empty()) return e;
  }
  return "";
}
void put_err(char *err, int len, const std::string &m) {
  if (err && len > 0) { strncpy(err, m.c_str(), len - 1); err[len - 1] = 0; }
}
}  // namespace

extern "C" {

void *csxo_from_csr(const int *rowptr, const int *colind, const double *values, int nrows, int ncols) {
  Handle *h = new Handle;
  h->coo.nrows = nrows; h->coo.ncols = ncols;
  size_t nnz = rowptr[nrows];
  h->coo.row.reserve(nnz); h->coo.col.reserve(nnz); h->coo.val.reserve(nnz);
  // Csr.hpp:256-373: zero-based arrays are exposed as 1-based elements
  for (int r = 0; r < nrows; r++)
    for (int k = rowptr[r]; k < rowptr[r + 1]; k++) {
      h->coo.row.push_back(r + 1); h->coo.col.push_back(colind[k] + 1); h->coo.val.push_back(values[k]);
    }
  return h;
}

void *csxo_from_mmf(const char *path, char *err, int errlen) {
  Handle *h = new Handle;
  std::string e = load_mmf(path, h->coo);
  if (!e.empty()) { put_err(err, errlen, e); delete h; return nullptr; }
  return h;
}

int csxo_tune(void *hv, const char *opts, char *err, int errlen) {
  Handle *h = (Handle *)hv;
  Options o;
  std::string e = parse_opts(opts, o);
  if (e.empty()) e = tune(h->coo, o, h->tuned);
  if (!e.empty()) { put_err(err, errlen, e); h->have_tuned = false; return -1; }
  h->have_tuned = true;
  return 0;
}

void csxo_free(void *hv) { delete (Handle *)hv; }

void csxo_dims(void *hv, long *nrows, long *ncols, long *nnz) {
  Handle *h = (Handle *)hv;
  *nrows = h->coo.nrows; *ncols = h->coo.ncols; *nnz = (long)h->coo.row.size();
}
// COO of the loaded input, 0-based
void csxo_coo(void *hv, int *rows, int *cols, double *vals) {
  Handle *h = (Handle *)hv;
  for (size_t i = 0; i < h->coo.row.size(); i++) { rows[i] = h->coo.row[i] - 1; cols[i] = h->coo.col[i] - 1; vals[i] = h->coo.val[i]; }
}
int csxo_nparts(void *hv) { return (int)((Handle *)hv)->tuned.parts.size(); }
const char *csxo_log(void *hv) { return ((Handle *)hv)->tuned.log.c_str(); }

// what: 0 nnz, 1 nrows, 2 ncols, 3 row_start, 4 ctl_size, 5 row_jumps, 6 id_map_len (incl. -1), 7 map_len, 8 dvalues_len
long csxo_part_info(void *hv, int part, int what) {
  const CsxPart &p = ((Handle *)hv)->tuned.parts[part];
  switch (what) {
    case 0: return p.nnz;
    case 1: return p.nrows;
    case 2: return p.ncols;
    case 3: return p.row_start;
    case 4: return (long)p.ctl.size();
    case 5: return p.row_jumps;
    case 6: return (long)p.id_map.size();
    case 7: return (long)p.map_cpus.size();
    case 8: return (long)p.dvalues.size();
  }
  return -1;
}
// what: 0 values(double) 1 ctl(u8) 2 id_map(long) 3 rows_info(int x3) 4 dvalues(double) 5 map_cpus(u32) 6 map_pos(u32)
void csxo_part_copy(void *hv, int part, int what, void *dst) {
  const CsxPart &p = ((Handle *)hv)->tuned.parts[part];
  switch (what) {
    case 0: memcpy(dst, p.values.data(), p.values.size() * 8); break;
    case 1: memcpy(dst, p.ctl.data(), p.ctl.size()); break;
    case 2: memcpy(dst, p.id_map.data(), p.id_map.size() * sizeof(long)); break;
    case 3: memcpy(dst, p.rows_info.data(), p.rows_info.size() * sizeof(RowInfo)); break;
    case 4: memcpy(dst, p.dvalues.data(), p.dvalues.size() * 8); break;
    case 5: memcpy(dst, p.map_cpus.data(), p.map_cpus.size() * 4); break;
    case 6: memcpy(dst, p.map_pos.data(), p.map_pos.size() * 4); break;
  }
}
void csxo_spmv(void *hv, double alpha, const double *x, double beta, double *y, int overwrite) {
  spmv(((Handle *)hv)->tuned, alpha, x, beta, y, overwrite != 0);
}
void csxo_decode(void *hv, int part, int *rows, int *cols) {
  std::vector<int> r, c;
  decode_coords(((Handle *)hv)->tuned, part, r, c);
  memcpy(rows, r.data(), r.size() * 4);
  memcpy(cols, c.data(), c.size() * 4);
}

// Timed loop for the "port" CPU baseline: `loops` back-to-back SpMVs, one
// std::thread per partition for the non-symmetric case (row-partitioned like
// CsxKernels.cpp:82-103).  Returns seconds.
double csxo_bench(void *hv, double alpha, const double *x, double *y, int loops);
}

namespace csxo {
void part_multiply_public(const CsxPart &csx, bool full_colind, const double *x, double *y, double scale_f);
}

// Persistent worker threads + barrier, following MatVecMult (CsxKernels.cpp:82-103): the main thread
// zeroes y serially (VecInit), releases the workers, runs partition 0 itself, joins at a barrier.
#include <pthread.h>
extern "C" double csxo_bench(void *hv, double alpha, const double *x, double *y, int loops) {
  Handle *h = (Handle *)hv;
  const Tuned &A = h->tuned;
  size_t nt = A.parts.size();
  if (A.symmetric || nt == 1) {
    auto t0 = std::chrono::steady_clock::now();
    for (int l = 0; l < loops; l++) spmv(A, alpha, x, 0.0, y, true);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, nullptr, (unsigned)nt);
  std::vector<std::thread> th;
  for (size_t t = 1; t < nt; t++)
    th.emplace_back([&, t]() {
      cpu_set_t set; CPU_ZERO(&set); CPU_SET((int)t, &set);
      pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
      for (int l = 0; l < loops; l++) {
        pthread_barrier_wait(&bar);
        part_multiply_public(A.parts[t], A.full_colind, x, y, alpha);
        pthread_barrier_wait(&bar);
      }
    });
  auto t0 = std::chrono::steady_clock::now();
  for (int l = 0; l < loops; l++) {
    for (long i = 0; i < A.nrows; i++) y[i] = 0;
    pthread_barrier_wait(&bar);
    part_multiply_public(A.parts[0], A.full_colind, x, y, alpha);
    pthread_barrier_wait(&bar);
  }
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (auto &t : th) t.join();
  pthread_barrier_destroy(&bar);
  return secs;
}

// Same persistent-thread protocol, but each partition runs a function compiled from the reference's own
// kernel templates (oracle/_ref, see refkernels.py).  fns[i] has the spmv_fn_t signature of
// SpmvMethod.hpp:25-26; csx[i] points to a csx_matrix_t; xvec/yvec point to vector_t structs.
typedef void (*ref_spmv_fn)(void *, void *, void *, double, void *);
struct ref_vector { double *elements; size_t size; int alloc_type; int vec_mode; };
extern "C" double csxo_ref_bench(int nparts, void **fns, void **csx, void *xvec, void *yvec, double alpha, int loops) {
  ref_vector *y = (ref_vector *)yvec;
  size_t nt = (size_t)nparts;
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, nullptr, (unsigned)nt);
  std::vector<std::thread> th;
  for (size_t t = 1; t < nt; t++)
    th.emplace_back([&, t]() {
      cpu_set_t set; CPU_ZERO(&set); CPU_SET((int)t, &set);
      pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
      for (int l = 0; l < loops; l++) {
        pthread_barrier_wait(&bar);
        ((ref_spmv_fn)fns[t])(csx[t], xvec, yvec, alpha, nullptr);
        pthread_barrier_wait(&bar);
      }
    });
  auto t0 = std::chrono::steady_clock::now();
  for (int l = 0; l < loops; l++) {
    for (size_t i = 0; i < y->size; i++) y->elements[i] = 0;   // VecInit(y, 0), CsxKernels.cpp:93
    pthread_barrier_wait(&bar);
    ((ref_spmv_fn)fns[0])(csx[0], xvec, yvec, alpha, nullptr);
    pthread_barrier_wait(&bar);
  }
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (auto &t : th) t.join();
  pthread_barrier_destroy(&bar);
  return secs;
}
