// Several GPUs behind one matrix handle, in ONE process (included at the end of engine.cu).
//
// The reference runs its nr_threads partitions on a pool of worker threads that share x and y
// (src/internals/CsxKernels.cpp:35-129: MatVecMult / MatVecMult_sym, do_mv_thread, the CSX-Sym local buffers and their
// map reduction, CsxSpmv.cpp:37-50).  Here the partitions of one tuned matrix are dealt out to the GPUs of the box in
// contiguous ranges ("members"): every member is a csxb_matrix of its own on its own device.  One SpMV =
//   1. every member receives the columns of x its partitions read (and, for beta != 0, its rows of y),
//   2. runs the ordinary kernels of csxb_spmv on its own stream,
//   3. CSX-Sym: the sums a member produced for rows of lower members (its halo pseudo-partition) travel to the owners,
//      which add them in member order (deterministic; the cross-device form of the reference's reduction phase),
//   4. every member returns its rows of y.
// x and y may live anywhere cudaMemcpyDefault can reach: plain or pinned host memory, managed memory, device memory.
// With plain host buffers on both sides and a non-symmetric matrix the members run their slab-pipelined
// csxb_spmv_host concurrently (one host thread per member) instead.
#include <thread>

struct csxb_group {
  struct Member {
    csxb_matrix *m = nullptr;
    int device = 0;
    int64_t row_lo = 0, row_hi = 0;     // rows this member returns
    int64_t col_lo = 0, col_hi = 0;     // columns of x it needs
    double *d_x = nullptr, *d_y = nullptr, *d_tmp = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
  };
  std::vector<Member> mem;
  int64_t nrows = 0, ncols = 0, nnz = 0;
  bool symmetric = false;
  int nparts_total = 0;
  std::vector<int32_t> permutation;
  ~csxb_group() {
    int cur = -1;
    cudaGetDevice(&cur);
    for (Member &g : mem) {
      cudaSetDevice(g.device);
      if (g.d_x) cudaFree(g.d_x);
      if (g.d_y) cudaFree(g.d_y);
      if (g.d_tmp) cudaFree(g.d_tmp);
      if (g.done) cudaEventDestroy(g.done);
      if (g.stream) cudaStreamDestroy(g.stream);
      delete g.m;
    }
    if (cur >= 0) cudaSetDevice(cur);
  }
};

static bool plain_host_pointer(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered || a.type == cudaMemoryTypeHost;
}

extern "C" {

csxb_group_t *csxb_group_create(csxb_matrix_t *whole, const int *devices, int ndev, int free_host, char *err, size_t errlen) {
  if (!whole || !devices || ndev < 1) { put_err(err, errlen, "invalid argument"); return nullptr; }
  CsxMatrix &H = whole->host;
  if (whole->uploaded || H.part_lo != 0 || (int)H.parts.size() != H.nparts_total) {
    put_err(err, errlen, "the matrix must hold all its partitions and must not be uploaded yet");
    return nullptr;
  }
  const int np = (int)H.parts.size();
  const int nm = std::max(1, std::min(ndev, np));   // never more members than partitions
  csxb_group *G = new csxb_group;
  G->nrows = H.nrows; G->ncols = H.ncols; G->nnz = H.nnz; G->symmetric = H.symmetric; G->nparts_total = H.nparts_total;
  G->permutation = H.permutation;
  G->mem.resize(nm);
  for (int g = 0; g < nm; g++) {
    const int p0 = (int)((int64_t)np * g / nm), p1 = (int)((int64_t)np * (g + 1) / nm);
    csxb_group::Member &M = G->mem[g];
    M.device = devices[g];
    M.m = new csxb_matrix;
    CsxMatrix &S = M.m->host;
    S.nrows = H.nrows; S.ncols = H.ncols; S.nnz = H.nnz; S.symmetric = H.symmetric; S.full_colind = H.full_colind;
    S.rows_per_thread = H.rows_per_thread; S.slice_elems = H.slice_elems; S.slab_rows = H.slab_rows;
    S.nparts_total = H.nparts_total; S.part_lo = p0;
    for (int p = p0; p < p1; p++) S.parts.push_back(std::move(H.parts[p]));
  }
  H.parts.clear();
  for (int g = 0; g < nm; g++) {
    csxb_group::Member &M = G->mem[g];
    auto bail = [&](const std::string &what) { put_err(err, errlen, what); delete G; return (csxb_group_t *)nullptr; };
    if (csxb_upload(M.m, M.device, free_host) != 0) return bail("member " + std::to_string(g) + ": " + g_last_error);
    const CsxMatrix &S = M.m->host;
    M.row_lo = S.parts.empty() ? 0 : S.parts.front().row_start;
    M.row_hi = M.m->covered_rows_end;
    if (S.parts.empty()) M.row_hi = M.row_lo;
    int64_t cmin = S.ncols, cmax = -1;
    for (auto &p : S.parts) if (p.col_max >= p.col_min) { cmin = std::min(cmin, p.col_min); cmax = std::max(cmax, p.col_max); }
    M.col_lo = cmax >= cmin ? cmin : 0;
    M.col_hi = cmax >= cmin ? cmax + 1 : 0;
    if (S.symmetric) {   // the diagonal and the transposed images read x at the member's own rows as well
      M.col_lo = std::min(M.col_lo, M.row_lo);
      M.col_hi = std::max(M.col_hi, M.row_hi);
      if (M.col_hi <= M.col_lo) { M.col_lo = M.row_lo; M.col_hi = M.row_hi; }
    }
    if (cudaSetDevice(M.device) != cudaSuccess) return bail("cudaSetDevice failed");
    const size_t nx = (size_t)std::max<int64_t>(S.ncols, 1), ny = (size_t)std::max<int64_t>(S.nrows, 1);
    if (cudaMalloc((void **)&M.d_x, nx * 8) != cudaSuccess || cudaMalloc((void **)&M.d_y, ny * 8) != cudaSuccess ||
        cudaMemset(M.d_x, 0, nx * 8) != cudaSuccess || cudaMemset(M.d_y, 0, ny * 8) != cudaSuccess ||
        (S.symmetric && cudaMalloc((void **)&M.d_tmp, (size_t)std::max<int64_t>(M.row_hi - M.row_lo, 1) * 8) != cudaSuccess) ||
        cudaStreamCreateWithFlags(&M.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&M.done, cudaEventDisableTiming) != cudaSuccess)
      return bail(std::string("member set-up: ") + cudaGetErrorString(cudaGetLastError()));
  }
  // direct NVLink copies between the members' buffers where the GPUs allow it (otherwise the driver stages them)
  for (int a = 0; a < nm; a++)
    for (int b = 0; b < nm; b++) {
      const int da = G->mem[a].device, db = G->mem[b].device;
      int can = 0;
      if (da == db || cudaDeviceCanAccessPeer(&can, da, db) != cudaSuccess || !can) { cudaGetLastError(); continue; }
      if (cudaSetDevice(da) == cudaSuccess) cudaDeviceEnablePeerAccess(db, 0);
      cudaGetLastError();   // already enabled is fine
    }
  delete whole;   // consumed: its partitions live in the members now
  return G;
}

void csxb_group_destroy(csxb_group_t *G) { delete G; }
int csxb_group_size(const csxb_group_t *G) { return G ? (int)G->mem.size() : 0; }
csxb_matrix_t *csxb_group_member(csxb_group_t *G, int i) { return (G && i >= 0 && (size_t)i < G->mem.size()) ? G->mem[i].m : nullptr; }
int csxb_group_device(const csxb_group_t *G, int i) { return (G && i >= 0 && (size_t)i < G->mem.size()) ? G->mem[i].device : -1; }

int csxb_group_spmv(csxb_group_t *G, double alpha, const double *x, double beta, double *y, int overwrite) {
  if (!G || !x || !y) return fail("invalid argument");
  const size_t nm = G->mem.size();
  int cur = -1;
  cudaGetDevice(&cur);
  struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{cur};

  bool all_slabbed = !G->symmetric;   // (a member without slabs would return rows it does not own)
  for (auto &M : G->mem) all_slabbed = all_slabbed && !M.m->slabs.empty();
  if (all_slabbed && plain_host_pointer(x) && plain_host_pointer(y)) {
    // user buffers on both sides: every member runs its slab-pipelined host-buffer path (uploads the columns it reads,
    // downloads its rows); the members are independent, so one host thread each
    std::vector<int> rc(nm, 0);
    std::vector<std::string> msg(nm);
    std::vector<std::thread> th;
    for (size_t g = 0; g < nm; g++)
      th.emplace_back([&, g]() {
        rc[g] = csxb_spmv_host(G->mem[g].m, alpha, x, beta, y, overwrite);
        if (rc[g]) msg[g] = g_last_error;
      });
    for (auto &t : th) t.join();
    for (size_t g = 0; g < nm; g++) if (rc[g]) return fail("member " + std::to_string(g) + ": " + msg[g]);
    return 0;
  }

  for (size_t g = 0; g < nm; g++) {
    csxb_group::Member &M = G->mem[g];
    CUDA_TRY(cudaSetDevice(M.device));
    if (M.col_hi > M.col_lo)
      CUDA_TRY(cudaMemcpyAsync(M.d_x + M.col_lo, x + M.col_lo, (size_t)(M.col_hi - M.col_lo) * 8, cudaMemcpyDefault, M.stream));
    if (!overwrite && M.row_hi > M.row_lo)
      CUDA_TRY(cudaMemcpyAsync(M.d_y + M.row_lo, y + M.row_lo, (size_t)(M.row_hi - M.row_lo) * 8, cudaMemcpyDefault, M.stream));
    if (csxb_spmv(M.m, alpha, M.d_x, beta, M.d_y, overwrite, M.stream)) return -1;
    CUDA_TRY(cudaEventRecord(M.done, M.stream));
  }
  if (G->symmetric) {
    // reduction phase: the owner of rows [lo, hi) adds what higher members summed up for them, lowest member first
    for (size_t o = 0; o < nm; o++) {
      csxb_group::Member &O = G->mem[o];
      bool touched = false;
      for (size_t g = o + 1; g < nm; g++) {
        csxb_group::Member &S = G->mem[g];
        const int64_t lo = std::max(S.m->sym_halo_lo, O.row_lo), hi = std::min(S.m->sym_halo_hi, O.row_hi);
        if (hi <= lo) continue;
        if (!touched) { CUDA_TRY(cudaSetDevice(O.device)); touched = true; }
        CUDA_TRY(cudaStreamWaitEvent(O.stream, S.done, 0));
        CUDA_TRY(cudaMemcpyPeerAsync(O.d_tmp + (lo - O.row_lo), O.device, S.d_y + lo, S.device, (size_t)(hi - lo) * 8, O.stream));
        if (csxb_vec_axpby(O.d_y + lo, O.d_y + lo, O.d_tmp + (lo - O.row_lo), 1.0, 1.0, hi - lo, O.stream)) return -1;
      }
    }
  }
  for (size_t g = 0; g < nm; g++) {
    csxb_group::Member &M = G->mem[g];
    // rows behind the last partition's last non-empty row belong to nobody: cleared by spx_matvec_mult, left alone by
    // spx_matvec_kernel (csxb_spmv did the clearing in the last member's d_y)
    const int64_t hi = (g + 1 == nm && overwrite) ? G->nrows : M.row_hi;
    if (hi <= M.row_lo) continue;
    CUDA_TRY(cudaSetDevice(M.device));
    CUDA_TRY(cudaMemcpyAsync(y + M.row_lo, M.d_y + M.row_lo, (size_t)(hi - M.row_lo) * 8, cudaMemcpyDefault, M.stream));
  }
  for (size_t g = 0; g < nm; g++) {
    CUDA_TRY(cudaSetDevice(G->mem[g].device));
    CUDA_TRY(cudaStreamSynchronize(G->mem[g].stream));
  }
  return 0;
}

// All partitions back in one container (csxb_save of a matrix that was never split reads the same).
int csxb_group_save(csxb_group_t *G, const char *path) {
  if (!G || !path) return fail("invalid argument");
  CsxMatrix W;
  for (size_t g = 0; g < G->mem.size(); g++) {
    csxb_matrix *m = G->mem[g].m;
    CsxMatrix &S = m->host;
    if (g == 0) {
      W.nrows = S.nrows; W.ncols = S.ncols; W.nnz = S.nnz; W.symmetric = S.symmetric; W.full_colind = S.full_colind;
      W.rows_per_thread = S.rows_per_thread; W.slice_elems = S.slice_elems; W.slab_rows = S.slab_rows;
      W.nparts_total = S.nparts_total; W.part_lo = 0;
    }
    for (size_t i = 0; i < S.parts.size(); i++) {
      CsxPartition p = S.parts[i];
      if ((int64_t)p.values.size() != p.nnz) {   // released at upload: back from the device
        CUDA_TRY(cudaSetDevice(m->device));
        p.values.resize((size_t)p.nnz);
        CUDA_TRY(cudaMemcpy(p.values.data(), m->d_values + m->layout.parts[i].val_base, (size_t)p.nnz * 8, cudaMemcpyDeviceToHost));
      }
      W.parts.push_back(std::move(p));
    }
  }
  W.permutation = G->permutation;
  std::string e = save_matrix(W, path);
  return e.empty() ? 0 : fail(e);
}

int csxb_group_get_entry(csxb_group_t *G, int64_t row, int64_t col, double *value) {
  if (!G) return fail("invalid argument");
  for (auto &M : G->mem) {
    int rc = csxb_get_entry(M.m, row, col, value);
    if (rc != 1) return rc;
  }
  return 1;
}
int csxb_group_set_entry(csxb_group_t *G, int64_t row, int64_t col, double value) {
  if (!G) return fail("invalid argument");
  // CSX-Sym stores A(r, c) = A(c, r) once, in the partition that owns the larger index
  for (auto &M : G->mem) {
    int rc = csxb_set_entry(M.m, row, col, value);
    if (rc != 1) return rc;
  }
  return 1;
}
int csxb_group_set_perm(csxb_group_t *G, const int32_t *perm, int64_t n) {
  if (!G || n < 0 || (n && (!perm || n != G->nrows))) return fail("invalid permutation");
  G->permutation.assign(perm, perm + n);
  return 0;
}

}  // extern "C"
