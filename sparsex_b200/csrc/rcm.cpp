// Reverse Cuthill-McKee reordering behind spx_mat_tune(input, SPX_MAT_REORDER).
//
// Reference: include/sparsex/internals/Rcm.hpp:116-340 (FindPerm, ConstructGraph_CSR/_MMF, ReorderMat_CSR/_MMF),
// which hands the work to the Boost Graph Library (boost/graph/cuthill_mckee_ordering.hpp,
// boost/graph/detail/sparse_ordering.hpp; an external dependency that is neither vendored in the reference tree nor
// present in this image).  What follows restates the BGL algorithm as published, step by step, on flat arrays:
//
//   graph       adjacency_list<vecS, vecS, undirectedS>: add_edge(u, v) appends v to u's list and u to v's list, so
//               the adjacency order is the order in which the matrix iterator delivered the elements; parallel edges
//               stay (a structurally symmetric CSR input yields every edge twice) and count in the degree;
//   components  cuthill_mckee_ordering(G, out, color, degree): one representative per connected component, lowest
//               vertex first (the depth_first_visit only colours the component);
//   start node  find_starting_node: pseudo_peripheral_pair(G, r) runs a BFS over an rcm_queue, which reports the
//               eccentricity of r and the "spouse" = the first vertex of minimal degree in the last BFS level;
//               x = spouse(r), y = spouse(x); while ecc(x) > ecc(r): r = x, x = y, y = spouse(x);
//   ordering    breadth_first_visit over a sparse_ordering_queue with bfs_rcm_visitor: a vertex is emitted when it is
//               popped, and when it is finished the vertices it has just pushed are sorted by degree with std::sort
//               (the same call on the same sequence here, so ties fall the same way under the same libstdc++);
//   reversal    the output iterator is inv_perm.rbegin(); perm[inv_perm[i]] = i (Rcm.hpp:131-142).
//
// Boost cannot be run here.  The pin against it is the sample output the BGL documentation prints for Boost's own example
// program (libs/graph/example/cuthill_mckee_ordering.cpp: 10 vertices, 14 edges; the orderings from vertex 6, from vertex 0
// and without a starting vertex, bandwidth 8 -> 4), which this file reproduces (tests/test_cpu_rcm.py:
// test_boost_documentation_example); beyond that it equals the structural restatement in oracle/rcm_oracle.cpp on seeded
// graphs, and the properties the reference relies on hold (perm is a bijection, B = P A P^T, SpMV results agree after
// spx_vec_reorder / spx_vec_inv_reorder).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include <csx_b200.h>

#include "csx_host.hpp"

namespace spxb {
namespace {

struct Graph {
  int64_t n = 0;
  std::vector<int64_t> off;   // n + 1
  std::vector<int32_t> adj;   // neighbours in insertion order
  int64_t degree(int32_t v) const { return off[v + 1] - off[v]; }
};

// Edge list (in iteration order) -> adjacency arrays that keep the per-vertex insertion order of add_edge.
void build_graph(int64_t n, const std::vector<int32_t> &eu, const std::vector<int32_t> &ev, Graph &g) {
  g.n = n;
  g.off.assign(n + 1, 0);
  for (size_t e = 0; e < eu.size(); e++) { g.off[eu[e] + 1]++; g.off[ev[e] + 1]++; }
  for (int64_t v = 0; v < n; v++) g.off[v + 1] += g.off[v];
  g.adj.resize(g.off[n]);
  std::vector<int64_t> fill(g.off.begin(), g.off.end() - 1);
  for (size_t e = 0; e < eu.size(); e++) {
    g.adj[fill[eu[e]]++] = ev[e];
    g.adj[fill[ev[e]]++] = eu[e];
  }
}

// pseudo_peripheral_pair (sparse_ordering.hpp): BFS from u through an rcm_queue.  `mark`/`stamp` stand in for the
// colour map that the BGL resets for every call.
int32_t pseudo_peripheral_pair(const Graph &g, int32_t u, int &ecc, std::vector<uint32_t> &mark, uint32_t stamp,
                               std::vector<int32_t> &queue) {
  queue.clear();
  queue.push_back(u);
  mark[u] = stamp;
  size_t head = 0, popped_in_level = 0, level_size = 1;
  int eccen = -1;
  int32_t w = u;
  while (head < queue.size()) {
    int32_t v = queue[head];
    // rcm_queue::front(): remember the first vertex of minimal degree of the level being emptied
    if (popped_in_level == 0) w = v;
    else if (g.degree(v) < g.degree(w)) w = v;
    // rcm_queue::pop(): the queue holds exactly one BFS level when its first vertex leaves
    if (popped_in_level == 0) level_size = queue.size() - head;
    head++;
    if (popped_in_level == level_size - 1) { popped_in_level = 0; eccen++; }
    else popped_in_level++;
    for (int64_t k = g.off[v]; k < g.off[v + 1]; k++) {
      int32_t t = g.adj[k];
      if (mark[t] != stamp) { mark[t] = stamp; queue.push_back(t); }
    }
  }
  ecc = eccen;
  return w;
}

int32_t find_starting_node(const Graph &g, int32_t r, std::vector<uint32_t> &mark, uint32_t &stamp, std::vector<int32_t> &queue) {
  int eccen_r, eccen_x;
  int32_t x = pseudo_peripheral_pair(g, r, eccen_r, mark, ++stamp, queue);
  int32_t y = pseudo_peripheral_pair(g, x, eccen_x, mark, ++stamp, queue);
  while (eccen_x > eccen_r) {
    r = x;
    eccen_r = eccen_x;
    x = y;
    y = pseudo_peripheral_pair(g, x, eccen_x, mark, ++stamp, queue);
  }
  return x;
}

int64_t bandwidth_of(const std::vector<int32_t> &eu, const std::vector<int32_t> &ev, const int32_t *perm) {
  int64_t b = 0;
  for (size_t e = 0; e < eu.size(); e++) {
    int64_t a = perm ? perm[eu[e]] : eu[e], c = perm ? perm[ev[e]] : ev[e];
    b = std::max<int64_t>(b, a > c ? a - c : c - a);
  }
  return b;
}

}  // namespace

// FindPerm (Rcm.hpp:116-153).  perm: old index -> new index; inv_perm: new -> old.  bandwidth[0/1]: before / after
// (what the reference logs).  Returns false when there is no edge ("no reordering available for this matrix").
bool rcm_find_perm(int64_t n, const std::vector<int32_t> &eu, const std::vector<int32_t> &ev, std::vector<int32_t> &perm,
                   std::vector<int32_t> &inv_perm, int64_t *bandwidth, int64_t start) {
  if (eu.empty() || n <= 0) return false;
  Graph g;
  build_graph(n, eu, ev, g);
  std::vector<uint32_t> mark(n, 0);
  uint32_t stamp = 0;
  std::vector<int32_t> queue;
  queue.reserve(n);

  // one representative per component, lowest vertex first (or the one starting vertex the caller names:
  // cuthill_mckee_ordering(G, s, ...), for connected graphs)
  std::vector<int32_t> starts;
  if (start >= 0) starts.push_back((int32_t)start);
  else {
    std::vector<int32_t> stack;
    ++stamp;
    for (int64_t v = 0; v < n; v++) {
      if (mark[v] == stamp) continue;
      starts.push_back((int32_t)v);
      mark[v] = stamp;
      stack.push_back((int32_t)v);
      while (!stack.empty()) {
        int32_t u = stack.back();
        stack.pop_back();
        for (int64_t k = g.off[u]; k < g.off[u + 1]; k++)
          if (mark[g.adj[k]] != stamp) { mark[g.adj[k]] = stamp; stack.push_back(g.adj[k]); }
      }
    }
  }
  if (start < 0)
    for (int32_t &s : starts)
      if (g.degree(s) > 0) s = find_starting_node(g, s, mark, stamp, queue);   // an isolated vertex is its own start

  // Cuthill-McKee: BFS per component, each vertex's newly discovered neighbours sorted by degree
  std::vector<int32_t> order;
  order.reserve(n);
  ++stamp;
  auto by_degree = [&g](int32_t a, int32_t b) { return g.degree(a) < g.degree(b); };
  for (int32_t s : starts) {
    queue.clear();
    queue.push_back(s);
    mark[s] = stamp;
    size_t head = 0;
    while (head < queue.size()) {
      int32_t u = queue[head++];
      order.push_back(u);                       // examine_vertex
      size_t index_begin = queue.size();        // == Q.size() after the pop, relative to the front
      for (int64_t k = g.off[u]; k < g.off[u + 1]; k++) {
        int32_t t = g.adj[k];
        if (mark[t] != stamp) { mark[t] = stamp; queue.push_back(t); }
      }
      std::sort(queue.begin() + index_begin, queue.end(), by_degree);   // finish_vertex
    }
  }
  if ((int64_t)order.size() != n) return false;   // a named starting vertex only reaches its own component
  // written through inv_perm.rbegin(): reversed
  inv_perm.assign(n, 0);
  perm.assign(n, 0);
  for (int64_t k = 0; k < n; k++) inv_perm[n - 1 - k] = order[k];
  for (int64_t i = 0; i < n; i++) perm[inv_perm[i]] = (int32_t)i;
  if (bandwidth) {
    bandwidth[0] = bandwidth_of(eu, ev, nullptr);
    bandwidth[1] = bandwidth_of(eu, ev, perm.data());
  }
  return true;
}

// ConstructGraph_CSR (Rcm.hpp:242-287).  The C API always wraps CSR inputs as non-symmetric (Facade.cpp:29-42), so
// every off-diagonal element is an edge; `symmetric` keeps the other branch (upper triangle only).
void rcm_edges_csr(const int32_t *rowptr, const int32_t *colind, int64_t nrows, bool symmetric, std::vector<int32_t> &eu,
                   std::vector<int32_t> &ev) {
  eu.clear();
  ev.clear();
  for (int64_t r = 0; r < nrows; r++)
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; k++) {
      int32_t c = colind[k];
      if (symmetric ? r < c : r != c) { eu.push_back((int32_t)r); ev.push_back(c); }
    }
}

// ConstructGraph_MMF (Rcm.hpp:155-204) for inputs the reader buffers (symmetric or column-wise files, Mmf.hpp:218):
// edges are the elements above the diagonal, in row-major order.  For a general row-wise file the reference's loop
// runs over an empty buffer (SetReordered(true) precedes begin(), Mmf.hpp:241-252, 303-308), finds no edge and
// leaves the matrix in its given order: the caller mirrors that.
void rcm_edges_coo(const CooHost &coo, std::vector<int32_t> &eu, std::vector<int32_t> &ev) {
  eu.clear();
  ev.clear();
  for (size_t i = 0; i < coo.row.size(); i++)
    if (coo.row[i] < coo.col[i]) { eu.push_back(coo.row[i] - 1); ev.push_back(coo.col[i] - 1); }
}

// ReorderMat_MMF (Rcm.hpp:206-217): both coordinates through perm, then the row-major sort.
void rcm_apply_coo(CooHost &coo, const std::vector<int32_t> &perm) {
  const size_t nnz = coo.row.size();
  std::vector<size_t> idx(nnz);
  std::iota(idx.begin(), idx.end(), (size_t)0);
  for (size_t i = 0; i < nnz; i++) { coo.row[i] = perm[coo.row[i] - 1] + 1; coo.col[i] = perm[coo.col[i] - 1] + 1; }
  std::sort(idx.begin(), idx.end(), [&coo](size_t a, size_t b) {
    return coo.row[a] < coo.row[b] || (coo.row[a] == coo.row[b] && coo.col[a] < coo.col[b]);
  });
  std::vector<int> r(nnz), c(nnz);
  std::vector<double> v(nnz);
  for (size_t i = 0; i < nnz; i++) { r[i] = coo.row[idx[i]]; c[i] = coo.col[idx[i]]; v[i] = coo.val[idx[i]]; }
  coo.row.swap(r); coo.col.swap(c); coo.val.swap(v);
}

}  // namespace spxb

extern "C" {

int csxb_rcm_csr(const int32_t *rowptr, const int32_t *colind, int64_t nrows, int64_t ncols, int32_t *perm,
                 int64_t *bandwidth) {
  if (!rowptr || !colind || !perm || nrows <= 0 || nrows != ncols) return -1;
  for (int64_t k = rowptr[0]; k < rowptr[nrows]; k++)
    if (colind[k] < 0 || colind[k] >= ncols) return -1;
  std::vector<int32_t> eu, ev, p, ip;
  spxb::rcm_edges_csr(rowptr, colind, nrows, false, eu, ev);
  if (!spxb::rcm_find_perm(nrows, eu, ev, p, ip, bandwidth)) return 1;
  std::memcpy(perm, p.data(), sizeof(int32_t) * (size_t)nrows);
  return 0;
}

// The ordering of an explicit undirected edge list (edges added in the given order): what csxb_rcm_csr runs after it
// has listed the off-diagonal elements.  For callers that hold a graph rather than a matrix, and for known-answer tests.
int csxb_rcm_edges(const int32_t *eu, const int32_t *ev, int64_t nedges, int64_t n, int64_t start, int32_t *perm,
                   int64_t *bandwidth) {
  if (!eu || !ev || !perm || n <= 0 || nedges < 0 || start >= n) return -1;
  for (int64_t e = 0; e < nedges; e++)
    if (eu[e] < 0 || eu[e] >= n || ev[e] < 0 || ev[e] >= n || eu[e] == ev[e]) return -1;
  std::vector<int32_t> u(eu, eu + nedges), v(ev, ev + nedges), p, ip;
  if (!spxb::rcm_find_perm(n, u, v, p, ip, bandwidth, start)) return 1;
  std::memcpy(perm, p.data(), sizeof(int32_t) * (size_t)n);
  return 0;
}

// ReorderMat_CSR + the reordered CSR iterator (Rcm.hpp:289-316, Csr.hpp:270-360): row i of the result is row
// inv_perm[i] of the input, its columns mapped through perm and sorted.
int csxb_permute_csr(const int32_t *rowptr, const int32_t *colind, const double *values, int64_t nrows, const int32_t *perm,
                     int32_t *out_rowptr, int32_t *out_colind, double *out_values) {
  if (!rowptr || !colind || !values || !perm || !out_rowptr || !out_colind || !out_values || nrows < 0) return -1;
  std::vector<int32_t> inv(nrows);
  for (int64_t i = 0; i < nrows; i++) {
    if (perm[i] < 0 || perm[i] >= nrows) return -1;
    inv[perm[i]] = (int32_t)i;
  }
  out_rowptr[0] = 0;
  for (int64_t i = 0; i < nrows; i++) out_rowptr[i + 1] = out_rowptr[i] + (rowptr[inv[i] + 1] - rowptr[inv[i]]);
  std::vector<std::pair<int32_t, double>> row;
  for (int64_t i = 0; i < nrows; i++) {
    const int64_t b = rowptr[inv[i]], e = rowptr[inv[i] + 1];
    row.clear();
    for (int64_t k = b; k < e; k++) row.emplace_back(perm[colind[k]], values[k]);
    std::sort(row.begin(), row.end(), [](const std::pair<int32_t, double> &l, const std::pair<int32_t, double> &r) { return l.first < r.first; });
    int64_t o = out_rowptr[i];
    for (auto &pr : row) { out_colind[o] = pr.first; out_values[o] = pr.second; o++; }
  }
  return 0;
}

}  // extern "C"
