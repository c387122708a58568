// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into or loaded by the product library).
//
// Reverse Cuthill-McKee as the reference computes it: include/sparsex/internals/Rcm.hpp:116-153 (FindPerm) and
// :242-287 (ConstructGraph_CSR) call boost::cuthill_mckee_ordering on an
// adjacency_list<vecS, vecS, undirectedS>.  The Boost Graph Library is an external dependency of the reference
// (configure.ac asks for Boost >= 1.48; no version is pinned, nothing is vendored) and is absent from this image, so
// this file restates the published BGL code — boost/graph/cuthill_mckee_ordering.hpp and
// boost/graph/detail/sparse_ordering.hpp — keeping its structure: an adjacency list filled by add_edge, a
// breadth_first_visit with visitor hooks, the rcm_queue of pseudo_peripheral_pair, the sparse_ordering_queue and the
// bfs_rcm_visitor of the ordering itself.  Boost cannot be run here; the pin is the sample output the BGL documentation
// prints for Boost's own example program (three orderings of a 10-vertex graph), reproduced in tests/test_cpu_rcm.py.
#include <algorithm>
#include <cstdint>
#include <deque>
#include <queue>
#include <vector>

namespace {

enum Color { WHITE, GRAY, BLACK };

struct Graph {   // adjacency_list<vecS, vecS, undirectedS>
  std::vector<std::vector<int>> out;
  explicit Graph(size_t n) : out(n) {}
  void add_edge(int u, int v) { out[u].push_back(v); out[v].push_back(u); }
  size_t degree(int v) const { return out[v].size(); }
  size_t num_vertices() const { return out.size(); }
};

// boost::breadth_first_visit(g, s, Q, vis, color)
template <class Queue, class Visitor>
void breadth_first_visit(const Graph &g, int s, Queue &Q, Visitor &vis, std::vector<Color> &color) {
  color[s] = GRAY;
  Q.push(s);
  while (!Q.empty()) {
    int u = Q.top();
    Q.pop();
    vis.examine_vertex(u);
    for (int v : g.out[u]) {
      if (color[v] == WHITE) {
        color[v] = GRAY;
        Q.push(v);
      }
    }
    color[u] = BLACK;
    vis.finish_vertex(u);
  }
}

struct NullVisitor {
  void examine_vertex(int) {}
  void finish_vertex(int) {}
};

// sparse::rcm_queue<Vertex, DegreeMap>
class RcmQueue : public std::queue<int> {
  typedef std::queue<int> base;
public:
  explicit RcmQueue(const Graph &g) : _size(0), Qsize(1), eccen(-1), w(0), g_(g) {}
  void pop() {
    if (!_size) Qsize = base::size();
    base::pop();
    if (_size == Qsize - 1) {
      _size = 0;
      ++eccen;
    } else {
      ++_size;
    }
  }
  int &front() {
    int &u = base::front();
    if (_size == 0) w = u;
    else if (g_.degree(u) < g_.degree(w)) w = u;
    return u;
  }
  int &top() { return front(); }
  int eccentricity() const { return eccen; }
  int spouse() const { return w; }
private:
  size_t _size, Qsize;
  int eccen;
  int w;
  const Graph &g_;
};

int pseudo_peripheral_pair(const Graph &G, int u, int &ecc, std::vector<Color> &color) {
  RcmQueue Q(G);
  for (size_t v = 0; v < G.num_vertices(); v++) color[v] = WHITE;
  NullVisitor vis;
  breadth_first_visit(G, u, Q, vis, color);
  ecc = Q.eccentricity();
  return Q.spouse();
}

int find_starting_node(const Graph &G, int r, std::vector<Color> &color) {
  int x, y, eccen_r, eccen_x;
  x = pseudo_peripheral_pair(G, r, eccen_r, color);
  y = pseudo_peripheral_pair(G, x, eccen_x, color);
  while (eccen_x > eccen_r) {
    r = x;
    eccen_r = eccen_x;
    x = y;
    y = pseudo_peripheral_pair(G, x, eccen_x, color);
  }
  return x;
}

// sparse_ordering_queue: a queue whose container can be indexed
struct OrderingQueue {
  std::deque<int> c;
  void push(int v) { c.push_back(v); }
  void pop() { c.pop_front(); }
  int top() const { return c.front(); }
  bool empty() const { return c.empty(); }
  size_t size() const { return c.size(); }
};

struct BfsRcmVisitor {   // detail::bfs_rcm_visitor
  std::vector<int> *permutation;
  OrderingQueue *Qptr;
  const Graph *g;
  size_t index_begin = 0;
  void examine_vertex(int u) {
    permutation->push_back(u);
    index_begin = Qptr->size();
  }
  void finish_vertex(int) {
    const Graph *gg = g;
    std::sort(Qptr->c.begin() + index_begin, Qptr->c.end(), [gg](int a, int b) { return gg->degree(a) < gg->degree(b); });
  }
};

void dfs_mark(const Graph &G, int s, std::vector<Color> &color) {   // depth_first_visit with a null visitor
  std::vector<int> stack{s};
  color[s] = BLACK;
  while (!stack.empty()) {
    int u = stack.back();
    stack.pop_back();
    for (int v : G.out[u])
      if (color[v] == WHITE) { color[v] = BLACK; stack.push_back(v); }
  }
}

}  // namespace

static int order_graph(Graph &graph, int64_t n, int32_t *perm, int start = -1) {
  std::vector<Color> color((size_t)n, WHITE);
  std::deque<int> vertex_queue;
  if (start >= 0) {
    // cuthill_mckee_ordering(G, s, permutation, color, degree): the caller names the starting vertex (connected graphs)
    vertex_queue.push_front(start);
  } else {
    // cuthill_mckee_ordering(G, permutation, color, degree)
    for (int64_t v = 0; v < n; v++)
      if (color[v] == WHITE) { dfs_mark(graph, (int)v, color); vertex_queue.push_back((int)v); }
    for (int &s : vertex_queue) s = find_starting_node(graph, s, color);
  }
  // cuthill_mckee_ordering(G, vertex_queue, permutation, color, degree)
  std::vector<int> visit_order;
  OrderingQueue Q;
  BfsRcmVisitor vis{&visit_order, &Q, &graph};
  for (int64_t v = 0; v < n; v++) color[v] = WHITE;
  while (!vertex_queue.empty()) {
    int s = vertex_queue.front();
    vertex_queue.pop_front();
    breadth_first_visit(graph, s, Q, vis, color);
  }
  // FindPerm: the output iterator is inv_perm.rbegin()
  std::vector<int> inv_perm((size_t)n);
  for (int64_t k = 0; k < n; k++) inv_perm[(size_t)(n - 1 - k)] = visit_order[(size_t)k];
  for (int64_t i = 0; i < n; i++) perm[inv_perm[(size_t)i]] = (int32_t)i;
  return 0;
}

// rowptr/colind zero-based, n x n.  perm[old] = new.  Returns 0, or 1 when no edge exists (Rcm.hpp:275-279).
extern "C" int rcm_oracle_csr(const int32_t *rowptr, const int32_t *colind, int64_t n, int symmetric, int32_t *perm) {
  Graph graph((size_t)n);
  size_t edges = 0;
  for (int64_t r = 0; r < n; r++)
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; k++)
      if (symmetric ? r < colind[k] : r != colind[k]) { graph.add_edge((int)r, colind[k]); edges++; }
  if (!edges) return 1;
  return order_graph(graph, n, perm);
}

// The same for an explicit edge list, added in the given order (known-answer graphs).
extern "C" int rcm_oracle_edges(const int32_t *eu, const int32_t *ev, int64_t ne, int64_t n, int start, int32_t *perm) {
  Graph graph((size_t)n);
  for (int64_t e = 0; e < ne; e++) graph.add_edge(eu[e], ev[e]);
  if (!ne) return 1;
  return order_graph(graph, n, perm, start);
}
