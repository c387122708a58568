// MatrixMarket-style reader behind spx_input_load_mmf.
// Behaviour follows include/sparsex/internals/Mmf.hpp:331-478 and Mmf.cpp:58-77:
//   * the "%%MatrixMarket matrix coordinate <field> general|symmetric" banner is
//     optional; extension tokens 0-base / 1-base / column / row may follow it;
//   * with a banner the file is column-wise by default and is loaded and sorted;
//     `symmetric` files are expanded to the full matrix first;
//   * without a banner the entries are streamed and must already be sorted
//     row-major ("indices are not sorted in MMF file").
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "csx_host.hpp"

namespace spxb {
namespace {
bool next_line_tokens(std::istream &in, std::vector<std::string> &tok) {
  std::string line;
  tok.clear();
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    std::string t;
    while (ss >> t) tok.push_back(t);
    if (!tok.empty()) return true;
  }
  return false;
}
struct Entry { int r, c; double v; };
}  // namespace

std::string read_mmf(const char *path, CooHost &out) {
  std::ifstream in(path);
  if (!in.is_open()) return "MMF file error";
  std::vector<std::string> tok;
  if (!next_line_tokens(in, tok)) return "MMF file error";
  bool symmetric = false, column_wise = true, zero_based = false, bare = false;
  if (tok[0] != "%%MatrixMarket") {
    if (tok[0].size() > 2 && tok[0][0] == '%' && tok[0][1] == '%') return "invalid header line in MMF file";
    bare = true;          // first line is already the size line (or a plain comment)
    column_wise = false;
  } else {
    if (tok.size() < 5) return "less arguments in header line of MMF file";
    for (auto &t : tok) for (auto &ch : t) ch = (char)std::tolower((unsigned char)ch);
    if (tok[1] != "matrix") return "unsupported object in header line of MMF file";
    if (tok[2] != "coordinate") return "unsupported matrix format in header line of MMF file";
    if (tok[4] == "symmetric") symmetric = true;
    else if (tok[4] != "general") return "unsupported symmetry in header line of MMF file";
    for (size_t i = 5; i < tok.size(); i++) {
      if (tok[i] == "0-base") zero_based = true;
      else if (tok[i] == "1-base") zero_based = false;
      else if (tok[i] == "column") column_wise = true;
      else if (tok[i] == "row") column_wise = false;
    }
  }
  if (!bare || tok[0][0] == '%') {
    while (in.peek() == '%') { std::string skip; std::getline(in, skip); }
    if (!next_line_tokens(in, tok)) return "size line error in MMF file";
  }
  if (tok.size() != 3) return "bad input, less arguments in line of MMF file";
  long nr = std::atol(tok[0].c_str()), nc = std::atol(tok[1].c_str()), nnz = std::atol(tok[2].c_str());
  if (nr < 0 || nc < 0 || nnz < 0) return "size line error in MMF file";
  std::vector<Entry> es;
  es.reserve(symmetric ? 2 * (size_t)nnz : (size_t)nnz);
  bool buffered = symmetric || column_wise;
  int prev_r = 0, prev_c = 0;
  for (long i = 0; i < nnz; i++) {
    if (!next_line_tokens(in, tok)) return "Requesting dereference, but mmf ended.";
    if (tok.size() != 3) return "bad input, less arguments in line of MMF file";
    Entry e{std::atoi(tok[0].c_str()), std::atoi(tok[1].c_str()), std::strtod(tok[2].c_str(), nullptr)};
    if (zero_based) { e.r++; e.c++; }
    if (buffered) {
      es.push_back(e);
      if (symmetric && e.r != e.c) es.push_back(Entry{e.c, e.r, e.v});
    } else {
      if (e.r < prev_r || (e.r == prev_r && e.c < prev_c)) return "indices are not sorted in MMF file";
      prev_c = (e.r == prev_r) ? e.c : 1;
      prev_r = e.r;
      es.push_back(e);
    }
  }
  if (buffered)
    std::sort(es.begin(), es.end(), [](const Entry &a, const Entry &b) { return a.r < b.r || (a.r == b.r && a.c < b.c); });
  out = CooHost();
  out.nrows = nr; out.ncols = nc;
  out.buffered = buffered;
  out.row.reserve(es.size()); out.col.reserve(es.size()); out.val.reserve(es.size());
  for (const Entry &e : es) {
    if (e.r < 1 || e.r > nr || e.c < 1 || e.c > nc) return "index out of bounds in MMF file";
    out.row.push_back(e.r); out.col.push_back(e.c); out.val.push_back(e.v);
  }
  return "";
}

}  // namespace spxb
