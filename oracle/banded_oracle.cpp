// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// CPU restatement ("oracle") of rust-bio 4.0.1 `bio::alignment::pairwise::banded::Aligner`
// and the pieces of `bio::alignment::sparse` / `data_structures::bit_tree` it calls:
//   banded.rs:406-869   compute_alignment (band-restricted fill, fix-ups, walk, leftovers)
//   banded.rs:872-1004  global / semiglobal / local presets
//   banded.rs:1047-1380 Band (add_kmer, add_entry, add_gap, set_boundaries, create*, num_cells)
//   sparse.rs:145-295   PrevPtr + sdpkpp;  sparse.rs:337-402 find_kmer_matches
//   bit_tree.rs:45-99   FenwickTree / MaxBitTree (incl. its "set at the last index is a no-op")
// Statement-by-statement: usize -> uint64_t, u32 arithmetic kept in uint32_t (wrapping like a
// release build), same loop order, same order of traceback writes (the eager get_mut writes that
// later set() calls overwrite matter), rolling arrays keep their leftovers exactly as there.
// Out-of-range indexing (a panic in the reference) is reported as an error, never masked.
//
// Parity is PINNED by tests/test_oracle_banded.py: the reference's own banded/sparse known-answer
// tests and doctests (banded.rs:29-90, 1470-1618, 1767-2416; sparse.rs:505-700; bit_tree.rs:122-141)
// and its differential tests banded == full (banded.rs:1621-1753) against the pinned full oracle.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <vector>

namespace {

constexpr int32_t MIN_SCORE = -858993459;
constexpr uint64_t MAX_CELLS = 5000000;      // banded.rs:104
constexpr int32_t DEFAULT_MATCH_SCORE = 2;   // banded.rs:105

constexpr uint16_t TB_START = 0, TB_INS = 1, TB_DEL = 2, TB_SUBST = 3, TB_MATCH = 4,
                   TB_XCLIP_PREFIX = 5, TB_XCLIP_SUFFIX = 6, TB_YCLIP_PREFIX = 7,
                   TB_YCLIP_SUFFIX = 8;

// A panic (or a traceback that never terminates) in the reference is recorded here instead of being
// thrown: C++ exceptions across this dlopen'ed, statically-linked-libstdc++ library are not reliable.
thread_local bool g_panic = false;
inline void panic(const char*) { g_panic = true; }

template <class T>
struct Vec {  // bounds-checked like a Rust Vec
  std::vector<T> v;
  mutable T dummy{};
  T& operator[](uint64_t i) {
    if (i >= v.size()) {
      panic("index out of bounds");
      return dummy;
    }
    return v[i];
  }
  const T& operator[](uint64_t i) const {
    if (i >= v.size()) {
      panic("index out of bounds");
      return dummy;
    }
    return v[i];
  }
  uint64_t len() const { return v.size(); }
};

struct Cell {
  uint16_t v = 0;
  void set_i(uint16_t x) { v = (uint16_t)((v & ~0x000F) | x); }
  void set_d(uint16_t x) { v = (uint16_t)((v & ~0x00F0) | (x << 4)); }
  void set_s(uint16_t x) { v = (uint16_t)((v & ~0x0F00) | (x << 8)); }
  void set_all(uint16_t x) { set_i(x); set_d(x); set_s(x); }
  uint16_t i() const { return v & 15; }
  uint16_t d() const { return (v >> 4) & 15; }
  uint16_t s() const { return (v >> 8) & 15; }
};

inline uint64_t sat_sub(uint64_t a, uint64_t b) { return a > b ? a - b : 0; }

}  // namespace

extern "C" {
struct orc_scoring {  // same layout as in pairwise_oracle.cpp / b2a_scoring
  int32_t gap_open, gap_extend;
  int32_t xclip_prefix, xclip_suffix, yclip_prefix, yclip_suffix;
  int32_t match_score, mismatch_score;
  int32_t has_match_scores;
  const int32_t* table;
  const uint8_t* alphabet;
  uint32_t alphabet_len;
};
struct orc_alignment {
  int32_t score;
  uint32_t ystart, xstart, yend, xend, ylen, xlen;
  uint32_t mode;
  uint32_t n_ops;
};
}

namespace {

using Match = std::pair<uint32_t, uint32_t>;

// ---------------------------------------------------------------- sparse::find_kmer_matches
// sparse.rs:337-402. Which side is hashed and the hash function cannot change the result: all (i,j)
// with x[i..i+k] == y[j..j+k], sorted ascending.
std::vector<Match> find_kmer_matches(const uint8_t* x, uint64_t m, const uint8_t* y, uint64_t n,
                                     uint64_t k) {
  std::vector<Match> out;
  std::unordered_map<std::string, std::vector<uint32_t>> set;
  // hash_kmers(seq2): for i in 0..(len+1).saturating_sub(k)   (sparse.rs:350-357)
  for (uint64_t i = 0; i < sat_sub(n + 1, k); ++i)
    set[std::string(reinterpret_cast<const char*>(y) + i, k)].push_back((uint32_t)i);
  for (uint64_t i = 0; i < sat_sub(m + 1, k); ++i) {
    auto it = set.find(std::string(reinterpret_cast<const char*>(x) + i, k));
    if (it != set.end())
      for (uint32_t pos : it->second) out.emplace_back((uint32_t)i, pos);
  }
  std::sort(out.begin(), out.end());
  return out;
}

// ---------------------------------------------------------------- bit_tree.rs:45-99
struct PrevPtr {  // sparse.rs:145-167, derived lexicographic Ord
  uint32_t plane = 0, score = 0, d = 0;
  uint64_t id = 0;
  uint32_t x = 0, y = 0;
  auto key() const { return std::make_tuple(plane, score, d, id, x, y); }
};
inline PrevPtr pmax(const PrevPtr& a, const PrevPtr& b) { return b.key() >= a.key() ? b : a; }

struct MaxBitTree {
  Vec<PrevPtr> tree;
  explicit MaxBitTree(uint64_t len) { tree.v.assign(len + 1, PrevPtr{}); }
  PrevPtr get(uint64_t idx) const {
    idx += 1;
    PrevPtr sum{};
    while (idx > 0) {
      sum = pmax(sum, tree[idx]);
      idx -= idx & (~idx + 1);
    }
    return sum;
  }
  void set(uint64_t idx, const PrevPtr& val) {
    idx += 1;
    while (idx < tree.len()) {
      tree[idx] = pmax(tree[idx], val);
      idx += idx & (~idx + 1);
    }
  }
};

// ---------------------------------------------------------------- sparse::sdpkpp, sparse.rs:188-295
struct SparseResult {
  std::vector<uint64_t> path;
  uint32_t score = 0;
};

SparseResult sdpkpp(const std::vector<Match>& matches, uint64_t k_, uint32_t match_score,
                    int32_t gap_open, int32_t gap_extend) {
  SparseResult res;
  if (matches.empty()) return res;
  const uint32_t k = (uint32_t)k_;
  if (!(gap_open <= 0 && gap_extend <= 0)) { panic("gap parameters cannot be positive"); return res; }
  const uint32_t go = (uint32_t)(-gap_open), ge = (uint32_t)(-gap_extend);
  for (size_t i = 1; i < matches.size(); ++i)
    if (!(matches[i - 1] < matches[i])) { panic("incoming matches must be sorted"); return res; }
  std::vector<std::tuple<uint32_t, uint32_t, uint32_t>> events;
  uint32_t n = 0;
  const uint32_t nm = (uint32_t)matches.size();
  for (uint32_t idx = 0; idx < nm; ++idx) {
    const uint32_t x = matches[idx].first, y = matches[idx].second;
    events.emplace_back(x, y, idx + nm);
    events.emplace_back(x + k, y + k, idx);
    n = std::max(n, x + k);
    n = std::max(n, y + k);
  }
  std::sort(events.begin(), events.end());
  MaxBitTree max_col_dp(n);
  Vec<std::pair<uint32_t, int32_t>> dp;
  dp.v.assign(events.size(), {0u, 0});
  std::pair<uint32_t, int32_t> best_dp{k, 0};
  for (const auto& ev : events) {
    const uint32_t e0 = std::get<0>(ev), e1 = std::get<1>(ev), e2 = std::get<2>(ev);
    const uint64_t p = e2 % nm;
    const bool is_start = e2 >= nm;
    if (is_start) {
      dp[p] = {k * match_score, -1};
      const PrevPtr best_prev = max_col_dp.get(e1);
      if (best_prev.score > 0) {
        const uint32_t gap = std::max(e0 - best_prev.x, e1 - best_prev.y);
        const uint32_t gap_penalty = gap > 0 ? go + gap * ge : 0;
        const uint32_t reward = k * match_score;
        const uint32_t sum = best_prev.score + reward;
        const uint32_t new_score = sum > gap_penalty ? sum - gap_penalty : 0;  // saturating_sub
        dp[p] = std::max(dp[p], std::make_pair(new_score, (int32_t)best_prev.id));
        best_dp = std::max(best_dp, std::make_pair(dp[p].first, (int32_t)p));
      }
    } else {
      if (e0 > k && e1 > k) {
        const Match want{e0 - k - 1, e1 - k - 1};
        auto it = std::lower_bound(matches.begin(), matches.end(), want);
        if (it != matches.end() && *it == want) {
          const uint64_t cont_idx = (uint64_t)(it - matches.begin());
          const auto cand = std::make_pair(dp[cont_idx].first + match_score, (int32_t)cont_idx);
          dp[p] = std::max(dp[p], cand);
          best_dp = std::max(best_dp, std::make_pair(dp[p].first, (int32_t)p));
        }
      }
      PrevPtr pf;
      pf.d = e0 + e1;
      pf.plane = dp[p].first + pf.d * ge;
      pf.score = dp[p].first;
      pf.id = p;
      pf.x = e0;
      pf.y = e1;
      max_col_dp.set(e1, pf);
    }
  }
  int32_t prev_match = best_dp.second;
  while (prev_match >= 0) {
    res.path.push_back((uint64_t)prev_match);
    prev_match = dp[(uint64_t)prev_match].second;
  }
  std::reverse(res.path.begin(), res.path.end());
  res.score = best_dp.first;
  return res;
}

// ---------------------------------------------------------------- sparse::lcskpp, sparse.rs:67-143
// MaxBitTree<(u32, u32)>: the same prefix-max tree over (score, position) tuples
struct MaxBitTreePair {
  Vec<std::pair<uint32_t, uint32_t>> tree;
  explicit MaxBitTreePair(uint64_t len) { tree.v.assign(len + 1, {0u, 0u}); }
  std::pair<uint32_t, uint32_t> get(uint64_t idx) const {
    idx += 1;
    std::pair<uint32_t, uint32_t> sum{0u, 0u};
    while (idx > 0) {
      sum = std::max(sum, tree[idx]);
      idx -= idx & (~idx + 1);
    }
    return sum;
  }
  void set(uint64_t idx, const std::pair<uint32_t, uint32_t>& val) {
    idx += 1;
    while (idx < tree.len()) {
      tree[idx] = std::max(tree[idx], val);
      idx += idx & (~idx + 1);
    }
  }
};

SparseResult lcskpp(const std::vector<Match>& matches, uint64_t k_) {
  SparseResult res;
  if (matches.empty()) return res;
  const uint32_t k = (uint32_t)k_;
  for (size_t i = 1; i < matches.size(); ++i)
    if (!(matches[i - 1] < matches[i])) { panic("incoming matches must be sorted."); return res; }
  std::vector<std::tuple<uint32_t, uint32_t, uint32_t>> events;
  uint32_t n = 0;
  const uint32_t nm = (uint32_t)matches.size();
  for (uint32_t idx = 0; idx < nm; ++idx) {
    const uint32_t x = matches[idx].first, y = matches[idx].second;
    events.emplace_back(x, y, idx + nm);
    events.emplace_back(x + k, y + k, idx);
    n = std::max(n, x + k);
    n = std::max(n, y + k);
  }
  std::sort(events.begin(), events.end());
  MaxBitTreePair max_col_dp(n);
  Vec<std::pair<uint32_t, int32_t>> dp;
  dp.v.assign(events.size(), {0u, 0});
  std::pair<uint32_t, int32_t> best_dp{k, 0};
  for (const auto& ev : events) {
    const uint32_t e0 = std::get<0>(ev), e1 = std::get<1>(ev), e2 = std::get<2>(ev);
    const uint64_t p = e2 % nm;
    const bool is_start = e2 >= nm;
    if (is_start) {
      dp[p] = {k, -1};
      const auto best = max_col_dp.get(e1);
      if (best.first > 0) {
        dp[p] = {k + best.first, (int32_t)best.second};
        best_dp = std::max(best_dp, std::make_pair(dp[p].first, (int32_t)p));
      }
    } else {
      if (e0 > k && e1 > k) {
        const Match want{e0 - k - 1, e1 - k - 1};
        auto it = std::lower_bound(matches.begin(), matches.end(), want);
        if (it != matches.end() && *it == want) {
          const uint64_t cont_idx = (uint64_t)(it - matches.begin());
          const auto cand = std::make_pair(dp[cont_idx].first + 1, (int32_t)cont_idx);
          dp[p] = std::max(dp[p], cand);
          best_dp = std::max(best_dp, std::make_pair(dp[p].first, (int32_t)p));
        }
      }
      max_col_dp.set(e1, {dp[p].first, (uint32_t)p});
    }
  }
  int32_t prev_match = best_dp.second;
  while (prev_match >= 0) {
    res.path.push_back((uint64_t)prev_match);
    prev_match = dp[(uint64_t)prev_match].second;
  }
  std::reverse(res.path.begin(), res.path.end());
  res.score = best_dp.first;
  return res;
}

// ---------------------------------------------------------------- sparse::sdpkpp_union_lcskpp_path, sparse.rs:297-330
std::vector<uint64_t> sdpkpp_union_lcskpp_path(const std::vector<Match>& matches, uint64_t k, uint32_t match_score,
                                               int32_t gap_open, int32_t gap_extend) {
  std::vector<uint64_t> out;
  if (matches.empty()) return out;
  const SparseResult l = lcskpp(matches, k);
  const SparseResult s = sdpkpp(matches, k, match_score, gap_open, gap_extend);
  if (g_panic || s.path.empty()) { panic("empty sdpkpp path"); return out; }
  // slice::binary_search on the (ascending) lcskpp path; any of several equal elements may be returned, but
  // path indices are distinct
  auto bsearch = [&](uint64_t v, bool& found) -> uint64_t {
    auto it = std::lower_bound(l.path.begin(), l.path.end(), v);
    found = it != l.path.end() && *it == v;
    return (uint64_t)(it - l.path.begin());
  };
  bool f0 = false, f1 = false;
  const uint64_t i0 = bsearch(s.path.front(), f0);
  const uint64_t pre = f0 ? i0 : 0;                  // .unwrap_or(0)
  const uint64_t i1 = bsearch(s.path.back(), f1);
  const uint64_t post = f1 ? i1 + 1 : l.path.size();
  for (uint64_t i = 0; i < pre; ++i) out.push_back(l.path[i]);
  for (uint64_t v : s.path) out.push_back(v);
  for (uint64_t i = post; i < l.path.size(); ++i) out.push_back(l.path[i]);
  return out;
}

// ---------------------------------------------------------------- sparse::expand_kmer_matches, sparse.rs:404-498
std::vector<Match> expand_kmer_matches(const uint8_t* seq1, uint64_t len1, const uint8_t* seq2, uint64_t len2,
                                       uint64_t k, const std::vector<Match>& sorted_matches,
                                       uint64_t allowed_mismatches) {
  std::vector<Match> none;
  for (size_t i = 1; i < sorted_matches.size(); ++i)
    if (!(sorted_matches[i - 1] < sorted_matches[i])) { panic("incoming matches must be sorted"); return none; }
  using P = std::pair<int32_t, int32_t>;
  std::unordered_map<int32_t, P> last_match_along_diagonal;
  std::vector<Match> left = sorted_matches;
  for (const Match& tm : sorted_matches) {
    const int32_t diag = (int32_t)tm.first - (int32_t)tm.second;
    const int32_t min_xy = (int32_t)std::min(tm.first, tm.second);
    const P def{(int32_t)tm.first - min_xy - 1, (int32_t)tm.second - min_xy - 1};
    auto it = last_match_along_diagonal.find(diag);
    const P last_match = it != last_match_along_diagonal.end() ? it->second : def;
    uint64_t n_mismatches = 0;
    P curr{(int32_t)tm.first - 1, (int32_t)tm.second - 1};
    for (;;) {
      if (last_match >= curr) break;
      if ((uint64_t)curr.first >= len1 || (uint64_t)curr.second >= len2 || curr.first < 0 || curr.second < 0) {
        panic("index out of bounds");
        return none;
      }
      n_mismatches += seq1[curr.first] == seq2[curr.second] ? 0 : 1;
      if (n_mismatches > allowed_mismatches) break;
      left.emplace_back((uint32_t)curr.first, (uint32_t)curr.second);
      curr = {curr.first - 1, curr.second - 1};
    }
    last_match_along_diagonal[diag] = {(int32_t)tm.first, (int32_t)tm.second};
  }
  std::sort(left.begin(), left.end());
  std::vector<Match> expanded = left;
  std::reverse(left.begin(), left.end());
  std::unordered_map<int32_t, Match> next_match_along_diagonal;
  for (const Match& tm : left) {
    const int32_t diag = (int32_t)tm.first - (int32_t)tm.second;
    // (min(len1 - x, len2 - y) as u32).saturating_sub(k as u32 - 1); `k as u32 - 1` underflows for k == 0
    if (k == 0) { panic("attempt to subtract with overflow"); return none; }
    const uint32_t a = (uint32_t)len1 - tm.first, b = (uint32_t)len2 - tm.second;
    const uint32_t mn = std::min(a, b), km1 = (uint32_t)k - 1;
    const uint32_t max_inc = mn > km1 ? mn - km1 : 0;
    auto it = next_match_along_diagonal.find(diag);
    const Match next_match = it != next_match_along_diagonal.end() ? it->second
                                                                   : Match{tm.first + max_inc, tm.second + max_inc};
    uint64_t n_mismatches = 0;
    Match curr{tm.first + 1, tm.second + 1};
    for (;;) {
      if (curr >= next_match) break;
      const uint64_t i1 = (uint64_t)curr.first + k - 1, i2 = (uint64_t)curr.second + k - 1;
      if (i1 >= len1 || i2 >= len2) { panic("index out of bounds"); return none; }
      n_mismatches += seq1[i1] == seq2[i2] ? 0 : 1;
      if (n_mismatches > allowed_mismatches) break;
      expanded.push_back(curr);
      curr = {curr.first + 1, curr.second + 1};
    }
    next_match_along_diagonal[diag] = tm;
  }
  std::sort(expanded.begin(), expanded.end());
  return expanded;
}

// ---------------------------------------------------------------- Band, banded.rs:1047-1380
struct Range {
  uint64_t start, end;
};

struct Band {
  uint64_t rows = 0, cols = 0;
  Vec<Range> ranges;

  static Band make(uint64_t m, uint64_t n) {  // Band::new, 1061-1067
    Band b;
    b.rows = m + 1;
    b.cols = n + 1;
    b.ranges.v.assign(n + 1, Range{m + 1, 0});
    return b;
  }
  void add_kmer(Match start, uint64_t k, uint64_t w) {  // 1071-1107
    const uint64_t r = start.first, c = start.second;
    if (k == 0) return;
    {
      const uint64_t i = sat_sub(r, w);
      for (uint64_t j = sat_sub(c, w); j < std::min(c + w + 1, cols); ++j)
        ranges[j].start = std::min(ranges[j].start, i);
    }
    {
      uint64_t i = sat_sub(r, w);
      for (uint64_t j = std::min(c + w, cols); j < std::min(c + k + w, cols); ++j) {
        ranges[j].start = std::min(ranges[j].start, i);
        i += 1;
      }
    }
    {
      uint64_t i = r + w + k;
      uint64_t j = sat_sub(c + k - 1, w);
      for (;;) {
        if (j <= sat_sub(c, w)) break;
        j -= 1;
        i -= 1;
        ranges[j].end = std::max(ranges[j].end, std::min(i, rows));
      }
    }
    {
      const uint64_t i = std::min(r + w + k, rows);
      for (uint64_t j = sat_sub(c + k - 1, w); j < std::min(c + k + w, cols); ++j)
        ranges[j].end = std::max(ranges[j].end, i);
    }
  }
  void add_entry(Match pos, uint64_t w) {  // 1111-1120
    const uint64_t r = pos.first, c = pos.second;
    const uint64_t istart = sat_sub(r, w), iend = std::min(r + w + 1, rows);
    for (uint64_t j = sat_sub(c, w); j < std::min(c + w + 1, cols); ++j) {
      ranges[j].start = std::min(ranges[j].start, istart);
      ranges[j].end = std::max(ranges[j].end, iend);
    }
  }
  void add_gap(Match start, Match end, uint64_t w) {  // 1123-1137 (u32 arithmetic)
    const uint32_t nrows = end.first - start.first;
    const uint32_t ncols = end.second - start.second;
    if (nrows > ncols) {
      for (uint32_t r = start.first; r < end.first; ++r) {
        const uint32_t den = end.first - start.first;
        if (den == 0) { panic("attempt to divide by zero"); return; }
        const uint32_t c = start.second + (end.second - start.second) * (r - start.first) / den;
        add_entry({r, c}, w);
      }
    } else {
      for (uint32_t c = start.second; c < end.second; ++c) {
        const uint32_t den = end.second - start.second;
        if (den == 0) { panic("attempt to divide by zero"); return; }
        const uint32_t r = start.first + (end.first - start.first) * (c - start.second) / den;
        add_entry({r, c}, w);
      }
    }
  }
  void set_boundaries(Match start, Match end, uint64_t k, uint64_t w, const orc_scoring& sc) {  // 1150-1276
    const uint64_t lazy_extend = 2 * k;
    {
      const uint64_t r = start.first, c = start.second;
      if (!(r == 0 && c == 0)) {
        int32_t score_to_start = r > 0 ? sc.xclip_prefix : 0;
        score_to_start += c > 0 ? sc.yclip_prefix : 0;
        if (score_to_start == 0) {
          const uint64_t d = std::min(lazy_extend, std::min(r, c));
          add_kmer({(uint32_t)(r - d), (uint32_t)(c - d)}, d, w);
          add_gap({(uint32_t)sat_sub(r, lazy_extend), (uint32_t)sat_sub(c, lazy_extend)},
                  {(uint32_t)(r - d), (uint32_t)(c - d)}, w);
        } else {
          const int32_t diagonal_score = r > c ? sc.xclip_prefix : (r < c ? sc.yclip_prefix : 0);
          if (diagonal_score == 0) {
            const uint64_t d = std::min(r, c);
            add_kmer({(uint32_t)(r - d), (uint32_t)(c - d)}, d, w);
            const Match s2{(uint32_t)sat_sub(r, lazy_extend), (uint32_t)sat_sub(c, lazy_extend)};
            const Match e2{(uint32_t)(r - d), (uint32_t)(c - d)};
            if (s2.first <= e2.first && s2.second <= e2.second) add_gap(s2, e2, w);
          } else {
            add_gap({0u, 0u}, start, w);
          }
        }
      }
    }
    {
      const uint64_t r = (uint64_t)end.first + k, c = (uint64_t)end.second + k;
      if (!(r == rows && c == cols)) {
        int32_t score_from_end = r == rows ? 0 : sc.xclip_suffix;
        score_from_end += c == cols ? 0 : sc.yclip_suffix;
        if (score_from_end == 0) {
          const uint64_t d = std::min(lazy_extend, std::min(rows - r, cols - c));
          add_kmer({(uint32_t)r, (uint32_t)c}, d, w);
          const uint64_t r1 = std::min(rows, r + d) - 1, c1 = std::min(cols, c + d) - 1;
          const uint64_t r2 = std::min(rows, r + lazy_extend), c2 = std::min(cols, c + lazy_extend);
          if (r1 <= r2 && c1 <= c2) add_gap({(uint32_t)r1, (uint32_t)c1}, {(uint32_t)r2, (uint32_t)c2}, w);
        } else {
          const uint64_t dr = rows - r, dc = cols - c;
          const int32_t diagonal_score = dr > dc ? sc.xclip_suffix : (dr < dc ? sc.yclip_suffix : 0);
          if (diagonal_score == 0) {
            const uint64_t d = std::min(dr, dc);
            add_kmer({(uint32_t)r, (uint32_t)c}, d, w);
            const uint64_t r1 = std::min(rows, r + d) - 1, c1 = std::min(cols, c + d) - 1;
            const uint64_t r2 = std::min(rows, r + lazy_extend), c2 = std::min(cols, c + lazy_extend);
            if (r1 <= r2 && c1 <= c2) add_gap({(uint32_t)r1, (uint32_t)c1}, {(uint32_t)r2, (uint32_t)c2}, w);
          } else {
            add_gap({(uint32_t)r, (uint32_t)c}, {(uint32_t)rows, (uint32_t)cols}, w);
          }
        }
      }
    }
  }
  void full_matrix() {  // 1369-1372
    ranges.v.assign(cols, Range{0, rows});
  }
  uint64_t num_cells() const {  // 1374-1380
    uint64_t cells = 0;
    for (uint64_t j = 0; j < ranges.len(); ++j) cells += sat_sub(ranges[j].end, ranges[j].start);
    return cells;
  }
  // create_from_match_path, 1330-1367
  static Band from_match_path(uint64_t m, uint64_t n, uint64_t k, uint64_t w, const orc_scoring& sc,
                              const std::vector<uint64_t>& path, const std::vector<Match>& matches) {
    Band band = Band::make(m, n);
    if (matches.empty()) {
      band.full_matrix();
      return band;
    }
    if (path.empty()) { panic("index out of bounds"); band.full_matrix(); return band; }
    const uint64_t ps = path.front(), pe = path.back();
    band.set_boundaries(matches[ps], matches[pe], k, w, sc);
    bool has_prev = false;
    Match prev{0, 0};
    for (uint64_t idx : path) {
      const Match curr = matches[idx];
      const bool cont = has_prev && curr.first == prev.first + 1 && curr.second == prev.second + 1;
      if (cont) {
        band.add_entry({prev.first + (uint32_t)k, prev.second + (uint32_t)k}, w);
      } else {
        if (has_prev)
          band.add_gap({prev.first + (uint32_t)(k - 1), prev.second + (uint32_t)(k - 1)}, curr, w);
        band.add_kmer(curr, k, w);
      }
      prev = curr;
      has_prev = true;
    }
    return band;
  }
  // create_with_matches, 1301-1328
  static Band with_matches(uint64_t m, uint64_t n, uint64_t k, uint64_t w, const orc_scoring& sc,
                           const std::vector<Match>& matches) {
    if (matches.empty()) {
      Band band = Band::make(m, n);
      band.full_matrix();
      return band;
    }
    const int32_t match_score = sc.has_match_scores ? sc.match_score : DEFAULT_MATCH_SCORE;
    const SparseResult res = sdpkpp(matches, k, (uint32_t)match_score, sc.gap_open, sc.gap_extend);
    return from_match_path(m, n, k, w, sc, res.path, matches);
  }
};

struct Op {
  uint32_t code, len;
};

// ---------------------------------------------------------------- banded::Aligner, banded.rs:122-1004
struct BandedAligner {
  Vec<int32_t> S[2], I[2], D[2];
  Vec<uint64_t> Lx, Ly;
  Vec<int32_t> Sn;
  Vec<Cell> tb;
  uint64_t tb_cols = 0;
  orc_scoring sc{};
  Band band;
  uint64_t k = 0, w = 0;

  int32_t score(uint8_t a, uint8_t b) const {
    if (sc.table) return sc.table[(size_t)a * 256 + b];
    return a == b ? sc.match_score : sc.mismatch_score;
  }
  Cell& at(uint64_t i, uint64_t j) { return tb[i * tb_cols + j]; }

  // compute_alignment, banded.rs:406-869
  void compute_alignment(const uint8_t* x, uint64_t m, const uint8_t* y, uint64_t n,
                         orc_alignment* out, std::vector<Op>& operations) {
    operations.clear();
    if (band.num_cells() > MAX_CELLS) {  // 407-420
      out->score = MIN_SCORE;
      out->ystart = out->xstart = out->yend = out->xend = out->ylen = out->xlen = 0;
      out->mode = 0;
      out->n_ops = 0;
      return;
    }
    tb_cols = n + 1;
    tb.v.assign((m + 1) * (n + 1), Cell{});  // traceback.init: the FULL matrix, banded.rs:423
    for (int kk = 0; kk < 2; ++kk) {
      I[kk].v.assign(m + 1, MIN_SCORE);
      D[kk].v.assign(m + 1, MIN_SCORE);
      S[kk].v.assign(m + 1, MIN_SCORE);
    }
    Lx.v.assign(n + 1, 0);
    Ly.v.assign(m + 1, 0);
    Sn.v.assign(m + 1, MIN_SCORE);
    {  // j = 0, 440-509
      const int curr = 0;
      const uint64_t i_start = band.ranges[0].start, i_end = band.ranges[0].end;
      if (i_start == 0) S[curr][0] = 0;
      for (uint64_t i = std::max<uint64_t>(1, i_start); i < i_end; ++i) {
        Cell c;
        c.set_all(TB_START);
        if (i == 1) {
          I[curr][i] = sc.gap_open;
          c.set_i(TB_START);
        } else {
          const int32_t i_score = sc.gap_open + sc.gap_extend * ((int32_t)i - 1);
          const int32_t c_score = sc.xclip_prefix + sc.gap_open;
          if (i_score > c_score) {
            I[curr][i] = i_score;
            c.set_i(TB_INS);
          } else {
            I[curr][i] = c_score;
            c.set_i(TB_XCLIP_PREFIX);
          }
        }
        if (i == m) c.set_s(TB_XCLIP_SUFFIX);
        if (I[curr][i] > S[curr][i]) {
          S[curr][i] = I[curr][i];
          c.set_s(TB_INS);
        }
        if (sc.xclip_prefix > S[curr][i]) {
          S[curr][i] = sc.xclip_prefix;
          c.set_s(TB_XCLIP_PREFIX);
        }
        if (S[curr][i] + sc.xclip_suffix > S[curr][m]) {
          S[curr][m] = S[curr][i] + sc.xclip_suffix;
          Lx[0] = m - i;
          at(m, 0).set_s(TB_XCLIP_SUFFIX);
        }
        at(i, 0) = c;
      }
      for (uint64_t i = i_end; i < std::min(m + 1, band.ranges[std::min<uint64_t>(n, 1)].end); ++i) {
        S[curr][i] = MIN_SCORE;
        I[curr][i] = MIN_SCORE;
      }
      if (i_end < m + 1) S[curr][m] = MIN_SCORE;
      if (sc.yclip_prefix > sc.yclip_suffix) {
        Sn[0] = sc.yclip_prefix;
        at(0, n).set_s(TB_YCLIP_PREFIX);
      } else {
        Sn[0] = sc.yclip_suffix;
        Ly[0] = n;
        at(0, n).set_s(TB_YCLIP_SUFFIX);
      }
    }
    for (uint64_t j = 1; j <= n; ++j) {  // 511-681
      const int curr = (int)(j % 2), prev = 1 - curr;
      const uint64_t i_start = band.ranges[j].start, i_end = band.ranges[j].end;
      if (i_start == 0) {
        Cell c;
        I[curr][0] = MIN_SCORE;
        if (j == 1) {
          D[curr][0] = sc.gap_open;
          c.set_d(TB_START);
        } else {
          const int32_t d_score = sc.gap_open + sc.gap_extend * ((int32_t)j - 1);
          const int32_t c_score = sc.yclip_prefix + sc.gap_open;
          if (d_score > c_score) {
            D[curr][0] = d_score;
            c.set_d(TB_DEL);
          } else {
            D[curr][0] = c_score;
            c.set_d(TB_YCLIP_PREFIX);
          }
        }
        if (D[curr][0] > sc.yclip_prefix) {
          S[curr][0] = D[curr][0];
          c.set_s(TB_DEL);
        } else {
          S[curr][0] = sc.yclip_prefix;
          c.set_s(TB_YCLIP_PREFIX);
        }
        if (S[curr][0] + sc.yclip_suffix > Sn[0]) {
          Sn[0] = S[curr][0] + sc.yclip_suffix;
          Ly[0] = n - j;
          at(0, n).set_s(TB_YCLIP_SUFFIX);
        }
        at(0, j) = c;
      }
      for (uint64_t i = sat_sub(i_start, 1); i < i_start; ++i) {
        S[curr][i] = MIN_SCORE;
        I[curr][i] = MIN_SCORE;
        D[curr][i] = MIN_SCORE;
      }
      S[curr][m] = MIN_SCORE;
      const uint8_t q = y[j - 1];
      const int32_t xclip_score =
          sc.xclip_prefix + std::max(j == n ? std::max(sc.yclip_prefix, Sn[0]) : sc.yclip_prefix,
                                     sc.gap_open + sc.gap_extend * ((int32_t)j - 1));
      for (uint64_t i = std::max<uint64_t>(1, i_start); i < i_end; ++i) {
        const uint8_t p = x[i - 1];
        Cell c;
        const int32_t m_score = S[prev][i - 1] + score(p, q);
        const int32_t i_score = I[curr][i - 1] + sc.gap_extend;
        int32_t s_score = S[curr][i - 1] + sc.gap_open;
        int32_t best_i_score;
        if (i_score > s_score) {
          best_i_score = i_score;
          c.set_i(TB_INS);
        } else {
          best_i_score = s_score;
          c.set_i(at(i - 1, j).s());
        }
        if (j == n) {
          const int32_t clip_score = Sn[i - 1] + sc.gap_open;
          if (clip_score > best_i_score) {
            best_i_score = clip_score;
            c.set_i(TB_YCLIP_SUFFIX);
          }
        }
        const int32_t d_score = D[prev][i] + sc.gap_extend;
        s_score = S[prev][i] + sc.gap_open;
        int32_t best_d_score;
        if (d_score > s_score) {
          best_d_score = d_score;
          c.set_d(TB_DEL);
        } else {
          best_d_score = s_score;
          c.set_d(at(i, j - 1).s());
        }
        if (i == m) {
          c.set_s(TB_XCLIP_SUFFIX);
        } else {
          S[curr][i] = MIN_SCORE;
        }
        int32_t best_s_score = S[curr][i];
        if (m_score > best_s_score) {
          best_s_score = m_score;
          c.set_s(p == q ? TB_MATCH : TB_SUBST);
        }
        if (best_i_score > best_s_score) {
          best_s_score = best_i_score;
          c.set_s(TB_INS);
        }
        if (best_d_score > best_s_score) {
          best_s_score = best_d_score;
          c.set_s(TB_DEL);
        }
        if (xclip_score > best_s_score) {
          best_s_score = xclip_score;
          c.set_s(TB_XCLIP_PREFIX);
        }
        const int32_t yclip_score = sc.yclip_prefix + sc.gap_open + sc.gap_extend * ((int32_t)i - 1);
        if (yclip_score > best_s_score) {
          best_s_score = yclip_score;
          c.set_s(TB_YCLIP_PREFIX);
        }
        S[curr][i] = best_s_score;
        I[curr][i] = best_i_score;
        D[curr][i] = best_d_score;
        if (S[curr][i] + sc.xclip_suffix > S[curr][m]) {
          S[curr][m] = S[curr][i] + sc.xclip_suffix;
          Lx[j] = m - i;
          at(m, j).set_s(TB_XCLIP_SUFFIX);
        }
        if (S[curr][i] + sc.yclip_suffix > Sn[i]) {
          Sn[i] = S[curr][i] + sc.yclip_suffix;
          Ly[i] = n - j;
          at(i, n).set_s(TB_YCLIP_SUFFIX);
        }
        at(i, j) = c;
      }
      if (S[curr][m] + sc.yclip_suffix > Sn[m]) {
        Sn[m] = S[curr][m] + sc.yclip_suffix;
        Ly[m] = n - j;
        at(m, n).set_s(TB_YCLIP_SUFFIX);
      }
      if (i_end < m + 1) {
        at(m, j).set_s(TB_XCLIP_SUFFIX);
        S[curr][m] = MIN_SCORE;
      }
      for (uint64_t i = i_end; i < std::min(m + 1, band.ranges[std::min(n, j + 1)].end); ++i) {
        S[curr][i] = MIN_SCORE;
        I[curr][i] = MIN_SCORE;
        D[curr][i] = MIN_SCORE;
      }
    }
    for (uint64_t i = 0; i <= m; ++i) {  // 684-701
      const uint64_t j = n;
      const int curr = (int)(j % 2);
      if (i != m && (i < band.ranges[j].start || i > band.ranges[j].end)) S[curr][i] = MIN_SCORE;
      if (Sn[i] > S[curr][i]) {
        S[curr][i] = Sn[i];
        at(i, j).set_s(TB_YCLIP_SUFFIX);
      }
      if (S[curr][i] + sc.xclip_suffix > S[curr][m]) {
        S[curr][m] = S[curr][i] + sc.xclip_suffix;
        Lx[j] = m - i;
        at(m, j).set_s(TB_XCLIP_SUFFIX);
      }
    }
    for (uint64_t i = std::max<uint64_t>(1, band.ranges[n].start); i < band.ranges[n].end; ++i) {  // 705-723
      const uint64_t j = n;
      const int curr = (int)(j % 2);
      const int32_t s_score = S[curr][i - 1] + sc.gap_open;
      if (s_score > I[curr][i]) {
        I[curr][i] = s_score;
        const uint16_t s_bit = at(i - 1, j).s();
        at(i, j).set_i(s_bit);
      }
      if (s_score > S[curr][i]) {
        S[curr][i] = s_score;
        at(i, j).set_s(TB_INS);
        if (S[curr][i] + sc.xclip_suffix > S[curr][m]) {
          S[curr][m] = S[curr][i] + sc.xclip_suffix;
          Lx[j] = m - i;
          at(m, j).set_s(TB_XCLIP_SUFFIX);
        }
      }
    }
    for (uint64_t j = 1; j <= n; ++j) {  // 725-744
      const int32_t d_score = sc.gap_open + sc.gap_extend * ((int32_t)j - 1);
      if (d_score > sc.yclip_prefix) {
        at(0, j).set_s(TB_DEL);
      } else {
        at(0, j).set_s(TB_YCLIP_PREFIX);
      }
      if (j == n) {
        int32_t best_score = std::max(d_score, sc.yclip_prefix);
        if (sc.yclip_suffix > best_score) {
          best_score = sc.yclip_suffix;
          at(0, j).set_s(TB_YCLIP_SUFFIX);
        }
        if (sc.xclip_suffix + best_score > S[n % 2][m]) {
          S[n % 2][m] = sc.xclip_suffix + best_score;
          Lx[n] = m;
          at(m, n).set_s(TB_XCLIP_SUFFIX);
        }
      }
    }
    for (uint64_t i = 1; i <= m; ++i) {  // 746-765
      const int32_t c_score = sc.gap_open + sc.gap_extend * ((int32_t)i - 1);
      if (c_score > sc.xclip_prefix) {
        at(i, 0).set_s(TB_INS);
      } else {
        at(i, 0).set_s(TB_XCLIP_PREFIX);
      }
      if (i == m) {
        int32_t best_score = std::max(c_score, sc.xclip_prefix);
        if (sc.xclip_suffix > best_score) {
          best_score = sc.xclip_suffix;
          at(i, 0).set_s(TB_XCLIP_SUFFIX);
        }
        if (sc.yclip_suffix + best_score > S[n % 2][m]) {
          S[n % 2][m] = sc.yclip_suffix + best_score;
          Ly[m] = n;
          at(m, n).set_s(TB_YCLIP_SUFFIX);
        }
      }
    }
    // walk, 767-831
    uint64_t i = m, j = n;
    uint64_t xstart = 0, ystart = 0, xend = m, yend = n;
    uint16_t last_layer = at(i, j).s();
    uint64_t guard = 4 * (m + n) + 64;
    for (;;) {
      uint16_t next_layer;
      if (last_layer == TB_START) break;
      if (guard-- == 0) { panic("traceback does not terminate"); break; }
      switch (last_layer) {
        case TB_INS:
          operations.push_back({3, 0});
          next_layer = at(i, j).i();
          if (i == 0) { panic("attempt to subtract with overflow"); last_layer = TB_START; continue; }
          i -= 1;
          break;
        case TB_DEL:
          operations.push_back({2, 0});
          next_layer = at(i, j).d();
          if (j == 0) { panic("attempt to subtract with overflow"); last_layer = TB_START; continue; }
          j -= 1;
          break;
        case TB_MATCH:
        case TB_SUBST:
          operations.push_back({last_layer == TB_MATCH ? 0u : 1u, 0});
          if (i == 0 || j == 0) { panic("attempt to subtract with overflow"); last_layer = TB_START; continue; }
          next_layer = at(i - 1, j - 1).s();
          i -= 1;
          j -= 1;
          break;
        case TB_XCLIP_PREFIX:
          operations.push_back({4, (uint32_t)i});
          xstart = i;
          i = 0;
          next_layer = at(0, j).s();
          break;
        case TB_XCLIP_SUFFIX:
          operations.push_back({4, (uint32_t)Lx[j]});
          if (Lx[j] > i) { panic("attempt to subtract with overflow"); last_layer = TB_START; continue; }
          i -= Lx[j];
          xend = i;
          next_layer = at(i, j).s();
          break;
        case TB_YCLIP_PREFIX:
          operations.push_back({5, (uint32_t)j});
          ystart = j;
          j = 0;
          next_layer = at(i, 0).s();
          break;
        case TB_YCLIP_SUFFIX:
          operations.push_back({5, (uint32_t)Ly[i]});
          if (Ly[i] > j) { panic("attempt to subtract with overflow"); last_layer = TB_START; continue; }
          j -= Ly[i];
          yend = j;
          next_layer = at(i, j).s();
          break;
        default:
          panic("Dint expect this!");
          last_layer = TB_START;
          continue;
      }
      last_layer = next_layer;
    }
    if (i != 0) {  // 834-844
      const int32_t i_score = sc.gap_open + sc.gap_extend * ((int32_t)i - 1);
      if (i_score > sc.xclip_prefix) {
        for (uint64_t t = 0; t < i; ++t) operations.push_back({3, 0});
        xstart = 0;
      } else {
        operations.push_back({4, (uint32_t)i});
        xstart = i;
      }
    }
    if (j != 0) {  // 845-855
      const int32_t d_score = sc.gap_open + sc.gap_extend * ((int32_t)j - 1);
      if (d_score > sc.yclip_prefix) {
        for (uint64_t t = 0; t < j; ++t) operations.push_back({2, 0});
        ystart = 0;
      } else {
        operations.push_back({5, (uint32_t)j});
        ystart = j;
      }
    }
    std::reverse(operations.begin(), operations.end());
    out->score = S[n % 2][m];
    out->ystart = (uint32_t)ystart;
    out->xstart = (uint32_t)xstart;
    out->yend = (uint32_t)yend;
    out->xend = (uint32_t)xend;
    out->ylen = (uint32_t)n;
    out->xlen = (uint32_t)m;
    out->mode = 0;
    out->n_ops = (uint32_t)operations.size();
  }

  // custom + presets, banded.rs:282-285, 872-1004
  void align(int mode, const uint8_t* x, uint64_t m, const uint8_t* y, uint64_t n, orc_alignment* out,
             std::vector<Op>& ops) {
    const int32_t saved[4] = {sc.xclip_prefix, sc.xclip_suffix, sc.yclip_prefix, sc.yclip_suffix};
    if (mode == 1) sc.xclip_prefix = sc.xclip_suffix = sc.yclip_prefix = sc.yclip_suffix = MIN_SCORE;
    if (mode == 2) {
      sc.xclip_prefix = sc.xclip_suffix = MIN_SCORE;
      sc.yclip_prefix = sc.yclip_suffix = 0;
    }
    if (mode == 3) sc.xclip_prefix = sc.xclip_suffix = sc.yclip_prefix = sc.yclip_suffix = 0;
    // Band::create, 1278-1287
    const std::vector<Match> matches = find_kmer_matches(x, m, y, n, k);
    band = Band::with_matches(m, n, k, w, sc, matches);
    compute_alignment(x, m, y, n, out, ops);
    out->mode = (uint32_t)mode;
    if (mode == 2 || mode == 3) {
      ops.erase(std::remove_if(ops.begin(), ops.end(), [](const Op& o) { return o.code >= 4; }), ops.end());
      out->n_ops = (uint32_t)ops.size();
    }
    sc.xclip_prefix = saved[0];
    sc.xclip_suffix = saved[1];
    sc.yclip_prefix = saved[2];
    sc.yclip_suffix = saved[3];
  }

  // The entry points that take the band's inputs from the caller, banded.rs:294-401 (+ semiglobal_with_prehash
  // 938-975, whose band is that of `semiglobal`: find_kmer_matches_seq2_hashed returns the same sorted matches).
  //   matches (have_matches)            custom_with_matches            -> Band::create_with_matches
  //   + allowed_mismatches / union      custom_with_expanded_matches   -> expand_kmer_matches, union path
  //   + path (have_path)                custom_with_match_path         -> Band::create_from_match_path
  // These are `custom` methods: the clip penalties are the scoring's own and clips stay in the operations.
  void align_hinted(const uint8_t* x, uint64_t m, const uint8_t* y, uint64_t n, bool have_matches,
                    const std::vector<Match>& matches_in, bool have_path, const std::vector<uint64_t>& path_in,
                    int64_t allowed_mismatches, bool use_lcskpp_union, orc_alignment* out, std::vector<Op>& ops) {
    if (!have_matches) {
      align(0, x, m, y, n, out, ops);
      return;
    }
    if (have_path) {  // custom_with_match_path, 391-401
      if (!matches_in.empty()) {
        if (path_in.empty()) { panic("index out of bounds: path[0]"); return; }
        for (uint64_t idx : path_in)
          if (idx >= matches_in.size()) { panic("index out of bounds: matches[idx]"); return; }
      }
      band = Band::from_match_path(m, n, k, w, sc, path_in, matches_in);
    } else if (allowed_mismatches >= 0 || use_lcskpp_union) {  // custom_with_expanded_matches, 338-375
      std::vector<Match> expanded = allowed_mismatches >= 0
                                        ? expand_kmer_matches(x, m, y, n, k, matches_in, (uint64_t)allowed_mismatches)
                                        : matches_in;
      if (g_panic) return;
      if (use_lcskpp_union) {
        const int32_t match_score = sc.has_match_scores ? sc.match_score : DEFAULT_MATCH_SCORE;
        const std::vector<uint64_t> path =
            sdpkpp_union_lcskpp_path(expanded, k, (uint32_t)match_score, sc.gap_open, sc.gap_extend);
        if (g_panic) return;
        band = Band::from_match_path(m, n, k, w, sc, path, expanded);
      } else {
        band = Band::with_matches(m, n, k, w, sc, expanded);
      }
    } else {  // custom_with_matches, 313-321
      band = Band::with_matches(m, n, k, w, sc, matches_in);
    }
    if (g_panic) return;
    compute_alignment(x, m, y, n, out, ops);
    out->mode = 0;
  }
};

}  // namespace

extern "C" {

// returns the number of matches (may exceed cap; only the first cap are written)
uint64_t orc_find_kmer_matches(const uint8_t* x, uint32_t m, const uint8_t* y, uint32_t n, uint32_t k,
                               uint32_t* out_xy, uint64_t cap) {
  const auto v = find_kmer_matches(x, m, y, n, k);
  for (uint64_t i = 0; i < v.size() && i < cap; ++i) {
    out_xy[2 * i] = v[i].first;
    out_xy[2 * i + 1] = v[i].second;
  }
  return v.size();
}

// sdpkpp over sorted matches; path indices written to out_path (cap >= n_matches). Returns -1 on a panic path.
int orc_sdpkpp(const uint32_t* xy, uint64_t n_matches, uint32_t k, uint32_t match_score, int32_t gap_open,
               int32_t gap_extend, uint64_t* out_path, uint64_t* n_path, uint32_t* score) {
  {
    g_panic = false;
    std::vector<Match> m(n_matches);
    for (uint64_t i = 0; i < n_matches; ++i) m[i] = {xy[2 * i], xy[2 * i + 1]};
    const SparseResult r = sdpkpp(m, k, match_score, gap_open, gap_extend);
    for (uint64_t i = 0; i < r.path.size(); ++i) out_path[i] = r.path[i];
    *n_path = r.path.size();
    *score = r.score;
    return g_panic ? -1 : 0;
  }
}

// Band::create(x, y, k, w, scoring) with the clip presets of `mode` applied; ranges as (start,end) pairs.
int orc_band_create(int mode, const orc_scoring* scoring, uint32_t k, uint32_t w, const uint8_t* x, uint32_t m,
                    const uint8_t* y, uint32_t n, uint64_t* ranges /*2*(n+1)*/, uint64_t* num_cells) {
  {
    g_panic = false;
    orc_scoring sc = *scoring;
    if (mode == 1) sc.xclip_prefix = sc.xclip_suffix = sc.yclip_prefix = sc.yclip_suffix = MIN_SCORE;
    if (mode == 2) {
      sc.xclip_prefix = sc.xclip_suffix = MIN_SCORE;
      sc.yclip_prefix = sc.yclip_suffix = 0;
    }
    if (mode == 3) sc.xclip_prefix = sc.xclip_suffix = sc.yclip_prefix = sc.yclip_suffix = 0;
    const auto matches = find_kmer_matches(x, m, y, n, k);
    const Band b = Band::with_matches(m, n, k, w, sc, matches);
    for (uint64_t j = 0; j <= n; ++j) {
      ranges[2 * j] = b.ranges[j].start;
      ranges[2 * j + 1] = b.ranges[j].end;
    }
    *num_cells = b.num_cells();
    return g_panic ? -1 : 0;
  }
}

// Band geometry primitives for the reference's unit tests (banded.rs:1470-1618): ops is a list of
// (kind, r, c, k, w): kind 0 add_entry, 1 add_kmer.
int orc_band_ops(uint32_t m, uint32_t n, const uint32_t* ops, uint32_t n_ops, uint64_t* ranges) {
  {
    g_panic = false;
    Band b = Band::make(m, n);
    for (uint32_t t = 0; t < n_ops; ++t) {
      const uint32_t* o = ops + 5 * t;
      if (o[0] == 0) b.add_entry({o[1], o[2]}, o[4]);
      else b.add_kmer({o[1], o[2]}, o[3], o[4]);
    }
    for (uint64_t j = 0; j <= n; ++j) {
      ranges[2 * j] = b.ranges[j].start;
      ranges[2 * j + 1] = b.ranges[j].end;
    }
    return g_panic ? -1 : 0;
  }
}

int orc_banded_align(int mode, const orc_scoring* scoring, uint32_t k, uint32_t w, const uint8_t* x, uint32_t m,
                     const uint8_t* y, uint32_t n, orc_alignment* out, uint32_t* ops) {
  g_panic = false;
  BandedAligner a;
  a.sc = *scoring;
  a.k = k;
  a.w = w;
  std::vector<Op> v;
  a.align(mode, x, m, y, n, out, v);
  if (g_panic) {
    out->n_ops = 0xFFFFFFFFu;
    return -1;
  }
  for (size_t t = 0; t < v.size(); ++t) ops[t] = v[t].code | (v[t].len << 3);
  return 0;
}

// lcskpp / union path / expand_kmer_matches for the reference's unit vectors (sparse.rs:518-760)
int orc_lcskpp(const uint32_t* xy, uint64_t n_matches, uint32_t k, uint64_t* out_path, uint64_t* n_path,
               uint32_t* score) {
  g_panic = false;
  std::vector<Match> m(n_matches);
  for (uint64_t i = 0; i < n_matches; ++i) m[i] = {xy[2 * i], xy[2 * i + 1]};
  const SparseResult r = lcskpp(m, k);
  for (uint64_t i = 0; i < r.path.size(); ++i) out_path[i] = r.path[i];
  *n_path = r.path.size();
  *score = r.score;
  return g_panic ? -1 : 0;
}

int orc_sdpkpp_union_lcskpp_path(const uint32_t* xy, uint64_t n_matches, uint32_t k, uint32_t match_score,
                                 int32_t gap_open, int32_t gap_extend, uint64_t* out_path, uint64_t* n_path) {
  g_panic = false;
  std::vector<Match> m(n_matches);
  for (uint64_t i = 0; i < n_matches; ++i) m[i] = {xy[2 * i], xy[2 * i + 1]};
  const std::vector<uint64_t> r = sdpkpp_union_lcskpp_path(m, k, match_score, gap_open, gap_extend);
  for (uint64_t i = 0; i < r.size(); ++i) out_path[i] = r[i];  // cap >= 2 * n_matches
  *n_path = r.size();
  return g_panic ? -1 : 0;
}

// returns the number of expanded matches (only the first cap are written), or ~0 on a panic path
uint64_t orc_expand_kmer_matches(const uint8_t* x, uint32_t m, const uint8_t* y, uint32_t n, uint32_t k,
                                 const uint32_t* xy, uint64_t n_matches, uint32_t allowed_mismatches,
                                 uint32_t* out_xy, uint64_t cap) {
  g_panic = false;
  std::vector<Match> mm(n_matches);
  for (uint64_t i = 0; i < n_matches; ++i) mm[i] = {xy[2 * i], xy[2 * i + 1]};
  const std::vector<Match> r = expand_kmer_matches(x, m, y, n, k, mm, allowed_mismatches);
  if (g_panic) return ~0ull;
  for (uint64_t i = 0; i < r.size() && i < cap; ++i) {
    out_xy[2 * i] = r[i].first;
    out_xy[2 * i + 1] = r[i].second;
  }
  return r.size();
}

// banded::Aligner::custom_with_{matches, expanded_matches, match_path} (have_path / allowed_mismatches < 0 /
// use_lcskpp_union select which, see BandedAligner::align_hinted)
int orc_banded_align_hinted(const orc_scoring* scoring, uint32_t k, uint32_t w, const uint8_t* x, uint32_t m,
                            const uint8_t* y, uint32_t n, const uint32_t* match_xy, uint64_t n_matches,
                            const uint64_t* path, uint64_t n_path, int32_t have_path, int32_t allowed_mismatches,
                            int32_t use_lcskpp_union, orc_alignment* out, uint32_t* ops, uint64_t* num_cells) {
  g_panic = false;
  BandedAligner a;
  a.sc = *scoring;
  a.k = k;
  a.w = w;
  std::vector<Match> mm(n_matches);
  for (uint64_t i = 0; i < n_matches; ++i) mm[i] = {match_xy[2 * i], match_xy[2 * i + 1]};
  std::vector<uint64_t> pp(path, path + (have_path ? n_path : 0));
  std::vector<Op> v;
  a.align_hinted(x, m, y, n, true, mm, have_path != 0, pp, allowed_mismatches, use_lcskpp_union != 0, out, v);
  if (g_panic) {
    out->n_ops = 0xFFFFFFFFu;
    return -1;
  }
  if (num_cells) *num_cells = a.band.num_cells();
  for (size_t t = 0; t < v.size(); ++t) ops[t] = v[t].code | (v[t].len << 3);
  return 0;
}

double orc_banded_align_batch(int mode, const orc_scoring* scoring, uint32_t k, uint32_t w, const uint8_t* blob,
                              const uint64_t* x_off, const uint32_t* x_len, const uint64_t* y_off,
                              const uint32_t* y_len, uint64_t n_pairs, orc_alignment* out, uint32_t* ops,
                              const uint64_t* ops_off, uint64_t* cells, int threads) {
  if (threads < 1) threads = 1;
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  std::vector<uint64_t> cell_acc(threads, 0);
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([=, &cell_acc]() {
      BandedAligner a;
      a.sc = *scoring;
      a.k = k;
      a.w = w;
      std::vector<Op> v;
      const uint64_t lo = n_pairs * (uint64_t)t / (uint64_t)threads;
      const uint64_t hi = n_pairs * (uint64_t)(t + 1) / (uint64_t)threads;
      for (uint64_t p = lo; p < hi; ++p) {
        g_panic = false;
        a.align(mode, blob + x_off[p], x_len[p], blob + y_off[p], y_len[p], &out[p], v);
        cell_acc[t] += a.band.num_cells();
        if (g_panic) {
          out[p].n_ops = 0xFFFFFFFFu;
        } else if (ops) {
          for (size_t q = 0; q < v.size(); ++q) ops[ops_off[p] + q] = v[q].code | (v[q].len << 3);
        }
      }
    });
  }
  for (auto& th : pool) th.join();
  if (cells) {
    *cells = 0;
    for (uint64_t c : cell_acc) *cells += c;
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
