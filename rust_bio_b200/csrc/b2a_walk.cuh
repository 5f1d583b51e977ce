// K2: row m, last-column fix-ups and the traceback walk -- one lane per pair.
//
// Reference rust-bio 4.0.1 src/alignment/pairwise/mod.rs, Aligner::custom:
//   row m of the fill (the only row whose S starts from the column tracker
//     S[curr][m] instead of MIN_SCORE)                      641-645, 757-758
//   "Handle suffix clipping in the j=n case"                809-821
//   "recompute the last column of I"                        825-843
//   the traceback state machine                             845-908
//   Alignment construction / clip filtering                 910-921, 974, 1006
// These parts are sequential per pair and O(m+n); they are replayed literally
// here on top of what K1 left in HBM: the boundary row m-1 (S, I, column
// tracker), the row trackers, the last column and the 4-bit traceback.
#pragma once
#include "b2a_common.cuh"

namespace b2a {

struct WalkParams {
  const Block* blocks;
  uint32_t nblocks;
  const uint32_t* pm;
  const uint32_t* pn;
  const uint32_t* order;  // sorted pair -> caller's pair index
  const uint8_t* seq;
  const uint8_t* bnd;
  uint8_t* rows;
  uint8_t* rowm;
  const uint8_t* tb;
  uint8_t* ops_scratch;
  const int32_t* lut;
  DevScoring sc;
  int32_t G, R;
  int32_t filter_clips;  // semiglobal / local: Alignment::filter_clip_operations
  int32_t packtrk;       // K1 ran with F_PACKTRK (how the column tracker in the boundary row is encoded)
  // outputs, indexed by the caller's pair index
  int32_t* score;
  uint32_t* xstart;
  uint32_t* xend;
  uint32_t* ystart;
  uint32_t* yend;
  uint32_t* n_ops;
  uint64_t* ops_src;   // where the pair's ops start inside ops_scratch
  uint32_t* clip_len;  // 4 per pair
  uint32_t* status;    // 0 ok, 1 = corrupt traceback (reference would panic, mod.rs:905)
  uint32_t* err_flag;  // set to 1 if any pair's status is non-zero
};

constexpr uint32_t LAZY = 15;  // "came from S of the neighbour": resolved when the walk needs it

struct PairView {
  DevScoring sc;
  const int32_t* lut;
  const uint32_t* xw;  // staged x words of this pair's task; word w at xw[w*P]
  const uint32_t* yw;
  int32_t P;
  int32_t m, n, pi;
  int32_t G, R, TBW, nstrips, K;
  const int4* bnd;     // [column][32]
  int32_t* rows;       // arrays of [rows_pad][32]
  int32_t rows_pad;
  uint16_t* rowm;      // [column][32]
  const uint32_t* tb;  // block base
  int32_t sub, g;      // task inside the block, slot inside the task
  int32_t packtrk;
  int32_t maxn;        // block maximum of n
  int64_t bnd_base;    // boundary row of this pair: bnd[bnd_base + j * bnd_stride] (see bnd_index)
  int32_t bnd_stride;

  B2A_HD int32_t xsym(int32_t i) const {  // x[i-1]
    const int32_t b = i - 1;
    return (int32_t)((xw[(b >> 2) * P] >> (8 * (b & 3))) & 0xffu);
  }
  B2A_HD int32_t ysym(int32_t j) const {
    const int32_t b = j - 1;
    return (int32_t)((yw[(b >> 2) * P] >> (8 * (b & 3))) & 0xffu);
  }
  B2A_HD int32_t score(int32_t p, int32_t q) const {
    if (sc.alpha) return lut[p * sc.alpha + q];
    return p == q ? sc.match_score : sc.mismatch_score;
  }
  B2A_HD int32_t& row(int arr, int32_t i) const { return rows[(arr * rows_pad + i) * 32 + pi]; }
  B2A_HD int4 load_bnd(int32_t j) const {
#if defined(__CUDA_ARCH__)
    return __ldg(&bnd[bnd_base + (int64_t)j * bnd_stride]);  // read-only path: K1 wrote it in an earlier launch
#else
    return bnd[bnd_base + (int64_t)j * bnd_stride];
#endif
  }
  // compressed traceback nibble of an interior cell 1 <= i <= m-1, 1 <= j <= n
  B2A_HD uint32_t nib(int32_t i, int32_t j) const {
    const int32_t GR = G * R;
    const int32_t s = (i - 1) / GR, rem = (i - 1) % GR;
    const int32_t l = rem / R, r = rem % R;
    const int32_t lane = g * G + l;
    const int32_t t = (j - 1) + l;
    const size_t word =
        ((((size_t)(sub * nstrips + s) * K + (t >> 3)) * TBW + (r >> 2)) * 32 + lane) * 4 + (r & 3);
    return (tb[word] >> (4 * (7 - (t & 7)))) & 15u;
  }
  B2A_HD uint32_t nib_scode(uint32_t nb, int32_t i, int32_t j) const {
    switch (nb & 3u) {
      case NB_DIAG: return xsym(i) == ysym(j) ? TB_MATCH : TB_SUBST;
      case NB_INS: return TB_INS;
      case NB_DEL: return TB_DEL;
      default: return TB_XCLIP_PREFIX;
    }
  }
};

// Boundary row m-1 as K1 leaves it (scaled domain, b2a_fill.cuh): x = 4*S, y = 4*I + 2,
// z/w = column tracker: packed key 4096*(max S) + (4095 - first row) [F_PACKTRK] or (4*(T), row).
struct Boundary {
  int32_t S, I, Tv, Ti;
};
B2A_HD Boundary decode_boundary(const int4 b, const bool packtrk, const int32_t xs, const int32_t m) {
  Boundary o;
  o.S = b.x >> 2;
  o.I = b.y >> 2;
  o.Tv = MIN_SCORE;
  o.Ti = m;
  if (packtrk) {
    if (b.z != (int32_t)0x80000000) {
      o.Tv = (b.z >> 12) + xs;
      o.Ti = 4095 - (b.z & 4095);
    }
  } else if (b.z > -(1 << 29)) {
    o.Tv = b.z >> 2;
    o.Ti = b.w;
  }
  return o;
}

// cells kept as the reference's u16: i | d << 4 | s << 8 (mod.rs:1031-1033)
B2A_HD uint32_t cell_make(uint32_t i, uint32_t d, uint32_t s) { return i | (d << 4) | (s << 8); }
B2A_HD uint32_t cell_i(uint32_t c) { return c & 15u; }
B2A_HD uint32_t cell_d(uint32_t c) { return (c >> 4) & 15u; }
B2A_HD uint32_t cell_s(uint32_t c) { return (c >> 8) & 15u; }
B2A_HD uint32_t cell_set_s(uint32_t c, uint32_t s) { return (c & ~0xF00u) | (s << 8); }
B2A_HD uint32_t cell_set_i(uint32_t c, uint32_t i) { return (c & ~0x00Fu) | i; }

struct WalkOut {
  int32_t score;
  uint32_t xstart, xend, ystart, yend, n_ops, status;
  uint32_t clip[4];
};

// ops are written backwards into ops_end[-1], ops_end[-2], ...
B2A_HD void walk_pair(const PairView& v, const bool filter_clips, uint8_t* ops_end, WalkOut& out) {
  const DevScoring& sc = v.sc;
  const int32_t m = v.m, n = v.n;
  const int32_t go = sc.gap_open, ge = sc.gap_extend;
  const int32_t xp = sc.xclip_prefix, xs = sc.xclip_suffix, yp = sc.yclip_prefix,
                ys = sc.yclip_suffix;

  // ------------------------------------------------------------------ row m
  int32_t SmN = 0, ImN = MIN_SCORE;   // S(m,n), I(m,n)
  uint32_t cmN = 0;                   // cell (m,n)
  int32_t Snm = MIN_SCORE, Lym = 0;   // Sn[m], Ly[m]
  int32_t Lx0 = 0, LxN = 0;           // Lx[0], Lx[n]
  if (m >= 1) {
    // column 0 (mod.rs:622-671 at i == m); the tracker over rows 1..m-1 first
    int32_t T = MIN_SCORE;
    for (int32_t i = 1; i < m; ++i) {
      const int32_t val = col0_S(sc, i) + xs;
      if (val > T) {
        T = val;
        Lx0 = m - i;
      }
    }
    int32_t Im = col0_I(sc, m);
    uint32_t ib = col0_ibits(sc, m);
    int32_t Sm = T;
    uint32_t sb = TB_XCLIP_SUFFIX;
    if (Im > Sm) {
      Sm = Im;
      sb = TB_INS;
    }
    if (xp > Sm) {
      Sm = xp;
      sb = TB_XCLIP_PREFIX;
    }
    if (Sm + ys > Snm) {
      Snm = Sm + ys;
      Lym = n;
    }
    int32_t Dm = MIN_SCORE;
    uint32_t cell = cell_make(ib, TB_START, sb);
    v.rowm[0 * 32 + v.pi] = (uint16_t)cell;
    LxN = Lx0;
    const int32_t p = v.xsym(m);
    const int32_t yclip_score = yp + go + ge * (m - 1);
    int32_t sdiag = (m == 1) ? 0 : col0_S(sc, m - 1);
    // columns are walked four at a time: the boundary row and y symbols of a group are loaded up front
    // (independent loads in flight), then the recurrence runs over them in order
    constexpr int UB = 4;
    for (int32_t j0 = 1; j0 <= n; j0 += UB) {
      int4 braw[UB];
      int32_t qv[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int32_t jj = j0 + u;
        braw[u] = make_int4(0, 0, 0, 0);
        qv[u] = 0;
        if (jj <= n) {
          if (m != 1) braw[u] = v.load_bnd(jj);
          qv[u] = v.ysym(jj);
        }
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
      const int32_t j = j0 + u;
      if (j > n) break;
      int32_t sup, iup, Tv, Ti;
      if (m == 1) {
        sup = row0_S(sc, j, n);
        iup = MIN_SCORE;
        Tv = MIN_SCORE;
        Ti = m;
      } else {
        const Boundary b = decode_boundary(braw[u], v.packtrk != 0, xs, m);
        sup = b.S;
        iup = b.I;
        Tv = b.Tv;
        Ti = b.Ti;
      }
      const int32_t q = qv[u];
      const int32_t m_score = sdiag + v.score(p, q);
      int32_t best_i, best_d;
      {
        const int32_t i_score = iup + ge, s_score = sup + go;
        if (i_score > s_score) {
          best_i = i_score;
          ib = TB_INS;
        } else {
          best_i = s_score;
          ib = LAZY;
        }
      }
      uint32_t db;
      {
        const int32_t d_score = Dm + ge, s_score = Sm + go;
        if (d_score > s_score) {
          best_d = d_score;
          db = TB_DEL;
        } else {
          best_d = s_score;
          db = sb;  // s_bits of (m, j-1), final for j-1 < n
        }
      }
      int32_t best = Tv;
      sb = TB_XCLIP_SUFFIX;
      if (m_score > best) {
        best = m_score;
        sb = (p == q) ? TB_MATCH : TB_SUBST;
      }
      if (best_i > best) {
        best = best_i;
        sb = TB_INS;
      }
      if (best_d > best) {
        best = best_d;
        sb = TB_DEL;
      }
      const int32_t xcs = xclip_score(sc, j);
      if (xcs > best) {
        best = xcs;
        sb = TB_XCLIP_PREFIX;
      }
      if (yclip_score > best) {
        best = yclip_score;
        sb = TB_YCLIP_PREFIX;
      }
      Sm = best;
      Im = best_i;
      Dm = best_d;
      if (Sm + ys > Snm) {
        Snm = Sm + ys;
        Lym = n - j;
      }
      if (j == n) {
        LxN = m - Ti;
        if (ib == LAZY) {  // i_bits captured before the fix-ups touch (m-1, n)
          ib = (m == 1) ? row0_sbits(sc, n, n)
                        : v.nib_scode((uint32_t)v.row(ROWS_NL, m - 1), m - 1, n);
        }
      }
      cell = cell_make(ib, db, sb);
      v.rowm[j * 32 + v.pi] = (uint16_t)cell;
      sdiag = sup;
          }
    }
    SmN = Sm;
    ImN = Im;
    cmN = cell;
  }

  // ------------------------------- column n: materialise the cells K1 left as nibbles and run fix-up 1
  // (mod.rs:809-821) in the same pass over the rows; rows 0..m-1 keep S in ROWS_SL, I in ROWS_IL, the
  // literal cell in ROWS_NL.  i_bits are captured from the PRE-fix-up s_bits of the row above (743),
  // exactly as the reference's fill did before its fix-up loops ran.
  // K1 only keeps the row trackers when yclip_suffix is live; a dead one can never win (Sn <= MIN/2 + S)
  const bool ys_live = ys > DEAD_CLIP;
  {
    int32_t s0, c0;
    if (n == 0) {
      s0 = 0;
      c0 = (int32_t)cell_make(TB_START, TB_START, TB_START);
    } else {
      s0 = row0_S(sc, n, n);
      c0 = (int32_t)cell_make(TB_START, row0_dbits(sc, n), row0_sbits(sc, n, n));
    }
    if (m == 0) {
      SmN = s0;
      cmN = (uint32_t)c0;
      Snm = ys;
      Lym = n;
    } else {
      uint32_t s_above = cell_s((uint32_t)c0);  // pre fix-up s_bits(i-1, n)
      {  // row 0: Sn[0] = yclip_suffix (mod.rs:618, 711)
        int32_t S = s0;
        uint32_t cell = (uint32_t)c0;
        if (ys > S) {
          S = ys;
          cell = cell_set_s(cell, TB_YCLIP_SUFFIX);
        }
        v.row(ROWS_SL, 0) = S;
        v.row(ROWS_IL, 0) = MIN_SCORE;
        v.row(ROWS_NL, 0) = (int32_t)cell;
        if (S + xs > SmN) {
          SmN = S + xs;
          LxN = m;
          cmN = cell_set_s(cmN, TB_XCLIP_SUFFIX);
        }
      }
      constexpr int UB = 4;
      for (int32_t i0 = 1; i0 < m; i0 += UB) {
        int32_t nbv[UB], Sv[UB], Snv[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int32_t i = i0 + u;
          nbv[u] = Sv[u] = 0;
          Snv[u] = MIN_SCORE;
          if (i < m && n != 0) {
            nbv[u] = v.row(ROWS_NL, i);
            Sv[u] = v.row(ROWS_SL, i);
            if (ys_live) Snv[u] = v.row(ROWS_SN, i);
          }
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int32_t i = i0 + u;
          if (i >= m) break;
          uint32_t cell;
          int32_t S, Sn;
          if (n == 0) {  // column n is column 0: closed forms (mod.rs:622-671)
            S = col0_S(sc, i);
            v.row(ROWS_IL, i) = col0_I(sc, i);
            cell = cell_make(col0_ibits(sc, i), TB_START, col0_sbits(sc, i));
            const int32_t val = S + ys;
            Sn = val > MIN_SCORE ? val : MIN_SCORE;
            if (!ys_live) Sn = MIN_SCORE;
            v.row(ROWS_LY, i) = 0;
          } else {
            const uint32_t nb = (uint32_t)nbv[u];
            cell = cell_make((nb & NB_IEXT) ? (uint32_t)TB_INS : s_above, (nb & NB_DEXT) ? (uint32_t)TB_DEL : LAZY,
                             v.nib_scode(nb, i, n));
            S = Sv[u];
            Sn = Snv[u];
          }
          s_above = cell_s(cell);
          if (Sn > S) {  // fix-up 1
            S = Sn;
            cell = cell_set_s(cell, TB_YCLIP_SUFFIX);
          }
          v.row(ROWS_SL, i) = S;
          v.row(ROWS_NL, i) = (int32_t)cell;
          if (S + xs > SmN) {
            SmN = S + xs;
            LxN = m - i;
            cmN = cell_set_s(cmN, TB_XCLIP_SUFFIX);
          }
        }
      }
      if (Snm > SmN) {  // i == m
        SmN = Snm;
        cmN = cell_set_s(cmN, TB_YCLIP_SUFFIX);
      }
      // ---------------------------------------------- fix-up 2, mod.rs:825-843
      int32_t S_prev = v.row(ROWS_SL, 0);
      uint32_t cell_prev = (uint32_t)v.row(ROWS_NL, 0);
      for (int32_t i0 = 1; i0 <= m; i0 += UB) {
        int32_t Iv[UB], Sv[UB], Cv[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int32_t i = i0 + u;
          Iv[u] = Sv[u] = Cv[u] = 0;
          if (i < m) {
            Iv[u] = v.row(ROWS_IL, i);
            Sv[u] = v.row(ROWS_SL, i);
            Cv[u] = v.row(ROWS_NL, i);
          }
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int32_t i = i0 + u;
          if (i > m) break;
          const int32_t s_score = S_prev + go;
          int32_t I = (i == m) ? ImN : Iv[u];
          int32_t S = (i == m) ? SmN : Sv[u];
          uint32_t cell = (i == m) ? cmN : (uint32_t)Cv[u];
          bool dirty = false;
          if (s_score > I) {
            I = s_score;
            cell = cell_set_i(cell, cell_s(cell_prev));
            dirty = true;
          }
          if (s_score > S) {
            S = s_score;
            cell = cell_set_s(cell, TB_INS);
            dirty = true;
            if (i != m && S + xs > SmN) {
              SmN = S + xs;
              LxN = m - i;
              cmN = cell_set_s(cmN, TB_XCLIP_SUFFIX);
            }
          }
          if (i == m) {
            ImN = I;
            SmN = S;
            cmN = cell;
          } else if (dirty) {
            v.row(ROWS_IL, i) = I;
            v.row(ROWS_SL, i) = S;
            v.row(ROWS_NL, i) = (int32_t)cell;
          }
          S_prev = S;
          cell_prev = cell;
        }
      }
    }
  }

  // ----------------------------------------------------- the walk, mod.rs:845-908
  auto get_cell_n = [&](int32_t i) -> uint32_t {  // column n, after the fix-ups
    return (i == m) ? cmN : (uint32_t)v.row(ROWS_NL, i);
  };
  auto get_s = [&](int32_t i, int32_t j) -> uint32_t {
    if (j == n) return cell_s(get_cell_n(i));
    if (i == 0) return row0_sbits(sc, j, n);
    if (i == m) return cell_s((uint32_t)v.rowm[j * 32 + v.pi]);
    if (j == 0) return col0_sbits(sc, i);
    return v.nib_scode(v.nib(i, j), i, j);
  };
  int32_t i = m, j = n;
  uint32_t xstart = 0, ystart = 0, xend = (uint32_t)m, yend = (uint32_t)n;
  uint32_t nops = 0, nclip = 0, status = 0;
  uint32_t clips[4] = {0, 0, 0, 0};
  uint32_t layer = cell_s(cmN);
  int32_t guard = m + n + 8;
  while (layer != TB_START) {
    if (--guard < 0) {
      status = 1;
      break;
    }
    uint32_t next;
    if (layer == TB_INS) {
      *(--ops_end) = 3;
      ++nops;
      uint32_t c;
      if (j == n) {
        c = cell_i(get_cell_n(i));
      } else if (i == m) {
        c = cell_i((uint32_t)v.rowm[j * 32 + v.pi]);
        if (c == LAZY) c = get_s(m - 1, j);
      } else if (j == 0) {
        c = col0_ibits(sc, i);
      } else {
        c = (v.nib(i, j) & NB_IEXT) ? (uint32_t)TB_INS : get_s(i - 1, j);
      }
      next = c;
      i -= 1;
    } else if (layer == TB_DEL) {
      *(--ops_end) = 2;
      ++nops;
      uint32_t c;
      if (i == 0) {
        c = row0_dbits(sc, j);
      } else if (i == m) {
        c = cell_d((uint32_t)v.rowm[j * 32 + v.pi]);
      } else if (j == n) {
        c = cell_d(get_cell_n(i));
        if (c == LAZY) c = get_s(i, n - 1);
      } else {
        c = (v.nib(i, j) & NB_DEXT) ? (uint32_t)TB_DEL : get_s(i, j - 1);
      }
      next = c;
      j -= 1;
    } else if (layer == TB_MATCH || layer == TB_SUBST) {
      *(--ops_end) = (layer == TB_MATCH) ? 0 : 1;
      ++nops;
      next = get_s(i - 1, j - 1);
      i -= 1;
      j -= 1;
    } else if (layer == TB_XCLIP_PREFIX) {
      if (!filter_clips) {
        *(--ops_end) = 4;
        ++nops;
        if (nclip < 4) clips[nclip] = (uint32_t)i;
        ++nclip;
      }
      xstart = (uint32_t)i;
      i = 0;
      next = get_s(0, j);
    } else if (layer == TB_XCLIP_SUFFIX) {
      int32_t lx;
      if (j == n) lx = LxN;
      else if (j == 0) lx = Lx0;
      else lx = (m >= 2) ? m - decode_boundary(v.load_bnd(j), v.packtrk != 0, xs, m).Ti : 0;
      if (!filter_clips) {
        *(--ops_end) = 4;
        ++nops;
        if (nclip < 4) clips[nclip] = (uint32_t)lx;
        ++nclip;
      }
      i -= lx;
      xend = (uint32_t)i;
      next = get_s(i, j);
    } else if (layer == TB_YCLIP_PREFIX) {
      if (!filter_clips) {
        *(--ops_end) = 5;
        ++nops;
        if (nclip < 4) clips[nclip] = (uint32_t)j;
        ++nclip;
      }
      ystart = (uint32_t)j;
      j = 0;
      next = get_s(i, 0);
    } else if (layer == TB_YCLIP_SUFFIX) {
      int32_t ly;
      if (i == m) ly = Lym;
      else if (i == 0 || n == 0) ly = n;
      else ly = n - v.row(ROWS_LY, i);
      if (!filter_clips) {
        *(--ops_end) = 5;
        ++nops;
        if (nclip < 4) clips[nclip] = (uint32_t)ly;
        ++nclip;
      }
      j -= ly;
      yend = (uint32_t)j;
      next = get_s(i, j);
    } else {
      status = 1;  // panic!("Dint expect this!") mod.rs:905
      break;
    }
    if (i < 0 || j < 0) {
      status = 1;
      break;
    }
    layer = next;
  }
  if (nclip > 4) status = 1;
  out.score = SmN;
  out.xstart = xstart;
  out.xend = xend;
  out.ystart = ystart;
  out.yend = yend;
  out.n_ops = nops;
  out.status = status;
  // clips were met end-to-start; report them in alignment order
  const uint32_t nc = nclip > 4 ? 4 : nclip;
  for (uint32_t k = 0; k < 4; ++k) out.clip[k] = (k < nc) ? clips[nc - 1 - k] : 0u;
}

#if defined(__CUDACC__)

// K2 for one pair (lane) of a block
__device__ __forceinline__ void walk_lane(const WalkParams& prm, const Block& blk, const int lane) {
  if ((uint32_t)lane >= blk.npairs) return;
  const uint32_t sp = blk.first + lane;
  const int32_t P = 32 / prm.G;
  PairView v;
  v.sc = prm.sc;
  v.lut = prm.lut;
  v.P = P;
  v.m = (int32_t)prm.pm[sp];
  v.n = (int32_t)prm.pn[sp];
  v.pi = lane;
  v.G = prm.G;
  v.R = prm.R;
  v.TBW = (prm.R + 3) / 4;
  v.nstrips = (int32_t)blk.nstrips;
  v.K = (int32_t)blk.K;
  v.sub = lane / P;
  v.g = lane % P;
  v.packtrk = prm.packtrk;
  v.maxn = (int32_t)blk.maxn;
  v.bnd_base = bnd_index(prm.G, 0, lane, v.maxn);
  v.bnd_stride = (int32_t)(bnd_index(prm.G, 1, lane, v.maxn) - v.bnd_base);
  const uint32_t* seqw = reinterpret_cast<const uint32_t*>(prm.seq + blk.seq_off);
  v.xw = seqw + (size_t)v.sub * blk.xwords * P + v.g;
  v.yw = seqw + (size_t)prm.G * blk.xwords * P + (size_t)v.sub * blk.ywords * P + v.g;
  v.bnd = reinterpret_cast<const int4*>(prm.bnd + blk.bnd_off);
  v.rows = reinterpret_cast<int32_t*>(prm.rows + blk.rows_off);
  v.rows_pad = (int32_t)blk.rows_pad;
  v.rowm = reinterpret_cast<uint16_t*>(prm.rowm + blk.rowm_off);
  v.tb = reinterpret_cast<const uint32_t*>(prm.tb + blk.tb_off);
  const uint32_t cap = blk.maxm + blk.maxn + 4;
  uint8_t* ops_end = prm.ops_scratch + blk.ops_off + (size_t)(lane + 1) * cap;
  WalkOut o;
  walk_pair(v, prm.filter_clips != 0, ops_end, o);
  if (o.status) {  // the reference panics on this pair (mod.rs:905): no alignment is reported for it
    o.score = MIN_SCORE;
    o.n_ops = 0;
    o.xstart = o.xend = o.ystart = o.yend = 0;
    o.clip[0] = o.clip[1] = o.clip[2] = o.clip[3] = 0;
  }
  const uint32_t dst = prm.order[sp];
  prm.score[dst] = o.score;
  prm.xstart[dst] = o.xstart;
  prm.xend[dst] = o.xend;
  prm.ystart[dst] = o.ystart;
  prm.yend[dst] = o.yend;
  prm.n_ops[dst] = o.n_ops;
  prm.ops_src[dst] = blk.ops_off + (uint64_t)(lane + 1) * cap - o.n_ops;
  prm.status[dst] = o.status;
  if (o.status) atomicOr(prm.err_flag, 1u);
#pragma unroll
  for (int k = 0; k < 4; ++k) prm.clip_len[4 * (size_t)dst + k] = o.clip[k];
}

#if defined(B2A_DEFINE_WALK_KERNEL)  // one translation unit (b2a_engine.cu) owns the stand-alone kernel
__global__ void __launch_bounds__(128, 8) walk_kernel(const WalkParams prm) {
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= prm.nblocks) return;
  const Block blk = prm.blocks[gw];
  walk_lane(prm, blk, lane);
}
#endif

#endif

}  // namespace b2a
