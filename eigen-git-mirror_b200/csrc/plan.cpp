// plan.cpp -- host-side planning (see plan.h).  Pure C++17, no CUDA.
#include "plan.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace b200s {

namespace {

inline int floor_log2(int v) {
  int l = 0;
  while ((2 << l) <= v) ++l;
  return l;
}
inline int ceil_log2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// Compress an uncompressed matrix and/or expand one stored triangle to the full symmetric pattern.
// Mirrors what the reference does lazily: SparseCompressedBase::InnerIterator skips the unused tail of each inner
// vector (SparseCompressedBase.h:189-194) and SparseSelfAdjointView reads one triangle and applies it twice
// (SparseSelfAdjointView.h:303-332; the explicit expansion is permute_symm_to_fullsymm, :427).
int canonicalise(int64_t rows, int64_t nnz, const int32_t* rowptr, const int32_t* colidx, const int32_t* inner_nnz,
                 int uplo, Plan& p, std::string& err) {
  auto row_end = [&](int64_t i) { return inner_nnz ? rowptr[i] + inner_nnz[i] : rowptr[i + 1]; };
  p.rowptr.assign(rows + 1, 0);
  if (uplo == B200S_BOTH) {
    if (!inner_nnz) {  // already canonical
      std::memcpy(p.rowptr.data(), rowptr, sizeof(int32_t) * (rows + 1));
      if (p.rowptr[0] != 0) {  // a Map/Ref of an inner panel may start past 0: rebase
        const int32_t base = p.rowptr[0];
        const int64_t cnt = static_cast<int64_t>(p.rowptr[rows]) - base;  // entries actually referenced
        if (base < 0 || cnt < 0) { err = "rowptr is not monotone"; return B200S_ERR_INVALID; }
        for (auto& v : p.rowptr) v -= base;
        p.src.resize(cnt);
        for (int64_t k = 0; k < cnt; ++k) p.src[k] = base + static_cast<int32_t>(k);
        p.colidx.assign(colidx + base, colidx + base + cnt);
        p.input_nnz = static_cast<int64_t>(base) + cnt;  // the caller's arrays are read up to slot base+cnt-1
      } else if (rows > 0 && p.rowptr[rows] > nnz) {
        err = "rowptr[rows] exceeds nnz";
        return B200S_ERR_INVALID;
      }
      return 0;
    }
    int64_t total = 0, span = 0;
    for (int64_t i = 0; i < rows; ++i) {
      if (inner_nnz[i] < 0) { err = "inner_nnz is negative"; return B200S_ERR_INVALID; }
      total += inner_nnz[i];
      span = std::max<int64_t>(span, row_end(i));
    }
    if (total >= (int64_t(1) << 31)) { err = "nnz does not fit int32"; return B200S_ERR_UNSUPPORTED; }
    p.input_nnz = span;  // one past the last slot any row references (the arrays have holes)
    p.src.resize(total);
    p.colidx.resize(total);
    int64_t o = 0;
    for (int64_t i = 0; i < rows; ++i) {
      p.rowptr[i] = static_cast<int32_t>(o);
      for (int32_t k = rowptr[i]; k < row_end(i); ++k, ++o) {
        p.src[o] = k;
        p.colidx[o] = colidx[k];
      }
    }
    p.rowptr[rows] = static_cast<int32_t>(o);
    return 0;
  }
  // one triangle -> full symmetric
  const bool lower = (uplo == B200S_LOWER);
  std::vector<int64_t> n_lo(rows, 0), n_di(rows, 0), n_hi(rows, 0);
  {
    int64_t span = 0;
    for (int64_t i = 0; i < rows; ++i) span = std::max<int64_t>(span, row_end(i));
    p.input_nnz = span;
  }
  for (int64_t i = 0; i < rows; ++i)
    for (int32_t k = rowptr[i]; k < row_end(i); ++k) {
      int64_t c = colidx[k];
      if (c == i) { n_di[i]++; continue; }
      if (lower ? (c < i) : (c > i)) {
        (c < i ? n_lo[i] : n_hi[i])++;   // the stored entry (i,c)
        (i < c ? n_lo[c] : n_hi[c])++;   // its mirror (c,i)
      }
    }
  int64_t total = 0;
  std::vector<int64_t> cur_lo(rows), cur_di(rows), cur_hi(rows);
  for (int64_t i = 0; i < rows; ++i) {
    p.rowptr[i] = static_cast<int32_t>(total);
    cur_lo[i] = total;
    cur_di[i] = total + n_lo[i];
    cur_hi[i] = cur_di[i] + n_di[i];
    total += n_lo[i] + n_di[i] + n_hi[i];
    if (total >= (int64_t(1) << 31)) { err = "expanded nnz does not fit int32"; return B200S_ERR_UNSUPPORTED; }
  }
  p.rowptr[rows] = static_cast<int32_t>(total);
  p.src.resize(total);
  p.colidx.resize(total);
  for (int64_t i = 0; i < rows; ++i)
    for (int32_t k = rowptr[i]; k < row_end(i); ++k) {
      int64_t c = colidx[k];
      if (c == i) {
        p.src[cur_di[i]] = k; p.colidx[cur_di[i]++] = static_cast<int32_t>(c);
      } else if (lower ? (c < i) : (c > i)) {
        int64_t& own = (c < i) ? cur_lo[i] : cur_hi[i];
        p.src[own] = k; p.colidx[own++] = static_cast<int32_t>(c);
        int64_t& mir = (i < c) ? cur_lo[c] : cur_hi[c];
        p.src[mir] = k; p.colidx[mir++] = static_cast<int32_t>(i);
      }
    }
  return 0;
}

// One stored triangle of a ROW-PARTITIONED symmetric matrix -> this rank's rows of the full matrix (global columns).
// The reference has no counterpart (it is single-process); the single-rank case is `canonicalise` above, and the
// operator is the one SparseSelfAdjointView defines (SparseSelfAdjointView.h:279-337): stored triangle + its mirror.
int canonicalise_distributed(const b200s_config& cfg, int64_t rows, const int32_t* rowptr, const int32_t* colidx,
                             const int32_t* inner_nnz, int uplo, Plan& p, std::string& err) {
  auto row_end = [&](int64_t i) { return inner_nnz ? rowptr[i] + inner_nnz[i] : rowptr[i + 1]; };
  const bool lower = (uplo == B200S_LOWER);
  const int W = p.world;
  const int64_t lo = p.row0, hi = p.row0 + rows;
  auto owner = [&](int64_t c) {
    int q = static_cast<int>(std::upper_bound(p.row_starts.begin(), p.row_starts.end(), c) - p.row_starts.begin()) - 1;
    return std::max(0, std::min(W - 1, q));
  };
  struct Ent { int64_t col; int64_t src; };
  std::vector<std::vector<Ent>> row_ents(rows);
  struct Exp { int64_t row, col; int32_t src; };
  std::vector<std::vector<Exp>> exports(W);
  int64_t span = 0;
  for (int64_t il = 0; il < rows; ++il) {
    const int64_t i = lo + il;
    span = std::max<int64_t>(span, row_end(il));
    for (int32_t k = rowptr[il]; k < row_end(il); ++k) {
      const int64_t c = colidx[k];
      if (c < 0 || c >= p.cols) { err = "column index out of range"; return B200S_ERR_INVALID; }
      if (c == i) { row_ents[il].push_back({c, k}); continue; }
      if (!(lower ? (c < i) : (c > i))) continue;  // the other triangle is ignored, as selfadjointView does
      row_ents[il].push_back({c, k});              // the stored entry (i, c)
      if (c >= lo && c < hi) row_ents[c - lo].push_back({i, k});   // its mirror (c, i) lives here too
      else exports[owner(c)].push_back({c, i, k});                  // ... or on the rank that owns row c
    }
  }
  p.input_nnz = span;
  // ---- who exports how much to whom ----
  std::vector<int64_t> mine(W, 0);
  for (int d = 0; d < W; ++d) mine[d] = static_cast<int64_t>(exports[d].size());
  p.tri_counts.assign(static_cast<size_t>(W) * W, 0);
  if (cfg.allgather(cfg.allgather_ctx, mine.data(), p.tri_counts.data(), sizeof(int64_t) * W)) {
    err = "allgather(mirror counts) failed"; return B200S_ERR_COMM;
  }
  int64_t max_total = 1;
  for (int q = 0; q < W; ++q) {
    int64_t t = 0;
    for (int d = 0; d < W; ++d) t += p.tri_counts[static_cast<size_t>(q) * W + d];
    max_total = std::max(max_total, t);
  }
  // ---- the (row, col) of every exported mirror, grouped by destination ----
  std::vector<int64_t> send(static_cast<size_t>(max_total) * 2, -1), recv(static_cast<size_t>(max_total) * 2 * W);
  p.export_src.clear();
  {
    size_t o = 0;
    for (int d = 0; d < W; ++d)
      for (const Exp& e : exports[d]) {
        send[2 * o] = e.row;
        send[2 * o + 1] = e.col;
        p.export_src.push_back(e.src);
        ++o;
      }
  }
  if (cfg.allgather(cfg.allgather_ctx, send.data(), recv.data(), sizeof(int64_t) * send.size())) {
    err = "allgather(mirror entries) failed"; return B200S_ERR_COMM;
  }
  p.n_import = 0;
  for (int q = 0; q < W; ++q) {
    int64_t off = 0;
    for (int d = 0; d < p.rank; ++d) off += p.tri_counts[static_cast<size_t>(q) * W + d];
    const int64_t cnt = p.tri_counts[static_cast<size_t>(q) * W + p.rank];
    const int64_t* lst = recv.data() + static_cast<size_t>(q) * max_total * 2;
    for (int64_t j = 0; j < cnt; ++j) {
      const int64_t r = lst[2 * (off + j)], c = lst[2 * (off + j) + 1];
      if (r < lo || r >= hi || c < 0 || c >= p.cols) { err = "mirror exchange inconsistent across ranks"; return B200S_ERR_COMM; }
      row_ents[r - lo].push_back({c, span + p.n_import});
      ++p.n_import;
    }
  }
  // ---- assemble: rows sorted by column (stable, so a sorted input stays sorted) ----
  int64_t total = 0;
  for (int64_t il = 0; il < rows; ++il) total += static_cast<int64_t>(row_ents[il].size());
  if (total >= (int64_t(1) << 31) - 64 || span + p.n_import >= (int64_t(1) << 31) - 64) {
    err = "expanded nnz does not fit int32"; return B200S_ERR_UNSUPPORTED;
  }
  p.rowptr.assign(rows + 1, 0);
  p.colidx.resize(total);
  p.src.resize(total);
  int64_t o = 0;
  for (int64_t il = 0; il < rows; ++il) {
    p.rowptr[il] = static_cast<int32_t>(o);
    auto& v = row_ents[il];
    std::stable_sort(v.begin(), v.end(), [](const Ent& a, const Ent& b) { return a.col < b.col; });
    for (const Ent& e : v) {
      p.colidx[o] = static_cast<int32_t>(e.col);
      p.src[o] = static_cast<int32_t>(e.src);
      ++o;
    }
  }
  p.rowptr[rows] = static_cast<int32_t>(o);
  return 0;
}

}  // namespace

static void build_tiles(Plan& p, const std::vector<uint8_t>& row_is_boundary) {
  const int64_t rows = p.rows;
  const int32_t* rp = p.rowptr.data();
  std::vector<Tile> interior, boundary;
  int64_t r = 0;
  while (r < rows) {
    int32_t len0 = rp[r + 1] - rp[r];
    Tile t{};
    bool is_b = false;
    if (len0 > p.tile_nnz) {  // a row longer than a stage: streamed straight from global memory by the whole CTA
      t.row0 = static_cast<int32_t>(r);
      t.nnz0 = rp[r];
      t.nnz = len0;
      is_b = !row_is_boundary.empty() && row_is_boundary[r];
      t.meta = 1 | (0 << 16) | ((kTileLong | (is_b ? kTileBoundary : 0)) << 24);
      p.n_long++;
      ++r;
    } else {
      int64_t start = r;
      int32_t nnz = 0, maxlen = 0;
      while (r < rows && (r - start) < p.tile_rows_cap) {
        int32_t len = rp[r + 1] - rp[r];
        if (nnz + len > p.tile_nnz) break;
        nnz += len;
        maxlen = std::max(maxlen, len);
        if (!row_is_boundary.empty() && row_is_boundary[r]) is_b = true;
        ++r;
      }
      int nrows = static_cast<int>(r - start);
      int mean = nrows ? (nnz + nrows - 1) / nrows : 0;
      // lanes per row: as many as keep every row of the tile busy in one sweep of the CTA, but no more than half
      // the row length that matters -- the mean, or a quarter of the longest row when the tile is imbalanced
      // (short rows are cheapest with one thread each; a long row among short ones needs lanes to keep the sweep
      // from waiting on it).
      static const int stream_factor = [] { const char* e = std::getenv("B200S_STREAM_FACTOR"); return e ? std::atoi(e) : 4; }();
      int eff = std::max(mean, maxlen / 4);
      int lg_fit = std::min(5, floor_log2(std::max(1, kSpmvThreads / std::max(1, nrows))));
      int lg_len = std::min(5, ceil_log2(std::max(1, eff / 2)));
      int lg = std::min(lg_fit, lg_len);
      int flags = is_b ? kTileBoundary : 0;
      // Two-phase (CSR-stream) tiles: products balanced over all threads first, then summed per row.  Measured
      // (profiles/r1_spmv_sweep.jsonl and the round-1 experiments in DESIGN.md): on balanced tiles it loses to row
      // lanes (27-point 0.53 vs 0.81 of the copy bandwidth, so it is never forced there), on strongly imbalanced
      // tiles (power-law rows, maxlen > 4*mean + 32) it wins by 10-45 %.  B200S_STREAM_FACTOR changes the 4; 0 = off.
      if (stream_factor > 0 && maxlen > stream_factor * mean + 32) {
        flags |= kTileStream;
        lg = lg_fit;
        p.n_stream++;
      } else {
        p.by_lanes[lg]++;
      }
      t.row0 = static_cast<int32_t>(start);
      t.nnz0 = rp[start];
      t.nnz = nnz;
      t.meta = nrows | (lg << 16) | (flags << 24);
    }
    (is_b ? boundary : interior).push_back(t);
  }
  p.n_boundary_tiles = static_cast<int32_t>(boundary.size());
  p.tiles = std::move(interior);
  p.tiles.insert(p.tiles.end(), boundary.begin(), boundary.end());
}

int build_plan(const b200s_config& cfg, int64_t rows, int64_t cols, int64_t nnz, const int32_t* rowptr,
               const int32_t* colidx, const int32_t* inner_nnz, int uplo, const int64_t* row_starts, Plan& p,
               std::string& err) {
  p = Plan();
  p.world = cfg.world > 0 ? cfg.world : 1;
  p.rank = cfg.rank;
  if (rows < 0 || cols < 0 || nnz < 0 || (rows > 0 && (!rowptr || (nnz > 0 && !colidx)))) {
    err = "analyze_pattern: null or negative argument";
    return B200S_ERR_INVALID;
  }
  if (uplo != B200S_LOWER && uplo != B200S_UPPER && uplo != B200S_BOTH) {
    err = "analyze_pattern: uplo must be 1 (Lower), 2 (Upper) or 3 (Lower|Upper)";
    return B200S_ERR_INVALID;
  }
  if (p.rank < 0 || p.rank >= p.world) { err = "rank outside [0, world)"; return B200S_ERR_INVALID; }
  if (rows >= (int64_t(1) << 31) - 64 || cols >= (int64_t(1) << 31) - 64 || nnz >= (int64_t(1) << 31) - 64) {
    err = "sizes must fit int32 (StorageIndex = int)";
    return B200S_ERR_UNSUPPORTED;
  }
  p.tile_nnz = cfg.tile_nnz > 0 ? cfg.tile_nnz : kDefaultTileNnz;
  p.tile_rows_cap = cfg.tile_rows > 0 ? cfg.tile_rows : kDefaultTileRows;
  p.tile_nnz = std::max(64, std::min(p.tile_nnz, 8192)) & ~7;
  p.tile_rows_cap = std::max(8, std::min(p.tile_rows_cap, 1024)) & ~7;
  p.rows = rows;
  p.cols = cols;
  p.input_nnz = nnz;
  if (p.world == 1) {
    // rectangular matrices are fine for the product (SparseDenseProduct.h:26-72); the solvers insist on square ones
    if (rows != cols && uplo != B200S_BOTH) { err = "a self-adjoint view needs a square matrix"; return B200S_ERR_INVALID; }
    p.row_starts = {0, rows};
  } else {
    if (!row_starts) { err = "row_starts is required when world > 1"; return B200S_ERR_INVALID; }
    if (!cfg.allgather) { err = "config.allgather is required when world > 1"; return B200S_ERR_COMM; }
    p.row_starts.assign(row_starts, row_starts + p.world + 1);
    if (p.row_starts[0] != 0 || p.row_starts[p.world] != cols) { err = "row_starts must span [0, cols]"; return B200S_ERR_INVALID; }
    for (int q = 0; q < p.world; ++q)
      if (p.row_starts[q] > p.row_starts[q + 1]) { err = "row_starts must be non-decreasing"; return B200S_ERR_INVALID; }
    if (p.row_starts[p.rank + 1] - p.row_starts[p.rank] != rows) { err = "rows does not match row_starts[rank]"; return B200S_ERR_INVALID; }
  }
  p.row0 = p.row_starts[p.rank];

  int rc = (p.world > 1 && uplo != B200S_BOTH)
               ? canonicalise_distributed(cfg, rows, rowptr, colidx, inner_nnz, uplo, p, err)
               : canonicalise(rows, nnz, rowptr, colidx, inner_nnz, uplo, p, err);
  if (rc) return rc;
  p.nnz = p.rowptr[rows];
  const bool own_cols = !p.colidx.empty() || p.nnz == 0;
  const int32_t* cin = own_cols ? p.colidx.data() : colidx;
  for (int64_t i = 0; i < rows; ++i)
    if (p.rowptr[i + 1] < p.rowptr[i]) { err = "rowptr is not monotone"; return B200S_ERR_INVALID; }
  for (int64_t k = 0; k < p.nnz; ++k)
    if (cin[k] < 0 || cin[k] >= cols) { err = "column index out of range"; return B200S_ERR_INVALID; }

  std::vector<uint8_t> row_is_boundary;
  p.recv_counts.assign(p.world, 0);
  p.recv_offsets.assign(p.world, 0);
  p.send_counts.assign(p.world, 0);
  p.send_offsets.assign(p.world, 0);
  p.send_slot0.assign(p.world, 0);
  if (p.world == 1) {
    if (!own_cols) p.alias_colidx = colidx;
  } else {
    // ---- ghosts: columns outside the owned range, sorted and unique ----
    const int64_t lo = p.row0, hi = p.row0 + rows;
    std::vector<int64_t> ext;
    for (int64_t k = 0; k < p.nnz; ++k)
      if (cin[k] < lo || cin[k] >= hi) ext.push_back(cin[k]);
    std::sort(ext.begin(), ext.end());
    ext.erase(std::unique(ext.begin(), ext.end()), ext.end());
    p.ghost_cols = std::move(ext);
    if (rows + static_cast<int64_t>(p.ghost_cols.size()) >= (int64_t(1) << 31) - 64) {
      err = "rows + ghosts does not fit int32"; return B200S_ERR_UNSUPPORTED;
    }
    // ---- remap columns: owned c -> c - row0, ghost g -> rows + g ----
    if (!own_cols) p.colidx.assign(colidx, colidx + p.nnz);
    row_is_boundary.assign(rows, 0);
    for (int64_t i = 0; i < rows; ++i)
      for (int32_t k = p.rowptr[i]; k < p.rowptr[i + 1]; ++k) {
        int64_t c = p.colidx[k];
        if (c >= lo && c < hi) {
          p.colidx[k] = static_cast<int32_t>(c - lo);
        } else {
          auto it = std::lower_bound(p.ghost_cols.begin(), p.ghost_cols.end(), c);
          p.colidx[k] = static_cast<int32_t>(rows + (it - p.ghost_cols.begin()));
          row_is_boundary[i] = 1;
        }
      }
    // ---- owners of my ghosts (contiguous ranges because ghost_cols is sorted) ----
    {
      int q = 0;
      for (size_t g = 0; g < p.ghost_cols.size(); ++g) {
        while (p.ghost_cols[g] >= p.row_starts[q + 1]) ++q;
        p.recv_counts[q]++;
      }
      int64_t off = 0;
      for (int q2 = 0; q2 < p.world; ++q2) { p.recv_offsets[q2] = off; off += p.recv_counts[q2]; }
    }
    // ---- exchange: everyone learns everyone's ghost lists; I extract what I must send to each peer ----
    const int W = p.world;
    std::vector<int64_t> all_counts(static_cast<size_t>(W) * W);
    if (cfg.allgather(cfg.allgather_ctx, p.recv_counts.data(), all_counts.data(), sizeof(int64_t) * W)) {
      err = "allgather(recv_counts) failed"; return B200S_ERR_COMM;
    }
    int64_t max_ghosts = 0;
    std::vector<int64_t> nghost(W, 0);
    for (int q = 0; q < W; ++q) {
      for (int o = 0; o < W; ++o) nghost[q] += all_counts[static_cast<size_t>(q) * W + o];
      max_ghosts = std::max(max_ghosts, nghost[q]);
    }
    std::vector<int64_t> mine(std::max<int64_t>(max_ghosts, 1), -1), all(static_cast<size_t>(W) * mine.size());
    std::copy(p.ghost_cols.begin(), p.ghost_cols.end(), mine.begin());
    if (cfg.allgather(cfg.allgather_ctx, mine.data(), all.data(), sizeof(int64_t) * mine.size())) {
      err = "allgather(ghost lists) failed"; return B200S_ERR_COMM;
    }
    int64_t soff = 0;
    for (int q = 0; q < W; ++q) {
      p.send_counts[q] = all_counts[static_cast<size_t>(q) * W + p.rank];
      p.send_offsets[q] = soff;
      int64_t slot0 = 0;
      for (int o = 0; o < p.rank; ++o) slot0 += all_counts[static_cast<size_t>(q) * W + o];
      p.send_slot0[q] = slot0;
      const int64_t* ql = all.data() + static_cast<size_t>(q) * mine.size();
      for (int64_t k = 0; k < p.send_counts[q]; ++k) {
        int64_t c = ql[slot0 + k];
        if (c < lo || c >= hi) { err = "halo plan inconsistent across ranks"; return B200S_ERR_COMM; }
        p.send_rows.push_back(static_cast<int32_t>(c - lo));
      }
      soff += p.send_counts[q];
    }
  }
  build_tiles(p, row_is_boundary);
  // Strongly irregular matrices (a quarter or more of the tiles two-phase) are bound by the x gathers, not by the
  // matrix stream: half-size tiles let more CTAs be resident per SM, which keeps the gather pipe busier while other
  // CTAs sit in their barriers (measured on power-law rows, profiles/r2_exp_powerlaw.txt: 0.225 -> 0.26 of the HBM
  // figure).  Only when the caller did not fix the geometry.
  static const int auto_small = [] { const char* e = std::getenv("B200S_AUTO_SMALL_TILES"); return e ? std::atoi(e) : 1; }();
  if (auto_small && cfg.tile_nnz <= 0 && p.tile_nnz == kDefaultTileNnz && !p.tiles.empty() &&
      static_cast<size_t>(p.n_stream) * 4 >= p.tiles.size()) {
    p.tile_nnz = kDefaultTileNnz / 2;
    p.n_stream = p.n_long = p.n_boundary_tiles = 0;
    for (int& b : p.by_lanes) b = 0;
    p.tiles.clear();
    build_tiles(p, row_is_boundary);
  }
  return 0;
}

int exchange_mirror_values(const b200s_config& cfg, const Plan& p, const void* values, size_t elem,
                           std::vector<unsigned char>& imports, std::string& err) {
  imports.clear();
  if (p.tri_counts.empty()) return 0;
  const int W = p.world;
  int64_t max_total = 1;
  for (int q = 0; q < W; ++q) {
    int64_t t = 0;
    for (int d = 0; d < W; ++d) t += p.tri_counts[static_cast<size_t>(q) * W + d];
    max_total = std::max(max_total, t);
  }
  // my exported values in export order (grouped by destination), padded to the longest list of any rank
  std::vector<unsigned char> send(static_cast<size_t>(max_total) * elem, 0), recv(send.size() * W);
  const unsigned char* v = static_cast<const unsigned char*>(values);
  for (size_t o = 0; o < p.export_src.size(); ++o)
    std::memcpy(send.data() + o * elem, v + static_cast<size_t>(p.export_src[o]) * elem, elem);
  if (!cfg.allgather || cfg.allgather(cfg.allgather_ctx, send.data(), recv.data(), send.size())) {
    err = "allgather(mirror values) failed";
    return B200S_ERR_COMM;
  }
  imports.resize(static_cast<size_t>(p.n_import) * elem);
  size_t o = 0;
  for (int q = 0; q < W; ++q) {  // same order as canonicalise_distributed numbered the imports
    int64_t off = 0;
    for (int d = 0; d < p.rank; ++d) off += p.tri_counts[static_cast<size_t>(q) * W + d];
    const int64_t cnt = p.tri_counts[static_cast<size_t>(q) * W + p.rank];
    if (cnt) std::memcpy(imports.data() + o, recv.data() + (static_cast<size_t>(q) * max_total + off) * elem, static_cast<size_t>(cnt) * elem);
    o += static_cast<size_t>(cnt) * elem;
  }
  if (o != imports.size()) { err = "mirror value exchange inconsistent with the plan"; return B200S_ERR_COMM; }
  return 0;
}

void fill_tile_stats(const Plan& p, b200s_stats* st) {
  if (!st) return;
  st->world = p.world;
  st->rank = p.rank;
  st->rows = p.rows;
  st->cols = p.cols;
  st->nnz = p.nnz;
  st->ghosts = static_cast<int64_t>(p.ghost_cols.size());
  st->halo_send = static_cast<int64_t>(p.send_rows.size());
  st->tiles = static_cast<int32_t>(p.tiles.size());
  st->tiles_boundary = p.n_boundary_tiles;
  for (int i = 0; i < 6; ++i) st->tiles_by_lanes[i] = p.by_lanes[i];
  st->tiles_stream = p.n_stream;
  st->tiles_long = p.n_long;
}

}  // namespace b200s
