// plan.h -- host-side (GPU-free) planning for the B200 sparse iterative-solve path.
//
// Turns the CSR arrays an Eigen solver holds after IterativeSolverBase::compute()
// (Ref<const SparseMatrix>: outerIndexPtr / innerIndexPtr / innerNonZeroPtr, IterativeSolverBase.h:77-91) into
//   1. a compressed, optionally symmetric-expanded local CSR block (UpLo handling of ConjugateGradient.h:202-213),
//   2. the row-block partition's halo plan (ghost columns, per-peer send lists),
//   3. the tile list the staged SpMV kernel walks, with a per-tile lanes-per-row choice taken from the tile's
//      nnz-per-row statistics.
// No CUDA call is made here, so the logic is testable on a CPU-only box (tests/test_plan.py, tests/test_dist_gloo.py).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/b200sparse.h"

namespace b200s {

#ifndef B200S_SPMV_THREADS
#define B200S_SPMV_THREADS 256
#endif
constexpr int kSpmvThreads = B200S_SPMV_THREADS;      // CTA size of the staged SpMV kernel
constexpr int kDefaultTileNnz = 8 * kSpmvThreads;     // shared-memory tile: non-zeros
constexpr int kDefaultTileRows = kSpmvThreads;        // shared-memory tile: rows (one thread per row at 1 lane/row)

// One unit of work of the staged SpMV kernel: a run of consecutive rows whose non-zeros fit one smem stage.
struct Tile {
  int32_t row0;  // first local row
  int32_t nnz0;  // rowptr[row0]
  int32_t meta;  // nrows | lanes_log2 << 16 | flags << 24
  int32_t nnz;   // non-zeros in the tile
};
enum : int32_t { kTileStream = 1, kTileLong = 2, kTileBoundary = 4 };
inline int tile_rows(const Tile& t) { return t.meta & 0xFFFF; }
inline int tile_lg(const Tile& t) { return (t.meta >> 16) & 0xFF; }
inline int tile_flags(const Tile& t) { return (t.meta >> 24) & 0xFF; }

struct Plan {
  int world = 1, rank = 0;
  int64_t rows = 0, cols = 0, nnz = 0, row0 = 0;
  std::vector<int64_t> row_starts;  // world+1
  // Local CSR.  rowptr is always owned; colidx is owned (remapped) unless `alias_colidx` is set, in which case the
  // caller's array is already local (world == 1, compressed, uplo == BOTH) and is uploaded directly.
  std::vector<int32_t> rowptr;
  std::vector<int32_t> colidx;
  const int32_t* alias_colidx = nullptr;
  // src[k] = index into the caller's value array of local entry k; empty = identity.
  std::vector<int32_t> src;
  int64_t input_nnz = 0;  // slots of the caller's value array that are read: one past the last referenced slot
  // One stored triangle on a row-partitioned matrix (uplo != BOTH, world > 1): the mirror image (c, i) of a stored entry
  // (i, c) belongs to the rank that owns row c.  export_src lists, grouped by destination rank, the value slots whose
  // mirrors other ranks need; tri_counts[q*world + d] = how many rank q exports to rank d (known to every rank);
  // the n_import values this rank receives (source rank ascending, then the source's export order) are staged behind
  // the caller's input_nnz values, and `src` refers to them as input_nnz + j.  factorize() exchanges the values.
  std::vector<int32_t> export_src;
  std::vector<int64_t> tri_counts;
  int64_t n_import = 0;
  // Halo plan.
  std::vector<int64_t> ghost_cols;     // sorted global ids; local column of ghost g is rows + g
  std::vector<int64_t> recv_counts;    // [world] ghosts owned by each peer
  std::vector<int64_t> recv_offsets;   // [world] first ghost slot of each peer's range
  std::vector<int64_t> send_counts;    // [world]
  std::vector<int64_t> send_offsets;   // [world] offset into send_rows
  std::vector<int64_t> send_slot0;     // [world] first ghost slot, on peer q, of the entries this rank sends to q
  std::vector<int32_t> send_rows;      // local rows to send, grouped by destination
  // Tiles: interior tiles first, then tiles touching ghosts.
  std::vector<Tile> tiles;
  int tile_nnz = kDefaultTileNnz, tile_rows_cap = kDefaultTileRows;
  int32_t n_boundary_tiles = 0, n_stream = 0, n_long = 0;
  int32_t by_lanes[6] = {0, 0, 0, 0, 0, 0};

  const int32_t* colidx_ptr() const { return alias_colidx ? alias_colidx : colidx.data(); }
};

// Returns 0 or a negative b200s_status; `err` receives the message.
int build_plan(const b200s_config& cfg, int64_t rows, int64_t cols, int64_t nnz, const int32_t* rowptr,
               const int32_t* colidx, const int32_t* inner_nnz, int uplo, const int64_t* row_starts, Plan& plan,
               std::string& err);

// Row-partitioned one-triangle input (plan.tri_counts non-empty): collects, through the config's allgather, the values
// of the mirror entries other ranks store for this rank's rows.  `values` is the caller's array (elements of `elem`
// bytes); `imports` receives plan.n_import elements, to be staged right behind the caller's plan.input_nnz values.
// Every rank must call it (it is a collective).  Returns 0 or a negative b200s_status.
int exchange_mirror_values(const b200s_config& cfg, const Plan& plan, const void* values, size_t elem,
                           std::vector<unsigned char>& imports, std::string& err);

void fill_tile_stats(const Plan& plan, b200s_stats* st);

}  // namespace b200s
