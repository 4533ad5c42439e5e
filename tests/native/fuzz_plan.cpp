// tests/native/fuzz_plan.cpp -- TEST INFRASTRUCTURE (CPU): one stored triangle on a ROW-PARTITIONED matrix
// (csrc/plan.cpp: canonicalise_distributed, exchange_mirror_values) under AddressSanitizer + UBSan.  W = 2..4 "ranks" are
// threads of one process with a barrier-based all-gather; random symmetric matrices, uneven partitions incl. empty
// ranks, Lower, Upper and fully stored; every rank must end up with exactly its rows of the full matrix (pattern and values).
// Built and run by tests/test_native_fuzz.py.
#include "plan.h"
#include <pthread.h>
#include <cstdio>
#include <cstring>
#include <random>
#include <thread>
#include <algorithm>
using namespace b200s;
struct Comm { int W; pthread_barrier_t bar; std::vector<unsigned char> buf; size_t bytes; };
struct Ctx { Comm* c; int rank; };
static int ag(void* vctx, const void* send, void* recv, size_t bytes) {
  Ctx* x = (Ctx*)vctx; Comm* c = x->c;
  if (x->rank == 0) { c->buf.assign(bytes * c->W, 0); c->bytes = bytes; }
  pthread_barrier_wait(&c->bar);
  std::memcpy(c->buf.data() + bytes * x->rank, send, bytes);
  pthread_barrier_wait(&c->bar);
  std::memcpy(recv, c->buf.data(), bytes * c->W);
  pthread_barrier_wait(&c->bar);
  return 0;
}
int main() {
  std::mt19937_64 rng(11);
  std::uniform_real_distribution<double> U(-1, 1);
  int bad = 0;
  for (int t = 0; t < 60; ++t) {
    const int W = 2 + t % 3;
    const int n = 5 + rng() % 200;
    // symmetric matrix, dense-ish storage of the full pattern
    std::vector<std::vector<double>> M(n, std::vector<double>(n, 0.0));
    for (int i = 0; i < n; ++i) { M[i][i] = 4 + U(rng); for (int j = 0; j < i; ++j) if (rng() % 100 < 8) M[i][j] = M[j][i] = U(rng); }
    std::vector<int64_t> starts(W + 1, 0);
    for (int q = 1; q < W; ++q) starts[q] = std::min<int64_t>(n, starts[q - 1] + rng() % (2 * n / W + 1));
    starts[W] = n; std::sort(starts.begin(), starts.end());
    for (int uplo = 1; uplo <= 3; ++uplo) {  // 3 = both triangles stored: the plain row-block plan
      Comm comm; comm.W = W; pthread_barrier_init(&comm.bar, nullptr, W);
      std::vector<int> fails(W, 0);
      std::vector<std::thread> th;
      for (int r = 0; r < W; ++r) th.emplace_back([&, r] {
        Ctx ctx{&comm, r};
        b200s_config cfg; std::memset(&cfg, 0, sizeof(cfg)); cfg.struct_size = sizeof(cfg); cfg.rank = r; cfg.world = W; cfg.allgather = ag; cfg.allgather_ctx = &ctx;
        const int64_t lo = starts[r], hi = starts[r + 1], rows = hi - lo;
        std::vector<int32_t> rp(rows + 1, 0), ci; std::vector<double> va;
        for (int64_t i = lo; i < hi; ++i) { for (int j = 0; j < n; ++j) if (M[i][j] != 0 && (uplo == 3 || (uplo == 1 ? j <= i : j >= i))) { ci.push_back(j); va.push_back(M[i][j]); } rp[i - lo + 1] = (int32_t)ci.size(); }
        Plan p; std::string err;
        int rc = build_plan(cfg, rows, n, (int64_t)ci.size(), rp.data(), ci.data(), nullptr, uplo, starts.data(), p, err);
        std::vector<unsigned char> imp;
        if (!rc) rc = exchange_mirror_values(cfg, p, va.data(), sizeof(double), imp, err);
        if (rc) { fails[r] = 1; std::printf("rank %d rc %d %s\n", r, rc, err.c_str()); return; }
        const double* im = (const double*)imp.data();
        for (int64_t il = 0; il < rows; ++il) {
          std::vector<double> row(n, 0.0);
          for (int32_t k = p.rowptr[il]; k < p.rowptr[il + 1]; ++k) {
            int32_t lc = p.colidx_ptr()[k]; int64_t gc = lc < rows ? lo + lc : p.ghost_cols[lc - rows];
            int64_t s = p.src.empty() ? k : p.src[k]; row[gc] += s < p.input_nnz ? va[s] : im[s - p.input_nnz];
          }
          for (int j = 0; j < n; ++j) if (row[j] != M[lo + il][j]) { fails[r] = 1; }
        }
      });
      for (auto& x : th) x.join();
      pthread_barrier_destroy(&comm.bar);
      for (int r = 0; r < W; ++r) bad += fails[r];
    }
  }
  std::printf("done, bad=%d\n", bad);
  return bad;
}
