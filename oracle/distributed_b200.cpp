// oracle/distributed_b200.cpp -- TEST INFRASTRUCTURE (built here, executed on a GPU box with >= 2 B200s).
//
// The row-partitioned path driven from C++ only: one process per GPU (forked before any CUDA call), the b200 binding
// classes configured with setDistributed(), a host all-gather over a shared-memory segment standing in for
// MPI_Allgather.  Every rank builds ITS row block of a 3-D Poisson / convection-diffusion matrix with global column
// indices, solves on the GPUs, and compares its segment of x with the reference's own CPU solver
// (Eigen::ConjugateGradient / Eigen::BiCGSTAB on the whole matrix) -- the reference's parity bar: iteration count
// within 2 %, x within 1e-8 norm-wise.  Also the distributed SpMV drop-in (b200::SparseOperator).
//   usage: distributed_b200 [world=2] [n=24]
#include <Eigen/IterativeLinearSolvers>
#include <Eigen/Sparse>
#include <b200/IterativeSolvers.h>
#include <b200/SparseOperator.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace Eigen;
typedef SparseMatrix<double, RowMajor> Csr;

struct Shared {  // lives in a MAP_SHARED segment
  std::atomic<int> arrived;
  std::atomic<int> generation;
  int world;
  size_t slot_bytes;
  unsigned char data[1];
};
struct Ctx {
  Shared* sh;
  int rank;
};

static void barrier(Shared* sh) {
  const int gen = sh->generation.load();
  if (sh->arrived.fetch_add(1) == sh->world - 1) {
    sh->arrived.store(0);
    sh->generation.fetch_add(1);
  } else {
    while (sh->generation.load() == gen) usleep(50);
  }
}
static int allgather(void* vctx, const void* send, void* recv, size_t bytes) {
  Ctx* c = static_cast<Ctx*>(vctx);
  if (bytes > c->sh->slot_bytes) return 1;
  std::memcpy(c->sh->data + c->rank * c->sh->slot_bytes, send, bytes);
  barrier(c->sh);
  for (int q = 0; q < c->sh->world; ++q)
    std::memcpy(static_cast<unsigned char*>(recv) + q * bytes, c->sh->data + q * c->sh->slot_bytes, bytes);
  barrier(c->sh);
  return 0;
}

static Csr stencil(int n, double lo, double up, int r0, int r1) {  // rows [r0, r1) of the n^3 operator, global columns
  const int N = n * n * n;
  std::vector<Triplet<double> > t;
  for (int r = r0; r < r1; ++r) {
    const int i = r % n, j = (r / n) % n, k = r / (n * n);
    if (k > 0) t.push_back(Triplet<double>(r - r0, r - n * n, lo));
    if (j > 0) t.push_back(Triplet<double>(r - r0, r - n, lo));
    if (i > 0) t.push_back(Triplet<double>(r - r0, r - 1, lo));
    t.push_back(Triplet<double>(r - r0, r, 6.0));
    if (i < n - 1) t.push_back(Triplet<double>(r - r0, r + 1, up));
    if (j < n - 1) t.push_back(Triplet<double>(r - r0, r + n, up));
    if (k < n - 1) t.push_back(Triplet<double>(r - r0, r + n * n, up));
  }
  Csr A(r1 - r0, N);
  A.setFromTriplets(t.begin(), t.end());
  A.makeCompressed();
  return A;
}

#define CHECK(cond)                                                                   \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      std::fprintf(stderr, "rank %d: CHECK failed: %s (line %d)\n", rank, #cond, __LINE__); \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

static int run_rank(int rank, int world, int n, Shared* sh) {
  Ctx ctx = {sh, rank};
  const int N = n * n * n, planes = n / world;
  std::vector<int64_t> starts(world + 1);
  for (int q = 0; q < world; ++q) starts[q] = static_cast<int64_t>(q) * planes * n * n;
  starts[world] = N;
  const int r0 = static_cast<int>(starts[rank]), r1 = static_cast<int>(starts[rank + 1]);
  srand(12345);
  VectorXd x_true = VectorXd::Random(N);

  // ---- CG on the Poisson operator ----
  {
    Csr full = stencil(n, -1.0, -1.0, 0, N), mine = stencil(n, -1.0, -1.0, r0, r1);
    VectorXd b = full * x_true;
    Eigen::ConjugateGradient<Csr, Lower | Upper> ref(full);
    ref.setTolerance(1e-10);
    VectorXd xr = ref.solve(b);
    b200::ConjugateGradient<Csr, Lower | Upper> cg;
    cg.setDistributed(rank, world, starts.data(), allgather, &ctx, rank);
    cg.compute(mine);
    CHECK(cg.info() == Success);
    cg.setTolerance(1e-10);
    VectorXd x = cg.solve(b.segment(r0, r1 - r0));
    CHECK(cg.info() == Success && x.size() == N);
    CHECK(std::abs(double(cg.iterations() - ref.iterations())) <= std::max(1.0, 0.02 * ref.iterations()));
    CHECK((x.segment(r0, r1 - r0) - xr.segment(r0, r1 - r0)).norm() <= 1e-8 * xr.norm());
    CHECK(x.head(r0).isZero(0) && x.tail(N - r1).isZero(0));
    std::printf("rank %d: distributed CG ok, iterations %d (reference %d), error %.3e\n", rank, int(cg.iterations()),
                int(ref.iterations()), cg.error());
    // the SpMV drop-in on the same partition
    b200::SparseOperator<double> op;
    op.setDistributed(rank, world, starts.data(), allgather, &ctx, rank);
    op.compute(mine);
    CHECK(op.info() == Success);
    VectorXd y(r1 - r0);
    VectorXd xl = x_true.segment(r0, r1 - r0);
    CHECK(op.multiply(xl.data(), y.data()));
    CHECK((y - b.segment(r0, r1 - r0)).norm() <= 1e-13 * b.norm());
    std::printf("rank %d: distributed SpMV ok\n", rank);
  }
  // ---- BiCGSTAB on convection-diffusion ----
  {
    Csr full = stencil(n, -1.5, -0.5, 0, N), mine = stencil(n, -1.5, -0.5, r0, r1);
    VectorXd b = full * x_true;
    Eigen::BiCGSTAB<Csr> ref(full);
    ref.setTolerance(1e-10);
    VectorXd xr = ref.solve(b);
    b200::BiCGSTAB<Csr> s;
    s.setDistributed(rank, world, starts.data(), allgather, &ctx, rank);
    s.compute(mine);
    s.setTolerance(1e-10);
    VectorXd x = s.solve(b.segment(r0, r1 - r0));
    CHECK(s.info() == Success);
    CHECK(std::abs(double(s.iterations() - ref.iterations())) <= std::max(2.0, 0.05 * ref.iterations()));
    CHECK((x.segment(r0, r1 - r0) - xr.segment(r0, r1 - r0)).norm() <= 1e-8 * xr.norm());
    std::printf("rank %d: distributed BiCGSTAB ok, iterations %d (reference %d)\n", rank, int(s.iterations()),
                int(ref.iterations()));
  }
  return 0;
}

int main(int argc, char** argv) {
  const int world = argc > 1 ? std::atoi(argv[1]) : 2;
  const int n = argc > 2 ? std::atoi(argv[2]) : 24;
  if (world < 1 || world > 8 || n % world != 0) {
    std::fprintf(stderr, "usage: distributed_b200 [world (divides n)] [n]\n");
    return 2;
  }
  const size_t slot = 1 << 22;
  void* mem = mmap(0, sizeof(Shared) + slot * world, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (mem == MAP_FAILED) return 3;
  Shared* sh = new (mem) Shared;
  sh->arrived.store(0);
  sh->generation.store(0);
  sh->world = world;
  sh->slot_bytes = slot;
  std::vector<pid_t> kids;
  for (int r = 0; r < world; ++r) {  // fork BEFORE any CUDA call: each child owns one GPU
    pid_t pid = fork();
    if (pid == 0) _exit(run_rank(r, world, n, sh));
    kids.push_back(pid);
  }
  int bad = 0;
  for (pid_t pid : kids) {
    int st = 0;
    waitpid(pid, &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) ++bad;
  }
  std::printf(bad ? "distributed_b200: %d rank(s) FAILED\n" : "distributed_b200: all %d ranks ok\n", bad ? bad : world);
  return bad ? 1 : 0;
}
