// tests/native/fuzz_factors.cpp -- TEST INFRASTRUCTURE (CPU): the host factorization / analysis code of the incomplete-
// factorization preconditioners (csrc/factors.cpp) under AddressSanitizer + UBSan on 300 random systems: symmetric and
// not, missing / negative / weak diagonals (shift retries, reported failures), random permutations, droptol / fillfactor
// variants, n = 0; every successful factor is pushed through the stage / level analysis as well.  Built and run by
// tests/test_native_fuzz.py.
#include "factors.h"
#include <cstdio>
#include <random>
#include <algorithm>
#include <numeric>
using namespace b200s;
int main() {
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> U(-1, 1);
  int bad = 0;
  for (int t = 0; t < 300; ++t) {
    int n = 1 + rng() % 150;
    double dens = 0.01 + 0.3 * (rng() % 100) / 100.0;
    std::vector<std::vector<std::pair<int,double>>> rows(n);
    bool sym = t % 2;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
      if (i == j) continue;
      if (sym && j > i) continue;
      if ((rng() % 10000) / 10000.0 < dens) { double v = U(rng); rows[i].push_back({j, v}); if (sym) rows[j].push_back({i, v}); }
    }
    for (int i = 0; i < n; ++i) { double s = 1e-3 + (t % 7 == 0 ? 0.0 : 1.0); for (auto& e : rows[i]) s += std::abs(e.second) * (t % 5 == 0 ? 0.3 : 1.0); if (!(t % 11 == 0 && i % 9 == 0)) rows[i].push_back({i, (t % 13 == 0 && i % 5 == 0) ? -s : s}); std::sort(rows[i].begin(), rows[i].end()); }
    std::vector<int32_t> rp(n + 1, 0), ci; std::vector<double> va;
    for (int i = 0; i < n; ++i) { for (auto& e : rows[i]) { ci.push_back(e.first); va.push_back(e.second); } rp[i + 1] = (int32_t)ci.size(); }
    std::vector<int32_t> perm(n); std::iota(perm.begin(), perm.end(), 0); std::shuffle(perm.begin(), perm.end(), rng);
    std::string err; Factors f;
    int rc = ilut_factorize(n, rp.data(), ci.data(), va.data(), t % 3 ? -1.0 : 0.05, t % 4 ? 0 : 1 + rng() % 20, t % 2 ? perm.data() : nullptr, f, err);
    if (rc == 0 && f.info == 0) { Factors g; rc = factors_from_ilut(n, f.outer.data(), f.inner.data(), f.vals.data(), f.perm.data(), g, err); if (rc) { ++bad; std::printf("from_ilut %d %s\n", t, err.c_str()); } }
    else if (rc) { /* missing diagonal etc. are reported, not crashes */ }
    for (int uplo = 1; uplo <= 2; ++uplo) {
      Factors h; rc = ichol_factorize(n, rp.data(), ci.data(), va.data(), uplo, -1.0, t % 3 ? perm.data() : nullptr, h, err);
      if (rc == 0 && h.info == 0) { Factors g; rc = factors_from_ichol(n, h.outer.data(), h.inner.data(), h.vals.data(), h.scale.data(), h.perm.empty() ? nullptr : h.perm.data(), g, err); if (rc) { ++bad; std::printf("from_ichol %d %s\n", t, err.c_str()); } }
    }
    std::vector<int32_t> mc(n); int nc = multicolor_ordering(n, rp.data(), ci.data(), mc.data(), err); if (nc < 0 || (n > 0 && nc == 0)) { ++bad; std::printf("mc %d\n", t); }
  }
  { Factors f; std::string err; int32_t rp0[1] = {0}; ilut_factorize(0, rp0, nullptr, nullptr, -1, 0, nullptr, f, err); ichol_factorize(0, rp0, nullptr, nullptr, 1, -1, nullptr, f, err); }
  std::printf("done, bad=%d\n", bad);
  return bad;
}
