// factors.cpp -- host-side factorization / analysis for the ILUT and incomplete-Cholesky preconditioners (see factors.h).
// Pure C++17, no CUDA.
#include "factors.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>

// The reference is normally compiled with FMA contraction (-O2/-O3 -march with FMA: `a -= b * c` becomes one fused
// operation); the factors are pinned against such a build of the unmodified reference.  This file is compiled by nvcc's host pass
// without -mfma, so every multiply-add site says which of the two it is (determined by probing that build, see
// tests/test_factors.py).
#define B200S_FUSED(a, b, c) std::fma((a), (b), (c))
#define B200S_UNFUSED(a, b, c) ((a) * (b) + (c))
#ifndef B200S_FMA_ILUT_NORM
#define B200S_FMA_ILUT_NORM B200S_FUSED
#endif
#ifndef B200S_FMA_ILUT_ELIM
#define B200S_FMA_ILUT_ELIM B200S_UNFUSED  // the product has two uses in the reference (fill-in or update): not contracted
#endif
#ifndef B200S_FMA_IC_SCALE
#define B200S_FMA_IC_SCALE B200S_FUSED
#endif
#ifndef B200S_FMA_IC_UPDATE
#define B200S_FMA_IC_UPDATE B200S_UNFUSED  // likewise (new entry or update share one product)
#endif
#ifndef B200S_FMA_IC_DIAG
#define B200S_FMA_IC_DIAG B200S_FUSED
#endif

namespace b200s {

namespace {

// internal::QuickSplit (IncompleteLUT.h:29-63): permutes row[0..n) so that the ncut entries of largest magnitude come
// first.  The loop structure fixes WHICH permutation results, and the reference's factors are stored in that order.
void quick_split(double* row, int32_t* ind, int64_t n, int64_t ncut) {
  ncut--;
  int64_t first = 0, last = n - 1;
  if (ncut < first || ncut > last) return;
  int64_t mid;
  do {
    mid = first;
    const double abskey = std::abs(row[mid]);
    for (int64_t j = first + 1; j <= last; j++)
      if (std::abs(row[j]) > abskey) {
        ++mid;
        std::swap(row[mid], row[j]);
        std::swap(ind[mid], ind[j]);
      }
    std::swap(row[mid], row[first]);
    std::swap(ind[mid], ind[first]);
    if (mid > ncut) last = mid - 1;
    else if (mid < ncut) first = mid + 1;
  } while (mid != ncut);
}

bool is_permutation(const int32_t* perm, int64_t n) {
  std::vector<uint8_t> seen(static_cast<size_t>(n), 0);
  for (int64_t i = 0; i < n; ++i) {
    if (perm[i] < 0 || perm[i] >= n || seen[perm[i]]) return false;
    seen[perm[i]] = 1;
  }
  return true;
}

// Dependency levels of one stage and its launch plan.  forward: rows depend on smaller rows, else on larger ones.
int analyse_levels(TriStage& s, int64_t n, bool forward, std::string& err) {
  std::vector<int32_t> level(static_cast<size_t>(n), 0);
  int32_t nlev = 0;
  for (int64_t step = 0; step < n; ++step) {
    const int64_t i = forward ? step : n - 1 - step;
    int32_t l = 0;
    for (int32_t k = s.rowptr[i]; k < s.rowptr[i + 1]; ++k) {
      const int64_t j = s.colidx[k];
      if (j < 0 || j >= n || (forward ? j >= i : j <= i)) { err = "factor is not triangular"; return B200S_ERR_INVALID; }
      l = std::max(l, level[j] + 1);
    }
    level[i] = l;
    nlev = std::max(nlev, l + 1);
  }
  s.level_ptr.assign(static_cast<size_t>(nlev) + 1, 0);
  for (int64_t i = 0; i < n; ++i) s.level_ptr[level[i] + 1]++;
  for (int32_t l = 0; l < nlev; ++l) s.level_ptr[l + 1] += s.level_ptr[l];
  s.level_rows.resize(static_cast<size_t>(n));
  std::vector<int32_t> cur(s.level_ptr.begin(), s.level_ptr.end() - 1);
  for (int64_t i = 0; i < n; ++i) s.level_rows[cur[level[i]]++] = static_cast<int32_t>(i);
  // launches: a wide level is one grid; a run of narrow levels is ONE single-CTA launch (block barrier between levels)
  s.launches.clear();
  for (int32_t l = 0; l < nlev;) {
    const int32_t rows = s.level_ptr[l + 1] - s.level_ptr[l];
    if (rows > kTriFusedBlock) {
      s.launches.push_back({l, l + 1, rows});
      ++l;
      continue;
    }
    int32_t e = l, widest = 0;
    while (e < nlev && s.level_ptr[e + 1] - s.level_ptr[e] <= kTriFusedBlock) {
      widest = std::max(widest, s.level_ptr[e + 1] - s.level_ptr[e]);
      ++e;
    }
    s.launches.push_back({l, e, widest});
    l = e;
  }
  return 0;
}

// The two solves of IncompleteLUT::_solve_impl on the row-major factor m_lu (TriangularSolver.h:26-102):
//   UnitLower: row i sums its entries in storage order and stops at the first one with column >= i (:42-50);
//   Upper    : skips the leading entries with column < i, takes the next one as the diagonal, sums ALL the rest (:78-93).
int stages_from_lu(Factors& f, std::string& err) {
  const int64_t n = f.n;
  TriStage &lo = f.first, &up = f.second;
  lo.rowptr.assign(n + 1, 0);
  up.rowptr.assign(n + 1, 0);
  up.diag.assign(n, 0.0);
  lo.diag.clear();
  lo.colidx.clear(); lo.vals.clear(); up.colidx.clear(); up.vals.clear();
  for (int64_t i = 0; i < n; ++i) {
    int32_t k = f.outer[i];
    const int32_t e = f.outer[i + 1];
    for (; k < e && f.inner[k] < i; ++k) { lo.colidx.push_back(f.inner[k]); lo.vals.push_back(f.vals[k]); }
    if (k >= e || f.inner[k] != i) { err = "ILUT factor: a row has no diagonal entry right after its lower part"; return B200S_ERR_INVALID; }
    up.diag[i] = f.vals[k++];
    for (; k < e; ++k) { up.colidx.push_back(f.inner[k]); up.vals.push_back(f.vals[k]); }
    lo.rowptr[i + 1] = static_cast<int32_t>(lo.colidx.size());
    up.rowptr[i + 1] = static_cast<int32_t>(up.colidx.size());
  }
  lo.fused = up.fused = false;  // both are row-wise loops of TriangularSolver.h (:42-50, :78-93)
  int rc = analyse_levels(lo, n, true, err);
  if (!rc) rc = analyse_levels(up, n, false, err);
  if (rc) return rc;
  // x = Pinv b ... z = P x  (IncompleteLUT.h:172-175) as gathers: (Pinv b)[k] = b[P[k]], (P x)[k] = x[Pinv[k]]
  f.pre_gather.assign(f.perm.begin(), f.perm.end());
  f.post_gather.resize(n);
  for (int64_t i = 0; i < n; ++i) f.post_gather[f.perm[i]] = static_cast<int32_t>(i);
  f.pre_scale.clear();
  f.post_scale.clear();
  return 0;
}

// The two solves of IncompleteCholesky::_solve_impl on the column-major lower factor m_L (IncompleteCholesky.h:149-157):
//   L.triangularView<Lower>()        column sweep (TriangularSolver.h:104-134): x[i] is finished by dividing by the
//                                    first entry >= i of column i, then pushed into the later rows -- row i therefore
//                                    receives its updates in ascending column order: the row-wise sum used here;
//   L.adjoint().triangularView<Upper>()  row-major backward substitution over column i of m_L read as row i (:64-102).
int stages_from_l(Factors& f, std::string& err) {
  const int64_t n = f.n;
  TriStage &lo = f.first, &up = f.second;
  lo.diag.assign(n, 0.0);
  up.diag.assign(n, 0.0);
  up.rowptr.assign(n + 1, 0);
  up.colidx.clear(); up.vals.clear();
  std::vector<int32_t> cnt(static_cast<size_t>(n) + 1, 0);
  std::vector<int32_t> dpos(static_cast<size_t>(n), 0);
  for (int64_t j = 0; j < n; ++j) {
    int32_t k = f.outer[j];
    const int32_t e = f.outer[j + 1];
    while (k < e && f.inner[k] < j) ++k;
    if (k >= e || f.inner[k] != j) { err = "incomplete Cholesky factor: a column has no diagonal entry"; return B200S_ERR_INVALID; }
    dpos[j] = k;
    lo.diag[j] = up.diag[j] = f.vals[k];
    for (++k; k < e; ++k) {
      if (f.inner[k] <= j || f.inner[k] >= n) { err = "incomplete Cholesky factor is not lower triangular"; return B200S_ERR_INVALID; }
      cnt[f.inner[k] + 1]++;
      up.colidx.push_back(f.inner[k]);
      up.vals.push_back(f.vals[k]);
    }
    up.rowptr[j + 1] = static_cast<int32_t>(up.colidx.size());
  }
  lo.rowptr.assign(n + 1, 0);
  for (int64_t i = 0; i < n; ++i) lo.rowptr[i + 1] = lo.rowptr[i] + cnt[i + 1];
  lo.colidx.resize(up.colidx.size());
  lo.vals.resize(up.vals.size());
  std::vector<int32_t> cur(lo.rowptr.begin(), lo.rowptr.end() - 1);
  for (int64_t j = 0; j < n; ++j)
    for (int32_t k = dpos[j] + 1; k < f.outer[j + 1]; ++k) {
      const int32_t i = f.inner[k];
      lo.colidx[cur[i]] = static_cast<int32_t>(j);
      lo.vals[cur[i]++] = f.vals[k];
    }
  lo.fused = true;   // the column sweep's update is one FMA (:129)
  up.fused = false;  // the row-wise backward loop is not contracted (:91)
  int rc = analyse_levels(lo, n, true, err);
  if (!rc) rc = analyse_levels(up, n, false, err);
  if (rc) return rc;
  // x = perm b; x = scale .* x; ...; x = scale .* x; z = perm^-1 x
  f.pre_scale = f.scale;
  if (f.perm.empty()) {
    f.pre_gather.clear();
    f.post_gather.clear();
    f.post_scale = f.scale;
  } else {
    f.pre_gather.resize(n);
    for (int64_t i = 0; i < n; ++i) f.pre_gather[f.perm[i]] = static_cast<int32_t>(i);   // (P b)[P[i]] = b[i]
    f.post_gather.assign(f.perm.begin(), f.perm.end());                                    // (P^-1 y)[k] = y[P[k]]
    f.post_scale.resize(n);
    for (int64_t k = 0; k < n; ++k) f.post_scale[k] = f.scale[f.perm[k]];
  }
  return 0;
}

}  // namespace

int multicolor_ordering(int64_t n, const int32_t* rowptr, const int32_t* colidx, int32_t* perm, std::string& err) {
  if (n < 0 || (n > 0 && (!rowptr || !colidx || !perm))) { err = "ordering: null or negative argument"; return B200S_ERR_INVALID; }
  // symmetrised adjacency (pattern of A + A^T without the diagonal)
  std::vector<int64_t> ptr(static_cast<size_t>(n) + 1, 0);
  for (int64_t i = 0; i < n; ++i)
    for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
      const int64_t j = colidx[k];
      if (j < 0 || j >= n) { err = "ordering: column index out of range"; return B200S_ERR_INVALID; }
      if (j != i) { ptr[i + 1]++; ptr[j + 1]++; }
    }
  for (int64_t i = 0; i < n; ++i) ptr[i + 1] += ptr[i];
  std::vector<int32_t> adj(static_cast<size_t>(ptr[n]));
  {
    std::vector<int64_t> cur(ptr.begin(), ptr.end() - 1);
    for (int64_t i = 0; i < n; ++i)
      for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const int64_t j = colidx[k];
        if (j != i) { adj[cur[i]++] = static_cast<int32_t>(j); adj[cur[j]++] = static_cast<int32_t>(i); }
      }
  }
  std::vector<int32_t> colour(static_cast<size_t>(n), -1), mark;
  int32_t ncol = 0;
  for (int64_t i = 0; i < n; ++i) {
    mark.assign(static_cast<size_t>(ncol) + 1, 0);
    for (int64_t k = ptr[i]; k < ptr[i + 1]; ++k)
      if (colour[adj[k]] >= 0) mark[colour[adj[k]]] = 1;
    int32_t c = 0;
    while (mark[c]) ++c;
    colour[i] = c;
    ncol = std::max(ncol, c + 1);
  }
  // stable counting sort by colour: new position of vertex i
  std::vector<int64_t> start(static_cast<size_t>(ncol) + 1, 0);
  for (int64_t i = 0; i < n; ++i) start[colour[i] + 1]++;
  for (int32_t c = 0; c < ncol; ++c) start[c + 1] += start[c];
  for (int64_t i = 0; i < n; ++i) perm[i] = static_cast<int32_t>(start[colour[i]]++);
  return ncol;
}

int factors_from_ilut(int64_t n, const int32_t* lu_rowptr, const int32_t* lu_colidx, const double* lu_vals,
                      const int32_t* perm, Factors& f, std::string& err) {
  if (n < 0 || (n > 0 && (!lu_rowptr || !lu_colidx || !lu_vals))) { err = "factors: null or negative argument"; return B200S_ERR_INVALID; }
  f = Factors();
  f.kind = B200S_FACTORS_ILUT;
  f.n = n;
  f.outer.assign(lu_rowptr, lu_rowptr + n + 1);
  const int64_t nz = n ? f.outer[n] : 0;
  for (int64_t i = 0; i < n; ++i)
    if (f.outer[i] > f.outer[i + 1] || f.outer[0] != 0) { err = "factors: row pointers must start at 0 and be monotone"; return B200S_ERR_INVALID; }
  f.inner.assign(lu_colidx, lu_colidx + nz);
  f.vals.assign(lu_vals, lu_vals + nz);
  f.perm.resize(n);
  if (perm) {
    if (!is_permutation(perm, n)) { err = "factors: perm is not a permutation"; return B200S_ERR_INVALID; }
    std::copy(perm, perm + n, f.perm.begin());
  } else {
    std::iota(f.perm.begin(), f.perm.end(), 0);
  }
  return stages_from_lu(f, err);
}

int factors_from_ichol(int64_t n, const int32_t* colptr, const int32_t* rowidx, const double* lvals, const double* scale,
                       const int32_t* perm, Factors& f, std::string& err) {
  if (n < 0 || (n > 0 && (!colptr || !rowidx || !lvals))) { err = "factors: null or negative argument"; return B200S_ERR_INVALID; }
  f = Factors();
  f.kind = B200S_FACTORS_ICHOL;
  f.n = n;
  f.outer.assign(colptr, colptr + n + 1);
  const int64_t nz = n ? f.outer[n] : 0;
  for (int64_t i = 0; i < n; ++i)
    if (f.outer[i] > f.outer[i + 1] || f.outer[0] != 0) { err = "factors: column pointers must start at 0 and be monotone"; return B200S_ERR_INVALID; }
  f.inner.assign(rowidx, rowidx + nz);
  f.vals.assign(lvals, lvals + nz);
  if (scale) f.scale.assign(scale, scale + n);
  else f.scale.assign(n, 1.0);
  if (perm) {
    if (!is_permutation(perm, n)) { err = "factors: perm is not a permutation"; return B200S_ERR_INVALID; }
    f.perm.assign(perm, perm + n);
  }
  return stages_from_l(f, err);
}

int ilut_factorize(int64_t n, const int32_t* rowptr, const int32_t* colidx, const double* vals, double droptol,
                   int fillfactor, const int32_t* perm, Factors& f, std::string& err) {
  if (n < 0 || (n > 0 && (!rowptr || !colidx || !vals))) { err = "ilut: null or negative argument"; return B200S_ERR_INVALID; }
  if (perm && !is_permutation(perm, n)) { err = "ilut: perm is not a permutation"; return B200S_ERR_INVALID; }
  if (droptol < 0) droptol = 1e-12;        // NumTraits<double>::dummy_precision(), IncompleteLUT.h:118
  if (fillfactor <= 0) fillfactor = 10;
  f = Factors();
  f.kind = B200S_FACTORS_ILUT;
  f.n = n;
  f.perm.resize(n);
  if (perm) std::copy(perm, perm + n, f.perm.begin());
  else std::iota(f.perm.begin(), f.perm.end(), 0);
  std::vector<int32_t> pinv(static_cast<size_t>(n));
  for (int64_t i = 0; i < n; ++i) pinv[f.perm[i]] = static_cast<int32_t>(i);
  const int64_t nnz_a = n ? static_cast<int64_t>(rowptr[n]) - rowptr[0] : 0;

  // mat = amat.twistedBy(m_Pinv) (:262): mat(Pinv[i], Pinv[j]) = amat(i, j); the evaluation goes through a temporary of
  // the other storage order (SparseSelfAdjointView.h:636-644), so the rows of `mat` come out sorted by column.
  std::vector<int32_t> mp(static_cast<size_t>(n) + 1, 0), mi(static_cast<size_t>(nnz_a));
  std::vector<double> mv(static_cast<size_t>(nnz_a));
  for (int64_t i = 0; i < n; ++i) mp[pinv[i] + 1] = rowptr[i + 1] - rowptr[i];
  for (int64_t i = 0; i < n; ++i) mp[i + 1] += mp[i];
  {
    std::vector<std::pair<int32_t, double>> row;
    for (int64_t i = 0; i < n; ++i) {
      row.clear();
      for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        if (colidx[k] < 0 || colidx[k] >= n) { err = "ilut: column index out of range"; return B200S_ERR_INVALID; }
        row.emplace_back(pinv[colidx[k]], vals[k]);
      }
      std::stable_sort(row.begin(), row.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
      int32_t o = mp[pinv[i]];
      for (const auto& e : row) { mi[o] = e.first; mv[o++] = e.second; }
    }
  }

  std::vector<double> u(static_cast<size_t>(n), 0.0);
  std::vector<int32_t> ju(static_cast<size_t>(n), 0), jr(static_cast<size_t>(n), -1);
  int64_t fill_in = n ? (nnz_a * fillfactor) / n + 1 : 1;    // :270-271
  if (fill_in > n) fill_in = n;
  const int64_t nnzL = fill_in / 2, nnzU = nnzL;             // :274-275
  f.outer.assign(1, 0);
  f.inner.clear();
  f.vals.clear();
  // the reference reserves n * (nnzL + nnzU + 1) (:276), an upper bound that is gigabytes beyond need for large n; the
  // vectors grow on demand here, a modest reservation only avoids the first reallocations
  f.inner.reserve(static_cast<size_t>(std::min<int64_t>(n * (nnzL + nnzU + 1), 2 * nnz_a + n)));
  f.vals.reserve(f.inner.capacity());
  std::vector<int32_t> diag_pos(static_cast<size_t>(n), 0);  // where row i of the factor holds its diagonal
  f.info = 0;

  for (int64_t ii = 0; ii < n; ii++) {                       // :279
    // 1 - copy the lower and the upper part of row ii into the working vector u (:281-318)
    int64_t sizeu = 1, sizel = 0;
    ju[ii] = static_cast<int32_t>(ii);
    u[ii] = 0;
    jr[ii] = static_cast<int32_t>(ii);
    double rownorm = 0;
    for (int32_t p = mp[ii]; p < mp[ii + 1]; ++p) {
      const int64_t k = mi[p];
      const double v = mv[p];
      if (k < ii) {
        ju[sizel] = static_cast<int32_t>(k);
        u[sizel] = v;
        jr[k] = static_cast<int32_t>(sizel);
        ++sizel;
      } else if (k == ii) {
        u[ii] = v;
      } else {
        const int64_t jpos = ii + sizeu;
        ju[jpos] = static_cast<int32_t>(k);
        u[jpos] = v;
        jr[k] = static_cast<int32_t>(jpos);
        ++sizeu;
      }
      rownorm = B200S_FMA_ILUT_NORM(v, v, rownorm);               // rownorm += abs2(value)
    }
    // 2 - zero row (:321-325)
    if (rownorm == 0) {
      f.info = 1;
      f.outer.resize(static_cast<size_t>(n) + 1, f.outer.back());
      return 0;  // NumericalIssue: the caller sees info, no stages are built
    }
    rownorm = std::sqrt(rownorm);

    // 3 - eliminate the previous rows (:330-396)
    int64_t jj = 0, len = 0;
    while (jj < sizel) {
      int64_t k = jj;                                        // smallest column index among ju(jj:sizel)
      for (int64_t q = jj + 1; q < sizel; ++q)
        if (ju[q] < ju[k]) k = q;
      const int32_t minrow = ju[k];
      if (minrow != ju[jj]) {
        const int32_t j = ju[jj];
        std::swap(ju[jj], ju[k]);
        jr[minrow] = static_cast<int32_t>(jj);
        jr[j] = static_cast<int32_t>(k);
        std::swap(u[jj], u[k]);
      }
      jr[minrow] = -1;
      int32_t ki = diag_pos[minrow];                         // :355-357: first entry of row minrow with column >= minrow
      const double fact = u[jj] / f.vals[ki];
      if (std::abs(fact) <= droptol) {                       // :361-365
        jj++;
        continue;
      }
      ++ki;
      for (; ki < f.outer[minrow + 1]; ++ki) {               // :368-391
        const int64_t j = f.inner[ki];
        const int32_t jpos = jr[j];
        if (jpos == -1) {                                    // fill-in
          int64_t newpos;
          if (j >= ii) { newpos = ii + sizeu; sizeu++; }
          else { newpos = sizel; sizel++; }
          ju[newpos] = static_cast<int32_t>(j);
          u[newpos] = -(fact * f.vals[ki]);
          jr[j] = static_cast<int32_t>(newpos);
        } else {
          u[jpos] = B200S_FMA_ILUT_ELIM(-fact, f.vals[ki], u[jpos]);   // u(jpos) -= fact * value
        }
      }
      u[len] = fact;                                         // :393-395
      ju[len] = minrow;
      ++len;
      jj++;
    }
    for (int64_t k = 0; k < sizeu; k++) jr[ju[ii + k]] = -1;  // :399

    // 4 - partial sort and insertion into the factor (:403-439)
    sizel = len;
    len = std::min(sizel, nnzL);
    quick_split(u.data(), ju.data(), sizel, len);
    for (int64_t k = 0; k < len; k++) { f.inner.push_back(ju[k]); f.vals.push_back(u[k]); }
    if (u[ii] == 0.0) u[ii] = std::sqrt(droptol) * rownorm;   // shifting rule (:417-418)
    diag_pos[ii] = static_cast<int32_t>(f.inner.size());
    f.inner.push_back(static_cast<int32_t>(ii));
    f.vals.push_back(u[ii]);
    len = 0;
    for (int64_t k = 1; k < sizeu; k++)                      // dropping rule (:423-432)
      if (std::abs(u[ii + k]) > droptol * rownorm) {
        ++len;
        u[ii + len] = u[ii + k];
        ju[ii + len] = ju[ii + k];
      }
    sizeu = len + 1;
    len = std::min(sizeu, nnzU);
    quick_split(u.data() + ii + 1, ju.data() + ii + 1, sizeu - 1, len);
    for (int64_t k = ii + 1; k < ii + len; k++) { f.inner.push_back(ju[k]); f.vals.push_back(u[k]); }
    if (f.inner.size() >= (size_t(1) << 31) - 64) { err = "ilut: factor does not fit int32"; return B200S_ERR_UNSUPPORTED; }
    f.outer.push_back(static_cast<int32_t>(f.inner.size()));
  }
  return stages_from_lu(f, err);
}

int ichol_factorize(int64_t n, const int32_t* rowptr, const int32_t* colidx, const double* a_vals, int uplo, double shift0,
                    const int32_t* perm, Factors& f, std::string& err) {
  if (n < 0 || (n > 0 && (!rowptr || !colidx || !a_vals))) { err = "ichol: null or negative argument"; return B200S_ERR_INVALID; }
  if (uplo != B200S_LOWER && uplo != B200S_UPPER) { err = "ichol: uplo must be 1 (Lower) or 2 (Upper): the triangle that is read"; return B200S_ERR_INVALID; }
  if (perm && !is_permutation(perm, n)) { err = "ichol: perm is not a permutation"; return B200S_ERR_INVALID; }
  if (shift0 < 0) shift0 = 1e-3;                             // m_initialShift, IncompleteCholesky.h:74
  f = Factors();
  f.kind = B200S_FACTORS_ICHOL;
  f.n = n;
  if (perm) f.perm.assign(perm, perm + n);
  f.info = 1;                                                // m_info = NumericalIssue until a factorization succeeds (:270)

  // m_L.selfadjointView<Lower>() = mat.selfadjointView<UpLo>()[.twistedBy(m_perm)] (:208-219): the lower triangle of
  // the (permuted) matrix, column-major.  Without a permutation the entries of a column keep the order in which the
  // row-major source yields them (SparseSelfAdjointView.h:516-577); with one the intermediate full matrix is sorted.
  std::vector<int32_t>& colPtr = f.outer;
  std::vector<int32_t>& rowIdx = f.inner;
  std::vector<double>& vals = f.vals;
  colPtr.assign(static_cast<size_t>(n) + 1, 0);
  auto keep = [&](int64_t r, int64_t c) { return uplo == B200S_LOWER ? c <= r : c >= r; };
  for (int64_t r = 0; r < n; ++r)
    for (int32_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
      const int64_t c = colidx[k];
      if (c < 0 || c >= n) { err = "ichol: column index out of range"; return B200S_ERR_INVALID; }
      if (!keep(r, c)) continue;
      const int64_t rp = perm ? perm[r] : r, cp = perm ? perm[c] : c;
      colPtr[std::min(rp, cp) + 1]++;
    }
  for (int64_t j = 0; j < n; ++j) colPtr[j + 1] += colPtr[j];
  const int64_t nnz = n ? colPtr[n] : 0;
  rowIdx.resize(static_cast<size_t>(nnz));
  vals.resize(static_cast<size_t>(nnz));
  {
    std::vector<int32_t> cur(colPtr.begin(), colPtr.end() - (n ? 1 : 0));
    for (int64_t r = 0; r < n; ++r)
      for (int32_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
        const int64_t c = colidx[k];
        if (!keep(r, c)) continue;
        const int64_t rp = perm ? perm[r] : r, cp = perm ? perm[c] : c;
        const int32_t o = cur[std::min(rp, cp)]++;
        rowIdx[o] = static_cast<int32_t>(std::max(rp, cp));
        vals[o] = a_vals[k];
      }
    if (perm) {
      std::vector<std::pair<int32_t, double>> col;
      for (int64_t j = 0; j < n; ++j) {
        col.clear();
        for (int32_t k = colPtr[j]; k < colPtr[j + 1]; ++k) col.emplace_back(rowIdx[k], vals[k]);
        std::stable_sort(col.begin(), col.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
        int32_t o = colPtr[j];
        for (const auto& e : col) { rowIdx[o] = e.first; vals[o++] = e.second; }
      }
    }
  }
  for (int64_t j = 0; j < n; ++j)
    if (colPtr[j] == colPtr[j + 1] || rowIdx[colPtr[j]] != j) {
      // the reference asserts this (:258); without the diagonal first in every column its loops read garbage
      err = "ichol: every column needs a stored diagonal entry (first in its column)";
      return B200S_ERR_INVALID;
    }

  std::vector<int32_t> firstElt(static_cast<size_t>(std::max<int64_t>(n, 1)), 0);
  std::vector<std::vector<int32_t>> listCol(static_cast<size_t>(n));
  std::vector<double> col_vals(static_cast<size_t>(n), 0.0);
  std::vector<int32_t> col_irow(static_cast<size_t>(n), 0), col_pattern(static_cast<size_t>(n), -1);

  // scaling factors (:232-250)
  std::vector<double>& scale = f.scale;
  scale.assign(static_cast<size_t>(n), 0.0);
  for (int64_t j = 0; j < n; j++)
    for (int32_t k = colPtr[j]; k < colPtr[j + 1]; k++) {
      scale[j] = B200S_FMA_IC_SCALE(vals[k], vals[k], scale[j]);
      if (rowIdx[k] != j) scale[rowIdx[k]] = B200S_FMA_IC_SCALE(vals[k], vals[k], scale[rowIdx[k]]);
    }
  for (int64_t j = 0; j < n; ++j) scale[j] = std::sqrt(std::sqrt(scale[j]));
  for (int64_t j = 0; j < n; ++j)
    scale[j] = scale[j] > (std::numeric_limits<double>::min)() ? 1.0 / scale[j] : 1.0;

  // scale the matrix and find the smallest diagonal entry (:254-261)
  double mindiag = (std::numeric_limits<double>::max)();
  for (int64_t j = 0; j < n; j++) {
    for (int32_t k = colPtr[j]; k < colPtr[j + 1]; k++) vals[k] *= (scale[j] * scale[rowIdx[k]]);
    mindiag = std::min(vals[colPtr[j]], mindiag);
  }
  const std::vector<int32_t> rowIdx_save = rowIdx;           // L_save (:263)
  const std::vector<double> vals_save = vals;
  double shift = 0;
  if (mindiag <= 0.0) shift = shift0 - mindiag;

  auto update_list = [&](int64_t col, int64_t jk) {          // updateList (:369-388)
    if (jk < colPtr[col + 1]) {
      int64_t minpos = jk;
      for (int64_t q = jk + 1; q < colPtr[col + 1]; ++q)
        if (rowIdx[q] < rowIdx[minpos]) minpos = q;
      if (rowIdx[minpos] != rowIdx[jk]) {
        std::swap(rowIdx[jk], rowIdx[minpos]);
        std::swap(vals[jk], vals[minpos]);
      }
      firstElt[col] = static_cast<int32_t>(jk);
      listCol[rowIdx[jk]].push_back(static_cast<int32_t>(col));
    }
  };

  int iter = 0;
  do {                                                       // :273
    for (int64_t j = 0; j < n; j++) vals[colPtr[j]] += shift;
    int64_t j = 0;
    for (; j < n; ++j) {
      double diag = vals[colPtr[j]];
      int32_t col_nnz = 0;
      for (int32_t i = colPtr[j] + 1; i < colPtr[j + 1]; i++) {
        const int32_t l = rowIdx[i];
        col_vals[col_nnz] = vals[i];
        col_irow[col_nnz] = l;
        col_pattern[l] = col_nnz;
        col_nnz++;
      }
      for (size_t kk = 0; kk < listCol[j].size(); ++kk) {    // previous columns that update column j (:296-318)
        const int32_t k = listCol[j][kk];
        int64_t jk = firstElt[k];
        const double v_j_jk = vals[jk];
        jk += 1;
        for (int64_t i = jk; i < colPtr[k + 1]; i++) {
          const int32_t l = rowIdx[i];
          if (col_pattern[l] < 0) {
            col_vals[col_nnz] = vals[i] * v_j_jk;
            col_irow[col_nnz] = l;
            col_pattern[l] = col_nnz;
            col_nnz++;
          } else {
            col_vals[col_pattern[l]] = B200S_FMA_IC_UPDATE(-vals[i], v_j_jk, col_vals[col_pattern[l]]);
          }
        }
        update_list(k, jk);
      }
      if (diag <= 0) {                                       // :322-338
        if (++iter >= 10) return 0;                          // info stays NumericalIssue, as in the reference
        shift = std::max(shift0, 2.0 * shift);
        vals = vals_save;
        rowIdx = rowIdx_save;
        std::fill(col_pattern.begin(), col_pattern.end(), -1);
        for (auto& l : listCol) l.clear();
        break;
      }
      const double rdiag = std::sqrt(diag);
      vals[colPtr[j]] = rdiag;
      for (int32_t k = 0; k < col_nnz; ++k) {                // :342-349
        const int32_t i = col_irow[k];
        col_vals[k] /= rdiag;
        vals[colPtr[i]] = B200S_FMA_IC_DIAG(-col_vals[k], col_vals[k], vals[colPtr[i]]);
      }
      const int64_t p = colPtr[j + 1] - colPtr[j] - 1;       // keep the p largest (:352-364)
      quick_split(col_vals.data(), col_irow.data(), col_nnz, p);
      int32_t cpt = 0;
      for (int32_t i = colPtr[j] + 1; i < colPtr[j + 1]; i++) {
        vals[i] = col_vals[cpt];
        rowIdx[i] = col_irow[cpt];
        col_pattern[col_irow[cpt]] = -1;
        cpt++;
      }
      update_list(j, colPtr[j] + 1);
    }
    if (j == n) f.info = 0;
  } while (f.info != 0);
  return stages_from_l(f, err);
}

}  // namespace b200s
