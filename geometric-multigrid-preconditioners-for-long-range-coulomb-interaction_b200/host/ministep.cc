// ministep implementation -- see ministep.h.  Host code, OpenMP over rows where it pays.
#include "ministep.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <stdexcept>

namespace ministep {

// =============================================================================== Forest
Forest::Forest(int reps_, double lo_, double hi_) : reps(reps_), lo(lo_), hi(hi_), H((hi_ - lo_) / reps_) {
  L.resize(1);
  auto &c = L[0];
  c.ijk.reserve((size_t)reps * reps * reps);
  for (int k = 0; k < reps; ++k)
    for (int j = 0; j < reps; ++j)
      for (int i = 0; i < reps; ++i) c.ijk.push_back({i, j, k});  // x fastest
  c.parent.assign(c.ijk.size(), -1);
  c.child0.assign(c.ijk.size(), -1);
}

bool Forest::inside(int l, int i, int j, int k) const {
  const int64_t n = cells_per_axis(l);
  return i >= 0 && j >= 0 && k >= 0 && i < n && j < n && k < n;
}

int Forest::lookup(int l, int i, int j, int k) const {
  if (!inside(l, i, j, k)) return -1;
  const int64_t n = cells_per_axis(l);
  const int64_t key = i + n * (j + n * (int64_t)k);
  if (l == 0) return (int)key;
  auto it = L[l].index.find(key);
  return it == L[l].index.end() ? -1 : it->second;
}

int64_t Forest::n_active_cells() const {
  int64_t n = 0;
  for (auto &lv : L)
    for (int c0 : lv.child0) n += c0 < 0;
  return n;
}

void Forest::refine(std::vector<std::vector<char>> flags) {
  const int nl = n_levels();
  flags.resize(nl);
  for (int l = 0; l < nl; ++l) flags[l].resize(n_cells(l), 0);
  // 2:1 balance over faces, edges and corners: a flagged level-l cell forces every coarser
  // (level l-1) active cell touching it to be refined as well; top-down so it cascades.
  for (int l = nl - 1; l >= 1; --l)
    for (int c = 0; c < n_cells(l); ++c) {
      if (!flags[l][c] || !active(l, c)) continue;
      const Int3 &p = L[l].ijk[c];
      for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            if (!dx && !dy && !dz) continue;
            const int i = p[0] + dx, j = p[1] + dy, k = p[2] + dz;
            if (!inside(l, i, j, k) || lookup(l, i, j, k) >= 0) continue;
            const int q = lookup(l - 1, i >> 1, j >> 1, k >> 1);
            if (q < 0 || !active(l - 1, q)) throw std::logic_error("mesh not 2:1 balanced");
            flags[l - 1][q] = 1;
          }
    }
  for (int l = 0; l < nl; ++l) {
    const int n_old = (int)flags[l].size();
    for (int c = 0; c < n_old; ++c) {
      if (!flags[l][c] || !active(l, c)) continue;
      if (l + 1 == n_levels()) L.emplace_back();
      auto &ch = L[l + 1];
      const int start = (int)ch.ijk.size();
      L[l].child0[c] = start;
      const Int3 p = L[l].ijk[c];
      const int64_t n = cells_per_axis(l + 1);
      for (int v = 0; v < NV; ++v) {
        const Int3 q = {2 * p[0] + vo(v, 0), 2 * p[1] + vo(v, 1), 2 * p[2] + vo(v, 2)};
        ch.index[q[0] + n * (q[1] + n * (int64_t)q[2])] = start + v;
        ch.ijk.push_back(q);
        ch.parent.push_back(c);
        ch.child0.push_back(-1);
      }
    }
  }
}

void Forest::refine_global(int times) {
  for (int t = 0; t < times; ++t) {
    std::vector<std::vector<char>> flags(n_levels());
    for (int l = 0; l < n_levels(); ++l) {
      flags[l].resize(n_cells(l));
      for (int c = 0; c < n_cells(l); ++c) flags[l][c] = active(l, c);
    }
    refine(flags);
  }
}

// =============================================================================== DoFs
int64_t DoFs::key(const Int3 &p) const {
  const int64_t N = f.points_per_axis();
  return p[0] + N * (p[1] + N * (int64_t)p[2]);
}
int DoFs::lookup(const Int3 &p) const {
  auto it = key2dof.find(key(p));
  return it == key2dof.end() ? -1 : it->second;
}
std::array<double, 3> DoFs::coords(int dof) const {
  const double s = f.H / (double)(1 << res);
  return {f.lo + xyz[dof][0] * s, f.lo + xyz[dof][1] * s, f.lo + xyz[dof][2] * s};
}

static inline Int3 vertex_xyz(const Int3 &ijk, int v, int shift) {
  return {(ijk[0] + vo(v, 0)) << shift, (ijk[1] + vo(v, 1)) << shift, (ijk[2] + vo(v, 2)) << shift};
}

DoFs::DoFs(const Forest &forest) : f(forest), res(forest.resolution()) {
  const int nl = f.n_levels();
  const int64_t N = f.points_per_axis();
  active_cells.resize(nl);
  active_pos.resize(nl);
  cell_dofs.resize(nl);
  size_t n_act = 0;
  for (int l = 0; l < nl; ++l) {
    active_pos[l].assign(f.n_cells(l), -1);
    for (int c = 0; c < f.n_cells(l); ++c)
      if (f.active(l, c)) {
        active_pos[l][c] = (int)active_cells[l].size();
        active_cells[l].push_back(c);
      }
    n_act += active_cells[l].size();
  }
  // distribute_dofs: first touch over active cells in (level, index) order, vertices 0..7
  key2dof.reserve(n_act * 2);
  for (int l = 0; l < nl; ++l) {
    cell_dofs[l].resize(active_cells[l].size());
    for (size_t p = 0; p < active_cells[l].size(); ++p) {
      const Int3 &ijk = f.L[l].ijk[active_cells[l][p]];
      for (int v = 0; v < NV; ++v) {
        const Int3 q = vertex_xyz(ijk, v, res - l);
        auto ins = key2dof.emplace(q[0] + N * (q[1] + N * (int64_t)q[2]), n);
        if (ins.second) {
          xyz.push_back(q);
          ++n;
        }
        cell_dofs[l][p][v] = ins.first->second;
      }
    }
  }
  auto on_boundary = [&](const Int3 &q) {
    return q[0] == 0 || q[1] == 0 || q[2] == 0 || q[0] == N - 1 || q[1] == N - 1 || q[2] == N - 1;
  };
  boundary.resize(n);
  for (int i = 0; i < n; ++i) boundary[i] = on_boundary(xyz[i]);
  // distribute_mg_dofs: per level, first touch over all cells of the level
  level_n.assign(nl, 0);
  level_cell_dofs.resize(nl);
  level_xyz.resize(nl);
  level_boundary.resize(nl);
  level_edge.resize(nl);
  for (int l = 0; l < nl; ++l) {
    std::unordered_map<int64_t, int> map;
    map.reserve((size_t)f.n_cells(l) * 2);
    level_cell_dofs[l].resize(f.n_cells(l));
    for (int c = 0; c < f.n_cells(l); ++c)
      for (int v = 0; v < NV; ++v) {
        const Int3 q = vertex_xyz(f.L[l].ijk[c], v, res - l);
        auto ins = map.emplace(q[0] + N * (q[1] + N * (int64_t)q[2]), level_n[l]);
        if (ins.second) {
          level_xyz[l].push_back(q);
          ++level_n[l];
        }
        level_cell_dofs[l][c][v] = ins.first->second;
      }
    level_boundary[l].resize(level_n[l]);
    for (int i = 0; i < level_n[l]; ++i) level_boundary[l][i] = on_boundary(level_xyz[l][i]);
    // refinement edge: faces of level-l cells whose neighbour across the face is coarser
    level_edge[l].assign(level_n[l], 0);
    if (l > 0)
      for (int c = 0; c < f.n_cells(l); ++c) {
        const Int3 &p = f.L[l].ijk[c];
        for (int a = 0; a < DIM; ++a)
          for (int side = 0; side < 2; ++side) {
            Int3 q = p;
            q[a] += side ? 1 : -1;
            if (!f.inside(l, q[0], q[1], q[2]) || f.lookup(l, q[0], q[1], q[2]) >= 0) continue;
            for (int v = 0; v < NV; ++v)
              if (vo(v, a) == side) level_edge[l][level_cell_dofs[l][c][v]] = 1;
          }
      }
  }
  // make_hanging_node_constraints
  std::vector<std::vector<std::pair<int, double>>> lists(n);
  hanging.assign(n, 0);
  for (int l = 0; l < nl; ++l) {
    if (res - l - 1 < 0) continue;
    const int half = 1 << (res - l - 1);
    for (int c : active_cells[l]) {
      const Int3 &p = f.L[l].ijk[c];
      for (int a = 0; a < DIM; ++a)
        for (int side = 0; side < 2; ++side) {
          Int3 q = p;
          q[a] += side ? 1 : -1;
          const int nb = f.lookup(l, q[0], q[1], q[2]);
          if (nb < 0 || f.active(l, nb)) continue;
          const int o0 = (a + 1) % 3, o1 = (a + 2) % 3;
          Int3 base = {p[0] << (res - l), p[1] << (res - l), p[2] << (res - l)};
          if (side) base[a] += 2 * half;
          for (int t0 = 0; t0 <= 2; ++t0)
            for (int t1 = 0; t1 <= 2; ++t1) {
              if (t0 != 1 && t1 != 1) continue;
              Int3 hp = base;
              hp[o0] += t0 * half;
              hp[o1] += t1 * half;
              const int hd = lookup(hp);
              if (hd < 0) throw std::logic_error("hanging vertex is not a dof");
              if (hanging[hd]) continue;
              hanging[hd] = 1;
              const int n0 = (t0 == 1) ? 2 : 1, n1 = (t1 == 1) ? 2 : 1;
              const double w = 1.0 / (n0 * n1);
              for (int s0 = 0; s0 < n0; ++s0)
                for (int s1 = 0; s1 < n1; ++s1) {
                  Int3 pp = hp;
                  if (t0 == 1) pp[o0] += s0 ? half : -half;
                  if (t1 == 1) pp[o1] += s1 ? half : -half;
                  const int pd = lookup(pp);
                  if (pd < 0) throw std::logic_error("hanging-node parent is not a dof");
                  lists[hd].push_back({pd, w});
                }
              std::sort(lists[hd].begin(), lists[hd].end());
            }
        }
    }
  }
  hang.n_rows = hang.n_cols = n;
  hang.rowptr.assign(n + 1, 0);
  for (int i = 0; i < n; ++i) hang.rowptr[i + 1] = hang.rowptr[i] + (int64_t)lists[i].size();
  hang.col.reserve(hang.rowptr[n]);
  hang.val.reserve(hang.rowptr[n]);
  for (int i = 0; i < n; ++i)
    for (auto &e : lists[i]) {
      if (hanging[e.first]) throw std::logic_error("chained hanging-node constraints");
      hang.col.push_back(e.first);
      hang.val.push_back(e.second);
    }
  dirichlet.resize(n);
  constrained.resize(n);
  for (int i = 0; i < n; ++i) {
    dirichlet[i] = boundary[i] && !hanging[i];  // interpolate_boundary_values skips constrained dofs
    constrained[i] = hanging[i] || dirichlet[i];
  }
  // copy indices: dofs of active level-l cells not on the refinement edge of level l
  copy_global.resize(nl);
  copy_level.resize(nl);
  for (int l = 0; l < nl; ++l) {
    std::vector<std::pair<int, int>> pairs;
    for (size_t p = 0; p < active_cells[l].size(); ++p) {
      const Dofs8 &lv = level_cell_dofs[l][active_cells[l][p]];
      for (int v = 0; v < NV; ++v)
        if (!level_edge[l][lv[v]]) pairs.push_back({lv[v], cell_dofs[l][p][v]});
    }
    std::sort(pairs.begin(), pairs.end());
    pairs.erase(std::unique(pairs.begin(), pairs.end(), [](auto &x, auto &y) { return x.first == y.first; }),
                pairs.end());
    for (auto &e : pairs) {
      copy_level[l].push_back(e.first);
      copy_global[l].push_back(e.second);
    }
  }
}

// =============================================================================== element
void gauss_unit(int n, std::vector<double> &pts, std::vector<double> &wts) {
  // Gauss-Legendre on [0,1] by Newton iteration on P_n
  pts.resize(n);
  wts.resize(n);
  for (int i = 0; i < n; ++i) {
    double x = std::cos(M_PI * (i + 0.75) / (n + 0.5));
    double dp = 1.0;
    for (int it = 0; it < 100; ++it) {
      double p0 = 1.0, p1 = x;
      for (int k = 2; k <= n; ++k) {
        const double p2 = ((2.0 * k - 1.0) * x * p1 - (k - 1.0) * p0) / k;
        p0 = p1;
        p1 = p2;
      }
      if (n == 1) { p0 = 1.0; p1 = x; }
      dp = n * (x * p1 - p0) / (x * x - 1.0);
      const double dx = p1 / dp;
      x -= dx;
      if (std::fabs(dx) < 1e-16) break;
    }
    double p0 = 1.0, p1 = x;
    for (int k = 2; k <= n; ++k) {
      const double p2 = ((2.0 * k - 1.0) * x * p1 - (k - 1.0) * p0) / k;
      p0 = p1;
      p1 = p2;
    }
    dp = n * (x * p1 - p0) / (x * x - 1.0);
    pts[n - 1 - i] = 0.5 * (x + 1.0);
    wts[n - 1 - i] = 1.0 / ((1.0 - x * x) * dp * dp);  // = 0.5 * 2/((1-x^2) P'^2)
  }
}

void unit_stiffness_q(double G[8][NV][NV], double pts[8][3]) {
  std::vector<double> gp, gw;
  gauss_unit(2, gp, gw);
  for (int q = 0; q < 8; ++q) {
    const int qi[3] = {q & 1, (q >> 1) & 1, (q >> 2) & 1};  // x fastest
    double x[3], w = 1.0;
    for (int d = 0; d < 3; ++d) {
      x[d] = gp[qi[d]];
      pts[q][d] = x[d];
      w *= gw[qi[d]];
    }
    double grad[NV][3];
    for (int v = 0; v < NV; ++v)
      for (int g = 0; g < 3; ++g) {
        double t = 1.0;
        for (int d = 0; d < 3; ++d)
          t *= (d == g) ? (vo(v, d) ? 1.0 : -1.0) : (vo(v, d) ? x[d] : 1.0 - x[d]);
        grad[v][g] = t;
      }
    for (int i = 0; i < NV; ++i)
      for (int j = 0; j < NV; ++j)
        G[q][i][j] = (grad[i][0] * grad[j][0] + grad[i][1] * grad[j][1] + grad[i][2] * grad[j][2]) * w;
  }
}

void unit_stiffness(double K[NV][NV]) {
  double G[8][NV][NV], pts[8][3];
  unit_stiffness_q(G, pts);
  for (int i = 0; i < NV; ++i)
    for (int j = 0; j < NV; ++j) {
      double s = 0.0;
      for (int q = 0; q < 8; ++q) s += G[q][i][j];
      K[i][j] = s;
    }
}

namespace {

// cell stiffness provider: h * Kref, or per-cell with a coefficient evaluated at the 8 Gauss points
struct CellK {
  double Kref[NV][NV];
  double G[8][NV][NV], pts[8][3];
  const Coefficient *coef;
  explicit CellK(const Coefficient &c) : coef(c ? &c : nullptr) {
    unit_stiffness(Kref);
    unit_stiffness_q(G, pts);
  }
  void get(const Forest &f, int l, int c, double K[NV][NV]) const {
    const double h = f.h(l);
    if (!coef) {
      for (int i = 0; i < NV; ++i)
        for (int j = 0; j < NV; ++j) K[i][j] = h * Kref[i][j];
      return;
    }
    const Int3 &p = f.L[l].ijk[c];
    double cq[8];
    for (int q = 0; q < 8; ++q)
      cq[q] = (*coef)(f.lo + (p[0] + pts[q][0]) * h, f.lo + (p[1] + pts[q][1]) * h, f.lo + (p[2] + pts[q][2]) * h);
    for (int i = 0; i < NV; ++i)
      for (int j = 0; j < NV; ++j) {
        double s = 0.0;
        for (int q = 0; q < 8; ++q) s += cq[q] * G[q][i][j];
        K[i][j] = h * s;
      }
  }
};

struct Entry {
  int col;
  double val;
};

// sort by column (stable) and merge duplicates, summing in encounter order
void compress_row(std::vector<Entry> &row) {
  std::stable_sort(row.begin(), row.end(), [](const Entry &a, const Entry &b) { return a.col < b.col; });
  size_t o = 0;
  for (size_t i = 0; i < row.size();) {
    double v = 0.0;
    size_t j = i;
    for (; j < row.size() && row[j].col == row[i].col; ++j) v += row[j].val;
    row[o++] = {row[i].col, v};
    i = j;
  }
  row.resize(o);
}

Csr rows_to_csr(int n_rows, int n_cols, std::vector<std::vector<Entry>> &rows) {
  Csr m;
  m.n_rows = n_rows;
  m.n_cols = n_cols;
  m.rowptr.assign(n_rows + 1, 0);
  for (int r = 0; r < n_rows; ++r) m.rowptr[r + 1] = m.rowptr[r] + (int64_t)rows[r].size();
  m.col.resize(m.rowptr[n_rows]);
  m.val.resize(m.rowptr[n_rows]);
#pragma omp parallel for schedule(static)
  for (int r = 0; r < n_rows; ++r) {
    int64_t at = m.rowptr[r];
    for (auto &e : rows[r]) {
      m.col[at] = e.col;
      m.val[at] = e.val;
      ++at;
    }
  }
  return m;
}

// incidence lists: for every dof the (cell, local index, weight) triples it receives contributions from
struct Incidence {
  std::vector<int64_t> ptr;
  std::vector<int> cell;   // flat cell index
  std::vector<signed char> local;
  std::vector<double> w;
};

}  // namespace

// =============================================================================== system matrix
Csr assemble_system_matrix(const Forest &f, const DoFs &d, const Coefficient &coef) {
  const int nl = f.n_levels();
  const int n = d.n;
  CellK cellK(coef);
  // flat list of active cells
  std::vector<int> lev, idx;
  std::vector<const Dofs8 *> dofs;
  for (int l = 0; l < nl; ++l)
    for (size_t p = 0; p < d.active_cells[l].size(); ++p) {
      lev.push_back(l);
      idx.push_back(d.active_cells[l][p]);
      dofs.push_back(&d.cell_dofs[l][p]);
    }
  const int nc = (int)lev.size();
  // resolved (dof, weight) lists per local dof: free -> itself; hanging -> its free parents; Dirichlet -> none
  auto resolved = [&](int dof, std::pair<int, double> out[4]) -> int {
    if (d.hanging[dof]) {
      int k = 0;
      for (int64_t e = d.hang.rowptr[dof]; e < d.hang.rowptr[dof + 1]; ++e)
        if (!d.constrained[d.hang.col[e]]) out[k++] = {d.hang.col[e], d.hang.val[e]};
      return k;
    }
    if (d.dirichlet[dof]) return 0;
    out[0] = {dof, 1.0};
    return 1;
  };
  // incidences: direct (dof is a vertex of the cell) and resolved (row receives a share)
  Incidence dir, rsv;
  dir.ptr.assign(n + 1, 0);
  rsv.ptr.assign(n + 1, 0);
  for (int c = 0; c < nc; ++c)
    for (int a = 0; a < NV; ++a) {
      const int dof = (*dofs[c])[a];
      dir.ptr[dof + 1]++;
      std::pair<int, double> r[4];
      const int k = resolved(dof, r);
      for (int t = 0; t < k; ++t) rsv.ptr[r[t].first + 1]++;
    }
  for (int i = 0; i < n; ++i) {
    dir.ptr[i + 1] += dir.ptr[i];
    rsv.ptr[i + 1] += rsv.ptr[i];
  }
  dir.cell.resize(dir.ptr[n]);
  dir.local.resize(dir.ptr[n]);
  rsv.cell.resize(rsv.ptr[n]);
  rsv.local.resize(rsv.ptr[n]);
  rsv.w.resize(rsv.ptr[n]);
  {
    std::vector<int64_t> cd(dir.ptr.begin(), dir.ptr.end() - 1), cr(rsv.ptr.begin(), rsv.ptr.end() - 1);
    for (int c = 0; c < nc; ++c)
      for (int a = 0; a < NV; ++a) {
        const int dof = (*dofs[c])[a];
        dir.cell[cd[dof]] = c;
        dir.local[cd[dof]++] = (signed char)a;
        std::pair<int, double> r[4];
        const int k = resolved(dof, r);
        for (int t = 0; t < k; ++t) {
          const int64_t at = cr[r[t].first]++;
          rsv.cell[at] = c;
          rsv.local[at] = (signed char)a;
          rsv.w[at] = r[t].second;
        }
      }
  }
  std::vector<std::vector<Entry>> rows(n);
#pragma omp parallel for schedule(dynamic, 1024)
  for (int i = 0; i < n; ++i) {
    std::vector<Entry> &row = rows[i];
    double K[NV][NV];
    // pattern of every cell containing i (keep_constrained_dofs = true): explicit zeros
    for (int64_t e = dir.ptr[i]; e < dir.ptr[i + 1]; ++e) {
      const int c = dir.cell[e];
      for (int b = 0; b < NV; ++b) row.push_back({(*dofs[c])[b], 0.0});
    }
    if (d.constrained[i]) {
      // constrained row: diagonal only, sum of |K_ii| over the cells (distribute_local_to_global)
      double diag = 0.0;
      for (int64_t e = dir.ptr[i]; e < dir.ptr[i + 1]; ++e) {
        const int c = dir.cell[e], a = dir.local[e];
        cellK.get(f, lev[c], idx[c], K);
        diag += std::fabs(K[a][a]);
      }
      row.push_back({i, diag});
    } else {
      for (int64_t e = rsv.ptr[i]; e < rsv.ptr[i + 1]; ++e) {
        const int c = rsv.cell[e], a = rsv.local[e];
        const double wa = rsv.w[e];
        cellK.get(f, lev[c], idx[c], K);
        for (int b = 0; b < NV; ++b) {
          std::pair<int, double> r[4];
          const int k = resolved((*dofs[c])[b], r);
          for (int t = 0; t < k; ++t) row.push_back({r[t].first, wa * r[t].second * K[a][b]});
        }
      }
    }
    compress_row(row);
  }
  return rows_to_csr(n, n, rows);
}

// =============================================================================== inputs of the device assembly
AssemblyInputs assembly_inputs_system(const Forest &f, const DoFs &d) {
  AssemblyInputs in;
  in.n_rows = d.n;
  for (int l = 0; l < f.n_levels(); ++l) in.n_cells += (int64_t)d.active_cells[l].size();
  in.cell_dofs.reserve(8 * in.n_cells);
  in.cell_h.reserve(in.n_cells);
  for (int l = 0; l < f.n_levels(); ++l)
    for (size_t p = 0; p < d.active_cells[l].size(); ++p) {
      for (int a = 0; a < NV; ++a) in.cell_dofs.push_back(d.cell_dofs[l][p][a]);
      in.cell_h.push_back(f.h(l));
    }
  in.flags.resize(d.n);
  for (int i = 0; i < d.n; ++i) {
    in.flags[i] = (uint8_t)((d.hanging[i] ? 2 : (d.dirichlet[i] ? 1 : 0)));
    if (d.hanging[i]) in.hanging = true;
  }
  return in;
}

AssemblyInputs assembly_inputs_level(const Forest &f, const DoFs &d, int l) {
  AssemblyInputs in;
  in.n_rows = d.level_n[l];
  in.n_cells = f.n_cells(l);
  in.cell_dofs.resize(8 * in.n_cells);
  for (int64_t c = 0; c < in.n_cells; ++c)
    for (int a = 0; a < NV; ++a) in.cell_dofs[8 * c + a] = d.level_cell_dofs[l][c][a];
  in.uniform_h = f.h(l);
  in.flags.resize(in.n_rows);
  for (int i = 0; i < in.n_rows; ++i) in.flags[i] = (uint8_t)((d.level_edge[l][i] || d.level_boundary[l][i]) ? 1 : 0);
  return in;
}

// =============================================================================== level operators
LevelOperators assemble_level_operators(const Forest &f, const DoFs &d, const Coefficient &coef, int first_level) {
  const int nl = f.n_levels();
  LevelOperators ops;
  ops.A.resize(nl);
  ops.I.resize(nl);
  ops.P.resize(nl > 0 ? nl - 1 : 0);
  CellK cellK(coef);
  for (int l = first_level; l < nl; ++l) {
    const int n = d.level_n[l], nc = f.n_cells(l);
    const auto &cd = d.level_cell_dofs[l];
    const auto &edge = d.level_edge[l];
    const auto &bd = d.level_boundary[l];
    // direct incidences
    std::vector<int64_t> ptr(n + 1, 0);
    for (int c = 0; c < nc; ++c)
      for (int a = 0; a < NV; ++a) ptr[cd[c][a] + 1]++;
    for (int i = 0; i < n; ++i) ptr[i + 1] += ptr[i];
    std::vector<int> icell(ptr[n]);
    std::vector<signed char> iloc(ptr[n]);
    {
      std::vector<int64_t> cur(ptr.begin(), ptr.end() - 1);
      for (int c = 0; c < nc; ++c)
        for (int a = 0; a < NV; ++a) {
          icell[cur[cd[c][a]]] = c;
          iloc[cur[cd[c][a]]++] = (signed char)a;
        }
    }
    std::vector<std::vector<Entry>> rowsA(n), rowsI(n);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int i = 0; i < n; ++i) {
      double K[NV][NV];
      const bool zi = edge[i] || bd[i];
      double diag = 0.0;
      for (int64_t e = ptr[i]; e < ptr[i + 1]; ++e) {
        const int c = icell[e], a = iloc[e];
        cellK.get(f, l, c, K);
        for (int b = 0; b < NV; ++b) {
          const int j = cd[c][b];
          const bool zj = edge[j] || bd[j];
          rowsA[i].push_back({j, (zi || zj) ? 0.0 : K[a][b]});
          // interface: i on the refinement edge, j not, neither on the boundary (src/step-50.cc:896-913)
          if (edge[i] && !bd[i] && !edge[j] && !bd[j]) rowsI[i].push_back({j, K[a][b]});
        }
        if (zi) diag += std::fabs(K[a][a]);
      }
      if (zi) rowsA[i].push_back({i, diag});
      compress_row(rowsA[i]);
      compress_row(rowsI[i]);
    }
    ops.A[l] = rows_to_csr(n, n, rowsA);
    ops.I[l] = rows_to_csr(n, n, rowsI);
  }
  // prolongation l -> l+1: Q1 embedding of the parent's dofs, `set` semantics, boundary columns zeroed
  for (int l = 0; l + 1 < nl; ++l) {
    const int nf = d.level_n[l + 1], ncoarse = d.level_n[l];
    std::vector<std::vector<Entry>> rows(nf);
    std::vector<char> done(nf, 0);
    for (int p = 0; p < f.n_cells(l); ++p) {
      const int c0 = f.L[l].child0[p];
      if (c0 < 0) continue;
      const Dofs8 &pd = d.level_cell_dofs[l][p];
      for (int c = 0; c < NV; ++c) {
        const Dofs8 &chd = d.level_cell_dofs[l + 1][c0 + c];
        for (int v = 0; v < NV; ++v) {
          const int row = chd[v];
          if (done[row]) continue;
          done[row] = 1;
          for (int w = 0; w < NV; ++w) {
            double wt = 1.0;
            for (int dd = 0; dd < DIM; ++dd) {
              const double t = 0.5 * (vo(c, dd) + vo(v, dd));
              wt *= vo(w, dd) ? t : 1.0 - t;
            }
            if (wt != 0.0 && !d.level_boundary[l][pd[w]]) rows[row].push_back({pd[w], wt});
          }
          compress_row(rows[row]);
        }
      }
    }
    ops.P[l] = rows_to_csr(nf, ncoarse, rows);
  }
  return ops;
}

// =============================================================================== constraints on vectors
std::vector<double> resolve_inhomogeneity(const DoFs &d, const std::vector<double> &g) {
  std::vector<double> ghat(g);
  for (int i = 0; i < d.n; ++i)
    if (d.hanging[i]) {
      double s = 0.0;
      for (int64_t e = d.hang.rowptr[i]; e < d.hang.rowptr[i + 1]; ++e) s += d.hang.val[e] * g[d.hang.col[e]];
      ghat[i] = s;
    }
  return ghat;
}

void distribute(const DoFs &d, const std::vector<double> &g, std::vector<double> &x) {
  for (int i = 0; i < d.n; ++i)
    if (d.dirichlet[i]) x[i] = g[i];
  for (int i = 0; i < d.n; ++i)
    if (d.hanging[i]) {
      double s = 0.0;
      for (int64_t e = d.hang.rowptr[i]; e < d.hang.rowptr[i + 1]; ++e) s += d.hang.val[e] * x[d.hang.col[e]];
      x[i] = s;
    }
}

// =============================================================================== error indicator
namespace {
// d/dx_a of the Q1 interpolant on a cube of edge h at tangential unit coordinates (s, t) of the
// two other axes (ordered by increasing axis index); constant along a.
inline double normal_derivative(const double U[NV], double h, int a, double s, double t) {
  const int o0 = (a == 0) ? 1 : 0, o1 = (a == 2) ? 1 : 2;
  double r = 0.0;
  for (int v = 0; v < NV; ++v) {
    double w = (vo(v, a) ? 1.0 : -1.0) / h;
    w = w * (vo(v, o0) ? s : 1.0 - s);
    w = w * (vo(v, o1) ? t : 1.0 - t);
    r += U[v] * w;
  }
  return r;
}
}  // namespace

IndicatorTopology indicator_topology(const Forest &f, const DoFs &d) {
  const int nl = f.n_levels();
  std::vector<int64_t> off(nl + 1, 0);
  for (int l = 0; l < nl; ++l) off[l + 1] = off[l] + (int64_t)d.active_cells[l].size();
  IndicatorTopology T;
  T.face_nb.assign(6 * (size_t)off[nl], -1);
  T.face_kind.assign(6 * (size_t)off[nl], 0);
  for (int l = 0; l < nl; ++l)
    for (size_t p = 0; p < d.active_cells[l].size(); ++p) {
      const int c = d.active_cells[l][p];
      const Int3 &ijk = f.L[l].ijk[c];
      const size_t me = 6 * (size_t)(off[l] + (int64_t)p);
      for (int a = 0; a < DIM; ++a) {
        const int o0 = (a == 0) ? 1 : 0, o1 = (a == 2) ? 1 : 2;
        for (int side = 0; side < 2; ++side) {
          Int3 q = ijk;
          q[a] += side ? 1 : -1;
          if (!f.inside(l, q[0], q[1], q[2])) continue;  // Dirichlet boundary face
          const int nb = f.lookup(l, q[0], q[1], q[2]);
          const size_t at = me + 2 * a + side;
          if (nb >= 0 && f.active(l, nb)) {
            T.face_nb[at] = (int32_t)(off[l] + d.active_pos[l][nb]);
            T.face_kind[at] = 0;
          } else if (nb >= 0) {  // refined neighbour: its four children on this face, ascending active position
            int32_t ch[4];
            int n = 0;
            for (int b1 = 0; b1 < 2; ++b1)
              for (int b0 = 0; b0 < 2; ++b0) {
                Int3 cq;
                cq[a] = 2 * q[a] + (side ? 0 : 1);
                cq[o0] = 2 * q[o0] + b0;
                cq[o1] = 2 * q[o1] + b1;
                const int cc = (l + 1 < nl) ? f.lookup(l + 1, cq[0], cq[1], cq[2]) : -1;
                if (cc < 0 || !f.active(l + 1, cc)) throw std::logic_error("unbalanced mesh in indicator_topology");
                ch[n++] = (int32_t)(off[l + 1] + d.active_pos[l + 1][cc]);
              }
            std::sort(ch, ch + 4);
            T.face_nb[at] = (int32_t)(T.hang_children.size() / 4);
            T.face_kind[at] = 2;
            T.hang_children.insert(T.hang_children.end(), ch, ch + 4);
          } else {  // coarser neighbour
            const int cc = f.lookup(l - 1, q[0] >> 1, q[1] >> 1, q[2] >> 1);
            if (cc < 0 || !f.active(l - 1, cc)) throw std::logic_error("unbalanced mesh in indicator_topology");
            const int sub0 = ijk[o0] - 2 * f.L[l - 1].ijk[cc][o0], sub1 = ijk[o1] - 2 * f.L[l - 1].ijk[cc][o1];
            T.face_nb[at] = (int32_t)(off[l - 1] + d.active_pos[l - 1][cc]);
            T.face_kind[at] = (uint8_t)(1 | (sub0 << 2) | (sub1 << 3));
          }
        }
      }
    }
  return T;
}

std::vector<std::vector<float>> error_indicator(const Forest &f, const DoFs &d, const std::vector<double> &u,
                                                const std::vector<double> &rho, int nq, bool residual_term) {
  const int nl = f.n_levels();
  std::vector<double> gp2, gw2, gp, gw;
  gauss_unit(2, gp2, gw2);
  gauss_unit(nq, gp, gw);
  std::vector<std::vector<std::array<double, 6>>> FI(nl);
  for (int l = 0; l < nl; ++l) FI[l].assign(d.active_cells[l].size(), std::array<double, 6>{0, 0, 0, 0, 0, 0});
  for (int l = 0; l < nl; ++l) {
    const double h = f.h(l);
    for (size_t p = 0; p < d.active_cells[l].size(); ++p) {
      const int c = d.active_cells[l][p];
      const Int3 &ijk = f.L[l].ijk[c];
      double U[NV];
      for (int v = 0; v < NV; ++v) U[v] = u[d.cell_dofs[l][p][v]];
      for (int a = 0; a < DIM; ++a) {
        const int o0 = (a == 0) ? 1 : 0, o1 = (a == 2) ? 1 : 2;
        for (int side = 0; side < 2; ++side) {
          Int3 q = ijk;
          q[a] += side ? 1 : -1;
          if (!f.inside(l, q[0], q[1], q[2])) continue;  // Dirichlet boundary face: no contribution
          const int nb = f.lookup(l, q[0], q[1], q[2]);
          if (nb >= 0 && !f.active(l, nb)) continue;  // coarse side: filled by the children below
          double Un[NV];
          double hn = h;
          int sub0 = 0, sub1 = 0, cpos = -1;
          if (nb >= 0) {
            const int np = d.active_pos[l][nb];
            for (int v = 0; v < NV; ++v) Un[v] = u[d.cell_dofs[l][np][v]];
          } else {
            const int cc = f.lookup(l - 1, q[0] >> 1, q[1] >> 1, q[2] >> 1);
            if (cc < 0 || !f.active(l - 1, cc)) throw std::logic_error("unbalanced mesh in error_indicator");
            cpos = d.active_pos[l - 1][cc];
            for (int v = 0; v < NV; ++v) Un[v] = u[d.cell_dofs[l - 1][cpos][v]];
            hn = 2.0 * h;
            sub0 = ijk[o0] - 2 * f.L[l - 1].ijk[cc][o0];
            sub1 = ijk[o1] - 2 * f.L[l - 1].ijk[cc][o1];
          }
          double I = 0.0;
          for (int t1 = 0; t1 < 2; ++t1)     // face rule QGauss<2>(2), first tangential axis fastest
            for (int t0 = 0; t0 < 2; ++t0) {
              const double s = gp2[t0], t = gp2[t1];
              const double own = normal_derivative(U, h, a, s, t);
              const double oth = (nb >= 0) ? normal_derivative(Un, hn, a, s, t)
                                           : normal_derivative(Un, hn, a, (sub0 + s) / 2.0, (sub1 + t) / 2.0);
              const double j = own - oth;
              I += (j * j) * (gw2[t0] * gw2[t1] * h * h);
            }
          FI[l][p][2 * a + side] = I;
          if (nb < 0) FI[l - 1][cpos][2 * a + (1 - side)] += I;  // sum over subfaces, child order
        }
      }
    }
  }
  std::vector<std::vector<float>> eta(nl);
  const int nq3 = nq * nq * nq;
  size_t off = 0;
  for (int l = 0; l < nl; ++l) {
    const double h = f.h(l);
    const double diam = std::sqrt(3.0 * h * h);
    eta[l].resize(d.active_cells[l].size());
    for (size_t p = 0; p < d.active_cells[l].size(); ++p, ++off) {
      float err = 0.0f;
      for (int face = 0; face < 6; ++face) err = (float)((double)err + FI[l][p][face] * diam);
      const float kelly = (float)std::sqrt((double)err);
      double resid = 0.0;
      const double *r = (rho.empty() || !residual_term) ? nullptr : &rho[off * nq3];
      if (r)
        for (int qz = 0, q = 0; qz < nq; ++qz)
          for (int qy = 0; qy < nq; ++qy)
            for (int qx = 0; qx < nq; ++qx, ++q) {
              const double t = 4.0 * M_PI * r[q];
              resid += (t * t) * (gw[qx] * gw[qy] * gw[qz] * h * h * h);
            }
      eta[l][p] = (float)std::sqrt((double)kelly * (double)kelly + diam * diam * resid);
    }
  }
  return eta;
}

double mark_cells(const Forest &f, const DoFs &d, const std::vector<std::vector<float>> &eta,
                  std::vector<std::vector<char>> &flags) {
  float mx = 0.0f;
  for (auto &e : eta)
    for (float v : e) mx = std::max(mx, v);
  const double threshold = 0.6 * (double)mx;
  flags.assign(f.n_levels(), {});
  for (int l = 0; l < f.n_levels(); ++l) {
    flags[l].assign(f.n_cells(l), 0);
    if (mx > 0.0f)
      for (size_t p = 0; p < d.active_cells[l].size(); ++p)
        flags[l][d.active_cells[l][p]] = ((double)eta[l][p] >= threshold);
  }
  return threshold;
}

// =============================================================================== solution transfer
std::vector<double> transfer_solution(int old_res, const DoFs &od, const std::vector<double> &u_old, const Forest &f,
                                      const DoFs &d) {
  const int shift = f.resolution() - old_res;
  std::vector<double> x(d.n, 0.0);
  std::vector<char> have(d.n, 0);
  for (int i = 0; i < od.n; ++i) {
    const Int3 q = {od.xyz[i][0] << shift, od.xyz[i][1] << shift, od.xyz[i][2] << shift};
    const int j = d.lookup(q);
    if (j >= 0) {
      x[j] = u_old[i];
      have[j] = 1;
    }
  }
  const int res = f.resolution();
  for (int l = 0; l + 1 < f.n_levels(); ++l) {
    const int half = 1 << (res - l - 1);
    for (int p = 0; p < f.n_cells(l); ++p) {
      if (f.L[l].child0[p] < 0) continue;
      const Int3 &ijk = f.L[l].ijk[p];
      const Int3 base = {ijk[0] << (res - l), ijk[1] << (res - l), ijk[2] << (res - l)};
      double Up[NV];
      bool ok = true;
      for (int v = 0; v < NV; ++v) {
        const int pd = d.lookup({base[0] + vo(v, 0) * 2 * half, base[1] + vo(v, 1) * 2 * half, base[2] + vo(v, 2) * 2 * half});
        if (pd < 0 || !have[pd]) { ok = false; break; }
        Up[v] = x[pd];
      }
      if (!ok) continue;
      for (int t2 = 0; t2 <= 2; ++t2)
        for (int t1 = 0; t1 <= 2; ++t1)
          for (int t0 = 0; t0 <= 2; ++t0) {
            if (t0 != 1 && t1 != 1 && t2 != 1) continue;
            const int dof = d.lookup({base[0] + t0 * half, base[1] + t1 * half, base[2] + t2 * half});
            if (dof < 0 || have[dof]) continue;
            const int t[3] = {t0, t1, t2};
            double val = 0.0;
            for (int v = 0; v < NV; ++v) {
              double w = 1.0;
              for (int k = 0; k < 3; ++k) w *= vo(v, k) ? t[k] / 2.0 : 1.0 - t[k] / 2.0;
              if (w != 0.0) val += w * Up[v];
            }
            x[dof] = val;
            have[dof] = 1;
          }
    }
  }
  for (int i = 0; i < d.n; ++i) {
    if (!have[i]) throw std::logic_error("solution transfer left a dof without a value");
    if (d.constrained[i]) x[i] = 0.0;
  }
  return x;
}

TransferTables transfer_tables(int old_res, const DoFs &od, const Forest &f, const DoFs &d) {
  TransferTables T;
  const int shift = f.resolution() - old_res;
  for (int i = 0; i < od.n; ++i) {
    const Int3 q = {od.xyz[i][0] << shift, od.xyz[i][1] << shift, od.xyz[i][2] << shift};
    const int j = d.lookup(q);
    if (j >= 0) {
      T.copy_old.push_back(i);
      T.copy_new.push_back(j);
    }
  }
  const int res = f.resolution();
  T.pass_ptr.push_back(0);
  for (int l = 0; l + 1 < f.n_levels(); ++l) {
    const int half = 1 << (res - l - 1);
    for (int p = 0; p < f.n_cells(l); ++p) {
      if (f.L[l].child0[p] < 0) continue;
      const Int3 &ijk = f.L[l].ijk[p];
      const Int3 base = {ijk[0] << (res - l), ijk[1] << (res - l), ijk[2] << (res - l)};
      for (int t2 = 0; t2 <= 2; ++t2)
        for (int t1 = 0; t1 <= 2; ++t1)
          for (int t0 = 0; t0 <= 2; ++t0)
            T.parent_dofs.push_back(d.lookup({base[0] + t0 * half, base[1] + t1 * half, base[2] + t2 * half}));
    }
    T.pass_ptr.push_back((int64_t)(T.parent_dofs.size() / 27));
  }
  return T;
}

void locate(const Forest &f, const DoFs &, const double X[3], int &level, int &cell, double xi[3]) {
  int ijk[3];
  for (int k = 0; k < 3; ++k) {
    ijk[k] = (int)std::floor((X[k] - f.lo) / f.H);
    ijk[k] = std::min(std::max(ijk[k], 0), f.reps - 1);
  }
  level = 0;
  cell = f.lookup(0, ijk[0], ijk[1], ijk[2]);
  while (!f.active(level, cell)) {
    const double h = f.h(level + 1);
    const Int3 &p = f.L[level].ijk[cell];
    int c[3];
    for (int k = 0; k < 3; ++k) {
      c[k] = (int)std::floor((X[k] - f.lo) / h);
      c[k] = std::min(std::max(c[k], 2 * p[k]), 2 * p[k] + 1);
    }
    cell = f.lookup(level + 1, c[0], c[1], c[2]);
    ++level;
  }
  const double h = f.h(level);
  const Int3 &p = f.L[level].ijk[cell];
  for (int k = 0; k < 3; ++k) xi[k] = (X[k] - (f.lo + p[k] * h)) / h;
}

}  // namespace ministep
