// sell_builder.cc -- host-side conversion of the caller's CSC constraint matrix
// (Eigen::SparseMatrix<double, ColMajor, int64_t> arrays, quadratic_program.h:138)
// into the two device images: SELL-32 over the rows of K and SELL-32 over the
// rows of K^T. This is the device counterpart of building the explicit
// transpose in ShardedQuadraticProgram (sharded_quadratic_program.cc:83).
//
// Ordering ("positions"): rows longer than `split_len` come first and are cut
// into virtual slots of at most `split_len` entries (their partial sums are
// combined by a fix-up kernel, in slot order, so the result is deterministic
// and needs no atomics). The remaining rows keep their original order except
// for a stable sort by descending length inside windows of `sigma` rows, which
// removes almost all slice padding while keeping gathers local.
#include <algorithm>
#include <cstdlib>
#include <numeric>
#include <stdexcept>

#include "device_ops.h"

namespace pdlp_b200 {
namespace {

int EnvInt(const char* name, int dflt) {
  const char* v = std::getenv(name);
  if (v == nullptr || *v == 0) return dflt;
  return std::atoi(v);
}

struct Csr {
  std::vector<int64_t> start;
  std::vector<int32_t> idx;
  std::vector<double> val;
};

int32_t ChooseSplitLen(int64_t nnz) {
  const int forced = EnvInt("PDLP_B200_SPLIT_LEN", 0);
  if (forced > 0) return forced;
  // Enough slots to fill 148 SMs x 2048 threads even when a few rows hold most
  // of the nonzeros.
  int64_t t = 64;
  while (t * 262144 < nnz) t *= 2;
  return static_cast<int32_t>(t);
}

// Fills positions: split rows first, then windows sorted by descending length.
void AssignPositions(const std::vector<int64_t>& len, int32_t split_len, int sigma, SellHost& s) {
  const int64_t rows = static_cast<int64_t>(len.size());
  s.num_rows = rows;
  s.split_len = split_len;
  s.row_of_pos.resize(rows);
  s.pos_of_row.resize(rows);
  std::vector<int32_t> split_rows, rest;
  for (int64_t r = 0; r < rows; ++r) (len[r] > split_len ? split_rows : rest).push_back(static_cast<int32_t>(r));
  if (sigma > 1) {
    for (size_t w = 0; w < rest.size(); w += sigma) {
      const size_t e = std::min(rest.size(), w + static_cast<size_t>(sigma));
      std::stable_sort(rest.begin() + w, rest.begin() + e, [&](int32_t a, int32_t b) { return len[a] > len[b]; });
    }
  }
  s.num_split = static_cast<int64_t>(split_rows.size());
  int64_t p = 0;
  for (int32_t r : split_rows) s.row_of_pos[p++] = r;
  for (int32_t r : rest) s.row_of_pos[p++] = r;
  for (int64_t q = 0; q < rows; ++q) s.pos_of_row[s.row_of_pos[q]] = static_cast<int32_t>(q);
}

void FillSell(const Csr& a, const std::vector<int32_t>& other_pos_of_row, int64_t num_cols, SellHost& s) {
  const int64_t rows = s.num_rows;
  const int32_t T = s.split_len;
  s.num_cols = num_cols;
  // virtual slots of the split rows
  s.split_first.assign(s.num_split + 1, 0);
  int64_t nv = 0;
  for (int64_t i = 0; i < s.num_split; ++i) {
    const int32_t r = s.row_of_pos[i];
    const int64_t len = a.start[r + 1] - a.start[r];
    s.split_first[i] = static_cast<int32_t>(nv);
    nv += (len + T - 1) / T;
  }
  s.split_first[s.num_split] = static_cast<int32_t>(nv);
  s.num_virtual = nv;
  s.num_virtual_padded = (nv + 31) / 32 * 32;
  s.num_slots = (s.num_virtual_padded + (rows - s.num_split) + 31) / 32 * 32;
  if (s.num_slots >= (int64_t{1} << 31)) throw std::runtime_error("too many rows for int32 slot indices");
  s.slot_len.assign(s.num_slots, 0);
  s.virt_pos.assign(s.num_virtual_padded, -1);
  // (source row, first entry) of every slot
  std::vector<int64_t> slot_src(s.num_slots, -1);
  for (int64_t i = 0; i < s.num_split; ++i) {
    const int32_t r = s.row_of_pos[i];
    const int64_t len = a.start[r + 1] - a.start[r];
    int64_t v = s.split_first[i];
    for (int64_t off = 0; off < len; off += T, ++v) {
      s.slot_len[v] = static_cast<int32_t>(std::min<int64_t>(T, len - off));
      s.virt_pos[v] = static_cast<int32_t>(i);
      slot_src[v] = a.start[r] + off;
    }
  }
  for (int64_t p = s.num_split; p < rows; ++p) {
    const int32_t r = s.row_of_pos[p];
    const int64_t slot = s.num_virtual_padded + (p - s.num_split);
    s.slot_len[slot] = static_cast<int32_t>(a.start[r + 1] - a.start[r]);
    slot_src[slot] = a.start[r];
  }
  const int64_t num_slices = s.num_slots / 32;
  s.slice_ptr.assign(num_slices + 1, 0);
  for (int64_t sl = 0; sl < num_slices; ++sl) {
    int32_t w = 0;
    for (int l = 0; l < 32; ++l) w = std::max(w, s.slot_len[sl * 32 + l]);
    s.slice_ptr[sl + 1] = s.slice_ptr[sl] + static_cast<int64_t>(w) * 32;
  }
  s.padded_nnz = s.slice_ptr[num_slices];
  s.col.assign(s.padded_nnz, 0);
  s.val.assign(s.padded_nnz, 0.0);
  for (int64_t slot = 0; slot < s.num_slots; ++slot) {
    const int64_t src = slot_src[slot];
    if (src < 0) continue;
    const int64_t base = s.slice_ptr[slot >> 5] + (slot & 31);
    for (int32_t j = 0; j < s.slot_len[slot]; ++j) {
      s.col[base + static_cast<int64_t>(j) * 32] = other_pos_of_row[a.idx[src + j]];
      s.val[base + static_cast<int64_t>(j) * 32] = a.val[src + j];
    }
  }
}

}  // namespace

QpHost BuildQpHost(const PdlpProblemView& v, int64_t row_begin, int64_t row_end, int sigma) {
  const int64_t n = v.num_variables, m_full = v.num_constraints;
  if (row_begin < 0 || row_end > m_full || row_begin > row_end) throw std::runtime_error("bad row range");
  if (n >= (int64_t{1} << 31) - 64 || m_full >= (int64_t{1} << 31) - 64) throw std::runtime_error("dimension exceeds int32 index range");
  const int64_t m = row_end - row_begin;
  sigma = EnvInt("PDLP_B200_SIGMA", sigma);
  // Column-major copy restricted to the row block (rows renumbered from 0).
  Csr kt;  // "rows" of K^T = columns of K
  kt.start.assign(n + 1, 0);
  const bool whole = (row_begin == 0 && row_end == m_full);
  for (int64_t c = 0; c < n; ++c) {
    const int64_t b = v.col_starts[c], e = v.col_starts[c + 1];
    if (e < b) throw std::runtime_error("col_starts is not monotone");
    int64_t cnt = 0;
    for (int64_t k = b; k < e; ++k) {
      const int64_t r = v.row_indices[k];
      if (r < 0 || r >= m_full) throw std::runtime_error("row index out of range");
      cnt += (whole || (r >= row_begin && r < row_end));
    }
    kt.start[c + 1] = kt.start[c] + cnt;
  }
  const int64_t nnz = kt.start[n];
  kt.idx.resize(nnz);
  kt.val.resize(nnz);
  std::vector<int64_t> row_len(m, 0), col_len(n, 0);
  {
    int64_t p = 0;
    for (int64_t c = 0; c < n; ++c) {
      for (int64_t k = v.col_starts[c]; k < v.col_starts[c + 1]; ++k) {
        const int64_t r = v.row_indices[k];
        if (!whole && (r < row_begin || r >= row_end)) continue;
        kt.idx[p] = static_cast<int32_t>(r - row_begin);
        kt.val[p] = v.values[k];
        ++row_len[r - row_begin];
        ++p;
      }
      col_len[c] = kt.start[c + 1] - kt.start[c];
    }
  }
  // Row-major copy by a stable counting sort (entries of a row end up in
  // ascending column order, like Eigen's transpose assignment).
  Csr k;
  k.start.assign(m + 1, 0);
  for (int64_t r = 0; r < m; ++r) k.start[r + 1] = k.start[r] + row_len[r];
  k.idx.resize(nnz);
  k.val.resize(nnz);
  {
    std::vector<int64_t> pos(k.start.begin(), k.start.end() - 1);
    for (int64_t c = 0; c < n; ++c)
      for (int64_t p = kt.start[c]; p < kt.start[c + 1]; ++p) {
        const int64_t q = pos[kt.idx[p]]++;
        k.idx[q] = static_cast<int32_t>(c);
        k.val[q] = kt.val[p];
      }
  }
  QpHost out;
  out.n = n;
  out.m = m;
  out.nnz = nnz;
  out.has_q = v.objective_matrix_diagonal != nullptr;
  const int32_t split_len = ChooseSplitLen(nnz);
  AssignPositions(row_len, split_len, sigma, out.rows);
  AssignPositions(col_len, split_len, sigma, out.cols);
  FillSell(k, out.cols.pos_of_row, n, out.rows);
  FillSell(kt, out.rows.pos_of_row, m, out.cols);
  return out;
}

}  // namespace pdlp_b200
