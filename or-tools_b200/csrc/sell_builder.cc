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
#include <exception>
#include <functional>
#include <numeric>
#include <stdexcept>
#include <thread>
#include <chrono>
#include <cstdio>

#include "device_ops.h"

namespace pdlp_b200 {
namespace {

int EnvInt(const char* name, int dflt) {
  const char* v = std::getenv(name);
  if (v == nullptr || *v == 0) return dflt;
  return std::atoi(v);
}

// Host threads used to build the device image (the conversion is the largest
// part of an end-to-end solve of a mid-sized problem, so it is parallel).
int BuildThreads() {
  const int forced = EnvInt("PDLP_B200_BUILD_THREADS", 0);
  if (forced > 0) return forced;
  const unsigned hc = std::thread::hardware_concurrency();
  return static_cast<int>(std::min<unsigned>(32u, std::max<unsigned>(1u, hc)));
}

// fn(begin, end, thread_index) over a contiguous partition of [0, count).
void ParallelFor(int64_t count, int threads, const std::function<void(int64_t, int64_t, int)>& fn) {
  threads = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(threads, count / 1024 + 1)));
  if (threads == 1) { fn(0, count, 0); return; }
  std::vector<std::thread> pool;
  std::vector<std::exception_ptr> errors(threads);
  for (int t = 0; t < threads; ++t) {
    const int64_t b = count * t / threads, e = count * (t + 1) / threads;
    pool.emplace_back([&, b, e, t] {
      try { fn(b, e, t); } catch (...) { errors[t] = std::current_exception(); }
    });
  }
  for (auto& th : pool) th.join();
  for (auto& e : errors) if (e) std::rethrow_exception(e);
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
    const int64_t windows = (static_cast<int64_t>(rest.size()) + sigma - 1) / sigma;
    ParallelFor(windows, BuildThreads(), [&](int64_t wb, int64_t we, int) {
      for (int64_t wi = wb; wi < we; ++wi) {
        const size_t w = static_cast<size_t>(wi) * sigma;
        const size_t e = std::min(rest.size(), w + static_cast<size_t>(sigma));
        std::stable_sort(rest.begin() + w, rest.begin() + e, [&](int32_t a, int32_t b) { return len[a] > len[b]; });
      }
    });
  }
  s.num_split = static_cast<int64_t>(split_rows.size());
  int64_t p = 0;
  for (int32_t r : split_rows) s.row_of_pos[p++] = r;
  for (int32_t r : rest) s.row_of_pos[p++] = r;
  ParallelFor(rows, BuildThreads(), [&](int64_t qb, int64_t qe, int) {
    for (int64_t q = qb; q < qe; ++q) s.pos_of_row[s.row_of_pos[q]] = static_cast<int32_t>(q);
  });
}

void FillSell(const Csr& a, const std::vector<int32_t>* other_pos_of_row, int64_t num_cols, SellHost& s) {
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
  // (first entry, distance between consecutive entries) of every slot. The virtual slots of a
  // split row that share a slice form a team and take the entries of their common range
  // round-robin (device_build.cu k_virtual_slots: same layout, bitwise the same products).
  std::vector<int64_t> slot_src(s.num_slots, -1);
  std::vector<int32_t> slot_stride(s.num_slots, 1);
  for (int64_t i = 0; i < s.num_split; ++i) {
    const int32_t r = s.row_of_pos[i];
    const int64_t len = a.start[r + 1] - a.start[r];
    const int64_t first = s.split_first[i], last = s.split_first[i + 1];
    for (int64_t v = first; v < last; ++v) {
      const bool teams = EnvInt("PDLP_B200_TEAM_SLOTS", 1) != 0;
      const int64_t ts = teams ? std::max<int64_t>(first, (v >> 5) << 5) : v, te = teams ? std::min<int64_t>(last, ((v >> 5) + 1) << 5) : v + 1;
      const int64_t t = te - ts, k = v - ts;
      const int64_t e0 = (ts - first) * T, e1 = std::min<int64_t>(len, (te - first) * T);
      const int64_t team_len = e1 - e0;
      s.slot_len[v] = team_len > k ? static_cast<int32_t>((team_len - k + t - 1) / t) : 0;
      s.virt_pos[v] = static_cast<int32_t>(i);
      slot_src[v] = a.start[r] + e0 + k;
      slot_stride[v] = static_cast<int32_t>(t);
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
  s.col.resize(s.padded_nnz);
  s.val.resize(s.padded_nnz);
  // One slice (32 slots, one contiguous block of the image) per work item:
  // padding and payload are written by the same thread.
  ParallelFor(num_slices, BuildThreads(), [&](int64_t sb, int64_t se, int) {
    for (int64_t sl = sb; sl < se; ++sl) {
      const int64_t w = (s.slice_ptr[sl + 1] - s.slice_ptr[sl]) / 32;
      for (int l = 0; l < 32; ++l) {
        const int64_t slot = sl * 32 + l;
        const int64_t src = slot_src[slot];
        const int32_t len = src < 0 ? 0 : s.slot_len[slot];
        const int64_t base = s.slice_ptr[sl] + l;
        const int64_t stride = slot_stride[slot];
        for (int32_t j = 0; j < len; ++j) {
          s.col[base + static_cast<int64_t>(j) * 32] = other_pos_of_row != nullptr ? (*other_pos_of_row)[a.idx[src + j * stride]] : a.idx[src + j * stride];
          s.val[base + static_cast<int64_t>(j) * 32] = a.val[src + j * stride];
        }
        for (int64_t j = len; j < w; ++j) {
          s.col[base + j * 32] = 0;
          s.val[base + j * 32] = 0.0;
        }
      }
    }
  });
}

}  // namespace

QpHost BuildQpHost(const PdlpProblemView& v, int64_t row_begin, int64_t row_end, int sigma, bool natural_primal_order) {
  const int64_t n = v.num_variables, m_full = v.num_constraints;
  if (row_begin < 0 || row_end > m_full || row_begin > row_end) throw std::runtime_error("bad row range");
  if (n >= (int64_t{1} << 31) - 64 || m_full >= (int64_t{1} << 31) - 64) throw std::runtime_error("dimension exceeds int32 index range");
  const int64_t m = row_end - row_begin;
  sigma = EnvInt("PDLP_B200_SIGMA", sigma);
  const int threads = BuildThreads();
  const bool trace = EnvInt("PDLP_B200_TRACE", 0) != 0;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[pdlp_b200 build] %-28s %.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
    t_last = now;
  };
  const bool whole = (row_begin == 0 && row_end == m_full);
  // ---- column-major copy restricted to the row block (rows renumbered from 0)
  Csr kt;  // "rows" of K^T = columns of K
  kt.start.assign(n + 1, 0);
  ParallelFor(n, threads, [&](int64_t cb, int64_t ce, int) {
    for (int64_t c = cb; c < ce; ++c) {
      const int64_t b = v.col_starts[c], e = v.col_starts[c + 1];
      if (e < b) throw std::runtime_error("col_starts is not monotone");
      int64_t cnt = 0;
      for (int64_t k = b; k < e; ++k) {
        const int64_t r = v.row_indices[k];
        if (r < 0 || r >= m_full) throw std::runtime_error("row index out of range");
        cnt += (whole || (r >= row_begin && r < row_end));
      }
      kt.start[c + 1] = cnt;
    }
  });
  lap("count columns");
  for (int64_t c = 0; c < n; ++c) kt.start[c + 1] += kt.start[c];
  const int64_t nnz = kt.start[n];
  kt.idx.resize(nnz);
  kt.val.resize(nnz);
  std::vector<int64_t> row_len(m, 0), col_len(n, 0);
  ParallelFor(n, threads, [&](int64_t cb, int64_t ce, int) {
    for (int64_t c = cb; c < ce; ++c) {
      int64_t p = kt.start[c];
      for (int64_t k = v.col_starts[c]; k < v.col_starts[c + 1]; ++k) {
        const int64_t r = v.row_indices[k];
        if (!whole && (r < row_begin || r >= row_end)) continue;
        kt.idx[p] = static_cast<int32_t>(r - row_begin);
        kt.val[p] = v.values[k];
        ++p;
      }
      col_len[c] = kt.start[c + 1] - kt.start[c];
    }
  });
  lap("column-major copy");
  // ---- row-major copy: every thread owns a contiguous range of rows and scans
  // the column-major entries in order, so the entries of a row end up in
  // ascending column order (like Eigen's transpose assignment) without atomics.
  Csr k;
  k.start.assign(m + 1, 0);
  k.idx.resize(nnz);
  k.val.resize(nnz);
  ParallelFor(m, threads, [&](int64_t rb, int64_t re, int) {
    for (int64_t p = 0; p < nnz; ++p) {
      const int64_t r = kt.idx[p];
      if (r >= rb && r < re) ++row_len[r];
    }
  });
  lap("row lengths");
  for (int64_t r = 0; r < m; ++r) k.start[r + 1] = k.start[r] + row_len[r];
  {
    std::vector<int64_t> pos(k.start.begin(), k.start.end() - 1);
    ParallelFor(m, threads, [&](int64_t rb, int64_t re, int) {
      for (int64_t c = 0; c < n; ++c)
        for (int64_t p = kt.start[c]; p < kt.start[c + 1]; ++p) {
          const int64_t r = kt.idx[p];
          if (r < rb || r >= re) continue;
          const int64_t q = pos[r]++;
          k.idx[q] = static_cast<int32_t>(c);
          k.val[q] = kt.val[p];
        }
    });
  }
  lap("row-major copy");
  QpHost out;
  out.n = n;
  out.m = m;
  out.nnz = nnz;
  out.has_q = v.objective_matrix_diagonal != nullptr;
  const int32_t split_len = ChooseSplitLen(nnz);
  AssignPositions(row_len, split_len, sigma, out.rows);
  AssignPositions(col_len, split_len, sigma, out.cols);
  // Column indices stored in one orientation are positions in the other
  // orientation's vector order. With natural_primal_order (row-sharded
  // multi-GPU) primal vectors stay in the caller's column order on every rank,
  // so K-by-rows stores plain column indices and the K^T product is scattered
  // through cols.row_of_pos by its epilogue.
  lap("assign positions");
  FillSell(k, natural_primal_order ? nullptr : &out.cols.pos_of_row, n, out.rows);
  lap("fill SELL rows");
  FillSell(kt, &out.rows.pos_of_row, m, out.cols);
  lap("fill SELL cols");
  return out;
}

}  // namespace pdlp_b200
