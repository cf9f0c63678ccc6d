// oracle/pdlp_cpu_core.h
//
// TEST INFRASTRUCTURE ONLY -- NOT PART OF THE PRODUCT.
// CPU restatement of the numerics of OR-Tools PDLP (reference checkout
// google/or-tools 9.15, ortools/pdlp/*). Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference leg may use it, and only as
// the checker / timed CPU baseline. The product (or-tools_b200/) never links,
// imports or calls anything in this directory.
//
// Parity status: PINNED against the reference's own known-answer tests
// (tests/test_kernel_goldens.py, test_solver_goldens.py, test_termination.py, test_params_validation.py and
// test_feasibility_polishing.py transcribe the vectors listed in SURVEY.md 8c). The
// reference itself cannot be built in this image (needs Eigen 3.4.0,
// abseil-cpp, protobuf + protoc, glop; none present, no network), so inner
// products follow the published semantics of Eigen 3.4.0 sparse^T * dense
// (storage-order accumulation per column) and shard-wise partial sums
// (sharder.cc:140-147); the exact SIMD summation order inside Eigen's dense
// reductions is not part of the contract (solvers.proto:287-296).
//
// Each function cites the reference file:line it restates.
#ifndef ORACLE_PDLP_CPU_CORE_H_
#define ORACLE_PDLP_CPU_CORE_H_

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <mutex>
#include <numeric>
#include <optional>
#include <random>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../include/pdlp_b200.h"  // POD structs of the boundary (interface only)

namespace pdlp_oracle {

using Vec = std::vector<double>;
constexpr double kInf = std::numeric_limits<double>::infinity();
inline double Sq(double v) { return v * v; }

// ---------------------------------------------------------------------------
// Thread pool: a barrier per ParallelFor, like GoogleThreadPoolScheduler
// (scheduler.h:50-73): the caller blocks until every index has run.
// ---------------------------------------------------------------------------
class ThreadPool {
 public:
  explicit ThreadPool(int num_threads) : num_threads_(std::max(1, num_threads)) {
    for (int t = 1; t < num_threads_; ++t) workers_.emplace_back([this] { WorkerLoop(); });
  }
  ~ThreadPool() {
    {
      std::lock_guard<std::mutex> lock(mu_);
      stop_ = true;
      ++generation_;
    }
    cv_start_.notify_all();
    for (auto& w : workers_) w.join();
  }
  int num_threads() const { return num_threads_; }

  void ParallelFor(int count, const std::function<void(int)>& fn) {
    if (count <= 0) return;
    if (num_threads_ == 1 || count == 1) {
      for (int i = 0; i < count; ++i) fn(i);
      return;
    }
    {
      std::lock_guard<std::mutex> lock(mu_);
      fn_ = &fn;
      count_ = count;
      next_.store(0, std::memory_order_relaxed);
      pending_workers_ = static_cast<int>(workers_.size());
      ++generation_;
    }
    cv_start_.notify_all();
    RunItems();
    std::unique_lock<std::mutex> lock(mu_);
    cv_done_.wait(lock, [this] { return pending_workers_ == 0; });
    fn_ = nullptr;
  }

 private:
  void RunItems() {
    for (;;) {
      const int i = next_.fetch_add(1, std::memory_order_relaxed);
      if (i >= count_) break;
      (*fn_)(i);
    }
  }
  void WorkerLoop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lock(mu_);
        cv_start_.wait(lock, [&] { return generation_ != seen; });
        seen = generation_;
        if (stop_) return;
      }
      RunItems();
      {
        std::lock_guard<std::mutex> lock(mu_);
        if (--pending_workers_ == 0) cv_done_.notify_one();
      }
    }
  }
  const int num_threads_;
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_start_, cv_done_;
  const std::function<void(int)>* fn_ = nullptr;
  int count_ = 0;
  std::atomic<int> next_{0};
  int pending_workers_ = 0;
  uint64_t generation_ = 0;
  bool stop_ = false;
};

// ---------------------------------------------------------------------------
// Sharder (sharder.h:34-231, sharder.cc:37-158)
// ---------------------------------------------------------------------------
class Sharder {
 public:
  Sharder() { starts_.push_back(0); }
  // Mass-balanced partition, sharder.cc:37-71.
  Sharder(int64_t num_elements, int num_shards, ThreadPool* pool,
          const std::function<int64_t(int64_t)>& element_mass)
      : pool_(pool) {
    if (num_elements == 0) {
      starts_.push_back(0);
      return;
    }
    int64_t overall_mass = 0;
    for (int64_t e = 0; e < num_elements; ++e) overall_mass += element_mass(e);
    starts_.push_back(0);
    int64_t this_shard_mass = element_mass(0);
    for (int64_t e = 1; e < num_elements; ++e) {
      const int64_t mass = element_mass(e);
      if (this_shard_mass + (mass / 2) >= overall_mass / num_shards) {
        masses_.push_back(this_shard_mass);
        starts_.push_back(e);
        this_shard_mass = mass;
      } else {
        this_shard_mass += mass;
      }
    }
    starts_.push_back(num_elements);
    masses_.push_back(this_shard_mass);
  }
  // Unit-mass partition, sharder.cc:73-102.
  Sharder(int64_t num_elements, int num_shards, ThreadPool* pool) : pool_(pool) {
    if (num_elements == 0) {
      starts_.push_back(0);
      return;
    }
    if (num_shards >= num_elements) {
      for (int64_t e = 0; e < num_elements; ++e) {
        starts_.push_back(e);
        masses_.push_back(1);
      }
    } else {
      for (int s = 0; s < num_shards; ++s) {
        const int64_t b = (num_elements * s) / num_shards;
        const int64_t e = (num_elements * (s + 1)) / num_shards;
        if (e - b > 0) {
          starts_.push_back(b);
          masses_.push_back(e - b);
        }
      }
    }
    starts_.push_back(num_elements);
  }
  // Same pool / shard count as `other`, different length (sharder.cc:104-108).
  Sharder(const Sharder& other, int64_t num_elements)
      : Sharder(num_elements, std::max(1, other.NumShards()), other.pool_) {}

  int NumShards() const { return static_cast<int>(starts_.size()) - 1; }
  int64_t NumElements() const { return starts_.back(); }
  int64_t ShardStart(int s) const { return starts_[s]; }
  int64_t ShardSize(int s) const { return starts_[s + 1] - starts_[s]; }
  int64_t ShardEnd(int s) const { return starts_[s + 1]; }
  int64_t ShardMass(int s) const { return masses_[s]; }
  const std::vector<int64_t>& starts() const { return starts_; }

  // fn(shard_index, begin, end); sharder.cc:110-138.
  void ForEachShard(const std::function<void(int, int64_t, int64_t)>& fn) const {
    const int n = NumShards();
    if (pool_ != nullptr && pool_->num_threads() > 1) {
      pool_->ParallelFor(n, [&](int s) { fn(s, starts_[s], starts_[s + 1]); });
    } else {
      for (int s = 0; s < n; ++s) fn(s, starts_[s], starts_[s + 1]);
    }
  }
  // Per-shard partials then a serial sum in shard order, sharder.cc:140-147.
  double SumOverShards(const std::function<double(int, int64_t, int64_t)>& fn) const {
    Vec partial(NumShards(), 0.0);
    ForEachShard([&](int s, int64_t b, int64_t e) { partial[s] = fn(s, b, e); });
    double sum = 0.0;
    for (double v : partial) sum += v;
    return sum;
  }
  double MaxOverShards(const std::function<double(int, int64_t, int64_t)>& fn) const {
    Vec partial(NumShards(), 0.0);
    ForEachShard([&](int s, int64_t b, int64_t e) { partial[s] = fn(s, b, e); });
    double mx = 0.0;
    for (double v : partial) mx = std::max(mx, std::abs(v));
    return mx;
  }
  bool TrueForAllShards(const std::function<bool(int, int64_t, int64_t)>& fn) const {
    std::vector<int> ok(NumShards(), 1);
    ForEachShard([&](int s, int64_t b, int64_t e) { ok[s] = fn(s, b, e) ? 1 : 0; });
    return std::all_of(ok.begin(), ok.end(), [](int v) { return v != 0; });
  }

 private:
  std::vector<int64_t> starts_;
  std::vector<int64_t> masses_;
  ThreadPool* pool_ = nullptr;
};

// Column-major sparse matrix with int64 indices: the layout of
// Eigen::SparseMatrix<double, ColMajor, int64_t> (quadratic_program.h:138).
struct SparseCsc {
  int64_t rows = 0, cols = 0;
  std::vector<int64_t> starts{0};  // [cols+1]
  std::vector<int64_t> index;      // row of each stored entry
  Vec value;
  int64_t nnz() const { return static_cast<int64_t>(value.size()); }
  int64_t ColNnz(int64_t c) const { return starts[c + 1] - starts[c]; }
};

// O(nnz) transpose (Eigen's SparseMatrix::transpose() assignment,
// sharded_quadratic_program.cc:83): entries of every output column end up
// sorted by row because the input is traversed column by column.
inline SparseCsc Transpose(const SparseCsc& a) {
  SparseCsc t;
  t.rows = a.cols;
  t.cols = a.rows;
  t.starts.assign(t.cols + 1, 0);
  for (int64_t k = 0; k < a.nnz(); ++k) t.starts[a.index[k] + 1]++;
  for (int64_t c = 0; c < t.cols; ++c) t.starts[c + 1] += t.starts[c];
  t.index.resize(a.nnz());
  t.value.resize(a.nnz());
  std::vector<int64_t> pos(t.starts.begin(), t.starts.end() - 1);
  for (int64_t c = 0; c < a.cols; ++c) {
    for (int64_t k = a.starts[c]; k < a.starts[c + 1]; ++k) {
      const int64_t p = pos[a.index[k]]++;
      t.index[p] = c;
      t.value[p] = a.value[k];
    }
  }
  return t;
}

inline Sharder MatrixSharder(const SparseCsc& m, int num_shards, ThreadPool* pool) {
  // sharder.h:161-165: element = column, mass = 1 + nnz(column).
  return Sharder(m.cols, num_shards, pool, [&m](int64_t c) { return 1 + m.ColNnz(c); });
}

// ---------------------------------------------------------------------------
// Vector operations (sharder.cc:160-332)
// ---------------------------------------------------------------------------
// answer = matrix^T * vector, one output per stored column; sharder.cc:160-173.
inline Vec TransposedMatrixVectorProduct(const SparseCsc& matrix, const Vec& vector,
                                         const Sharder& sharder) {
  Vec answer(matrix.cols);
  sharder.ForEachShard([&](int, int64_t b, int64_t e) {
    for (int64_t c = b; c < e; ++c) {
      double sum = 0.0;
      for (int64_t k = matrix.starts[c]; k < matrix.starts[c + 1]; ++k) {
        sum += matrix.value[k] * vector[matrix.index[k]];
      }
      answer[c] = sum;
    }
  });
  return answer;
}
inline void SetZero(const Sharder& sharder, Vec& dest) {
  dest.resize(sharder.NumElements());
  sharder.ForEachShard([&](int, int64_t b, int64_t e) { std::fill(dest.begin() + b, dest.begin() + e, 0.0); });
}
inline Vec ZeroVector(const Sharder& sharder) { Vec v; SetZero(sharder, v); return v; }
inline Vec OnesVector(const Sharder& sharder) {
  Vec v(sharder.NumElements());
  sharder.ForEachShard([&](int, int64_t b, int64_t e) { std::fill(v.begin() + b, v.begin() + e, 1.0); });
  return v;
}
inline void AddScaledVector(double scale, const Vec& increment, const Sharder& sharder, Vec& dest) {
  sharder.ForEachShard([&](int, int64_t b, int64_t e) { for (int64_t i = b; i < e; ++i) dest[i] += scale * increment[i]; });
}
inline void AssignVector(const Vec& vec, const Sharder& sharder, Vec& dest) {
  dest.resize(vec.size());
  sharder.ForEachShard([&](int, int64_t b, int64_t e) { std::copy(vec.begin() + b, vec.begin() + e, dest.begin() + b); });
}
inline Vec CloneVector(const Vec& vec, const Sharder& sharder) { Vec d; AssignVector(vec, sharder, d); return d; }
inline void CoefficientWiseProductInPlace(const Vec& scale, const Sharder& sharder, Vec& dest) {
  sharder.ForEachShard([&](int, int64_t b, int64_t e) { for (int64_t i = b; i < e; ++i) dest[i] = dest[i] * scale[i]; });
}
inline void CoefficientWiseQuotientInPlace(const Vec& scale, const Sharder& sharder, Vec& dest) {
  sharder.ForEachShard([&](int, int64_t b, int64_t e) { for (int64_t i = b; i < e; ++i) dest[i] = dest[i] / scale[i]; });
}
inline double Dot(const Vec& a, const Vec& b2, const Sharder& sharder) {
  return sharder.SumOverShards([&](int, int64_t b, int64_t e) { double s = 0; for (int64_t i = b; i < e; ++i) s += a[i] * b2[i]; return s; });
}
inline double LInfNorm(const Vec& v, const Sharder& sharder) {
  return sharder.MaxOverShards([&](int, int64_t b, int64_t e) { double m = 0; for (int64_t i = b; i < e; ++i) m = std::max(m, std::abs(v[i])); return m; });
}
inline double L1Norm(const Vec& v, const Sharder& sharder) {
  return sharder.SumOverShards([&](int, int64_t b, int64_t e) { double s = 0; for (int64_t i = b; i < e; ++i) s += std::abs(v[i]); return s; });
}
inline double SquaredNorm(const Vec& v, const Sharder& sharder) {
  return sharder.SumOverShards([&](int, int64_t b, int64_t e) { double s = 0; for (int64_t i = b; i < e; ++i) s += v[i] * v[i]; return s; });
}
inline double Norm(const Vec& v, const Sharder& sharder) { return std::sqrt(SquaredNorm(v, sharder)); }
inline double SquaredDistance(const Vec& a, const Vec& b2, const Sharder& sharder) {
  return sharder.SumOverShards([&](int, int64_t b, int64_t e) { double s = 0; for (int64_t i = b; i < e; ++i) s += Sq(a[i] - b2[i]); return s; });
}
inline double Distance(const Vec& a, const Vec& b, const Sharder& sharder) { return std::sqrt(SquaredDistance(a, b, sharder)); }
inline double ScaledLInfNorm(const Vec& v, const Vec& scale, const Sharder& sharder) {
  return sharder.MaxOverShards([&](int, int64_t b, int64_t e) { double m = 0; for (int64_t i = b; i < e; ++i) m = std::max(m, std::abs(v[i] * scale[i])); return m; });
}
inline double ScaledSquaredNorm(const Vec& v, const Vec& scale, const Sharder& sharder) {
  return sharder.SumOverShards([&](int, int64_t b, int64_t e) { double s = 0; for (int64_t i = b; i < e; ++i) s += Sq(v[i] * scale[i]); return s; });
}
inline double ScaledNorm(const Vec& v, const Vec& scale, const Sharder& sharder) { return std::sqrt(ScaledSquaredNorm(v, scale, sharder)); }

// sharder.cc:288-308
inline Vec ScaledColLInfNorm(const SparseCsc& matrix, const Vec& row_scaling, const Vec& col_scaling, const Sharder& sharder) {
  Vec answer(matrix.cols);
  sharder.ForEachShard([&](int, int64_t b, int64_t e) {
    for (int64_t c = b; c < e; ++c) {
      double mx = 0.0;
      for (int64_t k = matrix.starts[c]; k < matrix.starts[c + 1]; ++k) mx = std::max(mx, std::abs(matrix.value[k] * row_scaling[matrix.index[k]]));
      answer[c] = mx * std::abs(col_scaling[c]);
    }
  });
  return answer;
}
// sharder.cc:310-332
inline Vec ScaledColL2Norm(const SparseCsc& matrix, const Vec& row_scaling, const Vec& col_scaling, const Sharder& sharder) {
  Vec answer(matrix.cols);
  sharder.ForEachShard([&](int, int64_t b, int64_t e) {
    for (int64_t c = b; c < e; ++c) {
      double ss = 0.0;
      for (int64_t k = matrix.starts[c]; k < matrix.starts[c + 1]; ++k) ss += Sq(matrix.value[k] * row_scaling[matrix.index[k]]);
      answer[c] = std::sqrt(ss) * std::abs(col_scaling[c]);
    }
  });
  return answer;
}

// ---------------------------------------------------------------------------
// QuadraticProgram + ShardedQuadraticProgram
// (quadratic_program.h:61-151, sharded_quadratic_program.{h,cc})
// ---------------------------------------------------------------------------
struct QuadraticProgram {
  Vec objective_vector;
  std::optional<Vec> objective_matrix;  // diagonal
  SparseCsc constraint_matrix;
  Vec constraint_lower_bounds, constraint_upper_bounds;
  Vec variable_lower_bounds, variable_upper_bounds;
  std::optional<std::string> problem_name;
  double objective_offset = 0.0;
  double objective_scaling_factor = 1.0;
  double ApplyObjectiveScalingAndOffset(double objective) const {  // quadratic_program.h:130-132
    return objective_scaling_factor * (objective + objective_offset);
  }
};
inline bool IsLinearProgram(const QuadraticProgram& qp) { return !qp.objective_matrix.has_value(); }

class ShardedQp {
 public:
  // sharded_quadratic_program.cc:79-107 (the imbalance warning is log-only).
  ShardedQp(QuadraticProgram qp, int num_threads, int num_shards)
      : qp_(std::move(qp)),
        transposed_(Transpose(qp_.constraint_matrix)),
        pool_(num_threads == 1 ? nullptr : new ThreadPool(num_threads)),
        matrix_sharder_(MatrixSharder(qp_.constraint_matrix, num_shards, pool_.get())),
        transposed_sharder_(MatrixSharder(transposed_, num_shards, pool_.get())),
        primal_sharder_(static_cast<int64_t>(qp_.variable_lower_bounds.size()), num_shards, pool_.get()),
        dual_sharder_(static_cast<int64_t>(qp_.constraint_lower_bounds.size()), num_shards, pool_.get()) {}

  const QuadraticProgram& Qp() const { return qp_; }
  QuadraticProgram& MutableQp() { return qp_; }
  const SparseCsc& TransposedConstraintMatrix() const { return transposed_; }
  const Sharder& ConstraintMatrixSharder() const { return matrix_sharder_; }
  const Sharder& TransposedConstraintMatrixSharder() const { return transposed_sharder_; }
  const Sharder& PrimalSharder() const { return primal_sharder_; }
  const Sharder& DualSharder() const { return dual_sharder_; }
  int64_t PrimalSize() const { return static_cast<int64_t>(qp_.variable_lower_bounds.size()); }
  int64_t DualSize() const { return static_cast<int64_t>(qp_.constraint_lower_bounds.size()); }

  // sharded_quadratic_program.cc:114-181
  void RescaleQuadraticProgram(const Vec& col_scaling, const Vec& row_scaling) {
    const bool is_lp = IsLinearProgram(qp_);
    primal_sharder_.ForEachShard([&](int, int64_t b, int64_t e) {
      for (int64_t i = b; i < e; ++i) {
        qp_.objective_vector[i] = qp_.objective_vector[i] * col_scaling[i];
        qp_.variable_lower_bounds[i] = qp_.variable_lower_bounds[i] / col_scaling[i];
        qp_.variable_upper_bounds[i] = qp_.variable_upper_bounds[i] / col_scaling[i];
        if (!is_lp) (*qp_.objective_matrix)[i] = (*qp_.objective_matrix)[i] * (col_scaling[i] * col_scaling[i]);
      }
    });
    dual_sharder_.ForEachShard([&](int, int64_t b, int64_t e) {
      for (int64_t i = b; i < e; ++i) {
        qp_.constraint_lower_bounds[i] = qp_.constraint_lower_bounds[i] * row_scaling[i];
        qp_.constraint_upper_bounds[i] = qp_.constraint_upper_bounds[i] * row_scaling[i];
      }
    });
    ScaleMatrix(col_scaling, row_scaling, matrix_sharder_, qp_.constraint_matrix);
    ScaleMatrix(row_scaling, col_scaling, transposed_sharder_, transposed_);
  }
  // sharded_quadratic_program.cc:133-146, 183-189
  void ReplaceLargeConstraintBoundsWithInfinity(double threshold) {
    auto fix = [&](Vec& v) {
      dual_sharder_.ForEachShard([&](int, int64_t b, int64_t e) {
        for (int64_t i = b; i < e; ++i) {
          if (v[i] <= -threshold) v[i] = -kInf;
          if (v[i] >= threshold) v[i] = kInf;
        }
      });
    };
    fix(qp_.constraint_lower_bounds);
    fix(qp_.constraint_upper_bounds);
  }

 private:
  static void ScaleMatrix(const Vec& col_scaling, const Vec& row_scaling, const Sharder& sharder, SparseCsc& m) {
    sharder.ForEachShard([&](int, int64_t b, int64_t e) {
      for (int64_t c = b; c < e; ++c)
        for (int64_t k = m.starts[c]; k < m.starts[c + 1]; ++k) m.value[k] *= row_scaling[m.index[k]] * col_scaling[c];
    });
  }
  QuadraticProgram qp_;
  SparseCsc transposed_;
  std::unique_ptr<ThreadPool> pool_;
  Sharder matrix_sharder_, transposed_sharder_, primal_sharder_, dual_sharder_;
};

// ---------------------------------------------------------------------------
// sharded_optimization_utils
// ---------------------------------------------------------------------------
// ShardedWeightedAverage, sou.cc:43-79.
class WeightedAverage {
 public:
  explicit WeightedAverage(const Sharder* sharder) : sharder_(sharder) { average_ = ZeroVector(*sharder_); }
  void Add(const Vec& datapoint, double weight) {
    if (weight > 0.0) {
      const double ratio = weight / (sum_weights_ + weight);
      sharder_->ForEachShard([&](int, int64_t b, int64_t e) {
        for (int64_t i = b; i < e; ++i) average_[i] += ratio * (datapoint[i] - average_[i]);
      });
      sum_weights_ += weight;
    }
    ++num_terms_;
  }
  void Clear() { SetZero(*sharder_, average_); sum_weights_ = 0.0; num_terms_ = 0; }
  bool HasNonzeroWeight() const { return sum_weights_ > 0.0; }
  double Weight() const { return sum_weights_; }
  Vec ComputeAverage() const { return CloneVector(average_, *sharder_); }
  int NumTerms() const { return num_terms_; }

 private:
  Vec average_;
  double sum_weights_ = 0.0;
  int num_terms_ = 0;
  const Sharder* sharder_;
};

// sou.cc:83-92
inline double CombineBounds(double v1, double v2) {
  double mx = 0.0;
  if (std::abs(v1) < kInf) mx = std::abs(v1);
  if (std::abs(v2) < kInf) mx = std::max(mx, std::abs(v2));
  return mx;
}

// VectorInfoAccumulator, sou.cc:94-177.
struct InfoAcc {
  int64_t num_infinite = 0, num_zero = 0, num_finite_nonzero = 0;
  double max = -kInf, min = kInf, sum = 0.0, sum_squared = 0.0;
  void Add(double value) {
    if (std::isinf(value)) {
      ++num_infinite;
    } else if (value == 0) {
      ++num_zero;
    } else {
      ++num_finite_nonzero;
      const double a = std::abs(value);
      max = std::max(max, a);
      min = std::min(min, a);
      sum += a;
      sum_squared += a * a;
    }
  }
  void Merge(const InfoAcc& o) {
    num_infinite += o.num_infinite; num_zero += o.num_zero; num_finite_nonzero += o.num_finite_nonzero;
    max = std::max(max, o.max); min = std::min(min, o.min); sum += o.sum; sum_squared += o.sum_squared;
  }
};
struct VectorInfo {
  int64_t num_finite_nonzero = 0, num_infinite = 0, num_zero = 0;
  double largest = 0, smallest = 0, average = 0, l2_norm = 0;
};
inline VectorInfo Finish(const std::vector<InfoAcc>& parts) {
  InfoAcc t;
  for (const auto& p : parts) t.Merge(p);
  VectorInfo r;
  r.num_finite_nonzero = t.num_finite_nonzero; r.num_infinite = t.num_infinite; r.num_zero = t.num_zero;
  r.largest = t.num_finite_nonzero > 0 ? t.max : 0.0;
  r.smallest = t.num_finite_nonzero > 0 ? t.min : 0.0;
  r.average = (t.num_finite_nonzero + t.num_zero > 0) ? t.sum / static_cast<double>(t.num_finite_nonzero + t.num_zero)
                                                       : std::numeric_limits<double>::quiet_NaN();
  r.l2_norm = std::sqrt(t.sum_squared);
  return r;
}
template <class F>
VectorInfo InfoOver(const Sharder& sharder, F element) {
  std::vector<InfoAcc> parts(sharder.NumShards());
  sharder.ForEachShard([&](int s, int64_t b, int64_t e) { InfoAcc a; for (int64_t i = b; i < e; ++i) a.Add(element(i)); parts[s] = a; });
  return Finish(parts);
}
inline VectorInfo MatrixAbsElementInfo(const SparseCsc& m, const Sharder& sharder) {
  std::vector<InfoAcc> parts(sharder.NumShards());
  sharder.ForEachShard([&](int s, int64_t b, int64_t e) {
    InfoAcc a;
    for (int64_t k = m.starts[b]; k < m.starts[e]; ++k) a.Add(m.value[k]);
    parts[s] = a;
  });
  return Finish(parts);
}

// ComputeStats, sou.cc:270-343.
inline PdlpQuadraticProgramStats ComputeStats(const ShardedQp& sqp) {
  const QuadraticProgram& qp = sqp.Qp();
  const Vec ones_p = OnesVector(sqp.PrimalSharder()), ones_d = OnesVector(sqp.DualSharder());
  const Vec row_norms = ScaledColLInfNorm(sqp.TransposedConstraintMatrix(), ones_p, ones_d, sqp.TransposedConstraintMatrixSharder());
  const Vec col_norms = ScaledColLInfNorm(qp.constraint_matrix, ones_d, ones_p, sqp.ConstraintMatrixSharder());
  const VectorInfo row_info = InfoOver(sqp.DualSharder(), [&](int64_t i) { return row_norms[i]; });
  const VectorInfo col_info = InfoOver(sqp.PrimalSharder(), [&](int64_t i) { return col_norms[i]; });
  const VectorInfo mat = MatrixAbsElementInfo(qp.constraint_matrix, sqp.ConstraintMatrixSharder());
  const VectorInfo bounds = InfoOver(sqp.DualSharder(), [&](int64_t i) { return CombineBounds(qp.constraint_upper_bounds[i], qp.constraint_lower_bounds[i]); });
  const VectorInfo var_bounds = InfoOver(sqp.PrimalSharder(), [&](int64_t i) { return CombineBounds(qp.variable_upper_bounds[i], qp.variable_lower_bounds[i]); });
  const VectorInfo obj = InfoOver(sqp.PrimalSharder(), [&](int64_t i) { return qp.objective_vector[i]; });
  const VectorInfo gaps = InfoOver(sqp.PrimalSharder(), [&](int64_t i) { return qp.variable_upper_bounds[i] - qp.variable_lower_bounds[i]; });
  PdlpQuadraticProgramStats s;
  std::memset(&s, 0, sizeof(s));
  s.num_variables = sqp.PrimalSize();
  s.num_constraints = sqp.DualSize();
  s.constraint_matrix_col_min_l_inf_norm = col_info.smallest;
  s.constraint_matrix_row_min_l_inf_norm = row_info.smallest;
  s.constraint_matrix_num_nonzeros = mat.num_finite_nonzero;
  s.constraint_matrix_abs_max = mat.largest; s.constraint_matrix_abs_min = mat.smallest;
  s.constraint_matrix_abs_avg = mat.average; s.constraint_matrix_l2_norm = mat.l2_norm;
  s.combined_bounds_max = bounds.largest; s.combined_bounds_min = bounds.smallest;
  s.combined_bounds_avg = bounds.average; s.combined_bounds_l2_norm = bounds.l2_norm;
  s.combined_variable_bounds_max = var_bounds.largest; s.combined_variable_bounds_min = var_bounds.smallest;
  s.combined_variable_bounds_avg = var_bounds.average; s.combined_variable_bounds_l2_norm = var_bounds.l2_norm;
  s.variable_bound_gaps_num_finite = gaps.num_finite_nonzero + gaps.num_zero;
  s.variable_bound_gaps_max = gaps.largest; s.variable_bound_gaps_min = gaps.smallest;
  s.variable_bound_gaps_avg = gaps.average; s.variable_bound_gaps_l2_norm = gaps.l2_norm;
  s.objective_vector_abs_max = obj.largest; s.objective_vector_abs_min = obj.smallest;
  s.objective_vector_abs_avg = obj.average; s.objective_vector_l2_norm = obj.l2_norm;
  if (IsLinearProgram(qp)) {
    s.objective_matrix_num_nonzeros = 0;
    s.objective_matrix_abs_max = 0; s.objective_matrix_abs_min = 0;
    s.objective_matrix_abs_avg = std::numeric_limits<double>::quiet_NaN();
    s.objective_matrix_l2_norm = 0;
  } else {
    const Vec& q = *qp.objective_matrix;
    const VectorInfo qi = InfoOver(sqp.PrimalSharder(), [&](int64_t i) { return q[i]; });
    s.objective_matrix_num_nonzeros = qi.num_finite_nonzero;
    s.objective_matrix_abs_max = qi.largest; s.objective_matrix_abs_min = qi.smallest;
    s.objective_matrix_abs_avg = qi.average; s.objective_matrix_l2_norm = qi.l2_norm;
  }
  return s;
}

// sou.cc:354-405
enum class ScalingNorm { kL2, kLInf };
inline void DivideBySquareRootOfDivisor(const Vec& divisor, const Sharder& sharder, Vec& vector) {
  sharder.ForEachShard([&](int, int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) if (divisor[i] != 0) vector[i] /= std::sqrt(divisor[i]);
  });
}
inline void ApplyScalingIterationsForNorm(const ShardedQp& sqp, int num_iterations, ScalingNorm norm, Vec& row_scaling, Vec& col_scaling) {
  const QuadraticProgram& qp = sqp.Qp();
  for (int it = 0; it < num_iterations; ++it) {
    Vec col_norm, row_norm;
    if (norm == ScalingNorm::kL2) {
      col_norm = ScaledColL2Norm(qp.constraint_matrix, row_scaling, col_scaling, sqp.ConstraintMatrixSharder());
      row_norm = ScaledColL2Norm(sqp.TransposedConstraintMatrix(), col_scaling, row_scaling, sqp.TransposedConstraintMatrixSharder());
    } else {
      col_norm = ScaledColLInfNorm(qp.constraint_matrix, row_scaling, col_scaling, sqp.ConstraintMatrixSharder());
      row_norm = ScaledColLInfNorm(sqp.TransposedConstraintMatrix(), col_scaling, row_scaling, sqp.TransposedConstraintMatrixSharder());
    }
    DivideBySquareRootOfDivisor(col_norm, sqp.PrimalSharder(), col_scaling);
    DivideBySquareRootOfDivisor(row_norm, sqp.DualSharder(), row_scaling);
  }
}
inline void LInfRuizRescaling(const ShardedQp& sqp, int iters, Vec& row_scaling, Vec& col_scaling) {
  ApplyScalingIterationsForNorm(sqp, iters, ScalingNorm::kLInf, row_scaling, col_scaling);
}
inline void L2NormRescaling(const ShardedQp& sqp, Vec& row_scaling, Vec& col_scaling) {
  ApplyScalingIterationsForNorm(sqp, 1, ScalingNorm::kL2, row_scaling, col_scaling);
}
struct ScalingVectors { Vec row_scaling_vec, col_scaling_vec; };
// sou.cc:423-444
inline ScalingVectors ApplyRescaling(int l_inf_ruiz_iterations, bool l2_norm_rescaling, ShardedQp& sqp) {
  ScalingVectors sv{OnesVector(sqp.DualSharder()), OnesVector(sqp.PrimalSharder())};
  bool do_rescale = false;
  if (l_inf_ruiz_iterations > 0) { do_rescale = true; LInfRuizRescaling(sqp, l_inf_ruiz_iterations, sv.row_scaling_vec, sv.col_scaling_vec); }
  if (l2_norm_rescaling) { do_rescale = true; L2NormRescaling(sqp, sv.row_scaling_vec, sv.col_scaling_vec); }
  if (do_rescale) sqp.RescaleQuadraticProgram(sv.col_scaling_vec, sv.row_scaling_vec);
  return sv;
}

struct LagrangianPart { double value = 0.0; Vec gradient; };
// sou.cc:446-474
inline LagrangianPart ComputePrimalGradient(const ShardedQp& sqp, const Vec& primal, const Vec& dual_product) {
  LagrangianPart r; r.gradient.resize(sqp.PrimalSize());
  const QuadraticProgram& qp = sqp.Qp();
  r.value = sqp.PrimalSharder().SumOverShards([&](int, int64_t b, int64_t e) {
    double v = 0.0;
    if (IsLinearProgram(qp)) {
      for (int64_t i = b; i < e; ++i) { r.gradient[i] = qp.objective_vector[i] - dual_product[i]; v += primal[i] * r.gradient[i]; }
    } else {
      const Vec& q = *qp.objective_matrix;
      for (int64_t i = b; i < e; ++i) {
        const double op = q[i] * primal[i];
        r.gradient[i] = qp.objective_vector[i] + op - dual_product[i];
        v += primal[i] * (r.gradient[i] - 0.5 * op);
      }
    }
    return v;
  });
  return r;
}
// sou.cc:476-500
inline double DualSubgradientCoefficient(double lb, double ub, double dual, double primal_product) {
  if (dual < 0.0) return ub;
  if (dual > 0.0) return lb;
  if (std::isfinite(lb) && std::isfinite(ub)) {
    if (primal_product < lb) return lb;
    if (primal_product > ub) return ub;
    return primal_product;
  }
  if (std::isfinite(lb)) return lb;
  if (std::isfinite(ub)) return ub;
  return 0.0;
}
// sou.cc:502-527
inline LagrangianPart ComputeDualGradient(const ShardedQp& sqp, const Vec& dual, const Vec& primal_product) {
  LagrangianPart r; r.gradient.resize(sqp.DualSize());
  const QuadraticProgram& qp = sqp.Qp();
  r.value = sqp.DualSharder().SumOverShards([&](int, int64_t b, int64_t e) {
    double v = 0.0;
    for (int64_t i = b; i < e; ++i) {
      r.gradient[i] = DualSubgradientCoefficient(qp.constraint_lower_bounds[i], qp.constraint_upper_bounds[i], dual[i], primal_product[i]);
      v += r.gradient[i] * dual[i];
    }
    for (int64_t i = b; i < e; ++i) r.gradient[i] -= primal_product[i];
    return v;
  });
  return r;
}

// Power method, sou.cc:529-699. The reference draws the start vector with
// absl::Gaussian(std::mt19937); absl is absent here so the stream differs
// (std::normal_distribution) -- parity of this estimate is pinned only to the
// reference test's tolerance (sou_test.cc:546-557), and it is used only by
// CONSTANT_STEP_SIZE_RULE.
struct SingularValueAndIterations { double singular_value; int num_iterations; double estimated_relative_error; };
inline double PowerMethodFailureProbability(int64_t dimension, double epsilon, int k) {
  if (k < 2 || epsilon <= 0.0) return 1.0;
  return std::min(0.824, 0.354 / std::sqrt(epsilon * (k - 1))) * std::sqrt(static_cast<double>(dimension)) * std::pow(1.0 - epsilon, k - 0.5);
}
inline double NormalizeVector(const Sharder& sharder, Vec& v) {
  const double norm = Norm(v, sharder);
  if (norm != 0.0) sharder.ForEachShard([&](int, int64_t b, int64_t e) { for (int64_t i = b; i < e; ++i) v[i] /= norm; });
  return norm;
}
inline SingularValueAndIterations EstimateMaximumSingularValueOfConstraintMatrix(
    const ShardedQp& sqp, const std::optional<Vec>& primal_solution, const std::optional<Vec>& dual_solution,
    double desired_relative_error, double failure_probability, std::mt19937& gen) {
  const QuadraticProgram& qp = sqp.Qp();
  std::optional<Vec> p_ind, d_ind;
  if (primal_solution) {  // sou.cc:618-643
    p_ind.emplace(sqp.PrimalSize());
    for (int64_t i = 0; i < sqp.PrimalSize(); ++i)
      (*p_ind)[i] = ((*primal_solution)[i] == qp.variable_lower_bounds[i] || (*primal_solution)[i] == qp.variable_upper_bounds[i]) ? 0.0 : 1.0;
  }
  if (dual_solution) {  // sou.cc:647-674
    d_ind.emplace(sqp.DualSize());
    for (int64_t i = 0; i < sqp.DualSize(); ++i)
      (*d_ind)[i] = ((*dual_solution)[i] == 0.0 && (std::isinf(qp.constraint_lower_bounds[i]) || std::isinf(qp.constraint_upper_bounds[i]))) ? 0.0 : 1.0;
  }
  const int64_t dimension = qp.constraint_matrix.cols;
  Vec eigenvector(dimension);
  std::normal_distribution<double> gauss(0.0, 1.0);
  for (double& v : eigenvector) v = gauss(gen);
  if (p_ind) CoefficientWiseProductInPlace(*p_ind, sqp.PrimalSharder(), eigenvector);
  NormalizeVector(sqp.PrimalSharder(), eigenvector);
  double eigenvalue_estimate = 0.0;
  int num_iterations = 0;
  const double epsilon = 1.0 - Sq(1.0 - desired_relative_error);
  while (PowerMethodFailureProbability(dimension, epsilon, num_iterations) > failure_probability) {
    Vec dual_ev = TransposedMatrixVectorProduct(sqp.TransposedConstraintMatrix(), eigenvector, sqp.TransposedConstraintMatrixSharder());
    if (d_ind) CoefficientWiseProductInPlace(*d_ind, sqp.DualSharder(), dual_ev);
    Vec next = TransposedMatrixVectorProduct(qp.constraint_matrix, dual_ev, sqp.ConstraintMatrixSharder());
    if (p_ind) CoefficientWiseProductInPlace(*p_ind, sqp.PrimalSharder(), next);
    eigenvalue_estimate = Dot(eigenvector, next, sqp.PrimalSharder());
    eigenvector = std::move(next);
    ++num_iterations;
    NormalizeVector(sqp.PrimalSharder(), eigenvector);
  }
  return {std::sqrt(eigenvalue_estimate), num_iterations, desired_relative_error};
}

// sou.cc:701-723
inline bool HasValidBounds(const ShardedQp& sqp) {
  const QuadraticProgram& qp = sqp.Qp();
  const bool c_ok = sqp.DualSharder().TrueForAllShards([&](int, int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
      if (!(qp.constraint_lower_bounds[i] <= qp.constraint_upper_bounds[i] && qp.constraint_lower_bounds[i] < kInf && qp.constraint_upper_bounds[i] > -kInf)) return false;
    return true;
  });
  const bool v_ok = sqp.PrimalSharder().TrueForAllShards([&](int, int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
      if (!(qp.variable_lower_bounds[i] <= qp.variable_upper_bounds[i] && qp.variable_lower_bounds[i] < kInf && qp.variable_upper_bounds[i] > -kInf)) return false;
    return true;
  });
  return c_ok && v_ok;
}
// sou.cc:725-747
inline void ProjectToPrimalVariableBounds(const ShardedQp& sqp, Vec& primal, bool use_feasibility_bounds = false) {
  const QuadraticProgram& qp = sqp.Qp();
  sqp.PrimalSharder().ForEachShard([&](int, int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) {
      double ub = qp.variable_upper_bounds[i], lb = qp.variable_lower_bounds[i];
      if (use_feasibility_bounds) { ub = std::isfinite(ub) ? 0.0 : ub; lb = std::isfinite(lb) ? 0.0 : lb; }
      primal[i] = std::max(std::min(primal[i], ub), lb);
    }
  });
}
// sou.cc:749-770
inline void ProjectToDualVariableBounds(const ShardedQp& sqp, Vec& dual) {
  const QuadraticProgram& qp = sqp.Qp();
  sqp.DualSharder().ForEachShard([&](int, int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) {
      if (!std::isfinite(qp.constraint_upper_bounds[i])) dual[i] = std::max(dual[i], 0.0);
      if (!std::isfinite(qp.constraint_lower_bounds[i])) dual[i] = std::min(dual[i], 0.0);
    }
  });
}

// ---------------------------------------------------------------------------
// iteration_stats
// ---------------------------------------------------------------------------
struct ResidualNorms {
  double objective_correction = 0, objective_full_correction = 0;
  double l_inf_residual = 0, l_2_residual = 0, l_inf_componentwise_residual = 0;
};
// iteration_stats.cc:66-134
inline ResidualNorms PrimalResidualNorms(const ShardedQp& sqp, const Vec& row_scaling, const Vec& scaled_primal,
                                         double componentwise_offset, bool homogeneous_bounds = false) {
  const QuadraticProgram& qp = sqp.Qp();
  const Vec product = TransposedMatrixVectorProduct(sqp.TransposedConstraintMatrix(), scaled_primal, sqp.TransposedConstraintMatrixSharder());
  const int ns = sqp.DualSharder().NumShards();
  Vec linf(ns, 0.0), sumsq(ns, 0.0), cw(ns, 0.0);
  sqp.DualSharder().ForEachShard([&](int s, int64_t b, int64_t e) {
    double l = 0, q = 0, c = 0;
    for (int64_t i = b; i < e; ++i) {
      const double ub = (homogeneous_bounds && std::isfinite(qp.constraint_upper_bounds[i])) ? 0.0 : qp.constraint_upper_bounds[i];
      const double lb = (homogeneous_bounds && std::isfinite(qp.constraint_lower_bounds[i])) ? 0.0 : qp.constraint_lower_bounds[i];
      double scaled_residual = 0.0, residual_bound = 0.0;
      if (product[i] > ub) { scaled_residual = product[i] - ub; residual_bound = ub; }
      else if (product[i] < lb) { scaled_residual = lb - product[i]; residual_bound = lb; }
      const double residual = scaled_residual / row_scaling[i];
      l = std::max(l, residual);
      q += residual * residual;
      if (residual > 0.0) c = std::max(c, residual / (componentwise_offset + std::abs(residual_bound / row_scaling[i])));
    }
    linf[s] = l; sumsq[s] = q; cw[s] = c;
  });
  ResidualNorms r;
  for (int s = 0; s < ns; ++s) { r.l_inf_residual = std::max(r.l_inf_residual, std::abs(linf[s])); r.l_inf_componentwise_residual = std::max(r.l_inf_componentwise_residual, std::abs(cw[s])); }
  double t = 0; for (double v : sumsq) t += v;
  r.l_2_residual = std::sqrt(t);
  return r;
}
// iteration_stats.cc:136-177
inline bool TreatVariableBoundAsFinite(bool handle_as_residuals, double primal_value, double bound) {
  if (handle_as_residuals) return std::abs(primal_value - bound) <= std::abs(primal_value);
  return std::isfinite(bound);
}
inline double VariableBoundForDualObjective(double gradient, double lb, double ub) {
  const double primary = gradient >= 0.0 ? lb : ub;
  const double secondary = gradient >= 0.0 ? ub : lb;
  if (std::isfinite(primary)) return primary;
  if (std::isfinite(secondary)) return secondary;
  return 0.0;
}
// iteration_stats.cc:189-270
inline ResidualNorms DualResidualNorms(bool handle_as_residuals, const ShardedQp& sqp, const Vec& col_scaling,
                                       const Vec& scaled_primal, const Vec& scaled_gradient, double componentwise_offset) {
  const QuadraticProgram& qp = sqp.Qp();
  const int ns = sqp.PrimalSharder().NumShards();
  Vec corr(ns, 0.0), full(ns, 0.0), linf(ns, 0.0), sumsq(ns, 0.0), cw(ns, 0.0);
  sqp.PrimalSharder().ForEachShard([&](int s, int64_t b, int64_t e) {
    double dc = 0, dfc = 0, l = 0, q = 0, c = 0;
    for (int64_t i = b; i < e; ++i) {
      const double g = scaled_gradient[i];
      if (g == 0.0) continue;
      const double ub = qp.variable_upper_bounds[i], lb = qp.variable_lower_bounds[i];
      const double bound_for_rc = g > 0.0 ? lb : ub;
      dfc += bound_for_rc * g;
      const double eff_lb = TreatVariableBoundAsFinite(handle_as_residuals, scaled_primal[i], lb) ? lb : -kInf;
      const double eff_ub = TreatVariableBoundAsFinite(handle_as_residuals, scaled_primal[i], ub) ? ub : kInf;
      dc += VariableBoundForDualObjective(g, eff_lb, eff_ub) * g;
      const double eff_for_res = g > 0.0 ? eff_lb : eff_ub;
      if (std::isinf(eff_for_res)) {
        const double residual = std::abs(g) / col_scaling[i];
        l = std::max(l, residual);
        q += residual * residual;
        if (residual > 0.0) c = std::max(c, residual / (componentwise_offset + std::abs(qp.objective_vector[i] / col_scaling[i])));
      }
    }
    corr[s] = dc; full[s] = dfc; linf[s] = l; sumsq[s] = q; cw[s] = c;
  });
  ResidualNorms r;
  double t = 0;
  for (int s = 0; s < ns; ++s) {
    r.objective_correction += corr[s]; r.objective_full_correction += full[s];
    r.l_inf_residual = std::max(r.l_inf_residual, std::abs(linf[s]));
    r.l_inf_componentwise_residual = std::max(r.l_inf_componentwise_residual, std::abs(cw[s]));
    t += sumsq[s];
  }
  r.l_2_residual = std::sqrt(t);
  return r;
}
// iteration_stats.cc:273-297
inline Vec ObjectiveProduct(const ShardedQp& sqp, const Vec& primal) {
  Vec r(primal.size());
  if (IsLinearProgram(sqp.Qp())) { SetZero(sqp.PrimalSharder(), r); return r; }
  const Vec& q = *sqp.Qp().objective_matrix;
  sqp.PrimalSharder().ForEachShard([&](int, int64_t b, int64_t e) { for (int64_t i = b; i < e; ++i) r[i] = q[i] * primal[i]; });
  return r;
}
inline double QuadraticObjective(const ShardedQp& sqp, const Vec& primal, const Vec& objective_product) {
  return 0.5 * Dot(objective_product, primal, sqp.PrimalSharder());
}
// iteration_stats.cc:302-323
inline Vec PrimalGradientFromObjectiveProduct(const ShardedQp& sqp, const Vec& dual, Vec objective_product, bool use_zero_primal_objective = false) {
  const QuadraticProgram& qp = sqp.Qp();
  const SparseCsc& k = qp.constraint_matrix;
  sqp.ConstraintMatrixSharder().ForEachShard([&](int, int64_t b, int64_t e) {
    for (int64_t c = b; c < e; ++c) {
      double kty = 0.0;
      for (int64_t p = k.starts[c]; p < k.starts[c + 1]; ++p) kty += k.value[p] * dual[k.index[p]];
      if (use_zero_primal_objective) objective_product[c] = -kty;
      else objective_product[c] += qp.objective_vector[c] - kty;
    }
  });
  return objective_product;
}
// iteration_stats.cc:328-350
inline double DualObjectiveBoundsTerm(const ShardedQp& sqp, const Vec& dual) {
  const QuadraticProgram& qp = sqp.Qp();
  return sqp.DualSharder().SumOverShards([&](int, int64_t b, int64_t e) {
    double s = 0.0;
    for (int64_t i = b; i < e; ++i) {
      if (dual[i] > 0.0) s += qp.constraint_lower_bounds[i] * dual[i];
      else if (dual[i] < 0.0) s += qp.constraint_upper_bounds[i] * dual[i];
    }
    return s;
  });
}
// iteration_stats.cc:383-453
inline PdlpConvergenceInformation ComputeConvergenceInformation(
    bool handle_as_residuals, const ShardedQp& sqp, const Vec& col_scaling, const Vec& row_scaling,
    const Vec& scaled_primal, const Vec& scaled_dual, double cw_primal_offset, double cw_dual_offset, int candidate_type) {
  const QuadraticProgram& qp = sqp.Qp();
  PdlpConvergenceInformation r;
  std::memset(&r, 0, sizeof(r));
  const ResidualNorms pr = PrimalResidualNorms(sqp, row_scaling, scaled_primal, cw_primal_offset);
  r.l_inf_primal_residual = pr.l_inf_residual;
  r.l2_primal_residual = pr.l_2_residual;
  r.l_inf_componentwise_primal_residual = pr.l_inf_componentwise_residual;
  r.l_inf_primal_variable = ScaledLInfNorm(scaled_primal, col_scaling, sqp.PrimalSharder());
  r.l2_primal_variable = ScaledNorm(scaled_primal, col_scaling, sqp.PrimalSharder());
  r.l_inf_dual_variable = ScaledLInfNorm(scaled_dual, row_scaling, sqp.DualSharder());
  r.l2_dual_variable = ScaledNorm(scaled_dual, row_scaling, sqp.DualSharder());
  Vec objective_product = ObjectiveProduct(sqp, scaled_primal);
  const double quadratic_objective = QuadraticObjective(sqp, scaled_primal, objective_product);
  const Vec gradient = PrimalGradientFromObjectiveProduct(sqp, scaled_dual, std::move(objective_product));
  r.primal_objective = qp.ApplyObjectiveScalingAndOffset(quadratic_objective + Dot(qp.objective_vector, scaled_primal, sqp.PrimalSharder()));
  const double dual_objective_piece = -quadratic_objective + DualObjectiveBoundsTerm(sqp, scaled_dual);
  const ResidualNorms dr = DualResidualNorms(handle_as_residuals, sqp, col_scaling, scaled_primal, gradient, cw_dual_offset);
  r.dual_objective = qp.ApplyObjectiveScalingAndOffset(dual_objective_piece + dr.objective_correction);
  r.corrected_dual_objective = qp.ApplyObjectiveScalingAndOffset(dual_objective_piece + dr.objective_full_correction);
  r.l_inf_dual_residual = dr.l_inf_residual;
  r.l2_dual_residual = dr.l_2_residual;
  r.l_inf_componentwise_dual_residual = dr.l_inf_componentwise_residual;
  r.candidate_type = candidate_type;
  return r;
}
// iteration_stats.cc:486-564
inline PdlpInfeasibilityInformation ComputeInfeasibilityInformation(
    bool handle_as_residuals, const ShardedQp& sqp, const Vec& col_scaling, const Vec& row_scaling,
    const Vec& scaled_primal_ray, const Vec& scaled_dual_ray, const Vec& primal_for_residual_tests, int candidate_type) {
  const QuadraticProgram& qp = sqp.Qp();
  const double l_inf_primal = ScaledLInfNorm(scaled_primal_ray, col_scaling, sqp.PrimalSharder());
  const double l_inf_dual = ScaledLInfNorm(scaled_dual_ray, row_scaling, sqp.DualSharder());
  PdlpInfeasibilityInformation r;
  std::memset(&r, 0, sizeof(r));
  const Vec gradient = PrimalGradientFromObjectiveProduct(sqp, scaled_dual_ray, ZeroVector(sqp.PrimalSharder()), /*use_zero_primal_objective=*/true);
  const ResidualNorms dr = DualResidualNorms(handle_as_residuals, sqp, col_scaling, primal_for_residual_tests, gradient, 0.0);
  const double dual_ray_objective = DualObjectiveBoundsTerm(sqp, scaled_dual_ray) + dr.objective_correction;
  if (l_inf_dual > 0) {
    r.dual_ray_objective = dual_ray_objective / l_inf_dual;
    r.max_dual_ray_infeasibility = dr.l_inf_residual / l_inf_dual;
  }
  const ResidualNorms pr = PrimalResidualNorms(sqp, row_scaling, scaled_primal_ray, 0.0, /*homogeneous_bounds=*/true);
  if (l_inf_primal > 0.0) {
    const Vec op = ObjectiveProduct(sqp, scaled_primal_ray);
    r.primal_ray_quadratic_norm = LInfNorm(op, sqp.PrimalSharder()) / l_inf_primal;
    r.max_primal_ray_infeasibility = pr.l_inf_residual / l_inf_primal;
    r.primal_ray_linear_objective = Dot(scaled_primal_ray, qp.objective_vector, sqp.PrimalSharder()) / l_inf_primal;
  }
  r.candidate_type = candidate_type;
  return r;
}
// iteration_stats.cc:579-593
inline Vec ReducedCosts(const ShardedQp& sqp, const Vec& primal, const Vec& dual, bool use_zero_primal_objective) {
  Vec op = use_zero_primal_objective ? ZeroVector(sqp.PrimalSharder()) : ObjectiveProduct(sqp, primal);
  return PrimalGradientFromObjectiveProduct(sqp, dual, std::move(op), use_zero_primal_objective);
}

// ---------------------------------------------------------------------------
// termination (termination.cc)
// ---------------------------------------------------------------------------
struct DetailedCriteria {
  double primal_abs, primal_rel, dual_abs, dual_rel, gap_abs, gap_rel;
};
// termination.cc:126-159
inline DetailedCriteria EffectiveOptimalityCriteria(const PdlpTerminationCriteria& c) {
  if (c.optimality_criteria_case == PDLP_DETAILED_OPTIMALITY_CRITERIA) {
    return {c.eps_optimal_primal_residual_absolute, c.eps_optimal_primal_residual_relative, c.eps_optimal_dual_residual_absolute,
            c.eps_optimal_dual_residual_relative, c.eps_optimal_objective_gap_absolute, c.eps_optimal_objective_gap_relative};
  }
  double a, r;
  if (c.optimality_criteria_case == PDLP_SIMPLE_OPTIMALITY_CRITERIA) { a = c.simple_eps_optimal_absolute; r = c.simple_eps_optimal_relative; }
  else { a = c.eps_optimal_absolute; r = c.eps_optimal_relative; }
  return {a, r, a, r, a, r};
}
// termination.cc:26-41
inline bool ObjectiveGapMet(const DetailedCriteria& oc, const PdlpConvergenceInformation& s) {
  if (std::isinf(oc.gap_abs) || std::isinf(oc.gap_rel)) return true;
  const double abs_obj = std::abs(s.primal_objective) + std::abs(s.dual_objective);
  const double gap = std::abs(s.primal_objective - s.dual_objective);
  return std::isfinite(abs_obj) && gap <= oc.gap_abs + oc.gap_rel * abs_obj;
}
// termination.cc:43-97
inline bool OptimalityCriteriaMet(const DetailedCriteria& oc, const PdlpConvergenceInformation& s, int norm, const PdlpBoundNorms& bn) {
  double perr = 0, pbase = 0, derr = 0, dbase = 0;
  double pabs = oc.primal_abs, dabs = oc.dual_abs;
  switch (norm) {
    case PDLP_OPTIMALITY_NORM_L_INF:
      perr = s.l_inf_primal_residual; pbase = bn.l_inf_norm_constraint_bounds; derr = s.l_inf_dual_residual; dbase = bn.l_inf_norm_primal_linear_objective; break;
    case PDLP_OPTIMALITY_NORM_L2:
      perr = s.l2_primal_residual; pbase = bn.l2_norm_constraint_bounds; derr = s.l2_dual_residual; dbase = bn.l2_norm_primal_linear_objective; break;
    case PDLP_OPTIMALITY_NORM_L_INF_COMPONENTWISE:
      perr = s.l_inf_componentwise_primal_residual; pbase = 1.0; pabs = 0.0; derr = s.l_inf_componentwise_dual_residual; dbase = 1.0; dabs = 0.0; break;
    default: break;
  }
  const bool p_ok = std::isinf(oc.primal_abs) || std::isinf(oc.primal_rel) || perr <= pabs + oc.primal_rel * pbase;
  const bool d_ok = std::isinf(oc.dual_abs) || std::isinf(oc.dual_rel) || derr <= dabs + oc.dual_rel * dbase;
  return p_ok && d_ok && ObjectiveGapMet(oc, s);
}
// termination.cc:104-122
inline bool PrimalInfeasibilityCriteriaMet(double eps, const PdlpInfeasibilityInformation& s) {
  if (s.dual_ray_objective <= 0.0) return false;
  return s.max_dual_ray_infeasibility / s.dual_ray_objective <= eps;
}
inline bool DualInfeasibilityCriteriaMet(double eps, const PdlpInfeasibilityInformation& s) {
  if (s.primal_ray_linear_objective >= 0.0) return false;
  return (s.max_primal_ray_infeasibility / -s.primal_ray_linear_objective <= eps) &&
         (s.primal_ray_quadratic_norm / -s.primal_ray_linear_objective <= eps);
}
struct TerminationReasonAndPointType { int reason; int type; };
// termination.cc:161-184
inline std::optional<TerminationReasonAndPointType> CheckSimpleTerminationCriteria(
    const PdlpTerminationCriteria& c, const PdlpIterationStats& stats, const volatile int32_t* interrupt) {
  if (stats.iteration_number >= c.iteration_limit) return TerminationReasonAndPointType{PDLP_TERMINATION_REASON_ITERATION_LIMIT, PDLP_POINT_TYPE_NONE};
  if (stats.cumulative_kkt_matrix_passes >= c.kkt_matrix_pass_limit) return TerminationReasonAndPointType{PDLP_TERMINATION_REASON_KKT_MATRIX_PASS_LIMIT, PDLP_POINT_TYPE_NONE};
  if (stats.cumulative_time_sec >= c.time_sec_limit) return TerminationReasonAndPointType{PDLP_TERMINATION_REASON_TIME_LIMIT, PDLP_POINT_TYPE_NONE};
  if (interrupt != nullptr && *interrupt != 0) return TerminationReasonAndPointType{PDLP_TERMINATION_REASON_INTERRUPTED_BY_USER, PDLP_POINT_TYPE_NONE};
  return std::nullopt;
}
// termination.cc:186-219
inline std::optional<TerminationReasonAndPointType> CheckIterateTerminationCriteria(
    const PdlpTerminationCriteria& c, const PdlpIterationStats& stats, const PdlpBoundNorms& bn, bool force_numerical_termination) {
  const DetailedCriteria oc = EffectiveOptimalityCriteria(c);
  for (int i = 0; i < stats.num_convergence_information; ++i) {
    if (OptimalityCriteriaMet(oc, stats.convergence_information[i], c.optimality_norm, bn))
      return TerminationReasonAndPointType{PDLP_TERMINATION_REASON_OPTIMAL, stats.convergence_information[i].candidate_type};
  }
  for (int i = 0; i < stats.num_infeasibility_information; ++i) {
    const auto& inf = stats.infeasibility_information[i];
    if (PrimalInfeasibilityCriteriaMet(c.eps_primal_infeasible, inf)) return TerminationReasonAndPointType{PDLP_TERMINATION_REASON_PRIMAL_INFEASIBLE, inf.candidate_type};
    if (DualInfeasibilityCriteriaMet(c.eps_dual_infeasible, inf)) return TerminationReasonAndPointType{PDLP_TERMINATION_REASON_DUAL_INFEASIBLE, inf.candidate_type};
  }
  if (force_numerical_termination) return TerminationReasonAndPointType{PDLP_TERMINATION_REASON_NUMERICAL_ERROR, PDLP_POINT_TYPE_NONE};
  return std::nullopt;
}
// termination.cc:221-237
inline PdlpBoundNorms BoundNormsFromProblemStats(const PdlpQuadraticProgramStats& s) {
  return {s.objective_vector_l2_norm, s.combined_bounds_l2_norm, s.objective_vector_abs_max, s.combined_bounds_max};
}
inline double EpsilonRatio(double eps_abs, double eps_rel) { return (eps_abs == eps_rel) ? 1.0 : eps_abs / eps_rel; }
struct RelativeConvergenceInformation {
  double relative_l_inf_primal_residual = 0, relative_l2_primal_residual = 0, relative_l_inf_dual_residual = 0,
         relative_l2_dual_residual = 0, relative_optimality_gap = 0;
};
// termination.cc:239-271
inline RelativeConvergenceInformation ComputeRelativeResiduals(const DetailedCriteria& oc, const PdlpConvergenceInformation& s, const PdlpBoundNorms& bn) {
  const double rp = EpsilonRatio(oc.primal_abs, oc.primal_rel), rd = EpsilonRatio(oc.dual_abs, oc.dual_rel), rg = EpsilonRatio(oc.gap_abs, oc.gap_rel);
  RelativeConvergenceInformation info;
  info.relative_l_inf_primal_residual = s.l_inf_primal_residual / (rp + bn.l_inf_norm_constraint_bounds);
  info.relative_l2_primal_residual = s.l2_primal_residual / (rp + bn.l2_norm_constraint_bounds);
  info.relative_l_inf_dual_residual = s.l_inf_dual_residual / (rd + bn.l_inf_norm_primal_linear_objective);
  info.relative_l2_dual_residual = s.l2_dual_residual / (rd + bn.l2_norm_primal_linear_objective);
  const double abs_obj = std::abs(s.primal_objective) + std::abs(s.dual_objective);
  info.relative_optimality_gap = (s.primal_objective - s.dual_objective) / (rg + abs_obj);
  return info;
}

// ---------------------------------------------------------------------------
// trust_region
// ---------------------------------------------------------------------------
// A trust-region problem is given by five accessors (trust_region.cc:40-55).
struct TrProblem {
  std::function<double(int64_t)> Objective, LowerBound, UpperBound, CenterPoint, NormWeight;
};
// trust_region.h:193-228
inline double DistanceAtCriticalStepSize(const TrProblem& p, int64_t i) {
  const double o = p.Objective(i);
  if (o == 0.0) return 0.0;
  return (o > 0.0 ? p.LowerBound(i) : p.UpperBound(i)) - p.CenterPoint(i);
}
inline double CriticalStepSize(const TrProblem& p, int64_t i) {
  const double o = p.Objective(i);
  if (o == 0.0) return kInf;
  return -p.NormWeight(i) * DistanceAtCriticalStepSize(p, i) / o;
}
inline double ProjectedValue(const TrProblem& p, int64_t i, double step_size) {
  const double full_step = p.CenterPoint(i) - step_size * p.Objective(i) / p.NormWeight(i);
  return std::min(std::max(full_step, p.LowerBound(i)), p.UpperBound(i));
}
// trust_region.h:234-246: median via nth_element at position size/2.
template <class T, class F>
double EasyMedian(std::vector<T> array, F value) {
  auto middle = array.begin() + (array.size() / 2);
  std::nth_element(array.begin(), middle, array.end(), [&](const T& l, const T& r) { return value(l) < value(r); });
  return value(*middle);
}
struct TrStepResult { double solution_step_size; double objective_value; };

// SolveTrustRegionStepSize, trust_region.cc:332-449 (median-of-shard-medians
// threshold search, then the closed-form step).
inline TrStepResult SolveTrustRegionStepSize(const TrProblem& problem, double target_radius, const Sharder& sharder) {
  if (target_radius == 0.0) return {0.0, 0.0};
  const bool all_zero = sharder.TrueForAllShards([&](int, int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) if (problem.Objective(i) != 0.0) return false;
    return true;
  });
  if (all_zero) return {0.0, 0.0};
  const int ns = sharder.NumShards();
  std::vector<std::vector<int64_t>> undecided(ns);
  // ComputeInitialState, trust_region.cc:210-226 + trust_region.h:252-272.
  double variable_radius_coefficient = sharder.SumOverShards([&](int s, int64_t b, int64_t e) {
    double coef = 0.0;
    undecided[s].clear();
    for (int64_t i = b; i < e; ++i) {
      if (std::isfinite(CriticalStepSize(problem, i))) undecided[s].push_back(i);
      else coef += Sq(problem.Objective(i)) / problem.NormWeight(i);
    }
    return coef;
  });
  double fixed_radius_squared = 0.0;
  auto num_undecided = [&] { int64_t t = 0; for (auto& u : undecided) t += static_cast<int64_t>(u.size()); return t; };
  while (num_undecided() > 0) {
    // MedianOfShardMedians, trust_region.cc:177-203.
    std::vector<std::optional<double>> shard_medians(ns);
    sharder.ForEachShard([&](int s, int64_t, int64_t) {
      if (!undecided[s].empty()) shard_medians[s] = EasyMedian(undecided[s], [&](int64_t i) { return CriticalStepSize(problem, i); });
    });
    std::vector<double> medians;
    for (auto& m : shard_medians) if (m.has_value()) medians.push_back(*m);
    const double threshold = EasyMedian(medians, [](double x) { return x; });
    const double radius_sq_undecided = sharder.SumOverShards([&](int s, int64_t, int64_t) {
      double sum = 0.0;
      for (int64_t i : undecided[s]) sum += problem.NormWeight(i) * Sq(ProjectedValue(problem, i, threshold) - problem.CenterPoint(i));
      return sum;
    });
    const double radius_sq_at_threshold = radius_sq_undecided + fixed_radius_squared + variable_radius_coefficient * Sq(threshold);
    if (radius_sq_at_threshold > Sq(target_radius)) {
      // RemoveCriticalStepsAboveThreshold, trust_region.h:290-311.
      variable_radius_coefficient += sharder.SumOverShards([&](int s, int64_t, int64_t) {
        double coef = 0.0;
        auto& u = undecided[s];
        for (int64_t i : u) if (CriticalStepSize(problem, i) >= threshold) coef += Sq(problem.Objective(i)) / problem.NormWeight(i);
        u.erase(std::remove_if(u.begin(), u.end(), [&](int64_t i) { return CriticalStepSize(problem, i) >= threshold; }), u.end());
        return coef;
      });
    } else {
      // RemoveCriticalStepsBelowThreshold, trust_region.h:317-337.
      fixed_radius_squared += sharder.SumOverShards([&](int s, int64_t, int64_t) {
        double rs = 0.0;
        auto& u = undecided[s];
        for (int64_t i : u) if (CriticalStepSize(problem, i) <= threshold) rs += problem.NormWeight(i) * Sq(DistanceAtCriticalStepSize(problem, i));
        u.erase(std::remove_if(u.begin(), u.end(), [&](int64_t i) { return CriticalStepSize(problem, i) <= threshold; }), u.end());
        return rs;
      });
    }
  }
  double step_size = 0.0;
  if (variable_radius_coefficient > 0.0) step_size = std::sqrt((Sq(target_radius) - fixed_radius_squared) / variable_radius_coefficient);
  else step_size = std::numeric_limits<double>::max();
  // ComputeObjectiveValue, trust_region.cc:288-304.
  const double objective_value = sharder.SumOverShards([&](int, int64_t b, int64_t e) {
    double v = 0.0;
    for (int64_t i = b; i < e; ++i) v += problem.Objective(i) * (ProjectedValue(problem, i, step_size) - problem.CenterPoint(i));
    return v;
  });
  return {step_size, objective_value};
}
inline Vec ComputeSolution(const TrProblem& problem, double step_size, const Sharder& sharder) {
  Vec sol(sharder.NumElements());
  sharder.ForEachShard([&](int, int64_t b, int64_t e) { for (int64_t i = b; i < e; ++i) sol[i] = ProjectedValue(problem, i, step_size); });
  return sol;
}
struct TrustRegionResult { double solution_step_size; double objective_value; Vec solution; };
inline TrProblem VectorTrProblem(const Vec& objective, const Vec& lb, const Vec& ub, const Vec& center, const Vec& weights) {
  return TrProblem{[&](int64_t i) { return objective[i]; }, [&](int64_t i) { return lb[i]; }, [&](int64_t i) { return ub[i]; },
                   [&](int64_t i) { return center[i]; }, [&](int64_t i) { return weights[i]; }};
}
// trust_region.cc:453-471
inline TrustRegionResult SolveTrustRegion(const Vec& objective, const Vec& lb, const Vec& ub, const Vec& center, const Vec& weights,
                                          double target_radius, const Sharder& sharder) {
  const TrProblem p = VectorTrProblem(objective, lb, ub, center, weights);
  const TrStepResult s = SolveTrustRegionStepSize(p, target_radius, sharder);
  return {s.solution_step_size, s.objective_value, ComputeSolution(p, s.solution_step_size, sharder)};
}

// Diagonal-QP trust region (bisection), trust_region.cc:611-753.
struct DiagTrProblem : TrProblem { std::function<double(int64_t)> ObjectiveMatrixDiagonalAt; };
inline double ProjectedValueOfScaledDifference(const DiagTrProblem& p, int64_t i, double scaling_factor) {
  const double w = p.NormWeight(i);
  return std::min(std::max((-p.Objective(i) / std::sqrt(w)) / (p.ObjectiveMatrixDiagonalAt(i) / w + scaling_factor),
                           std::sqrt(w) * (p.LowerBound(i) - p.CenterPoint(i))),
                  std::sqrt(w) * (p.UpperBound(i) - p.CenterPoint(i)));
}
inline double NormOfDeltaProjection(const DiagTrProblem& p, const Sharder& sharder, double scaling_factor) {
  return std::sqrt(sharder.SumOverShards([&](int, int64_t b, int64_t e) {
    double s = 0.0;
    for (int64_t i = b; i < e; ++i) s += Sq(ProjectedValueOfScaledDifference(p, i, scaling_factor));
    return s;
  }));
}
inline double FindScalingFactor(const DiagTrProblem& p, const Sharder& sharder, double target_radius, double solve_tol) {
  double lo = 0.0, hi = 1.0;
  while (NormOfDeltaProjection(p, sharder, hi) >= target_radius) { lo = hi; hi *= 2; }
  while ((hi - lo) >= solve_tol * std::max(1.0, lo)) {
    const double mid = (lo + hi) / 2.0;
    if (NormOfDeltaProjection(p, sharder, mid) <= target_radius) hi = mid; else lo = mid;
  }
  return (hi + lo) / 2.0;
}
inline TrustRegionResult SolveDiagonalTrustRegionProblem(const DiagTrProblem& p, const Sharder& sharder, double target_radius, double solve_tol) {
  const int64_t n = sharder.NumElements();
  if (target_radius == 0.0) {
    Vec sol(n);
    for (int64_t i = 0; i < n; ++i) sol[i] = p.CenterPoint(i);
    return {0.0, 0.0, std::move(sol)};
  }
  const double scaling = FindScalingFactor(p, sharder, target_radius, solve_tol);
  Vec sol(n);
  sharder.ForEachShard([&](int, int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) {
      const double w = p.NormWeight(i);
      sol[i] = p.CenterPoint(i) + std::sqrt(1 / w) * ProjectedValueOfScaledDifference(p, i, scaling);
    }
  });
  const double value = sharder.SumOverShards([&](int, int64_t b, int64_t e) {
    double s = 0.0;
    for (int64_t i = b; i < e; ++i) { const double d = sol[i] - p.CenterPoint(i); s += 0.5 * d * p.ObjectiveMatrixDiagonalAt(i) * d + d * p.Objective(i); }
    return s;
  });
  return {scaling, value, std::move(sol)};
}
inline TrustRegionResult SolveDiagonalTrustRegion(const Vec& objective, const Vec& qdiag, const Vec& lb, const Vec& ub, const Vec& center,
                                                  const Vec& weights, double target_radius, const Sharder& sharder, double tol) {
  DiagTrProblem p;
  static_cast<TrProblem&>(p) = VectorTrProblem(objective, lb, ub, center, weights);
  p.ObjectiveMatrixDiagonalAt = [&](int64_t i) { return qdiag[i]; };
  return SolveDiagonalTrustRegionProblem(p, sharder, target_radius, tol);
}

// JointTrustRegionProblem / DiagonalTrustRegionProblemFromQp,
// trust_region.cc:115-162, 538-607.
inline DiagTrProblem JointTrProblem(const QuadraticProgram& qp, const Vec& primal, const Vec& dual, const Vec& primal_gradient,
                                    const Vec& dual_gradient, double primal_weight) {
  const int64_t n = static_cast<int64_t>(primal.size());
  DiagTrProblem p;
  p.Objective = [&, n](int64_t i) { return i < n ? primal_gradient[i] : -dual_gradient[i - n]; };
  p.LowerBound = [&, n](int64_t i) { return i < n ? qp.variable_lower_bounds[i] : (std::isfinite(qp.constraint_upper_bounds[i - n]) ? -kInf : 0.0); };
  p.UpperBound = [&, n](int64_t i) { return i < n ? qp.variable_upper_bounds[i] : (std::isfinite(qp.constraint_lower_bounds[i - n]) ? kInf : 0.0); };
  p.CenterPoint = [&, n](int64_t i) { return i < n ? primal[i] : dual[i - n]; };
  p.NormWeight = [n, primal_weight](int64_t i) { return i < n ? 0.5 * primal_weight : 0.5 / primal_weight; };
  p.ObjectiveMatrixDiagonalAt = [&, n](int64_t i) { return (qp.objective_matrix.has_value() && i < n) ? (*qp.objective_matrix)[i] : 0.0; };
  return p;
}

struct LocalizedLagrangianBounds { double lagrangian_value, lower_bound, upper_bound, radius; };
inline double BoundGap(const LocalizedLagrangianBounds& b) { return b.upper_bound - b.lower_bound; }
enum class PrimalDualNorm { kMaxNorm, kEuclideanNorm };

// trust_region.cc:755-1016
inline LocalizedLagrangianBounds ComputeLocalizedLagrangianBounds(
    const ShardedQp& sqp, const Vec& primal, const Vec& dual, PrimalDualNorm norm, double primal_weight, double radius,
    const Vec* primal_product, const Vec* dual_product, bool use_diagonal_qp_solver, double diagonal_tol) {
  const QuadraticProgram& qp = sqp.Qp();
  Vec pp_store, dp_store;
  if (primal_product == nullptr) {
    pp_store = TransposedMatrixVectorProduct(sqp.TransposedConstraintMatrix(), primal, sqp.TransposedConstraintMatrixSharder());
    primal_product = &pp_store;
  }
  if (dual_product == nullptr) {
    dp_store = TransposedMatrixVectorProduct(qp.constraint_matrix, dual, sqp.ConstraintMatrixSharder());
    dual_product = &dp_store;
  }
  const LagrangianPart primal_part = ComputePrimalGradient(sqp, primal, *dual_product);
  const LagrangianPart dual_part = ComputeDualGradient(sqp, dual, *primal_product);
  const double lagrangian_value = primal_part.value + dual_part.value;
  if (norm == PrimalDualNorm::kMaxNorm) {  // trust_region.cc:855-884
    const double primal_radius = std::sqrt(2) * radius / std::sqrt(primal_weight);
    const double dual_radius = std::sqrt(2) * radius * std::sqrt(primal_weight);
    TrProblem pprob{[&](int64_t i) { return primal_part.gradient[i]; }, [&](int64_t i) { return qp.variable_lower_bounds[i]; },
                    [&](int64_t i) { return qp.variable_upper_bounds[i]; }, [&](int64_t i) { return primal[i]; }, [](int64_t) { return 1.0; }};
    TrProblem dprob{[&](int64_t i) { return -dual_part.gradient[i]; },
                    [&](int64_t i) { return std::isfinite(qp.constraint_upper_bounds[i]) ? -kInf : 0.0; },
                    [&](int64_t i) { return std::isfinite(qp.constraint_lower_bounds[i]) ? kInf : 0.0; },
                    [&](int64_t i) { return dual[i]; }, [](int64_t) { return 1.0; }};
    const TrStepResult pr = SolveTrustRegionStepSize(pprob, primal_radius, sqp.PrimalSharder());
    const TrStepResult dr = SolveTrustRegionStepSize(dprob, dual_radius, sqp.DualSharder());
    return {lagrangian_value, lagrangian_value + pr.objective_value, lagrangian_value - dr.objective_value, radius};
  }
  // Euclidean, trust_region.cc:886-974
  const int64_t n = sqp.PrimalSize(), m = sqp.DualSize();
  const Sharder joint_sharder(sqp.PrimalSharder(), n + m);
  const DiagTrProblem joint = JointTrProblem(qp, primal, dual, primal_part.gradient, dual_part.gradient, primal_weight);
  Vec solution;
  if (use_diagonal_qp_solver) {
    solution = SolveDiagonalTrustRegionProblem(joint, joint_sharder, radius, diagonal_tol).solution;
  } else {
    const TrStepResult r = SolveTrustRegionStepSize(joint, radius, joint_sharder);
    solution = ComputeSolution(joint, r.solution_step_size, joint_sharder);
  }
  double primal_delta = sqp.PrimalSharder().SumOverShards([&](int, int64_t b, int64_t e) {
    double s = 0; for (int64_t i = b; i < e; ++i) s += primal_part.gradient[i] * (solution[i] - primal[i]); return s;
  });
  if (use_diagonal_qp_solver && qp.objective_matrix.has_value()) {
    primal_delta += sqp.PrimalSharder().SumOverShards([&](int, int64_t b, int64_t e) {
      double s = 0; for (int64_t i = b; i < e; ++i) s += 0.5 * (*qp.objective_matrix)[i] * Sq(solution[i] - primal[i]); return s;
    });
  }
  const double dual_delta = sqp.DualSharder().SumOverShards([&](int, int64_t b, int64_t e) {
    double s = 0; for (int64_t i = b; i < e; ++i) s += dual_part.gradient[i] * (solution[n + i] - dual[i]); return s;
  });
  return {lagrangian_value, lagrangian_value + primal_delta, lagrangian_value + dual_delta, radius};
}

// ---------------------------------------------------------------------------
// Params: defaults and validation (solvers.proto, solvers_proto_validation.cc)
// ---------------------------------------------------------------------------
inline void SetDefaultParams(PdlpParams* p) {
  std::memset(p, 0, sizeof(*p));
  PdlpTerminationCriteria& t = p->termination_criteria;
  t.optimality_norm = PDLP_OPTIMALITY_NORM_L2;
  t.optimality_criteria_case = PDLP_OPTIMALITY_CRITERIA_NOT_SET;
  t.simple_eps_optimal_absolute = t.simple_eps_optimal_relative = 1e-6;
  t.eps_optimal_primal_residual_absolute = t.eps_optimal_primal_residual_relative = 1e-6;
  t.eps_optimal_dual_residual_absolute = t.eps_optimal_dual_residual_relative = 1e-6;
  t.eps_optimal_objective_gap_absolute = t.eps_optimal_objective_gap_relative = 1e-6;
  t.eps_optimal_absolute = t.eps_optimal_relative = 1e-6;
  t.eps_primal_infeasible = t.eps_dual_infeasible = 1e-8;
  t.time_sec_limit = kInf;
  t.iteration_limit = std::numeric_limits<int32_t>::max();
  t.kkt_matrix_pass_limit = kInf;
  p->num_threads = 1; p->num_shards = 0; p->scheduler_type = PDLP_SCHEDULER_TYPE_GOOGLE_THREADPOOL;
  p->major_iteration_frequency = 64; p->termination_check_frequency = 64;
  p->restart_strategy = PDLP_ADAPTIVE_HEURISTIC; p->primal_weight_update_smoothing = 0.5;
  p->l_inf_ruiz_iterations = 5; p->l2_norm_rescaling = 1;
  p->sufficient_reduction_for_restart = 0.1; p->necessary_reduction_for_restart = 0.9;
  p->linesearch_rule = PDLP_ADAPTIVE_LINESEARCH_RULE;
  p->adaptive_step_size_reduction_exponent = 0.3; p->adaptive_step_size_growth_exponent = 0.6;
  p->malitsky_pock_step_size_downscaling_factor = 0.7; p->malitsky_pock_linesearch_contraction_factor = 0.99;
  p->malitsky_pock_step_size_interpolation = 1.0;
  p->initial_step_size_scaling = 1.0; p->infinite_constraint_bound_threshold = kInf;
  p->handle_some_primal_gradients_on_finite_bounds_as_residuals = 1;
  p->diagonal_qp_trust_region_solver_tolerance = 1e-8;
}

inline std::string FormatDouble(double v) {  // absl::StrCat(double) uses %g-style
  char buf[64]; std::snprintf(buf, sizeof(buf), "%g", v); return buf;
}
// Returns "" if valid, else the reference's message (solvers_proto_validation.cc:33-298).
inline std::string ValidateParams(const PdlpParams& p) {
  auto non_negative = [](double v, const char* name) -> std::string {
    if (std::isnan(v)) return std::string(name) + " is NAN";
    if (v < 0) return std::string(name) + " must be non-negative";
    return "";
  };
  const PdlpTerminationCriteria& c = p.termination_criteria;
  auto criteria = [&]() -> std::string {
    if (c.optimality_norm != PDLP_OPTIMALITY_NORM_L_INF && c.optimality_norm != PDLP_OPTIMALITY_NORM_L2 &&
        c.optimality_norm != PDLP_OPTIMALITY_NORM_L_INF_COMPONENTWISE) return "invalid value for optimality_norm";
    if (c.optimality_criteria_case != PDLP_OPTIMALITY_CRITERIA_NOT_SET) {
      if (c.has_eps_optimal_absolute) return "eps_optimal_absolute should not be set if detailed_optimality_criteria or simple_optimality_criteria is used";
      if (c.has_eps_optimal_relative) return "eps_optimal_relative should not be set if detailed_optimality_criteria or simple_optimality_criteria is used";
    }
    std::string e;
    if (c.optimality_criteria_case == PDLP_DETAILED_OPTIMALITY_CRITERIA) {
      if (!(e = non_negative(c.eps_optimal_primal_residual_absolute, "detailed_optimality_criteria.eps_optimal_primal_residual_absolute")).empty()) return e;
      if (!(e = non_negative(c.eps_optimal_primal_residual_relative, "detailed_optimality_criteria.eps_optimal_primal_residual_relative")).empty()) return e;
      if (!(e = non_negative(c.eps_optimal_dual_residual_absolute, "detailed_optimality_criteria.eps_optimal_dual_residual_absolute")).empty()) return e;
      if (!(e = non_negative(c.eps_optimal_dual_residual_relative, "detailed_optimality_criteria.eps_optimal_dual_residual_relative")).empty()) return e;
      if (!(e = non_negative(c.eps_optimal_objective_gap_absolute, "detailed_optimality_criteria.eps_optimal_objective_gap_absolute")).empty()) return e;
      if (!(e = non_negative(c.eps_optimal_objective_gap_relative, "detailed_optimality_criteria.eps_optimal_objective_gap_relative")).empty()) return e;
    } else if (c.optimality_criteria_case == PDLP_SIMPLE_OPTIMALITY_CRITERIA) {
      if (!(e = non_negative(c.simple_eps_optimal_absolute, "simple_optimality_criteria.eps_optimal_absolute")).empty()) return e;
      if (!(e = non_negative(c.simple_eps_optimal_relative, "simple_optimality_criteria.eps_optimal_relative")).empty()) return e;
    } else {
      if (!(e = non_negative(c.eps_optimal_absolute, "eps_optimal_absolute")).empty()) return e;
      if (!(e = non_negative(c.eps_optimal_relative, "eps_optimal_relative")).empty()) return e;
    }
    if (!(e = non_negative(c.eps_primal_infeasible, "eps_primal_infeasible")).empty()) return e;
    if (!(e = non_negative(c.eps_dual_infeasible, "eps_dual_infeasible")).empty()) return e;
    if (!(e = non_negative(c.time_sec_limit, "time_sec_limit")).empty()) return e;
    if (c.iteration_limit < 0) return "iteration_limit must be non-negative";
    if (!(e = non_negative(c.kkt_matrix_pass_limit, "kkt_matrix_pass_limit")).empty()) return e;
    return "";
  };
  std::string e = criteria();
  if (!e.empty()) return e + "; termination_criteria invalid";
  if (p.num_threads <= 0) return "num_threads must be positive";
  if (p.verbosity_level < 0) return "verbosity_level must be non-negative";
  if (p.log_interval_seconds < 0.0) return "log_interval_seconds must be non-negative";
  if (std::isnan(p.log_interval_seconds)) return "log_interval_seconds is NAN";
  if (p.major_iteration_frequency <= 0) return "major_iteration_frequency must be positive";
  if (p.termination_check_frequency <= 0) return "termination_check_frequency must be positive";
  if (p.restart_strategy != PDLP_NO_RESTARTS && p.restart_strategy != PDLP_EVERY_MAJOR_ITERATION &&
      p.restart_strategy != PDLP_ADAPTIVE_HEURISTIC && p.restart_strategy != PDLP_ADAPTIVE_DISTANCE_BASED) return "invalid restart_strategy";
  if (std::isnan(p.primal_weight_update_smoothing)) return "primal_weight_update_smoothing is NAN";
  if (p.primal_weight_update_smoothing < 0 || p.primal_weight_update_smoothing > 1) return "primal_weight_update_smoothing must be between 0 and 1 inclusive";
  if (std::isnan(p.initial_primal_weight)) return "initial_primal_weight is NAN";
  if (p.has_initial_primal_weight && (p.initial_primal_weight <= 1.0e-50 || p.initial_primal_weight >= 1.0e50))
    return "initial_primal_weight must be between " + FormatDouble(1.0e-50) + " and " + FormatDouble(1.0e50) + " if specified";
  if (p.l_inf_ruiz_iterations < 0) return "l_inf_ruiz_iterations must be non-negative";
  if (p.l_inf_ruiz_iterations > 100) return "l_inf_ruiz_iterations must be at most 100";
  if (std::isnan(p.sufficient_reduction_for_restart)) return "sufficient_reduction_for_restart is NAN";
  if (p.sufficient_reduction_for_restart <= 0 || p.sufficient_reduction_for_restart >= 1) return "sufficient_reduction_for_restart must be between 0 and 1 exclusive";
  if (std::isnan(p.necessary_reduction_for_restart)) return "necessary_reduction_for_restart is NAN";
  if (p.necessary_reduction_for_restart < p.sufficient_reduction_for_restart || p.necessary_reduction_for_restart >= 1)
    return "necessary_reduction_for_restart must be in the interval [sufficient_reduction_for_restart, 1)";
  if (p.linesearch_rule != PDLP_ADAPTIVE_LINESEARCH_RULE && p.linesearch_rule != PDLP_MALITSKY_POCK_LINESEARCH_RULE &&
      p.linesearch_rule != PDLP_CONSTANT_STEP_SIZE_RULE) return "invalid linesearch_rule";
  {
    std::string a;
    if (std::isnan(p.adaptive_step_size_reduction_exponent)) a = "step_size_reduction_exponent is NAN";
    else if (p.adaptive_step_size_reduction_exponent < 0.1 || p.adaptive_step_size_reduction_exponent > 1.0) a = "step_size_reduction_exponent must be between 0.1 and 1.0 inclusive";
    else if (std::isnan(p.adaptive_step_size_growth_exponent)) a = "step_size_growth_exponent is NAN";
    else if (p.adaptive_step_size_growth_exponent < 0.1 || p.adaptive_step_size_growth_exponent > 1.0) a = "step_size_growth_exponent must be between 0.1 and 1.0 inclusive";
    if (!a.empty()) return a + "; adaptive_linesearch_parameters invalid";
  }
  {
    std::string a;
    if (std::isnan(p.malitsky_pock_step_size_downscaling_factor)) a = "step_size_downscaling_factor is NAN";
    else if (p.malitsky_pock_step_size_downscaling_factor <= 1.0e-50 || p.malitsky_pock_step_size_downscaling_factor >= 1)
      a = "step_size_downscaling_factor must be between " + FormatDouble(1.0e-50) + " and 1 exclusive";
    else if (std::isnan(p.malitsky_pock_linesearch_contraction_factor)) a = "linesearch_contraction_factor is NAN";
    else if (p.malitsky_pock_linesearch_contraction_factor <= 0 || p.malitsky_pock_linesearch_contraction_factor >= 1) a = "linesearch_contraction_factor must be between 0 and 1 exclusive";
    else if (std::isnan(p.malitsky_pock_step_size_interpolation)) a = "step_size_interpolation is NAN";
    else if (p.malitsky_pock_step_size_interpolation < 0 || p.malitsky_pock_step_size_interpolation >= 1.0e50)
      a = "step_size_interpolation must be non-negative and less than " + FormatDouble(1.0e50);
    if (!a.empty()) return a + "; malitsky_pock_parameters invalid";
  }
  if (std::isnan(p.initial_step_size_scaling)) return "initial_step_size_scaling is NAN";
  if (p.initial_step_size_scaling <= 1.0e-50 || p.initial_step_size_scaling >= 1.0e50)
    return "initial_step_size_scaling must be between " + FormatDouble(1.0e-50) + " and " + FormatDouble(1.0e50);
  if (std::isnan(p.infinite_constraint_bound_threshold)) return "infinite_constraint_bound_threshold is NAN";
  if (p.infinite_constraint_bound_threshold <= 0.0) return "infinite_constraint_bound_threshold must be positive";
  if (std::isnan(p.diagonal_qp_trust_region_solver_tolerance)) return "diagonal_qp_trust_region_solver_tolerance is NAN";
  if (p.diagonal_qp_trust_region_solver_tolerance < 10 * std::numeric_limits<double>::epsilon())
    return "diagonal_qp_trust_region_solver_tolerance must be at least " + FormatDouble(10 * std::numeric_limits<double>::epsilon());
  if (p.use_feasibility_polishing && p.handle_some_primal_gradients_on_finite_bounds_as_residuals)
    return "use_feasibility_polishing requires !handle_some_primal_gradients_on_finite_bounds_as_residuals";
  if (p.use_feasibility_polishing && p.presolve_use_glop) return "use_feasibility_polishing and glop presolve can not be used together.";
  return "";
}

}  // namespace pdlp_oracle

#endif  // ORACLE_PDLP_CPU_CORE_H_
