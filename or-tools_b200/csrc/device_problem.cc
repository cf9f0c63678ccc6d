// device_problem.cc -- see device_problem.h.
#include "device_problem.h"

#include "comm.h"

#include <algorithm>
#include <atomic>
#include <thread>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <stdexcept>

namespace pdlp_b200 {

namespace {
constexpr double kInf = std::numeric_limits<double>::infinity();
}

// Contiguous row blocks of (roughly) equal nnz, the mass rule of sharder.cc:51-70
// applied to the rows of K: block g ends at the first row where the running nnz
// reaches (g + 1) / G of the total.
void ComputeRowBlock(const PdlpProblemView& v, int rank, int world, int64_t* begin, int64_t* end) {
  const int64_t m = v.num_constraints, nnz = v.col_starts[v.num_variables];
  std::vector<int64_t> cum(m + 1, 0);
  // Row histogram of K. Every rank of a sharded solve computes it, so the threads
  // available to one rank are the host's divided by the world size; counts are
  // integers, so the (relaxed atomic) order of the increments does not matter.
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const int threads = nnz < (int64_t{1} << 21) ? 1 : static_cast<int>(std::min<unsigned>(16u, std::max(1u, hw / static_cast<unsigned>(std::max(world, 1)))));
  std::atomic<bool> bad{false};
  auto count = [&](int64_t k0, int64_t k1, bool shared) {
    int64_t* c = cum.data();
    for (int64_t k = k0; k < k1; ++k) {
      const int64_t r = v.row_indices[k];
      if (r < 0 || r >= m) {
        bad.store(true, std::memory_order_relaxed);
        return;
      }
      if (shared) __atomic_fetch_add(&c[r + 1], int64_t{1}, __ATOMIC_RELAXED);
      else ++c[r + 1];
    }
  };
  if (threads <= 1) {
    count(0, nnz, false);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(count, nnz * t / threads, nnz * (t + 1) / threads, true);
    for (std::thread& th : pool) th.join();
  }
  if (bad.load()) throw std::runtime_error("row index out of range");
  for (int64_t r = 0; r < m; ++r) cum[r + 1] += cum[r];
  auto boundary = [&](int g) -> int64_t {
    if (g <= 0) return 0;
    if (g >= world) return m;
    // balance nnz + rows so that empty / very sparse tails are still spread
    const double target = (static_cast<double>(nnz) + static_cast<double>(m)) * g / world;
    int64_t lo = 0, hi = m;
    while (lo < hi) {
      const int64_t mid = (lo + hi) / 2;
      if (static_cast<double>(cum[mid]) + static_cast<double>(mid) < target) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  *begin = boundary(rank);
  *end = boundary(rank + 1);
}

DeviceProblem::DeviceProblem(const PdlpProblemView& view, int cuda_device, Comm* comm) : dev_(new Device(cuda_device)), comm_(comm) {
  m_global_ = view.num_constraints;
  int64_t row_end = m_global_;
  if (comm_ != nullptr) {
    ComputeRowBlock(view, comm_->rank(), comm_->world_size(), &row_begin_, &row_end);
    dev_->SetComm(comm_);
  }
  // cheap host-side sanity of the column starts (everything else is validated on the device)
  for (int64_t c = 0; c < view.num_variables; ++c)
    if (view.col_starts[c + 1] < view.col_starts[c] || view.col_starts[c] < 0) throw std::runtime_error("col_starts is not monotone");
  const char* host_build = std::getenv("PDLP_B200_HOST_BUILD");
  device_built_ = !(host_build != nullptr && host_build[0] == '1');
  QpHost h;
  if (device_built_) {
    dev_->BuildSellPair(view, row_begin_, row_end, 4096, /*natural_primal_order=*/comm_ != nullptr, &rows_, &cols_, &dual_perm_, &primal_perm_, &build_info_);
    n_ = build_info_.n;
    m_ = build_info_.m;
    nnz_ = build_info_.nnz;
  } else {
    h = BuildQpHost(view, row_begin_, row_end, 4096, /*natural_primal_order=*/comm_ != nullptr);
    n_ = h.n;
    m_ = h.m;
    nnz_ = h.nnz;
  }
  if (comm_ != nullptr) {
    exchange_ = dev_->AllocF64(n_ + 2);  // K^T y' partial + {||dy||^2, (K dx) . dy}
    layout_ = PeerLayout::For(n_, m_global_, comm_->world_size());
    slice_begin_ = std::min<int64_t>(n_, layout_.stride * comm_->rank());
    slice_end_ = std::min<int64_t>(n_, slice_begin_ + layout_.stride);
    dev_->SetPrimalSlice(n_, slice_begin_, slice_end_);
    // PDLP_B200_EXCHANGE=nccl keeps the all-reduce exchange (A/B and boxes without peer mapping)
    const char* ex = std::getenv("PDLP_B200_EXCHANGE");
    if (!(ex != nullptr && std::strcmp(ex, "nccl") == 0)) {
      const auto t0 = std::chrono::steady_clock::now();
      arena_ = comm_->AcquirePeerArena(layout_.doubles * static_cast<int64_t>(sizeof(double)), dev_->stream());
      dev_->SetPeerArena(arena_, n_, m_global_);
      if (const char* t = std::getenv("PDLP_B200_TRACE"); t != nullptr && t[0] == '1')
        std::fprintf(stderr, "[pdlp_b200 trace] rank %d peer arena %s in %.3f s\n", comm_->rank(), arena_ != nullptr ? "mapped" : "unavailable (NCCL exchange)",
                     std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
  }
  objective_offset_ = view.objective_offset;
  objective_scaling_factor_ = view.objective_scaling_factor;
  if (!device_built_) {
    rows_ = dev_->UploadSell(h.rows);
    cols_ = dev_->UploadSell(h.cols);
    primal_perm_ = dev_->UploadI32(h.cols.row_of_pos);
    dual_perm_ = dev_->UploadI32(h.rows.row_of_pos);
    col_starts_.assign(view.col_starts, view.col_starts + n_ + 1);
    cols_meta_ = std::move(h.cols);
    std::vector<int32_t>().swap(cols_meta_.col);
    std::vector<double>().swap(cols_meta_.val);
  }
  auto up_primal = [&](const double* src) { double* d = NewPrimal(); UploadPrimal(d, src); return d; };
  auto up_dual = [&](const double* src) { double* d = NewDual(); UploadDual(d, src); return d; };
  c_ = up_primal(view.objective_vector);
  if (view.objective_matrix_diagonal != nullptr) q_ = up_primal(view.objective_matrix_diagonal);
  lv_ = up_primal(view.variable_lower_bounds);
  uv_ = up_primal(view.variable_upper_bounds);
  lc_ = up_dual(view.constraint_lower_bounds);
  uc_ = up_dual(view.constraint_upper_bounds);
  for (int k = 0; k < 4; ++k) {
    tmp_n_[k] = NewPrimal();
    tmp_m_[k] = NewDual();
  }
  ones_n_ = NewPrimal();
  ones_m_ = NewDual();
  dev_->Fill(ones_n_, 1.0, n_);
  dev_->Fill(ones_m_, 1.0, m_);
  if (arena_ != nullptr) {
    // Which exchange the step loop uses: "peer-d" all-gathers x~ and y' and needs
    // the image of this rank's column slice over ALL rows (8 (n + m) bytes on the
    // wire per step); "peer-s" all-gathers x~ and reduce-scatters the K^T y'
    // partials of the row block (16 n bytes). Default: whichever moves less.
    const char* ex = std::getenv("PDLP_B200_EXCHANGE");
    bool all_gather = m_global_ <= n_;
    if (ex != nullptr && std::strcmp(ex, "peer-s") == 0) all_gather = false;
    if (ex != nullptr && std::strcmp(ex, "peer-d") == 0) all_gather = true;
    if (all_gather) {
      // box-wide dual order: the block of rank g starts at its first row, positions inside as in its row image
      double* full = dev_->AllocF64(m_global_);
      dev_->Fill(full, 0.0, m_global_);
      dev_->WriteGlobalRowPositions(full, dual_perm_, row_begin_, m_);
      dev_->AllReduceSumVec(full, m_global_);
      double* gpos_buf = dev_->AllocF64(m_global_ / 2 + 2);
      int32_t* gpos = reinterpret_cast<int32_t*>(gpos_buf);
      dev_->DoublesToI32(gpos, full, m_global_);
      dev_->BuildColumnSliceImage(view, slice_begin_, slice_end_, gpos, 4096, &cols_slice_, &slice_perm_);
      has_cols_slice_ = true;
      dev_->Free(full);
      dev_->Free(gpos_buf);
    }
  }
  dev_->Sync();
}

DeviceProblem::~DeviceProblem() {
  if (arena_ != nullptr) comm_->ReleasePeerArena(arena_, dev_->stream());
  for (double* p : {c_, q_, lv_, uv_, lc_, uc_, ones_n_, ones_m_, exchange_}) dev_->Free(p);
  for (int k = 0; k < 4; ++k) { dev_->Free(tmp_n_[k]); dev_->Free(tmp_m_[k]); }
  dev_->Free(primal_perm_);
  dev_->Free(dual_perm_);
  dev_->FreeSell(rows_);
  dev_->FreeSell(cols_);
  if (has_cols_slice_) dev_->FreeSell(cols_slice_);
  dev_->Free(slice_perm_);
  dev_->FreeBuildInfo(build_info_);
}

void DeviceProblem::RescaleQuadraticProgram(const double* col_scaling, const double* row_scaling) {
  Device& d = *dev_;
  d.Mul(c_, col_scaling, n_);
  d.Div(lv_, col_scaling, n_);
  d.Div(uv_, col_scaling, n_);
  if (q_ != nullptr) d.MulSq(q_, col_scaling, n_);
  d.Mul(lc_, row_scaling, m_);
  d.Mul(uc_, row_scaling, m_);
  d.ScaleMatrix(rows_, row_scaling, col_scaling);
  d.ScaleMatrix(cols_, col_scaling, row_scaling, sharded() ? primal_perm_ : nullptr);
  if (has_cols_slice_) {
    // the slice image gathers in the box-wide dual order: it needs every rank's row scaling
    double* dr_full = d.AllocF64(m_global_);
    d.Fill(dr_full, 0.0, m_global_);
    d.CopyD2D(dr_full + row_begin_, row_scaling, m_);
    d.AllReduceSumVec(dr_full, m_global_);
    d.ScaleMatrix(cols_slice_, col_scaling + slice_begin_, dr_full, slice_perm_);
    d.Sync();
    d.Free(dr_full);
  }
}

void DeviceProblem::KTy(const double* y, double* out) {
  if (!sharded()) { dev_->SpMV(cols_, y, out); return; }
  // local partial in column order, then the exchange step
  dev_->SpMVScatter(cols_, y, primal_perm_, exchange_);
  dev_->AllReduceSumVec(exchange_, n_);
  dev_->CopyD2D(out, exchange_, n_);
}

void DeviceProblem::DownloadDual(double* host, const double* src) {
  if (!sharded()) { dev_->DownloadPermuted(host, src, dual_perm_, m_); return; }
  // every rank receives the whole dual vector: blocks are written into a
  // zeroed full-length buffer and summed across ranks
  double* full = dev_->AllocF64(m_global_);
  dev_->Fill(full, 0.0, m_global_);
  dev_->ScatterInto(full + row_begin_, src, dual_perm_, m_);
  dev_->AllReduceSumVec(full, m_global_);
  dev_->Download(host, full, m_global_);
  dev_->Free(full);
}

// Column norms of D_r K D_c (ScaledColLInfNorm / ScaledColL2Norm of the
// reference applied to K): on a row block they are partial and are completed
// by an all-reduce (max for LInf, sum of squares for L2).
void DeviceProblem::ColumnNorms(int norm, const double* row_scaling, const double* col_scaling, double* out) {
  Device& d = *dev_;
  if (!sharded()) { d.ScaledRowNorm(cols_, norm, row_scaling, col_scaling, out); return; }
  d.RowNormRawScatter(cols_, norm, row_scaling, primal_perm_, out);
  if (norm == 0) d.AllReduceMaxVec(out, n_); else d.AllReduceSumVec(out, n_);
  d.FinishRowNorm(out, norm, col_scaling, n_);
}

void DeviceProblem::ReplaceLargeConstraintBoundsWithInfinity(double threshold) {
  dev_->ReplaceLargeWithInf(lc_, threshold, m_);
  dev_->ReplaceLargeWithInf(uc_, threshold, m_);
}

bool DeviceProblem::HasValidBounds() { return dev_->BoundsValid(lc_, uc_, m_, sharded()) && dev_->BoundsValid(lv_, uv_, n_); }
bool DeviceProblem::ObjectiveMatrixIsNonNegative() { return q_ == nullptr || dev_->AllNonNegative(q_, n_); }

namespace {
struct Info { double largest, smallest, average, l2; int64_t nfn, nzero; };
Info Finish(const VectorInfoDev& v) {  // VectorInfoAccumulator::operator VectorInfo, sou.cc:155-167
  Info r;
  r.nfn = static_cast<int64_t>(v.num_finite_nonzero);
  r.nzero = static_cast<int64_t>(v.num_zero);
  r.largest = r.nfn > 0 ? v.largest : 0.0;
  r.smallest = r.nfn > 0 ? v.smallest : 0.0;
  r.average = (r.nfn + r.nzero > 0) ? v.sum / static_cast<double>(r.nfn + r.nzero) : std::numeric_limits<double>::quiet_NaN();
  r.l2 = std::sqrt(v.sumsq);
  return r;
}
}  // namespace

PdlpQuadraticProgramStats DeviceProblem::ComputeStats() {
  Device& d = *dev_;
  // row / column LInf norms with unit scaling (sou.cc:240-266)
  d.ScaledRowNorm(rows_, 0, ones_n_, ones_m_, tmp_m_[0]);
  ColumnNorms(0, ones_m_, ones_n_, tmp_n_[0]);
  const Info row_info = Finish(d.VectorInfo(tmp_m_[0], m_, sharded()));
  const Info col_info = Finish(d.VectorInfo(tmp_n_[0], n_));
  const Info mat = Finish(d.MatrixInfo(cols_));
  const Info bounds = Finish(d.CombinedBoundsInfo(uc_, lc_, m_, sharded()));
  const Info var_bounds = Finish(d.CombinedBoundsInfo(uv_, lv_, n_));
  const Info obj = Finish(d.VectorInfo(c_, n_));
  const Info gaps = Finish(d.GapInfo(lv_, uv_, n_));
  PdlpQuadraticProgramStats s;
  std::memset(&s, 0, sizeof(s));
  s.num_variables = n_;
  s.num_constraints = m_global_;
  s.constraint_matrix_col_min_l_inf_norm = col_info.smallest;
  s.constraint_matrix_row_min_l_inf_norm = row_info.smallest;
  s.constraint_matrix_num_nonzeros = mat.nfn;
  s.constraint_matrix_abs_max = mat.largest; s.constraint_matrix_abs_min = mat.smallest;
  s.constraint_matrix_abs_avg = mat.average; s.constraint_matrix_l2_norm = mat.l2;
  s.combined_bounds_max = bounds.largest; s.combined_bounds_min = bounds.smallest;
  s.combined_bounds_avg = bounds.average; s.combined_bounds_l2_norm = bounds.l2;
  s.combined_variable_bounds_max = var_bounds.largest; s.combined_variable_bounds_min = var_bounds.smallest;
  s.combined_variable_bounds_avg = var_bounds.average; s.combined_variable_bounds_l2_norm = var_bounds.l2;
  s.variable_bound_gaps_num_finite = gaps.nfn + gaps.nzero;
  s.variable_bound_gaps_max = gaps.largest; s.variable_bound_gaps_min = gaps.smallest;
  s.variable_bound_gaps_avg = gaps.average; s.variable_bound_gaps_l2_norm = gaps.l2;
  s.objective_vector_abs_max = obj.largest; s.objective_vector_abs_min = obj.smallest;
  s.objective_vector_abs_avg = obj.average; s.objective_vector_l2_norm = obj.l2;
  if (q_ == nullptr) {
    s.objective_matrix_abs_avg = std::numeric_limits<double>::quiet_NaN();
  } else {
    const Info qi = Finish(d.VectorInfo(q_, n_));
    s.objective_matrix_num_nonzeros = qi.nfn;
    s.objective_matrix_abs_max = qi.largest; s.objective_matrix_abs_min = qi.smallest;
    s.objective_matrix_abs_avg = qi.average; s.objective_matrix_l2_norm = qi.l2;
  }
  return s;
}

// sou.cc:367-405: both norms are computed from the same (old) scaling vectors.
void DeviceProblem::ApplyScalingIterationsForNorm(int num_iterations, int norm, double* row_scaling, double* col_scaling) {
  Device& d = *dev_;
  for (int it = 0; it < num_iterations; ++it) {
    ColumnNorms(norm, row_scaling, col_scaling, tmp_n_[0]);             // column norms of D_r K D_c
    d.ScaledRowNorm(rows_, norm, col_scaling, row_scaling, tmp_m_[0]);  // row norms
    d.DivideBySqrt(col_scaling, tmp_n_[0], n_);
    d.DivideBySqrt(row_scaling, tmp_m_[0], m_);
  }
}

void DeviceProblem::ApplyRescaling(int l_inf_ruiz_iterations, bool l2_norm_rescaling, double** row_scaling, double** col_scaling) {
  Device& d = *dev_;
  double* r = NewDual();
  double* c = NewPrimal();
  d.Fill(r, 1.0, m_);
  d.Fill(c, 1.0, n_);
  bool do_rescale = false;
  if (l_inf_ruiz_iterations > 0) { do_rescale = true; ApplyScalingIterationsForNorm(l_inf_ruiz_iterations, 0, r, c); }
  if (l2_norm_rescaling) { do_rescale = true; ApplyScalingIterationsForNorm(1, 1, r, c); }
  if (do_rescale) RescaleQuadraticProgram(c, r);
  *row_scaling = r;
  *col_scaling = c;
}

PdlpConvergenceInformation DeviceProblem::ComputeConvergenceInformation(bool handle_as_residuals, const double* dc, const double* dr, const double* x,
                                                                        const double* y, const double* kty_or_null, double cw_primal_offset,
                                                                        double cw_dual_offset, int candidate_type, const double* kx_or_null) {
  Device& d = *dev_;
  PdlpConvergenceInformation r;
  std::memset(&r, 0, sizeof(r));
  const double* kx = kx_or_null;
  if (kx == nullptr) { Kx(x, tmp_m_[0]); kx = tmp_m_[0]; }
  const double* kty = kty_or_null;
  if (kty == nullptr) { KTy(y, tmp_n_[0]); kty = tmp_n_[0]; }
  d.BeginBatch();  // both sides in one host round trip
  const int om = d.DualSideStatsLaunch(y, kx, lc_, uc_, dr, cw_primal_offset, /*homogeneous=*/false, m_);
  const int on = d.PrimalSideStatsLaunch(x, x, kty, c_, q_, lv_, uv_, dc, cw_dual_offset, /*zero_objective=*/false, handle_as_residuals, n_);
  d.EndBatch();
  const MSideStats ms = d.ReadDualSideStats(om);
  const NSideStats ns = d.ReadPrimalSideStats(on);
  r.l_inf_primal_residual = ms.linf_residual;
  r.l2_primal_residual = std::sqrt(ms.sumsq_residual);
  r.l_inf_componentwise_primal_residual = ms.cw_residual;
  r.l_inf_primal_variable = ns.linf_scaled;
  r.l2_primal_variable = std::sqrt(ns.sumsq_scaled);
  r.l_inf_dual_variable = ms.linf_scaled;
  r.l2_dual_variable = std::sqrt(ms.sumsq_scaled);
  const double quadratic_objective = 0.5 * ns.quadratic;
  r.primal_objective = ApplyObjectiveScalingAndOffset(quadratic_objective + ns.objective_dot);
  const double dual_objective_piece = -quadratic_objective + ms.bounds_term;
  r.dual_objective = ApplyObjectiveScalingAndOffset(dual_objective_piece + ns.correction);
  r.corrected_dual_objective = ApplyObjectiveScalingAndOffset(dual_objective_piece + ns.full_correction);
  r.l_inf_dual_residual = ns.linf_residual;
  r.l2_dual_residual = std::sqrt(ns.sumsq_residual);
  r.l_inf_componentwise_dual_residual = ns.cw_residual;
  r.candidate_type = candidate_type;
  return r;
}

PdlpInfeasibilityInformation DeviceProblem::ComputeInfeasibilityInformation(bool handle_as_residuals, const double* dc, const double* dr,
                                                                            const double* primal_ray, const double* dual_ray,
                                                                            const double* primal_for_residual_tests,
                                                                            const double* kty_of_dual_ray_or_null, int candidate_type) {
  Device& d = *dev_;
  PdlpInfeasibilityInformation r;
  std::memset(&r, 0, sizeof(r));
  const double* kty = kty_of_dual_ray_or_null;
  if (kty == nullptr) { KTy(dual_ray, tmp_n_[1]); kty = tmp_n_[1]; }
  Kx(primal_ray, tmp_m_[1]);
  // dual ray side: gradient = -K^T ray; DualResidualNorms against the bounds at
  // `primal_for_residual_tests`; primal ray side: objective terms + norms.
  d.BeginBatch();  // both sides in one host round trip
  const int on = d.PrimalSideStatsLaunch(primal_ray, primal_for_residual_tests, kty, c_, q_, lv_, uv_, dc, 0.0, /*zero_objective=*/true,
                                         handle_as_residuals, n_);
  const int om = d.DualSideStatsLaunch(dual_ray, tmp_m_[1], lc_, uc_, dr, 0.0, /*homogeneous=*/true, m_);
  d.EndBatch();
  const NSideStats ns_dual = d.ReadPrimalSideStats(on);
  const MSideStats ms = d.ReadDualSideStats(om);
  const double l_inf_primal = ns_dual.linf_scaled;
  const double l_inf_dual = ms.linf_scaled;
  const double dual_ray_objective = ms.bounds_term + ns_dual.correction;
  if (l_inf_dual > 0) {
    r.dual_ray_objective = dual_ray_objective / l_inf_dual;
    r.max_dual_ray_infeasibility = ns_dual.linf_residual / l_inf_dual;
  }
  if (l_inf_primal > 0.0) {
    r.primal_ray_quadratic_norm = ns_dual.linf_qx / l_inf_primal;
    r.max_primal_ray_infeasibility = ms.linf_residual / l_inf_primal;
    r.primal_ray_linear_objective = ns_dual.objective_dot / l_inf_primal;
  }
  r.candidate_type = candidate_type;
  return r;
}

void DeviceProblem::ReducedCosts(const double* x, const double* y, bool use_zero_primal_objective, double* out) {
  KTy(y, tmp_n_[0]);
  dev_->PrimalGradient(x, tmp_n_[0], c_, q_, use_zero_primal_objective, out, n_);
}

void DeviceProblem::ComputeLocalizedLagrangianBounds(const double* x, const double* y, double primal_weight, double radius, const double* kx,
                                                     const double* kty, bool use_diagonal_solver, double diagonal_tol, double out[4],
                                                     const double* x0, const double* y0, double* dist_sq) {
  if (kx == nullptr) { Kx(x, tmp_m_[2]); kx = tmp_m_[2]; }
  if (kty == nullptr) { KTy(y, tmp_n_[2]); kty = tmp_n_[2]; }
  double r3[3], extra[3];
  dev_->LocalizedLagrangianBounds(x, y, kx, kty, c_, q_, lv_, uv_, lc_, uc_, primal_weight, radius, use_diagonal_solver, diagonal_tol, n_, m_, r3, x0, y0,
                                  extra);
  out[0] = r3[0]; out[1] = r3[1]; out[2] = r3[2]; out[3] = extra[0];
  if (dist_sq != nullptr) { dist_sq[0] = extra[1]; dist_sq[1] = extra[2]; }
}

bool DeviceProblem::ComputeLocalizedLagrangianBoundsPair(const double* const x[2], const double* const y[2], const double* const kx[2],
                                                         const double* const kty[2], double primal_weight, bool use_diagonal_solver, const double* x0,
                                                         const double* y0, double out[2][4], double dist_sq[2][2]) {
  if (use_diagonal_solver) return false;
  double r3[2][3], extra[2][3];
  if (!dev_->LocalizedLagrangianBoundsPair(x, y, kx, kty, c_, q_, lv_, uv_, lc_, uc_, primal_weight, n_, m_, x0, y0, r3, extra)) return false;
  for (int k = 0; k < 2; ++k) {
    out[k][0] = r3[k][0]; out[k][1] = r3[k][1]; out[k][2] = r3[k][2]; out[k][3] = extra[k][0];
    dist_sq[k][0] = extra[k][1]; dist_sq[k][1] = extra[k][2];
  }
  return true;
}

void DeviceProblem::ComputeLocalizedLagrangianBoundsMaxNorm(const double* x, const double* y, double primal_weight, double radius, const double* kx,
                                                            const double* kty, double out[4]) {
  if (sharded()) throw std::runtime_error("max-norm localized Lagrangian bounds are not available on a row-sharded problem");
  Device& d = *dev_;
  if (kx == nullptr) { Kx(x, tmp_m_[2]); kx = tmp_m_[2]; }
  if (kty == nullptr) { KTy(y, tmp_n_[2]); kty = tmp_n_[2]; }
  double* gx = tmp_n_[0];
  double* gy = tmp_m_[0];
  const double primal_value = d.LagrangianPrimalGradient(x, kty, c_, q_, gx, n_);
  const double dual_value = d.LagrangianDualGradient(y, kx, lc_, uc_, gy, m_);
  const double lagrangian = primal_value + dual_value;
  double step = 0.0, primal_objective = 0.0, dual_objective = 0.0;
  d.SolveTrustRegion(gx, lv_, uv_, x, ones_n_, std::sqrt(2.0) * radius / std::sqrt(primal_weight), n_, tmp_n_[1], &step, &primal_objective);
  d.DualTrustRegionProblem(gy, lc_, uc_, tmp_m_[1], tmp_m_[2], tmp_m_[3], m_);
  d.SolveTrustRegion(tmp_m_[1], tmp_m_[2], tmp_m_[3], y, ones_m_, std::sqrt(2.0) * radius * std::sqrt(primal_weight), m_, tmp_m_[0], &step, &dual_objective);
  out[0] = lagrangian;
  out[1] = lagrangian + primal_objective;
  out[2] = lagrangian - dual_objective;
  out[3] = radius;
}

void DeviceProblem::DownloadValuesCsc(double* values) {
  if (sharded()) throw std::runtime_error("DownloadValuesCsc is not available on a row-sharded problem");
  if (device_built_) { dev_->DownloadValuesCscFromSell(cols_, build_info_, values); return; }
  std::vector<double> sell;
  dev_->DownloadSellValues(cols_, sell);
  const SellHost& s = cols_meta_;
  const int32_t T = s.split_len;
  for (int64_t pos = 0; pos < n_; ++pos) {
    const int32_t col = s.row_of_pos[pos];
    const int64_t dst = col_starts_[col];
    if (pos < s.num_split) {
      // virtual slots are filled in teams (sell_builder.cc FillSell): member k of a team of t holds entries e0 + k + t j
      const int64_t first = s.split_first[pos], last = s.split_first[pos + 1];
      for (int64_t v = first; v < last; ++v) {
        const char* team_env = std::getenv("PDLP_B200_TEAM_SLOTS");
        const bool teams = !(team_env != nullptr && team_env[0] == '0');
        const int64_t ts = teams ? std::max<int64_t>(first, (v >> 5) << 5) : v, te = teams ? std::min<int64_t>(last, ((v >> 5) + 1) << 5) : v + 1;
        const int64_t t = te - ts, k = v - ts, e0 = (ts - first) * T;
        const int64_t base = s.slice_ptr[v >> 5] + (v & 31);
        for (int32_t j = 0; j < s.slot_len[v]; ++j) values[dst + e0 + k + t * j] = sell[base + static_cast<int64_t>(j) * 32];
      }
    } else {
      const int64_t slot = s.num_virtual_padded + (pos - s.num_split);
      const int64_t base = s.slice_ptr[slot >> 5] + (slot & 31);
      for (int32_t j = 0; j < s.slot_len[slot]; ++j) values[dst + j] = sell[base + static_cast<int64_t>(j) * 32];
    }
  }
}

}  // namespace pdlp_b200
