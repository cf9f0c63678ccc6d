// device_problem.h -- the device-resident QP: B200 counterpart of
// ShardedQuadraticProgram (sharded_quadratic_program.h:37-124) plus the
// operators of sharded_optimization_utils / iteration_stats / trust_region on
// device vectors. All device vectors are in "position order" (device_ops.h).
#ifndef PDLP_B200_DEVICE_PROBLEM_H_
#define PDLP_B200_DEVICE_PROBLEM_H_

#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "device_ops.h"

namespace pdlp_b200 {

// The block of constraint rows rank `rank` of `world` keeps (host-only helper).
void ComputeRowBlock(const PdlpProblemView& view, int rank, int world, int64_t* begin, int64_t* end);

class DeviceProblem {
 public:
  // Uploads the QP and builds both sparse orientations. Throws
  // std::runtime_error on CUDA failure / no device.
  // With a communicator (one process per GPU) this rank keeps
  // the contiguous block of constraint rows [row_begin, row_end) chosen by the
  // nnz-balanced rule of sharder.cc:51-70; dual-length vectors are that block,
  // primal-length vectors are replicated in the caller's column order.
  DeviceProblem(const PdlpProblemView& view, int cuda_device, Comm* comm = nullptr);
  ~DeviceProblem();
  DeviceProblem(const DeviceProblem&) = delete;
  DeviceProblem& operator=(const DeviceProblem&) = delete;

  Device& dev() { return *dev_; }
  int64_t n() const { return n_; }
  int64_t m() const { return m_; }                 // rows held by this rank
  int64_t m_global() const { return m_global_; }   // rows of the whole problem
  int64_t row_begin() const { return row_begin_; }
  bool sharded() const { return comm_ != nullptr; }
  const int32_t* primal_scatter() const { return primal_perm_; }
  double* exchange() const { return exchange_; }
  // Peer-memory exchange of the step loop (nullptr: NCCL all-reduce exchange).
  const PeerArena* arena() const { return arena_; }
  int64_t slice_stride() const { return layout_.stride; }
  int64_t slice_begin() const { return slice_begin_; }
  int64_t slice_end() const { return slice_end_; }
  // All-gather exchange (chosen when m_global <= n): image of (K[:, slice])^T; nullptr otherwise.
  const SellDev* cols_slice() const { return has_cols_slice_ ? &cols_slice_ : nullptr; }
  const int32_t* slice_perm() const { return slice_perm_; }
  int64_t nnz() const { return nnz_; }
  bool is_lp() const { return q_ == nullptr; }

  // problem data (device, position order)
  double *c() const { return c_; }
  double *q() const { return q_; }
  double *lv() const { return lv_; }
  double *uv() const { return uv_; }
  double *lc() const { return lc_; }
  double *uc() const { return uc_; }
  const SellDev& rows() const { return rows_; }
  const SellDev& cols() const { return cols_; }
  double objective_offset() const { return objective_offset_; }
  double objective_scaling_factor() const { return objective_scaling_factor_; }
  double ApplyObjectiveScalingAndOffset(double v) const { return objective_scaling_factor_ * (v + objective_offset_); }

  // vectors
  // (row-sharded: padded to world * slice stride so that slices can be all-gathered in place)
  double* NewPrimal() { return dev_->AllocF64(sharded() ? layout_.n_pad : n_); }
  double* NewDual() { return dev_->AllocF64(m_); }
  // host vectors are always full length in the caller's order
  void UploadPrimal(double* dst, const double* host) {
    if (sharded()) dev_->Upload(dst, host, n_); else dev_->UploadPermuted(dst, host, primal_perm_, n_);
  }
  void UploadDual(double* dst, const double* host) { dev_->UploadPermuted(dst, host + row_begin_, dual_perm_, m_); }
  void DownloadPrimal(double* host, const double* src) {
    if (sharded()) dev_->Download(host, src, n_); else dev_->DownloadPermuted(host, src, primal_perm_, n_);
  }
  void DownloadDual(double* host, const double* src);

  void Kx(const double* x, double* out) { dev_->SpMV(rows_, x, out); }    // K x    (pdhg.cc:1912-1916)
  void KTy(const double* y, double* out);                                  // K^T y  (sharder.cc:160-173)

  // SwapObjectiveVector / SwapVariableBounds / SwapConstraintBounds (sharded_quadratic_program.h:88-109):
  // exchange the device vectors of the working problem with the caller's (feasibility polishing).
  void SwapObjectiveVector(double** objective) { std::swap(c_, *objective); }
  void SwapVariableBounds(double** lower, double** upper) { std::swap(lv_, *lower); std::swap(uv_, *upper); }
  void SwapConstraintBounds(double** lower, double** upper) { std::swap(lc_, *lower); std::swap(uc_, *upper); }
  // sharded_quadratic_program.cc:148-189
  void RescaleQuadraticProgram(const double* col_scaling, const double* row_scaling);
  void ReplaceLargeConstraintBoundsWithInfinity(double threshold);
  // sharded_optimization_utils.cc
  PdlpQuadraticProgramStats ComputeStats();
  void ApplyScalingIterationsForNorm(int num_iterations, int norm /*0 LInf, 1 L2*/, double* row_scaling, double* col_scaling);
  // ApplyRescaling: allocates and returns the scaling vectors (device).
  void ApplyRescaling(int l_inf_ruiz_iterations, bool l2_norm_rescaling, double** row_scaling, double** col_scaling);
  bool HasValidBounds();
  bool ObjectiveMatrixIsNonNegative();

  // iteration_stats.cc; dc / dr may be null (ones). tmp vectors are owned here.
  PdlpConvergenceInformation ComputeConvergenceInformation(bool handle_as_residuals, const double* dc, const double* dr, const double* x,
                                                           const double* y, const double* kty_or_null, double cw_primal_offset,
                                                           double cw_dual_offset, int candidate_type, const double* kx_or_null = nullptr);
  // primal_ray must already be projected to the feasibility bounds and dual_ray
  // to the dual bounds where required (pdhg.cc:1704-1722).
  PdlpInfeasibilityInformation ComputeInfeasibilityInformation(bool handle_as_residuals, const double* dc, const double* dr,
                                                               const double* primal_ray, const double* dual_ray,
                                                               const double* primal_for_residual_tests, const double* kty_of_dual_ray_or_null,
                                                               int candidate_type);
  void ReducedCosts(const double* x, const double* y, bool use_zero_primal_objective, double* out);
  // trust_region.cc:978-1016 (Euclidean). kx / kty may be null (computed).
  void ComputeLocalizedLagrangianBounds(const double* x, const double* y, double primal_weight, double radius, const double* kx,
                                        const double* kty, bool use_diagonal_solver, double diagonal_tol, double out[4],
                                        const double* x0 = nullptr, const double* y0 = nullptr,   // radius < 0: distance to (x0, y0)
                                        double* dist_sq = nullptr);                                 // then also {||x - x0||^2, ||y - y0||^2}

  // trust_region.cc:855-884 (max norm): the primal and the dual trust-region problems are solved
  // separately with radii sqrt(2) r / sqrt(w) and sqrt(2) r sqrt(w). Not used by the solver
  // (pdhg.cc asks for the Euclidean norm); single GPU only.
  // Both points of the restart test in one go (Device::LocalizedLagrangianBoundsPair); kx / kty must be given.
  // out[k] = {lagrangian, lower, upper, radius}, dist_sq[k] = {||x - x0||^2, ||y - y0||^2}. false: not applicable.
  bool ComputeLocalizedLagrangianBoundsPair(const double* const x[2], const double* const y[2], const double* const kx[2], const double* const kty[2],
                                            double primal_weight, bool use_diagonal_solver, const double* x0, const double* y0, double out[2][4],
                                            double dist_sq[2][2]);
  void ComputeLocalizedLagrangianBoundsMaxNorm(const double* x, const double* y, double primal_weight, double radius, const double* kx,
                                               const double* kty, double out[4]);

  // download of the (possibly rescaled) problem in the caller's CSC order
  void DownloadValuesCsc(double* values);

  // scratch vectors (position order)
  double* tmp_n(int k) { return tmp_n_[k]; }
  double* tmp_m(int k) { return tmp_m_[k]; }

 private:
  void ColumnNorms(int norm, const double* row_scaling, const double* col_scaling, double* out);
  std::unique_ptr<Device> dev_;
  Comm* comm_ = nullptr;
  int64_t n_ = 0, m_ = 0, nnz_ = 0;
  int64_t m_global_ = 0, row_begin_ = 0;
  double* exchange_ = nullptr;  // [n + 1], row-sharded solves only
  PeerArena* arena_ = nullptr;  // row-sharded solves on one NVLink box
  PeerLayout layout_;
  int64_t slice_begin_ = 0, slice_end_ = 0;
  SellDev cols_slice_;
  bool has_cols_slice_ = false;
  int32_t* slice_perm_ = nullptr;
  double objective_offset_ = 0, objective_scaling_factor_ = 1;
  double *c_ = nullptr, *q_ = nullptr, *lv_ = nullptr, *uv_ = nullptr, *lc_ = nullptr, *uc_ = nullptr;
  SellDev rows_, cols_;
  bool device_built_ = false;
  DeviceBuildInfo build_info_;         // device builder: what the value download needs
  SellHost cols_meta_;                 // host builder: structure only (col/val released)
  std::vector<int64_t> col_starts_;    // caller's CSC column starts
  int32_t *primal_perm_ = nullptr, *dual_perm_ = nullptr;
  double* tmp_n_[4] = {nullptr, nullptr, nullptr, nullptr};
  double* tmp_m_[4] = {nullptr, nullptr, nullptr, nullptr};
  double *ones_n_ = nullptr, *ones_m_ = nullptr;
};

}  // namespace pdlp_b200

#endif  // PDLP_B200_DEVICE_PROBLEM_H_
