// device_ops.h -- C++ interface between the host PDHG driver (solver.cc) and
// the hand-written sm_100a kernels (device_ops.cu). Plain pointers + PODs; no
// CUDA types leak into the host translation units.
//
// Data layout in HBM (DESIGN.md "Data layout"):
//  * K is stored twice, as sliced-ELL with 32-row slices ("SELL-32"): once
//    row-major (rows of K; used for K x) and once column-major (rows of K^T;
//    used for K^T y), so neither product needs atomics -- the device
//    equivalent of ShardedQuadraticProgram's matrix + explicit transpose
//    (sharded_quadratic_program.cc:79-107).
//  * Every vector lives in "position order": dual-length vectors in the slot
//    order of the row-major copy, primal-length vectors in the slot order of
//    the column-major copy. Column indices stored in one copy are positions of
//    the other copy's vector order, so all epilogue accesses are coalesced.
//    Permutations are applied only at upload / download.
#ifndef PDLP_B200_DEVICE_OPS_H_
#define PDLP_B200_DEVICE_OPS_H_

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/pdlp_b200.h"

namespace pdlp_b200 {

class Comm;        // comm.h
struct PeerArena;  // comm.h

// ---------------------------------------------------------------------------
// Host-side image of one SELL-32 orientation (built by sell_builder.cc).
// ---------------------------------------------------------------------------
struct SellHost {
  int64_t num_rows = 0;      // logical rows of this orientation (= positions)
  int64_t num_cols = 0;      // length of the gathered vector
  int64_t num_split = 0;     // rows longer than split_len, cut into virtual slots
  int64_t num_virtual = 0;   // virtual slots in use
  int64_t num_virtual_padded = 0;  // rounded up to 32
  int64_t num_slots = 0;     // all slots, multiple of 32
  int64_t padded_nnz = 0;
  int32_t split_len = 0;
  std::vector<int64_t> slice_ptr;    // [num_slots/32 + 1]
  std::vector<int32_t> slot_len;     // [num_slots]
  std::vector<int32_t> col;          // [padded_nnz] position in the other order
  std::vector<double> val;           // [padded_nnz]
  std::vector<int32_t> split_first;  // [num_split+1] first virtual slot of split row
  std::vector<int32_t> virt_pos;     // [num_virtual_padded] position of the row a virtual slot belongs to (-1 pad)
  std::vector<int32_t> row_of_pos;   // [num_rows] original row index at a position
  std::vector<int32_t> pos_of_row;   // [num_rows]
};

struct QpHost {  // device-ready problem image
  int64_t n = 0, m = 0, nnz = 0;
  SellHost rows;  // K   : m logical rows, gathers primal positions
  SellHost cols;  // K^T : n logical rows, gathers dual positions
  bool has_q = false;
};

// Builds both SELL copies from the CSC view. `row_begin/row_end` restrict the
// image to a contiguous block of constraint rows (multi-GPU row sharding);
// pass 0 / m for the whole problem. Throws std::runtime_error on bad input.
QpHost BuildQpHost(const PdlpProblemView& view, int64_t row_begin, int64_t row_end,
                   int sigma = 4096, bool natural_primal_order = false);

// ---------------------------------------------------------------------------
// Device objects
// ---------------------------------------------------------------------------
struct SellDev {
  int64_t num_rows = 0, num_cols = 0, num_split = 0, num_virtual_padded = 0, num_slots = 0, padded_nnz = 0;
  int64_t* slice_ptr = nullptr;
  int32_t* slot_len = nullptr;
  int32_t* col = nullptr;
  double* val = nullptr;
  int32_t* split_first = nullptr;
  int32_t* virt_pos = nullptr;
  double* virt_partial = nullptr;  // [num_virtual_padded] scratch for split rows
};

// State of the device-resident adaptive / constant step loop
// (Solver::TakeAdaptiveStep / TakeConstantSizeStep, pdhg.cc:2558-2675). It
// lives in device memory; kernels read it at launch so that a rejected step
// needs no host round trip. The host mirrors it at checkpoints.
struct StepState {
  double step_size;
  double primal_weight;
  int32_t cur, prev, cand;          // buffer indices of x/y/K^T y
  int32_t iterations_completed;
  int32_t num_rejected_steps;
  int32_t inner_iterations;
  int32_t halt;                     // kHalt*
  int32_t k_stop;
  double kkt_pass_limit;
  double avg_weight_sum;
  int32_t avg_num_terms;
  int32_t rule;                     // PDLP_ADAPTIVE_LINESEARCH_RULE / PDLP_CONSTANT_STEP_SIZE_RULE
  double pending_ratio;             // deferred average update of the current iterate
  double reduction_exponent, growth_exponent;
  double last_dx2, last_dy2, last_nonlinearity, last_movement;
  int64_t attempts;                 // attempts actually executed (not no-ops)
  // (total + 1)^-reduction_exponent and ^-growth_exponent of the adaptive rule for the attempt
  // count `pow_total`: evaluated by the primal-step kernel, off the critical path of the decision
  // (two double-precision pow() in one thread are ~5 us). Ignored unless pow_total matches.
  double pow_reduction, pow_growth, pow_total;
  // Malitsky-Pock rule on the device (pdhg.cc:2463-2556): one attempt = one inner iteration.
  double mp_new_tau;                // trial primal step size of this attempt
  double mp_ratio;                  // ratio_last_two_step_sizes_
  double mp_interpolation, mp_downscaling, mp_contraction;
  int32_t mp_skip_primal;           // != 0: a retry (x' and K x' of this iteration exist) or halted
  int32_t avg_num_terms_primal;     // under Malitsky-Pock the primal average has its own weight / count
  double avg_weight_sum_primal;     //   (the starting point enters it once after every restart)
  double pending_ratio0;            // deferred update of the primal average with the PREVIOUS iterate (that entry)
  double pending_ratio_dual;        // deferred update of the dual average (= pending_ratio under the other rules)
};
enum { kHaltNone = 0, kHaltCheckpoint = 1, kHaltZeroMovement = 2, kHaltDivergent = 3, kHaltInnerLimit = 4, kHaltPeerTimeout = 5 };

// Layout (in doubles) of the peer arena of the row-sharded step loop; see
// PeerPtrs in device_ops.cu. The primal vector is cut into `world` slices of
// `stride` (even) columns; n_pad = world * stride is also the allocated length
// of every primal vector so that slices can be all-gathered in place.
struct PeerLayout {
  int64_t stride = 0, n_pad = 0, xt_off = 0, partial_off = 0, y_off = 0, scal_off = 0, flags_off = 0, epoch_off = 0, tr_off = 0, cand_off = 0, doubles = 0;
  int64_t tr2_off = 0, cand2_off = 0;  // second set of trust-region segments (barrier 4): the two solves of a restart test run concurrently
  static constexpr int64_t kTrCandCap = 4096;                  // = kTrFinishCap of the trust-region solve
  static constexpr int64_t kTrCandSegment = 3 * kTrCandCap + 8;  // per rank: keys, a, b of its candidates + their count
  static PeerLayout For(int64_t n, int64_t m_global, int world) {
    PeerLayout l;
    l.stride = 2 * ((n + 2 * world - 1) / (2 * world));
    if (l.stride == 0) l.stride = 2;
    l.n_pad = l.stride * world;
    l.xt_off = 0;
    l.partial_off = l.n_pad;
    l.y_off = 2 * l.n_pad;
    l.scal_off = l.y_off + 2 * ((m_global + 1) / 2) + 2;
    l.flags_off = l.scal_off + 4 * 8;
    l.epoch_off = l.flags_off + 8 * 8;   // barriers 0..7 (0 / 1: step loop A / B, 3 / 4: trust-region solves)
    l.tr_off = l.epoch_off + 8;
    l.cand_off = l.tr_off + 2 * 48 * 8;
    l.tr2_off = l.cand_off + kTrCandSegment * 8;
    l.cand2_off = l.tr2_off + 2 * 48 * 8;
    l.doubles = l.cand2_off + kTrCandSegment * 8;
    return l;
  }
};

// What the device builder keeps besides the two SELL images (device pointers).
struct DeviceBuildInfo {
  int64_t n = 0, m = 0, nnz = 0;
  int32_t* col_slot_row = nullptr;  // [cols.num_slots] original column of a slot (-1 = padding)
  int64_t* col_slot_off = nullptr;  // [cols.num_slots] first element of the slot inside its column
  int32_t* col_slot_stride = nullptr;  // [cols.num_slots] distance between consecutive elements of the slot (team size of a virtual slot, else 1)
  int64_t* col_start = nullptr;     // [n + 1] starts of the (row-block) columns in CSC order
};

struct MSideStats {  // reductions over the dual (row) side
  double linf_residual, sumsq_residual, cw_residual;  // PrimalResidualNorms
  double bounds_term;                                 // DualObjectiveBoundsTerm
  double linf_scaled, sumsq_scaled;                   // ||y o d_r||
};
struct NSideStats {  // reductions over the primal (column) side
  double correction, full_correction, linf_residual, sumsq_residual, cw_residual;  // DualResidualNorms
  double objective_dot, quadratic, linf_scaled, sumsq_scaled, linf_qx;
};
struct VectorInfoDev { double num_finite_nonzero, num_infinite, num_zero, largest, smallest, sum, sumsq; };

class Device {
 public:
  // Picks the CUDA device; throws std::runtime_error if none is usable.
  explicit Device(int cuda_device);
  ~Device();
  Device(const Device&) = delete;
  Device& operator=(const Device&) = delete;

  static int DeviceCount();  // 0 if no driver / device

  int64_t launches() const { return launches_; }
  // Row-sharded solves (one process per GPU): reductions over dual-side
  // (row-sharded) data are completed by an all-reduce on this communicator;
  // primal-side data is replicated. Not owned.
  void SetComm(Comm* comm) { comm_ = comm; }
  Comm* comm() const { return comm_; }
  bool count_primal() const;  // true on rank 0 / single GPU
  // Row-sharded solves keep primal-length vectors replicated; reductions over
  // them are split across ranks: this rank reduces [begin, end) of a vector of
  // length n and the partial results are all-reduced.
  // The mapped peer arenas of a row-sharded solve on one NVLink box (not owned);
  // the trust-region search exchanges its bin totals through them.
  void SetPeerArena(const PeerArena* arena, int64_t n, int64_t m_global) { peer_arena_ = arena; peer_arena_n_ = n; peer_arena_m_ = m_global; }
  void SetPrimalSlice(int64_t n, int64_t begin, int64_t end) { pslice_n_ = n; pslice_begin_ = begin; pslice_end_ = end; }
  bool PrimalSliced(int64_t n) const { return comm_ != nullptr && pslice_n_ == n && pslice_n_ >= 0; }
  int64_t PrimalSliceBegin(int64_t n) const { return PrimalSliced(n) ? pslice_begin_ : 0; }
  int64_t PrimalSliceEnd(int64_t n) const { return PrimalSliced(n) ? pslice_end_ : n; }
  // Rank 0's value on every rank (time limits / interrupt flags must lead to
  // the same control flow everywhere); identity without a communicator.
  double RootValue(double v);
  double MaxOverRanks(double v);
  void AllReduceSumVec(double* buf, int64_t n);
  void AllReduceMaxVec(double* buf, int64_t n);
  void* stream() const { return stream_; }
  void Sync();

  // raw memory
  double* AllocF64(int64_t n);
  void Free(void* p);
  void Upload(double* dst, const double* src, int64_t n);
  void Download(double* dst, const double* src, int64_t n);
  void CopyD2D(double* dst, const double* src, int64_t n);
  void Fill(double* dst, double value, int64_t n);
  // dst[p] = src_host[row_of_pos[p]] and inverse; perm lives on device.
  void UploadPermuted(double* dst, const double* src_host, const int32_t* row_of_pos_dev, int64_t n);
  void DownloadPermuted(double* dst_host, const double* src, const int32_t* row_of_pos_dev, int64_t n);
  int32_t* UploadI32(const std::vector<int32_t>& v);
  void ScatterInto(double* dst, const double* src, const int32_t* row_of_pos_dev, int64_t n);  // dst[row_of_pos[p]] = src[p]

  // Builds both SELL-32 images on the device from the caller's CSC arrays
  // (device_build.cu); *dual_perm / *primal_perm are the row_of_pos maps of the
  // row / column copy (device, owned by the caller).
  void BuildSellPair(const PdlpProblemView& view, int64_t row_begin, int64_t row_end, int sigma, bool natural_primal_order,
                     SellDev* rows, SellDev* cols, int32_t** dual_perm, int32_t** primal_perm, DeviceBuildInfo* info);
  // Column-slice image for the all-gather exchange of a row-sharded solve:
  // (K[:, col_begin:col_end])^T with indices into the box-wide dual order (device_build.cu).
  void BuildColumnSliceImage(const PdlpProblemView& view, int64_t col_begin, int64_t col_end, const int32_t* row_pos_global_dev, int sigma,
                             SellDev* out, int32_t** row_of_pos_out);
  void FreeBuildInfo(DeviceBuildInfo& info);
  void DownloadValuesCscFromSell(const SellDev& cols, const DeviceBuildInfo& info, double* values_host);
  SellDev UploadSell(const SellHost& h);
  void FreeSell(SellDev& s);
  void DownloadSellValues(const SellDev& s, std::vector<double>& out);

  // ---- SpMV + generic vector kernels (positions order) -------------------
  void SpMV(const SellDev& a, const double* x, double* out);           // out = A x
  // out[perm[pos]] = (A x)[pos]: the product scattered into another order
  // (row-sharded K^T y partials go to the caller's column order before the all-reduce).
  void SpMVScatter(const SellDev& a, const double* x, const int32_t* perm, double* out);
  // Raw row-wise max|a_ij s_j| (norm 0) or sum (a_ij s_j)^2 (norm 1) scattered through perm,
  // and the finishing step out = (norm ? sqrt(out) : out) * |own| after the all-reduce.
  void RowNormRawScatter(const SellDev& a, int norm, const double* other_scale, const int32_t* perm, double* out);
  void FinishRowNorm(double* out, int norm, const double* own_scale, int64_t n);
  // out[pos] = norm over the row of |a_ij * other_scale[col]| * |own_scale[pos]|; norm 0 LInf, 1 L2
  // (ScaledColLInfNorm / ScaledColL2Norm, sharder.cc:288-332).
  void ScaledRowNorm(const SellDev& a, int norm, const double* other_scale, const double* own_scale, double* out);
  // a_ij *= own[i]*other[j]; own is indexed by position, or by own_perm[position] when given
  void ScaleMatrix(SellDev& a, const double* own_scale, const double* other_scale, const int32_t* own_perm = nullptr);
  // row-sharded helpers for the all-gather exchange: out_full[row_begin + perm[p]] = row_begin + p
  // (global position of every row; other blocks zero) and out_full[row_begin + p] = v[p]
  void WriteGlobalRowPositions(double* out_full, const int32_t* row_of_pos, int64_t row_begin, int64_t m);
  void DoublesToI32(int32_t* dst, const double* src, int64_t n);
  void DivideBySqrt(double* vec, const double* divisor, int64_t n);                   // skip zeros (sou.cc:354-365)
  void Mul(double* dst, const double* a, int64_t n);                                  // dst *= a
  void Div(double* dst, const double* a, int64_t n);                                  // dst /= a
  void MulSq(double* dst, const double* a, int64_t n);                                // dst *= a*a
  void Axpy(double* dst, double s, const double* a, int64_t n);                       // dst += s*a
  void Sub(double* dst, const double* a, const double* b, int64_t n);                 // dst = a - b
  void ReplaceLargeWithInf(double* v, double threshold, int64_t n);
  void MapFiniteValuesToZero(double* dst, const double* src, int64_t n);             // pdhg.cc:2940-2949
  // DualTrustRegionProblem (trust_region.h:386-417): objective = -gradient, y_i >= 0 unless the
  // constraint has a finite upper bound, y_i <= 0 unless it has a finite lower bound.
  void DualTrustRegionProblem(const double* dual_gradient, const double* lc, const double* uc, double* objective, double* lb, double* ub, int64_t m);
  void ClampPrimal(double* x, const double* lb, const double* ub, bool feasibility_bounds, int64_t n);
  void ClampDual(double* y, const double* lc, const double* uc, int64_t m);
  void WeightedAverageAdd(double* avg, const double* v, double ratio, int64_t n);     // avg += ratio*(v-avg)

  // reductions (deterministic: fixed grid + fixed-order final pass); host result
  // `sharded`: the vector is row-sharded across ranks (all-reduced result)
  double Dot(const double* a, const double* b, int64_t n, bool sharded = false);
  double SumSq(const double* a, int64_t n, bool sharded = false);
  double SumSqDiff(const double* a, const double* b, int64_t n, bool sharded = false);
  double LInf(const double* a, int64_t n, bool sharded = false);
  double L1(const double* a, int64_t n, bool sharded = false);
  double ScaledLInf(const double* a, const double* s, int64_t n, bool sharded = false);
  double ScaledSumSq(const double* a, const double* s, int64_t n, bool sharded = false);
  void DistancesSq(const double* x, const double* x0, int64_t n, const double* y, const double* y0, int64_t m, double out[2]);
  VectorInfoDev VectorInfo(const double* v, int64_t n, bool sharded = false);                       // ComputeVectorInfo (sou.cc:179-191)
  VectorInfoDev CombinedBoundsInfo(const double* a, const double* b, int64_t n, bool sharded = false);  // sou.cc:223-238
  VectorInfoDev GapInfo(const double* lb, const double* ub, int64_t n);       // sou.cc:193-205
  VectorInfoDev MatrixInfo(const SellDev& a);                                 // sou.cc:207-221
  bool BoundsValid(const double* lb, const double* ub, int64_t n, bool sharded = false);           // HasValidBounds
  bool AllNonNegative(const double* v, int64_t n);

  // Several reductions, ONE device->host copy and synchronisation: between BeginBatch and EndBatch
  // the *Launch functions only enqueue and return where their results will be; Read* after EndBatch.
  void BeginBatch();
  void EndBatch();
  int DualSideStatsLaunch(const double* y, const double* kx, const double* lc, const double* uc, const double* dr, double cw_offset,
                          bool homogeneous_bounds, int64_t m);
  int PrimalSideStatsLaunch(const double* x, const double* x_for_bounds, const double* kty, const double* c, const double* q, const double* lv,
                            const double* uv, const double* dc, double cw_offset, bool zero_objective, bool handle_as_residuals, int64_t n);
  int ActiveSetPrimalLaunch(const double* x, const double* x0, const double* lv, const double* uv, int64_t n);
  int ActiveSetDualLaunch(const double* y, const double* y0, const double* lc, const double* uc, int64_t m);
  MSideStats ReadDualSideStats(int off) const;
  NSideStats ReadPrimalSideStats(int off) const;
  void ReadCounts(int off, int64_t out[2]) const;
  // KKT reductions (iteration_stats.cc:66-350). dr/dc may be null (= ones).
  MSideStats DualSideStats(const double* y, const double* kx, const double* lc, const double* uc, const double* dr,
                           double cw_offset, bool homogeneous_bounds, int64_t m);
  NSideStats PrimalSideStats(const double* x, const double* x_for_bounds, const double* kty, const double* c, const double* q,
                             const double* lv, const double* uv, const double* dc, double cw_offset, bool zero_objective,
                             bool handle_as_residuals, int64_t n);
  // out = (zero_objective ? 0 : q*x + c) - kty   (PrimalGradientFromObjectiveProduct)
  void PrimalGradient(const double* x, const double* kty, const double* c, const double* q, bool zero_objective, double* out, int64_t n);
  // ComputePrimalGradient / ComputeDualGradient (sou.cc:446-527): gradient + value
  double LagrangianPrimalGradient(const double* x, const double* kty, const double* c, const double* q, double* grad, int64_t n);
  double LagrangianDualGradient(const double* y, const double* kx, const double* lc, const double* uc, double* grad, int64_t m);
  // SetActiveSetInformation (pdhg.cc:1476-1545): out = {count, change}
  void ActiveSetPrimal(const double* x, const double* x0, const double* lv, const double* uv, int64_t n, int64_t out[2]);
  void ActiveSetDual(const double* y, const double* y0, const double* lc, const double* uc, int64_t m, int64_t out[2]);
  double RandomProjection(const double* v, int64_t n, uint32_t seed, uint32_t stream_id, bool sharded = false, int64_t index_offset = 0);

  // ---- trust region (trust_region.cc) ------------------------------------
  // Joint problem (trust_region.cc:115-162) over (x, y): returns lagrangian
  // value, lower, upper (ComputeLocalizedLagrangianBounds, Euclidean norm). radius < 0 with x0 / y0
  // given: the radius is the weighted distance of (x, y) to (x0, y0) (pdhg.cc:1998-2007), computed in the
  // same launch; extra_out (3 doubles) receives {radius, ||x - x0||^2, ||y - y0||^2} (the last two -1 when
  // the radius was given).
  void LocalizedLagrangianBounds(const double* x, const double* y, const double* kx, const double* kty, const double* c, const double* q,
                                 const double* lv, const double* uv, const double* lc, const double* uc, double primal_weight,
                                 double radius, bool use_diagonal_solver, double diagonal_tol, int64_t n, int64_t m, double out[3],
                                 const double* x0 = nullptr, const double* y0 = nullptr, double* extra_out = nullptr);
  // The same for TWO points at once (index 0 / 1; radius = weighted distance to (x0, y0)): concurrent launches,
  // one host synchronisation. false: not applicable here, solve them one by one. extra_out as above.
  bool LocalizedLagrangianBoundsPair(const double* const x[2], const double* const y[2], const double* const kx[2], const double* const kty[2],
                                     const double* c, const double* q, const double* lv, const double* uv, const double* lc, const double* uc,
                                     double primal_weight, int64_t n, int64_t m, const double* x0, const double* y0, double out[2][3],
                                     double extra_out[2][3]);
  // Explicit-vector problems (SolveTrustRegion / SolveDiagonalTrustRegion).
  void SolveTrustRegion(const double* obj, const double* lb, const double* ub, const double* center, const double* w, double radius,
                        int64_t n, double* solution, double* step_size, double* objective_value);
  void SolveDiagonalTrustRegion(const double* obj, const double* qdiag, const double* lb, const double* ub, const double* center,
                                const double* w, double radius, double tol, int64_t n, double* solution, double* step_size,
                                double* objective_value);

  // ---- PDHG step (hot path) ----------------------------------------------
  struct StepBuffers {
    int64_t n = 0, m = 0;
    double* x[3] = {nullptr, nullptr, nullptr};
    double* y[3] = {nullptr, nullptr, nullptr};
    double* kty[3] = {nullptr, nullptr, nullptr};
    double* kx[3] = {nullptr, nullptr, nullptr};  // K x of the iterates (maintained by the dual kernel)
    double* x_tilde = nullptr;
    double* avg_x = nullptr;
    double* avg_y = nullptr;
    // peer-memory exchange only (else nullptr): K x / K^T y of the average, maintained by the step kernels
    // alongside the averages (avg_kty on this rank's primal slice; all-gathered with avg_x)
    double* avg_kx = nullptr;
    double* avg_kty = nullptr;
    const double *c = nullptr, *q = nullptr, *lv = nullptr, *uv = nullptr, *lc = nullptr, *uc = nullptr;
    StepState* state = nullptr;  // device, TWO slots: an attempt reads one, its decision writes the other
    // row-sharded solve only: [n + 1] exchange buffer for the K^T y' partial
    // (+ this rank's ||dy||^2 in the last slot) and the scatter permutation
    double* exchange = nullptr;
    const int32_t* primal_scatter = nullptr;
    // peer-memory exchange (row-sharded solve on one NVLink box): the mapped
    // arenas and this rank's slice [slice_begin, slice_end) of the primal vector
    const PeerArena* arena = nullptr;
    int64_t slice_begin = 0, slice_end = 0, slice_stride = 0;
    int64_t m_global = 0, row_begin = 0;
    // all-gather exchange (chosen when the dual side is the shorter one): image
    // of (K[:, slice])^T gathering from the all-gathered y', and the column
    // (relative to slice_begin) at each of its positions
    const SellDev* cols_slice = nullptr;
    const int32_t* slice_perm = nullptr;
  };
  StepState* AllocState();  // two slots
  void UploadState(StepState* dev, const StepState& host);
  void DownloadState(StepState& host, const StepState* dev);
  // Both slots; returns the index of the one the last executed decision wrote (preferred_slot on a tie).
  int DownloadLatestState(StepState& host, const StepState* slots, int preferred_slot);
  // Enqueues `count` attempts of the fused 3-kernel step, the first one reading state slot
  // `first_slot`; attempts after the device sets `halt` are no-ops. Does not synchronise.
  void EnqueueSteps(const StepBuffers& b, const SellDev& rows, const SellDev& cols, int count, int first_slot);
  // Per-kernel CUDA-event sampling of the step loop (bench / profiling only).
  // While enabled, every `stride`-th attempt of an EnqueueSteps batch is
  // bracketed by events on the launching stream; CollectStepTimings (call it
  // after the batch has been synchronised) adds the samples whose attempt
  // index is < `executed_attempts` (later ones were halted no-ops).
  // Kernel classes: 0 primal step, 1 K x~ + dual epilogue (+fix-up),
  // 2 K^T y' (+fix-up; its block 0 takes the step decision), 3 row-sharded solves only: sums + barrier.
  struct StepTimings { double ms[4] = {0, 0, 0, 0}; int64_t samples[4] = {0, 0, 0, 0}; };
  void EnableStepTiming(bool on, int stride = 8);
  void CollectStepTimings(int64_t executed_attempts);
  const StepTimings& step_timings() const { return step_timings_; }
  // Device-timeline stopwatch on the launching stream (CUDA events).
  void TimelineStart(int id);        // id in {0, 1}
  double TimelineStopMs(int id);     // synchronises
  void TimelineStop(int id);         // records only; TimelineCollectMs (which synchronises) returns the time later
  double TimelineCollectMs(int id);  // 0 if nothing is pending
  // Applies the deferred average update (if any) for both averages.
  void FlushAverages(const StepBuffers& b, int slot);
  // Peer exchange only: inside the step loop every rank advances just its slice
  // of x, K^T y and the primal average; this all-gathers the slices of
  // x[cur], x[prev], kty[cur] and avg_x so that the replicated primal side is
  // whole again before restart / termination work reads it.
  void GatherPrimalSlices(const StepBuffers& b, int cur, int prev);

  // Device-resident Malitsky-Pock loop: `count` inner iterations (attempts), five launches each
  // (x' once per iteration | K x' once | dual trial | K^T y' + ||K^T y - K^T y'||^2 | decision).
  void EnqueueMalitskyPockSteps(const StepBuffers& b, const SellDev& rows, const SellDev& cols, int count, int first_slot);
  // Unfused pieces for the Malitsky-Pock rule (pdhg.cc:2463-2556): row-sharded solves drive it from the host.
  void PrimalStep(const double* x, const double* kty, const double* c, const double* q, const double* lv, const double* uv,
                  double tau, double* x_next, int64_t n);
  void DualStepFromProducts(const double* y, const double* kx_cur, const double* kx_next, const double* lc, const double* uc,
                            double sigma, double theta, double* y_next, int64_t m);

 private:
  friend struct DeviceImpl;
  int device_ = 0;
  Comm* comm_ = nullptr;
  void* stream_ = nullptr;
  void* stream2_ = nullptr;                 // second stream of LocalizedLagrangianBoundsPair (created on first use)
  void* pair_ev_[2] = {nullptr, nullptr};   // fork / join events of the pair
  double* partials_ = nullptr;   // reduction scratch
  double* results_ = nullptr;    // small device result vector
  double* host_results_ = nullptr;  // pinned
  int64_t launches_ = 0;
  int64_t pslice_n_ = -1, pslice_begin_ = 0, pslice_end_ = 0;
  const PeerArena* peer_arena_ = nullptr;
  int64_t peer_arena_n_ = 0, peer_arena_m_ = 0;
  int32_t* tr_peer_error_ = nullptr;  // device flag: a peer did not arrive at a barrier of the trust-region search
  unsigned int* tile_counters_ = nullptr;  // tile queues of the persistent SpMV pair of the step loop (one per kernel)
  unsigned int* loop_sync_ = nullptr;      // k_peer_loop: [0..2] grid-barrier words, [16..32) phase-time sums (u64)
  bool loop_traced_ = false;               // the last EnqueueSteps launched k_peer_loop with the phase trace on
  unsigned long long* loop_trace_host_ = nullptr;  // pinned copy of the phase-time sums of the last launch
  int num_sms_ = 148;
  // trust-region scratch (grown on demand)
  double* tr_scratch_ = nullptr;
  int64_t tr_scratch_size_ = 0;
  double* step_partials_ = nullptr;
  int64_t step_partials_size_ = 0;
  double* TrScratch(int64_t doubles);
  double* ReduceTarget(int count);
  bool batch_active_ = false;
  int batch_off_ = 0, last_result_off_ = 0;
  // step timing
  bool step_timing_ = false;
  int step_timing_stride_ = 8;
  static constexpr int kEvPerSlot = 8;
  std::vector<void*> timing_events_;     // kEvPerSlot per sample slot
  bool timing_peer_ = false;
  double detail_ms_[7] = {0, 0, 0, 0, 0, 0, 0};  // PDLP_B200_TRACE: per sub-phase
  int64_t detail_samples_ = 0;
  std::vector<int> timing_attempt_idx_;  // attempt index of each used slot in the last batch
  StepTimings step_timings_;
  void* timeline_ev_[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  bool timeline_pending_[2] = {false, false};
};

}  // namespace pdlp_b200

#endif  // PDLP_B200_DEVICE_OPS_H_
