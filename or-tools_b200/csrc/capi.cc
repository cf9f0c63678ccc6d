// capi.cc -- the extern "C" boundary declared in include/pdlp_b200.h.
// No exception crosses the boundary: CUDA / allocation failures become
// PDLP_B200_STATUS_* codes (and TERMINATION_REASON_OTHER for the solve).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>

#include "comm.h"
#include "device_problem.h"
#include "solver.h"

using namespace pdlp_b200;

struct PdlpDeviceProblem {
  std::unique_ptr<DeviceProblem> p;
  PdlpParams default_params;
};
struct PdlpSolveSession {
  std::unique_ptr<SolveSession> s;
  PdlpMessageCallback msg_cb = nullptr;
  PdlpIterationStatsCallback stats_cb = nullptr;
  void* user = nullptr;
};
static int g_default_device = 0;
struct PdlpDistributedContext {
  std::unique_ptr<Comm> comm;
};

namespace {

char* DupString(const std::string& s) {
  char* p = static_cast<char*>(std::malloc(s.size() + 1));
  std::memcpy(p, s.c_str(), s.size() + 1);
  return p;
}
void FillResult(SolverResultCpp&& r, PdlpResult* out) {
  std::memset(out, 0, sizeof(*out));
  out->primal_size = static_cast<int64_t>(r.primal_solution.size());
  out->dual_size = static_cast<int64_t>(r.dual_solution.size());
  out->primal_solution = r.primal_solution.release();  // (malloc storage handed over, no copy)
  out->dual_solution = r.dual_solution.release();
  out->reduced_costs = r.reduced_costs.release();
  const SolveLogCpp& l = r.solve_log;
  out->instance_name = l.instance_name ? DupString(*l.instance_name) : nullptr;
  out->termination_reason = l.termination_reason;
  out->termination_string = l.termination_string ? DupString(*l.termination_string) : nullptr;
  out->iteration_count = l.iteration_count;
  out->solve_time_sec = l.solve_time_sec;
  out->preprocessing_time_sec = l.preprocessing_time_sec;
  out->solution_type = l.solution_type;
  out->has_solution_stats = l.has_solution_stats;
  out->solution_stats = l.solution_stats;
  out->has_original_problem_stats = l.has_original_stats;
  out->has_preprocessed_problem_stats = l.has_preprocessed_stats;
  out->original_problem_stats = l.original_stats;
  out->preprocessed_problem_stats = l.preprocessed_stats;
  out->num_iteration_stats = static_cast<int64_t>(l.iteration_stats.size());
  if (!l.iteration_stats.empty()) {
    out->iteration_stats = static_cast<PdlpIterationStats*>(std::malloc(l.iteration_stats.size() * sizeof(PdlpIterationStats)));
    std::memcpy(out->iteration_stats, l.iteration_stats.data(), l.iteration_stats.size() * sizeof(PdlpIterationStats));
  }
  out->params = l.params;
  out->num_feasibility_polishing_details = static_cast<int64_t>(l.feasibility_polishing_details.size());
  if (!l.feasibility_polishing_details.empty()) {
    out->feasibility_polishing_details =
        static_cast<PdlpFeasibilityPolishingDetails*>(std::calloc(l.feasibility_polishing_details.size(), sizeof(PdlpFeasibilityPolishingDetails)));
    for (size_t k = 0; k < l.feasibility_polishing_details.size(); ++k) {
      const PolishingDetailsCpp& d = l.feasibility_polishing_details[k];
      PdlpFeasibilityPolishingDetails& o = out->feasibility_polishing_details[k];
      o.polishing_phase_type = d.polishing_phase_type;
      o.main_iteration_count = d.main_iteration_count;
      o.params = d.params;
      o.termination_reason = d.termination_reason;
      o.iteration_count = d.iteration_count;
      o.solve_time_sec = d.solve_time_sec;
      o.solution_stats = d.solution_stats;
      o.solution_type = d.solution_type;
      o.num_iteration_stats = static_cast<int64_t>(d.iteration_stats.size());
      if (!d.iteration_stats.empty()) {
        o.iteration_stats = static_cast<PdlpIterationStats*>(std::malloc(d.iteration_stats.size() * sizeof(PdlpIterationStats)));
        std::memcpy(o.iteration_stats, d.iteration_stats.data(), d.iteration_stats.size() * sizeof(PdlpIterationStats));
      }
    }
  }
  out->gpu_kernel_launches = l.gpu_kernel_launches;
  out->device_iteration_time_sec = l.device_iteration_time_sec;
}

// Runs `body`, mapping exceptions to status codes.
template <class F>
int32_t Guard(F body) {
  try {
    if (Device::DeviceCount() <= 0) return PDLP_B200_STATUS_NO_DEVICE;
    body();
    return PDLP_B200_STATUS_OK;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "pdlp_b200: %s\n", e.what());
    return PDLP_B200_STATUS_CUDA_ERROR;
  }
}

// RAII device vector uploaded from a host pointer (position order of `p`).
struct PrimalVec {
  DeviceProblem& p; double* d;
  PrimalVec(DeviceProblem& p_, const double* host) : p(p_), d(p_.NewPrimal()) { if (host != nullptr) p.UploadPrimal(d, host); }
  ~PrimalVec() { p.dev().Free(d); }
};
struct DualVec {
  DeviceProblem& p; double* d;
  DualVec(DeviceProblem& p_, const double* host) : p(p_), d(p_.NewDual()) { if (host != nullptr) p.UploadDual(d, host); }
  ~DualVec() { p.dev().Free(d); }
};
struct PlainVec {  // natural order
  Device& dev; double* d; int64_t n;
  PlainVec(Device& dv, const double* host, int64_t n_) : dev(dv), d(dv.AllocF64(n_)), n(n_) { if (host != nullptr) dev.Upload(d, host, n); }
  ~PlainVec() { dev.Free(d); }
};

}  // namespace

extern "C" {

void pdlp_b200_params_set_defaults(PdlpParams* params) { SetDefaultParams(params); }

int32_t pdlp_b200_params_validate(const PdlpParams* params, char* message, int64_t capacity) {
  const std::string e = ValidateParams(*params);
  if (message != nullptr && capacity > 0) {
    std::strncpy(message, e.c_str(), static_cast<size_t>(capacity - 1));
    message[capacity - 1] = 0;
  }
  return e.empty() ? 1 : 0;
}

}  // extern "C"
namespace {
template <class T>
T* CopyOut(const std::vector<T>& v) {
  T* p = static_cast<T*>(std::malloc(std::max<size_t>(1, v.size()) * sizeof(T)));
  if (p != nullptr && !v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(T));
  return p;
}
void FillLayout(const SellHost& h, PdlpSellLayout* o) {
  o->num_rows = h.num_rows;
  o->num_cols = h.num_cols;
  o->num_split = h.num_split;
  o->num_virtual = h.num_virtual;
  o->num_virtual_padded = h.num_virtual_padded;
  o->num_slots = h.num_slots;
  o->padded_nnz = h.padded_nnz;
  o->split_len = h.split_len;
  o->slice_ptr = CopyOut(h.slice_ptr);
  o->slot_len = CopyOut(h.slot_len);
  o->col = CopyOut(h.col);
  o->val = CopyOut(h.val);
  o->split_first = CopyOut(h.split_first);
  o->virt_pos = CopyOut(h.virt_pos);
  o->row_of_pos = CopyOut(h.row_of_pos);
  o->pos_of_row = CopyOut(h.pos_of_row);
}
}  // namespace
extern "C" {

int32_t pdlp_b200_host_sell_layout(const PdlpProblemView* qp, int64_t row_begin, int64_t row_end, int32_t sigma, int32_t natural_primal_order,
                                   PdlpSellLayout* out_rows, PdlpSellLayout* out_cols) {
  if (qp == nullptr || out_rows == nullptr || out_cols == nullptr) return PDLP_B200_STATUS_BAD_ARGUMENT;
  std::memset(out_rows, 0, sizeof *out_rows);
  std::memset(out_cols, 0, sizeof *out_cols);
  try {
    const QpHost h = BuildQpHost(*qp, row_begin, row_end, sigma, natural_primal_order != 0);
    FillLayout(h.rows, out_rows);
    FillLayout(h.cols, out_cols);
    return PDLP_B200_STATUS_OK;
  } catch (const std::exception&) {
    return PDLP_B200_STATUS_BAD_ARGUMENT;
  }
}

void pdlp_b200_sell_layout_free(PdlpSellLayout* layout) {
  if (layout == nullptr) return;
  std::free(layout->slice_ptr);
  std::free(layout->slot_len);
  std::free(layout->col);
  std::free(layout->val);
  std::free(layout->split_first);
  std::free(layout->virt_pos);
  std::free(layout->row_of_pos);
  std::free(layout->pos_of_row);
  std::memset(layout, 0, sizeof *layout);
}

int32_t pdlp_b200_device_count(void) { return Device::DeviceCount(); }
const char* pdlp_b200_version(void) { return "pdlp_b200 0.1.0 (sm_100a)"; }
int64_t pdlp_b200_sizeof(int32_t index) {
  switch (index) {
    case 0: return sizeof(PdlpTerminationCriteria);
    case 1: return sizeof(PdlpParams);
    case 2: return sizeof(PdlpProblemView);
    case 3: return sizeof(PdlpQuadraticProgramStats);
    case 4: return sizeof(PdlpConvergenceInformation);
    case 5: return sizeof(PdlpInfeasibilityInformation);
    case 6: return sizeof(PdlpPointMetadata);
    case 7: return sizeof(PdlpIterationStats);
    case 8: return sizeof(PdlpBoundNorms);
    case 9: return sizeof(PdlpIterationCallbackInfo);
    case 10: return sizeof(PdlpResult);
    case 11: return sizeof(PdlpSessionStatus);
    default: return -1;
  }
}

int32_t pdlp_b200_primal_dual_hybrid_gradient(const PdlpProblemView* qp, const PdlpParams* params, const double* initial_primal,
                                              int64_t initial_primal_size, const double* initial_dual, int64_t initial_dual_size,
                                              const volatile int32_t* interrupt_solve, PdlpMessageCallback message_callback,
                                              PdlpIterationStatsCallback stats_callback, void* user_data, PdlpResult* result) {
  if (qp == nullptr || params == nullptr || result == nullptr) return PDLP_B200_STATUS_BAD_ARGUMENT;
  std::memset(result, 0, sizeof(*result));
  if (Device::DeviceCount() <= 0) return PDLP_B200_STATUS_NO_DEVICE;
  Logger logger{message_callback, user_data};
  std::optional<InitialSolution> init;
  if (initial_primal != nullptr || initial_dual != nullptr) {
    init.emplace();
    if (initial_primal != nullptr) init->primal.assign(initial_primal, initial_primal + initial_primal_size);
    if (initial_dual != nullptr) init->dual.assign(initial_dual, initial_dual + initial_dual_size);
  }
  StatsCallback cb;
  if (stats_callback != nullptr) cb = [=](const PdlpIterationCallbackInfo& info) { stats_callback(&info, user_data); };
  try {
    FillResult(PrimalDualHybridGradient(*qp, *params, std::move(init), interrupt_solve, logger, std::move(cb), g_default_device), result);
    return PDLP_B200_STATUS_OK;
  } catch (const std::exception& e) {
    SolverResultCpp r;
    r.solve_log.termination_reason = PDLP_TERMINATION_REASON_OTHER;
    r.solve_log.termination_string = std::string("device failure: ") + e.what();
    FillResult(std::move(r), result);
    return PDLP_B200_STATUS_CUDA_ERROR;
  }
}

void pdlp_b200_result_free(PdlpResult* r) {
  if (r == nullptr) return;
  std::free(r->primal_solution); std::free(r->dual_solution); std::free(r->reduced_costs);
  std::free(r->instance_name); std::free(r->termination_string); std::free(r->iteration_stats);
  for (int64_t k = 0; k < r->num_feasibility_polishing_details; ++k) std::free(r->feasibility_polishing_details[k].iteration_stats);
  std::free(r->feasibility_polishing_details);
  std::memset(r, 0, sizeof(*r));
}

int32_t pdlp_b200_set_default_device(int32_t cuda_device) {
  if (cuda_device < 0 || cuda_device >= Device::DeviceCount()) return PDLP_B200_STATUS_BAD_ARGUMENT;
  g_default_device = cuda_device;
  return PDLP_B200_STATUS_OK;
}

// ---- sessions -----------------------------------------------------------------
int32_t pdlp_b200_session_create(const PdlpProblemView* qp, const PdlpParams* params, const double* initial_primal, int64_t initial_primal_size,
                                 const double* initial_dual, int64_t initial_dual_size, PdlpMessageCallback message_callback,
                                 PdlpIterationStatsCallback stats_callback, void* user_data, int32_t cuda_device, PdlpSolveSession** out) {
  if (qp == nullptr || params == nullptr || out == nullptr) return PDLP_B200_STATUS_BAD_ARGUMENT;
  *out = nullptr;
  return Guard([&] {
    auto h = std::make_unique<PdlpSolveSession>();
    h->msg_cb = message_callback;
    h->stats_cb = stats_callback;
    h->user = user_data;
    Logger logger{message_callback, user_data};
    std::optional<InitialSolution> init;
    if (initial_primal != nullptr || initial_dual != nullptr) {
      init.emplace();
      if (initial_primal != nullptr) init->primal.assign(initial_primal, initial_primal + initial_primal_size);
      if (initial_dual != nullptr) init->dual.assign(initial_dual, initial_dual + initial_dual_size);
    }
    StatsCallback cb;
    if (stats_callback != nullptr) cb = [=](const PdlpIterationCallbackInfo& info) { stats_callback(&info, user_data); };
    h->s = SolveSession::Create(*qp, *params, std::move(init), logger, std::move(cb), cuda_device);
    *out = h.release();
  });
}
int32_t pdlp_b200_session_advance(PdlpSolveSession* h, int32_t target_iterations, const volatile int32_t* interrupt_solve, PdlpSessionStatus* out) {
  if (h == nullptr) return PDLP_B200_STATUS_BAD_ARGUMENT;
  return Guard([&] {
    h->s->Advance(target_iterations, interrupt_solve);
    if (out != nullptr) h->s->Status(out);
  });
}
int32_t pdlp_b200_session_enable_timing(PdlpSolveSession* h, int32_t enable, int32_t stride) {
  if (h == nullptr) return PDLP_B200_STATUS_BAD_ARGUMENT;
  return Guard([&] { h->s->EnableTiming(enable != 0, stride); });
}
int32_t pdlp_b200_session_status(PdlpSolveSession* h, PdlpSessionStatus* out) {
  if (h == nullptr || out == nullptr) return PDLP_B200_STATUS_BAD_ARGUMENT;
  return Guard([&] { h->s->Status(out); });
}
int32_t pdlp_b200_session_finish(PdlpSolveSession* h, PdlpResult* result) {
  if (h == nullptr || result == nullptr) return PDLP_B200_STATUS_BAD_ARGUMENT;
  std::memset(result, 0, sizeof(*result));
  return Guard([&] { FillResult(h->s->Finish(), result); });
}
void pdlp_b200_session_destroy(PdlpSolveSession* h) { delete h; }

// ---- multi-GPU: row-sharded solve over NCCL ---------------------
int32_t pdlp_b200_nccl_unique_id(const char* nccl_library_path, uint8_t out_id[128]) {
  if (out_id == nullptr) return PDLP_B200_STATUS_BAD_ARGUMENT;
  try {
    Comm::UniqueId(nccl_library_path, out_id);
    return PDLP_B200_STATUS_OK;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "pdlp_b200: %s\n", e.what());
    return PDLP_B200_STATUS_CUDA_ERROR;
  }
}
int32_t pdlp_b200_distributed_init(const char* nccl_library_path, int32_t rank, int32_t world_size, int32_t cuda_device,
                                   const uint8_t nccl_unique_id[128], PdlpDistributedContext** out) {
  if (out == nullptr || nccl_unique_id == nullptr) return PDLP_B200_STATUS_BAD_ARGUMENT;
  *out = nullptr;
  return Guard([&] {
    auto h = std::make_unique<PdlpDistributedContext>();
    h->comm.reset(new Comm(nccl_library_path, rank, world_size, cuda_device, nccl_unique_id));
    *out = h.release();
  });
}
void pdlp_b200_distributed_destroy(PdlpDistributedContext* c) { delete c; }
int32_t pdlp_b200_row_block(const PdlpProblemView* qp, int32_t rank, int32_t world_size, int64_t* row_begin, int64_t* row_end) {
  if (qp == nullptr || row_begin == nullptr || row_end == nullptr || world_size < 1 || rank < 0 || rank >= world_size) return PDLP_B200_STATUS_BAD_ARGUMENT;
  try {
    ComputeRowBlock(*qp, rank, world_size, row_begin, row_end);
    return PDLP_B200_STATUS_OK;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "pdlp_b200: %s\n", e.what());
    return PDLP_B200_STATUS_BAD_ARGUMENT;
  }
}
int32_t pdlp_b200_peer_arena_layout(int64_t num_variables, int64_t num_constraints, int32_t world_size, int64_t out[13]) {
  if (out == nullptr || num_variables < 0 || num_constraints < 0 || world_size < 1 || world_size > kMaxPeers) return PDLP_B200_STATUS_BAD_ARGUMENT;
  const PeerLayout l = PeerLayout::For(num_variables, num_constraints, world_size);
  const int64_t v[13] = {l.stride, l.n_pad, l.xt_off, l.partial_off, l.y_off, l.scal_off, l.flags_off, l.epoch_off, l.tr_off, l.cand_off, l.tr2_off, l.cand2_off, l.doubles};
  for (int k = 0; k < 13; ++k) out[k] = v[k];
  return PDLP_B200_STATUS_OK;
}
int32_t pdlp_b200_primal_dual_hybrid_gradient_distributed(PdlpDistributedContext* ctx, const PdlpProblemView* qp, const PdlpParams* params,
                                                          const double* initial_primal, int64_t initial_primal_size, const double* initial_dual,
                                                          int64_t initial_dual_size, const volatile int32_t* interrupt_solve,
                                                          PdlpMessageCallback message_callback, PdlpIterationStatsCallback stats_callback,
                                                          void* user_data, PdlpResult* result) {
  if (ctx == nullptr || qp == nullptr || params == nullptr || result == nullptr) return PDLP_B200_STATUS_BAD_ARGUMENT;
  std::memset(result, 0, sizeof(*result));
  if (Device::DeviceCount() <= 0) return PDLP_B200_STATUS_NO_DEVICE;
  Logger logger{message_callback, user_data};
  std::optional<InitialSolution> init;
  if (initial_primal != nullptr || initial_dual != nullptr) {
    init.emplace();
    if (initial_primal != nullptr) init->primal.assign(initial_primal, initial_primal + initial_primal_size);
    if (initial_dual != nullptr) init->dual.assign(initial_dual, initial_dual + initial_dual_size);
  }
  StatsCallback cb;
  if (stats_callback != nullptr) cb = [=](const PdlpIterationCallbackInfo& info) { stats_callback(&info, user_data); };
  try {
    FillResult(PrimalDualHybridGradient(*qp, *params, std::move(init), interrupt_solve, logger, std::move(cb), ctx->comm->cuda_device(), ctx->comm.get()),
               result);
    return PDLP_B200_STATUS_OK;
  } catch (const std::exception& e) {
    SolverResultCpp r;
    r.solve_log.termination_reason = PDLP_TERMINATION_REASON_OTHER;
    r.solve_log.termination_string = std::string("device failure: ") + e.what();
    FillResult(std::move(r), result);
    return PDLP_B200_STATUS_CUDA_ERROR;
  }
}
int32_t pdlp_b200_session_create_distributed(PdlpDistributedContext* ctx, const PdlpProblemView* qp, const PdlpParams* params,
                                             PdlpSolveSession** out) {
  if (ctx == nullptr || qp == nullptr || params == nullptr || out == nullptr) return PDLP_B200_STATUS_BAD_ARGUMENT;
  *out = nullptr;
  return Guard([&] {
    auto h = std::make_unique<PdlpSolveSession>();
    Logger logger{nullptr, nullptr};
    h->s = SolveSession::Create(*qp, *params, std::nullopt, logger, StatsCallback(), ctx->comm->cuda_device(), ctx->comm.get());
    *out = h.release();
  });
}

// ---- device-resident problem -----------------------------------------------
int32_t pdlp_b200_problem_create(const PdlpProblemView* qp, int32_t cuda_device, PdlpDeviceProblem** out) {
  if (qp == nullptr || out == nullptr) return PDLP_B200_STATUS_BAD_ARGUMENT;
  *out = nullptr;
  return Guard([&] {
    auto h = std::make_unique<PdlpDeviceProblem>();
    h->p.reset(new DeviceProblem(*qp, cuda_device));
    SetDefaultParams(&h->default_params);
    *out = h.release();
  });
}
void pdlp_b200_problem_destroy(PdlpDeviceProblem* h) { delete h; }

int32_t pdlp_b200_transposed_matrix_vector_product(PdlpDeviceProblem* h, const double* y, double* out) {
  return Guard([&] {
    DeviceProblem& p = *h->p;
    DualVec yv(p, y);
    PrimalVec o(p, nullptr);
    p.KTy(yv.d, o.d);
    p.DownloadPrimal(out, o.d);
  });
}
int32_t pdlp_b200_matrix_vector_product(PdlpDeviceProblem* h, const double* x, double* out) {
  return Guard([&] {
    DeviceProblem& p = *h->p;
    PrimalVec xv(p, x);
    DualVec o(p, nullptr);
    p.Kx(xv.d, o.d);
    p.DownloadDual(out, o.d);
  });
}
int32_t pdlp_b200_apply_rescaling(PdlpDeviceProblem* h, int32_t ruiz, int32_t l2, double* row_scaling, double* col_scaling) {
  return Guard([&] {
    DeviceProblem& p = *h->p;
    double *r = nullptr, *c = nullptr;
    p.ApplyRescaling(ruiz, l2 != 0, &r, &c);
    p.DownloadDual(row_scaling, r);
    p.DownloadPrimal(col_scaling, c);
    p.dev().Free(r);
    p.dev().Free(c);
  });
}
int32_t pdlp_b200_scaling_iterations(PdlpDeviceProblem* h, int32_t norm, int32_t iters, double* row_scaling, double* col_scaling) {
  return Guard([&] {
    DeviceProblem& p = *h->p;
    DualVec r(p, row_scaling);
    PrimalVec c(p, col_scaling);
    p.ApplyScalingIterationsForNorm(iters, norm, r.d, c.d);
    p.DownloadDual(row_scaling, r.d);
    p.DownloadPrimal(col_scaling, c.d);
  });
}
int32_t pdlp_b200_scaled_col_norm(PdlpDeviceProblem* h, int32_t norm, const double* row_scaling, const double* col_scaling, double* out) {
  return Guard([&] {
    DeviceProblem& p = *h->p;
    DualVec r(p, row_scaling);
    PrimalVec c(p, col_scaling), o(p, nullptr);
    p.dev().ScaledRowNorm(p.cols(), norm, r.d, c.d, o.d);
    p.DownloadPrimal(out, o.d);
  });
}
int32_t pdlp_b200_scaled_row_norm(PdlpDeviceProblem* h, int32_t norm, const double* row_scaling, const double* col_scaling, double* out) {
  return Guard([&] {
    DeviceProblem& p = *h->p;
    DualVec r(p, row_scaling), o(p, nullptr);
    PrimalVec c(p, col_scaling);
    p.dev().ScaledRowNorm(p.rows(), norm, c.d, r.d, o.d);
    p.DownloadDual(out, o.d);
  });
}
int32_t pdlp_b200_rescale_quadratic_program(PdlpDeviceProblem* h, const double* col_scaling, const double* row_scaling) {
  return Guard([&] {
    DeviceProblem& p = *h->p;
    PrimalVec c(p, col_scaling);
    DualVec r(p, row_scaling);
    p.RescaleQuadraticProgram(c.d, r.d);
    p.dev().Sync();
  });
}
int32_t pdlp_b200_problem_download(PdlpDeviceProblem* h, double* values, double* objective_vector, double* objective_matrix_diagonal,
                                   double* clb, double* cub, double* vlb, double* vub) {
  return Guard([&] {
    DeviceProblem& p = *h->p;
    if (values) p.DownloadValuesCsc(values);
    if (objective_vector) p.DownloadPrimal(objective_vector, p.c());
    if (objective_matrix_diagonal && p.q() != nullptr) p.DownloadPrimal(objective_matrix_diagonal, p.q());
    if (clb) p.DownloadDual(clb, p.lc());
    if (cub) p.DownloadDual(cub, p.uc());
    if (vlb) p.DownloadPrimal(vlb, p.lv());
    if (vub) p.DownloadPrimal(vub, p.uv());
  });
}
int32_t pdlp_b200_compute_stats(PdlpDeviceProblem* h, PdlpQuadraticProgramStats* out) {
  return Guard([&] { *out = h->p->ComputeStats(); });
}
int32_t pdlp_b200_project_to_primal_variable_bounds(PdlpDeviceProblem* h, double* primal, int32_t use_feasibility_bounds) {
  return Guard([&] {
    DeviceProblem& p = *h->p;
    PrimalVec x(p, primal);
    p.dev().ClampPrimal(x.d, p.lv(), p.uv(), use_feasibility_bounds != 0, p.n());
    p.DownloadPrimal(primal, x.d);
  });
}
int32_t pdlp_b200_project_to_dual_variable_bounds(PdlpDeviceProblem* h, double* dual) {
  return Guard([&] {
    DeviceProblem& p = *h->p;
    DualVec y(p, dual);
    p.dev().ClampDual(y.d, p.lc(), p.uc(), p.m());
    p.DownloadDual(dual, y.d);
  });
}
int32_t pdlp_b200_compute_primal_gradient(PdlpDeviceProblem* h, const double* primal, const double* dual_product, double* gradient, double* value) {
  return Guard([&] {
    DeviceProblem& p = *h->p;
    PrimalVec x(p, primal), dp(p, dual_product), g(p, nullptr);
    *value = p.dev().LagrangianPrimalGradient(x.d, dp.d, p.c(), p.q(), g.d, p.n());
    p.DownloadPrimal(gradient, g.d);
  });
}
int32_t pdlp_b200_compute_dual_gradient(PdlpDeviceProblem* h, const double* dual, const double* primal_product, double* gradient, double* value) {
  return Guard([&] {
    DeviceProblem& p = *h->p;
    DualVec y(p, dual), pp(p, primal_product), g(p, nullptr);
    *value = p.dev().LagrangianDualGradient(y.d, pp.d, p.lc(), p.uc(), g.d, p.m());
    p.DownloadDual(gradient, g.d);
  });
}
int32_t pdlp_b200_compute_convergence_information(PdlpDeviceProblem* h, const PdlpParams* params, const double* col_scaling, const double* row_scaling,
                                                  const double* primal, const double* dual, double cw_primal_offset, double cw_dual_offset,
                                                  int32_t candidate_type, PdlpConvergenceInformation* out) {
  return Guard([&] {
    DeviceProblem& p = *h->p;
    PrimalVec x(p, primal), cs(p, col_scaling);
    DualVec y(p, dual), rs(p, row_scaling);
    *out = p.ComputeConvergenceInformation(params->handle_some_primal_gradients_on_finite_bounds_as_residuals != 0, col_scaling ? cs.d : nullptr,
                                           row_scaling ? rs.d : nullptr, x.d, y.d, nullptr, cw_primal_offset, cw_dual_offset, candidate_type);
  });
}
int32_t pdlp_b200_compute_infeasibility_information(PdlpDeviceProblem* h, const PdlpParams* params, const double* col_scaling,
                                                    const double* row_scaling, const double* primal_ray, const double* dual_ray,
                                                    const double* primal_for_residual_tests, int32_t candidate_type, PdlpInfeasibilityInformation* out) {
  return Guard([&] {
    DeviceProblem& p = *h->p;
    PrimalVec x(p, primal_ray), xr(p, primal_for_residual_tests), cs(p, col_scaling);
    DualVec y(p, dual_ray), rs(p, row_scaling);
    *out = p.ComputeInfeasibilityInformation(params->handle_some_primal_gradients_on_finite_bounds_as_residuals != 0, col_scaling ? cs.d : nullptr,
                                             row_scaling ? rs.d : nullptr, x.d, y.d, xr.d, nullptr, candidate_type);
  });
}
int32_t pdlp_b200_reduced_costs(PdlpDeviceProblem* h, const PdlpParams*, const double* primal, const double* dual, int32_t use_zero_primal_objective,
                                double* out) {
  return Guard([&] {
    DeviceProblem& p = *h->p;
    PrimalVec x(p, primal), o(p, nullptr);
    DualVec y(p, dual);
    p.ReducedCosts(x.d, y.d, use_zero_primal_objective != 0, o.d);
    p.DownloadPrimal(out, o.d);
  });
}
int32_t pdlp_b200_compute_localized_lagrangian_bounds(PdlpDeviceProblem* h, const double* primal, const double* dual, double primal_weight,
                                                      double radius, const double* primal_product, const double* dual_product,
                                                      int32_t use_diagonal_solver, double diagonal_tol, double out[4]) {
  return Guard([&] {
    DeviceProblem& p = *h->p;
    PrimalVec x(p, primal), dp(p, dual_product);
    DualVec y(p, dual), pp(p, primal_product);
    p.ComputeLocalizedLagrangianBounds(x.d, y.d, primal_weight, radius, primal_product ? pp.d : nullptr, dual_product ? dp.d : nullptr,
                                       use_diagonal_solver != 0, diagonal_tol, out);
  });
}

int32_t pdlp_b200_compute_localized_lagrangian_bounds_max_norm(PdlpDeviceProblem* h, const double* primal, const double* dual, double primal_weight,
                                                               double radius, const double* primal_product, const double* dual_product, double out[4]) {
  return Guard([&] {
    DeviceProblem& p = *h->p;
    PrimalVec x(p, primal), dp(p, dual_product);
    DualVec y(p, dual), pp(p, primal_product);
    p.ComputeLocalizedLagrangianBoundsMaxNorm(x.d, y.d, primal_weight, radius, primal_product ? pp.d : nullptr, dual_product ? dp.d : nullptr, out);
  });
}

// ---- problem-free vector entry points ---------------------------------------
}  // extern "C"
namespace {
// Preconditions of SolveTrustRegion / SolveDiagonalTrustRegion (trust_region.h:58-89): a radius that is
// not negative (nor NaN) and strictly positive norm weights. Checked on the host, before any device work.
bool TrustRegionArgumentsOk(int64_t size, const double* weights, double target_radius) {
  if (size < 0 || (size > 0 && weights == nullptr) || !(target_radius >= 0.0)) return false;
  for (int64_t i = 0; i < size; ++i)
    if (!(weights[i] > 0.0)) return false;
  return true;
}
}  // namespace
extern "C" {
int32_t pdlp_b200_solve_trust_region(int32_t cuda_device, int64_t size, const double* objective, const double* lb, const double* ub,
                                     const double* center, const double* weights, double target_radius, double* solution, double* step_size,
                                     double* objective_value) {
  // the reference CHECK-fails on these (trust_region.cc: "target_radius >= 0.0", "norm_weights_are_positive"); here they are a status
  if (!TrustRegionArgumentsOk(size, weights, target_radius) || step_size == nullptr || objective_value == nullptr ||
      (size > 0 && (objective == nullptr || lb == nullptr || ub == nullptr || center == nullptr || solution == nullptr)))
    return PDLP_B200_STATUS_BAD_ARGUMENT;
  return Guard([&] {
    Device dev(cuda_device);
    PlainVec o(dev, objective, size), l(dev, lb, size), u(dev, ub, size), c(dev, center, size), w(dev, weights, size), s(dev, nullptr, size);
    dev.SolveTrustRegion(o.d, l.d, u.d, c.d, w.d, target_radius, size, s.d, step_size, objective_value);
    dev.Download(solution, s.d, size);
  });
}
int32_t pdlp_b200_solve_diagonal_trust_region(int32_t cuda_device, int64_t size, const double* objective, const double* qdiag, const double* lb,
                                              const double* ub, const double* center, const double* weights, double target_radius, double tol,
                                              double* solution, double* step_size, double* objective_value) {
  if (!TrustRegionArgumentsOk(size, weights, target_radius) || step_size == nullptr || objective_value == nullptr ||
      (size > 0 && (objective == nullptr || qdiag == nullptr || lb == nullptr || ub == nullptr || center == nullptr || solution == nullptr)))
    return PDLP_B200_STATUS_BAD_ARGUMENT;
  return Guard([&] {
    Device dev(cuda_device);
    PlainVec o(dev, objective, size), q(dev, qdiag, size), l(dev, lb, size), u(dev, ub, size), c(dev, center, size), w(dev, weights, size),
        s(dev, nullptr, size);
    dev.SolveDiagonalTrustRegion(o.d, q.d, l.d, u.d, c.d, w.d, target_radius, tol, size, s.d, step_size, objective_value);
    dev.Download(solution, s.d, size);
  });
}
int32_t pdlp_b200_weighted_average(int32_t cuda_device, int64_t size, int64_t count, const double* datapoints, const double* weights,
                                   double* out_average, double* out_sum_weights, int32_t* out_num_terms) {
  return Guard([&] {
    Device dev(cuda_device);
    PlainVec avg(dev, nullptr, size), v(dev, nullptr, size);
    dev.Fill(avg.d, 0.0, size);
    double sum = 0.0;
    int32_t terms = 0;
    for (int64_t k = 0; k < count; ++k) {  // ShardedWeightedAverage::Add, sou.cc:54-66
      const double w = weights[k];
      if (w > 0.0) {
        dev.Upload(v.d, datapoints + k * size, size);
        dev.WeightedAverageAdd(avg.d, v.d, w / (sum + w), size);
        sum += w;
      }
      ++terms;
    }
    dev.Download(out_average, avg.d, size);
    if (out_sum_weights) *out_sum_weights = sum;
    if (out_num_terms) *out_num_terms = terms;
  });
}
int32_t pdlp_b200_vector_reduce(int32_t cuda_device, int32_t op, int64_t size, const double* a, const double* b, double* out) {
  int32_t bad = 0;
  const int32_t rc = Guard([&] {
    Device dev(cuda_device);
    PlainVec va(dev, a, size), vb(dev, b, size);
    switch (op) {
      case PDLP_VECOP_DOT: *out = dev.Dot(va.d, vb.d, size); break;
      case PDLP_VECOP_LINF_NORM: *out = dev.LInf(va.d, size); break;
      case PDLP_VECOP_L1_NORM: *out = dev.L1(va.d, size); break;
      case PDLP_VECOP_SQUARED_NORM: *out = dev.SumSq(va.d, size); break;
      case PDLP_VECOP_NORM: *out = std::sqrt(dev.SumSq(va.d, size)); break;
      case PDLP_VECOP_SQUARED_DISTANCE: *out = dev.SumSqDiff(va.d, vb.d, size); break;
      case PDLP_VECOP_DISTANCE: *out = std::sqrt(dev.SumSqDiff(va.d, vb.d, size)); break;
      case PDLP_VECOP_SCALED_LINF_NORM: *out = dev.ScaledLInf(va.d, vb.d, size); break;
      case PDLP_VECOP_SCALED_SQUARED_NORM: *out = dev.ScaledSumSq(va.d, vb.d, size); break;
      case PDLP_VECOP_SCALED_NORM: *out = std::sqrt(dev.ScaledSumSq(va.d, vb.d, size)); break;
      default: bad = 1;
    }
  });
  return bad ? PDLP_B200_STATUS_BAD_ARGUMENT : rc;
}

}  // extern "C"
