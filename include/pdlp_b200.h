/* pdlp_b200.h -- C ABI of the B200-native PDLP hot path.
 *
 * Drop-in boundary for OR-Tools' PDLP entry point
 *   SolverResult PrimalDualHybridGradient(QuadraticProgram, const
 *       PrimalDualHybridGradientParams&, optional<PrimalAndDualSolution>,
 *       const atomic<bool>* interrupt, message_callback, iteration_stats_callback)
 *   (reference: ortools/pdlp/primal_dual_hybrid_gradient.h:151-169).
 *
 * Everything here is plain C: POD structs, pointers and sizes. No torch, Eigen
 * or protobuf types cross this boundary. The POD structs mirror the reference
 * protos field by field (ortools/pdlp/solvers.proto, ortools/pdlp/solve_log.proto)
 * and every enum keeps the proto's numeric value.
 *
 * The library has NO CPU fallback: every compute entry point returns
 * PDLP_B200_STATUS_NO_DEVICE if no CUDA device is usable.
 */
#ifndef PDLP_B200_H_
#define PDLP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- return codes of the C functions (NOT termination reasons) ---------- */
#define PDLP_B200_STATUS_OK 0
#define PDLP_B200_STATUS_NO_DEVICE 1      /* no usable CUDA device           */
#define PDLP_B200_STATUS_CUDA_ERROR 2     /* a CUDA / NCCL call failed       */
#define PDLP_B200_STATUS_BAD_ARGUMENT 3   /* null pointer / bad handle       */

/* ---- enums: numeric values identical to the protos ---------------------- */
/* solvers.proto:24-41 */
enum { PDLP_OPTIMALITY_NORM_UNSPECIFIED = 0, PDLP_OPTIMALITY_NORM_L_INF = 1,
       PDLP_OPTIMALITY_NORM_L2 = 2, PDLP_OPTIMALITY_NORM_L_INF_COMPONENTWISE = 3 };
/* solvers.proto:44-51 */
enum { PDLP_SCHEDULER_TYPE_UNSPECIFIED = 0, PDLP_SCHEDULER_TYPE_GOOGLE_THREADPOOL = 1,
       PDLP_SCHEDULER_TYPE_EIGEN_THREADPOOL = 3 };
/* solvers.proto:239-277 */
enum { PDLP_RESTART_STRATEGY_UNSPECIFIED = 0, PDLP_NO_RESTARTS = 1,
       PDLP_EVERY_MAJOR_ITERATION = 2, PDLP_ADAPTIVE_HEURISTIC = 3,
       PDLP_ADAPTIVE_DISTANCE_BASED = 4 };
enum { PDLP_LINESEARCH_RULE_UNSPECIFIED = 0, PDLP_ADAPTIVE_LINESEARCH_RULE = 1,
       PDLP_MALITSKY_POCK_LINESEARCH_RULE = 2, PDLP_CONSTANT_STEP_SIZE_RULE = 3 };
/* solve_log.proto:105-117 */
enum { PDLP_RESTART_CHOICE_UNSPECIFIED = 0, PDLP_RESTART_CHOICE_NO_RESTART = 1,
       PDLP_RESTART_CHOICE_WEIGHTED_AVERAGE_RESET = 2,
       PDLP_RESTART_CHOICE_RESTART_TO_AVERAGE = 3 };
/* solve_log.proto:121-135 */
enum { PDLP_POINT_TYPE_UNSPECIFIED = 0, PDLP_POINT_TYPE_CURRENT_ITERATE = 1,
       PDLP_POINT_TYPE_ITERATE_DIFFERENCE = 2, PDLP_POINT_TYPE_AVERAGE_ITERATE = 3,
       PDLP_POINT_TYPE_NONE = 4, PDLP_POINT_TYPE_PRESOLVER_SOLUTION = 5,
       PDLP_POINT_TYPE_FEASIBILITY_POLISHING_SOLUTION = 6 };
/* solve_log.proto:336-360 */
enum { PDLP_TERMINATION_REASON_UNSPECIFIED = 0, PDLP_TERMINATION_REASON_OPTIMAL = 1,
       PDLP_TERMINATION_REASON_PRIMAL_INFEASIBLE = 2,
       PDLP_TERMINATION_REASON_DUAL_INFEASIBLE = 3,
       PDLP_TERMINATION_REASON_TIME_LIMIT = 4,
       PDLP_TERMINATION_REASON_ITERATION_LIMIT = 5,
       PDLP_TERMINATION_REASON_NUMERICAL_ERROR = 6, PDLP_TERMINATION_REASON_OTHER = 7,
       PDLP_TERMINATION_REASON_KKT_MATRIX_PASS_LIMIT = 8,
       PDLP_TERMINATION_REASON_INVALID_PROBLEM = 9,
       PDLP_TERMINATION_REASON_INVALID_PARAMETER = 10,
       PDLP_TERMINATION_REASON_PRIMAL_OR_DUAL_INFEASIBLE = 11,
       PDLP_TERMINATION_REASON_INTERRUPTED_BY_USER = 12,
       PDLP_TERMINATION_REASON_INVALID_INITIAL_SOLUTION = 13 };
/* primal_dual_hybrid_gradient.h:75-88 (IterationType, declaration order) */
enum { PDLP_ITERATION_TYPE_NORMAL = 0, PDLP_ITERATION_TYPE_PRIMAL_FEASIBILITY = 1,
       PDLP_ITERATION_TYPE_DUAL_FEASIBILITY = 2,
       PDLP_ITERATION_TYPE_PRESOLVE_TERMINATION = 3,
       PDLP_ITERATION_TYPE_NORMAL_TERMINATION = 4,
       PDLP_ITERATION_TYPE_FEASIBILITY_POLISHING_TERMINATION = 5 };
/* which member of TerminationCriteria.optimality_criteria (oneof) is set;
 * values are the proto tags (solvers.proto:66-187). */
enum { PDLP_OPTIMALITY_CRITERIA_NOT_SET = 0, PDLP_SIMPLE_OPTIMALITY_CRITERIA = 9,
       PDLP_DETAILED_OPTIMALITY_CRITERIA = 10 };

#define PDLP_MAX_RANDOM_PROJECTION_SEEDS 8

/* ---- TerminationCriteria (solvers.proto:66-187) ------------------------- */
typedef struct PdlpTerminationCriteria {
  int32_t optimality_norm;             /* default L2                          */
  int32_t optimality_criteria_case;    /* PDLP_OPTIMALITY_CRITERIA_*          */
  /* SimpleOptimalityCriteria (tag 9)  defaults 1e-6 / 1e-6                  */
  double simple_eps_optimal_absolute;
  double simple_eps_optimal_relative;
  /* DetailedOptimalityCriteria (tag 10) all default 1e-6                    */
  double eps_optimal_primal_residual_absolute;
  double eps_optimal_primal_residual_relative;
  double eps_optimal_dual_residual_absolute;
  double eps_optimal_dual_residual_relative;
  double eps_optimal_objective_gap_absolute;
  double eps_optimal_objective_gap_relative;
  /* deprecated top-level fields (tags 2, 3) with proto2 presence            */
  int32_t has_eps_optimal_absolute;
  int32_t has_eps_optimal_relative;
  double eps_optimal_absolute;         /* default 1e-6                        */
  double eps_optimal_relative;         /* default 1e-6                        */
  double eps_primal_infeasible;        /* default 1e-8                        */
  double eps_dual_infeasible;          /* default 1e-8                        */
  double time_sec_limit;               /* default +inf                        */
  int32_t iteration_limit;             /* default INT32_MAX                   */
  double kkt_matrix_pass_limit;        /* default +inf                        */
} PdlpTerminationCriteria;

/* ---- PrimalDualHybridGradientParams (solvers.proto:238-497) ------------- */
typedef struct PdlpParams {
  PdlpTerminationCriteria termination_criteria;
  int32_t num_threads;                 /* 1; accepted, unused on the device   */
  int32_t num_shards;                  /* 0; accepted, unused on the device   */
  int32_t scheduler_type;              /* GOOGLE_THREADPOOL; unused on device */
  int32_t record_iteration_stats;      /* false                               */
  int32_t verbosity_level;             /* 0                                   */
  double log_interval_seconds;         /* 0.0                                 */
  int32_t major_iteration_frequency;   /* 64                                  */
  int32_t termination_check_frequency; /* 64                                  */
  int32_t restart_strategy;            /* ADAPTIVE_HEURISTIC                  */
  double primal_weight_update_smoothing;    /* 0.5                            */
  int32_t has_initial_primal_weight;   /* false                               */
  double initial_primal_weight;
  int32_t l_inf_ruiz_iterations;       /* 5                                   */
  int32_t l2_norm_rescaling;           /* true                                */
  double sufficient_reduction_for_restart;  /* 0.1                            */
  double necessary_reduction_for_restart;   /* 0.9                            */
  int32_t linesearch_rule;             /* ADAPTIVE_LINESEARCH_RULE            */
  /* AdaptiveLinesearchParams (solvers.proto:189-204)                        */
  double adaptive_step_size_reduction_exponent;  /* 0.3                       */
  double adaptive_step_size_growth_exponent;     /* 0.6                       */
  /* MalitskyPockParams (solvers.proto:206-226)                              */
  double malitsky_pock_step_size_downscaling_factor;   /* 0.7                 */
  double malitsky_pock_linesearch_contraction_factor;  /* 0.99                */
  double malitsky_pock_step_size_interpolation;        /* 1.0                 */
  double initial_step_size_scaling;    /* 1.0                                 */
  double infinite_constraint_bound_threshold;    /* +inf                      */
  int32_t handle_some_primal_gradients_on_finite_bounds_as_residuals; /* true */
  int32_t use_diagonal_qp_trust_region_solver;   /* false                     */
  double diagonal_qp_trust_region_solver_tolerance; /* 1e-8                   */
  int32_t num_random_projection_seeds; /* <= PDLP_MAX_RANDOM_PROJECTION_SEEDS */
  int32_t random_projection_seeds[PDLP_MAX_RANDOM_PROJECTION_SEEDS];
  /* presolve_options.use_glop (tag 16) is host-side glop presolve, which this
   * library does not run: setting it yields TERMINATION_REASON_INVALID_PARAMETER
   * with an explanatory termination_string (never a silent fallback).
   * Feasibility polishing (tags 30, 33, 34; LPs only) runs on the device.     */
  int32_t presolve_use_glop;           /* false                               */
  int32_t use_feasibility_polishing;   /* false                               */
  int32_t apply_feasibility_polishing_after_limits_reached;      /* false    */
  int32_t apply_feasibility_polishing_if_solver_is_interrupted;  /* false    */
} PdlpParams;

/* ---- QuadraticProgram view (quadratic_program.h:134-150) ----------------- *
 * K is compressed sparse COLUMN with int64 indices, exactly the arrays of
 * Eigen::SparseMatrix<double, ColMajor, int64_t> (outerIndexPtr,
 * innerIndexPtr, valuePtr). All pointers are HOST pointers; the library never
 * writes through them (the reference takes the QP by value).                */
typedef struct PdlpProblemView {
  int64_t num_variables;               /* n = cols of K                       */
  int64_t num_constraints;             /* m = rows of K                       */
  int64_t num_nonzeros;
  const int64_t* col_starts;           /* [n+1]                               */
  const int64_t* row_indices;          /* [nnz], sorted within a column       */
  const double* values;                /* [nnz]                               */
  const double* objective_vector;      /* [n]                                 */
  const double* objective_matrix_diagonal; /* [n] or NULL for an LP           */
  const double* constraint_lower_bounds;   /* [m]                             */
  const double* constraint_upper_bounds;   /* [m]                             */
  const double* variable_lower_bounds;     /* [n]                             */
  const double* variable_upper_bounds;     /* [n]                             */
  double objective_offset;
  double objective_scaling_factor;
  const char* problem_name;            /* NUL-terminated or NULL              */
  /* Lengths as the caller's vectors actually have them, so that
   * ValidateQuadraticProgramDimensions (quadratic_program.cc:38-97) can be
   * mirrored. A negative value means "consistent with n / m".              */
  int64_t objective_vector_size, objective_matrix_size;
  int64_t constraint_lower_bounds_size, constraint_upper_bounds_size;
  int64_t variable_lower_bounds_size, variable_upper_bounds_size;
} PdlpProblemView;

/* ---- solve_log.proto messages ------------------------------------------- */
typedef struct PdlpQuadraticProgramStats {   /* solve_log.proto:28-102        */
  int64_t num_variables, num_constraints;
  double constraint_matrix_col_min_l_inf_norm, constraint_matrix_row_min_l_inf_norm;
  int64_t constraint_matrix_num_nonzeros;
  double constraint_matrix_abs_max, constraint_matrix_abs_min, constraint_matrix_abs_avg,
      constraint_matrix_l2_norm;
  double combined_bounds_max, combined_bounds_min, combined_bounds_avg, combined_bounds_l2_norm;
  double combined_variable_bounds_max, combined_variable_bounds_min,
      combined_variable_bounds_avg, combined_variable_bounds_l2_norm;
  int64_t variable_bound_gaps_num_finite;
  double variable_bound_gaps_max, variable_bound_gaps_min, variable_bound_gaps_avg,
      variable_bound_gaps_l2_norm;
  double objective_vector_abs_max, objective_vector_abs_min, objective_vector_abs_avg,
      objective_vector_l2_norm;
  int64_t objective_matrix_num_nonzeros;
  double objective_matrix_abs_max, objective_matrix_abs_min, objective_matrix_abs_avg,
      objective_matrix_l2_norm;
} PdlpQuadraticProgramStats;

typedef struct PdlpConvergenceInformation {  /* solve_log.proto:139-205       */
  int32_t candidate_type;
  double primal_objective, dual_objective, corrected_dual_objective;
  double l_inf_primal_residual, l2_primal_residual, l_inf_componentwise_primal_residual;
  double l_inf_dual_residual, l2_dual_residual, l_inf_componentwise_dual_residual;
  double l_inf_primal_variable, l2_primal_variable, l_inf_dual_variable, l2_dual_variable;
} PdlpConvergenceInformation;

typedef struct PdlpInfeasibilityInformation { /* solve_log.proto:209-249      */
  int32_t candidate_type;
  double max_primal_ray_infeasibility, primal_ray_linear_objective, primal_ray_quadratic_norm;
  double max_dual_ray_infeasibility, dual_ray_objective;
} PdlpInfeasibilityInformation;

typedef struct PdlpPointMetadata {            /* solve_log.proto:251-274      */
  int32_t point_type;
  int32_t num_random_projections;
  double random_primal_projections[PDLP_MAX_RANDOM_PROJECTION_SEEDS];
  double random_dual_projections[PDLP_MAX_RANDOM_PROJECTION_SEEDS];
  int32_t has_active_set_information;   /* false for ITERATE_DIFFERENCE      */
  int64_t active_primal_variable_count, active_dual_variable_count;
  int64_t active_primal_variable_change, active_dual_variable_change;
} PdlpPointMetadata;

typedef struct PdlpIterationStats {           /* solve_log.proto:281-334      */
  int32_t iteration_number;
  int32_t num_convergence_information;        /* <= 3                         */
  PdlpConvergenceInformation convergence_information[3];
  int32_t num_infeasibility_information;      /* <= 3                         */
  PdlpInfeasibilityInformation infeasibility_information[3];
  int32_t num_point_metadata;                 /* <= 3                         */
  PdlpPointMetadata point_metadata[3];
  double cumulative_kkt_matrix_passes;
  int32_t cumulative_rejected_steps;
  double cumulative_time_sec;
  int32_t restart_used;
  double step_size;
  double primal_weight;
} PdlpIterationStats;

/* QuadraticProgramBoundNorms (termination.h:30-37) */
typedef struct PdlpBoundNorms {
  double l2_norm_primal_linear_objective, l2_norm_constraint_bounds;
  double l_inf_norm_primal_linear_objective, l_inf_norm_constraint_bounds;
} PdlpBoundNorms;

/* IterationCallbackInfo (primal_dual_hybrid_gradient.h:90-98) */
typedef struct PdlpIterationCallbackInfo {
  int32_t iteration_type;
  const PdlpTerminationCriteria* termination_criteria;
  const PdlpIterationStats* iteration_stats;
  PdlpBoundNorms bound_norms;
} PdlpIterationCallbackInfo;

/* SolverResult + SolveLog (primal_dual_hybrid_gradient.h:60-71,
 * solve_log.proto:385-459). Vectors are for the ORIGINAL (unscaled) problem.
 * Buffers are owned by the library: release with pdlp_b200_result_free().   */
/* FeasibilityPolishingDetails (solve_log.proto:362-383): one entry per primal /
 * dual feasibility polishing phase, in the order they ran.                   */
enum { PDLP_POLISHING_PHASE_TYPE_UNSPECIFIED = 0, PDLP_POLISHING_PHASE_TYPE_PRIMAL_FEASIBILITY = 1,
       PDLP_POLISHING_PHASE_TYPE_DUAL_FEASIBILITY = 2 };
typedef struct PdlpFeasibilityPolishingDetails {
  int32_t polishing_phase_type;
  int32_t main_iteration_count;               /* main iterations done when the phase started */
  PdlpParams params;                          /* parameters of the phase      */
  int32_t termination_reason;
  int32_t iteration_count;
  double solve_time_sec;
  PdlpIterationStats solution_stats;
  int32_t solution_type;
  int64_t num_iteration_stats;                /* record_iteration_stats       */
  PdlpIterationStats* iteration_stats;
} PdlpFeasibilityPolishingDetails;

typedef struct PdlpResult {
  int64_t primal_size, dual_size;             /* 0 for INVALID_* results      */
  double* primal_solution;                    /* [primal_size]                */
  double* dual_solution;                      /* [dual_size]                  */
  double* reduced_costs;                      /* [primal_size]                */
  /* SolveLog */
  char* instance_name;                        /* NULL if unset                */
  int32_t termination_reason;
  char* termination_string;                   /* NULL if unset                */
  int32_t iteration_count;
  double solve_time_sec;
  double preprocessing_time_sec;
  int32_t solution_type;
  int32_t has_solution_stats;
  PdlpIterationStats solution_stats;
  int32_t has_original_problem_stats, has_preprocessed_problem_stats;
  PdlpQuadraticProgramStats original_problem_stats, preprocessed_problem_stats;
  int64_t num_iteration_stats;                /* record_iteration_stats       */
  PdlpIterationStats* iteration_stats;
  PdlpParams params;                          /* SolveLog.params              */
  int64_t num_feasibility_polishing_details;
  PdlpFeasibilityPolishingDetails* feasibility_polishing_details;
  /* Device-side accounting for this solve (not in the reference log).       */
  int64_t gpu_kernel_launches;
  double device_iteration_time_sec;           /* CUDA-event time in PDHG steps*/
} PdlpResult;

typedef void (*PdlpMessageCallback)(const char* message, void* user_data);
typedef void (*PdlpIterationStatsCallback)(const PdlpIterationCallbackInfo* info,
                                           void* user_data);

/* ---- parameters ---------------------------------------------------------- */
/* Fills *params with the proto defaults (solvers.proto:66-497). */
void pdlp_b200_params_set_defaults(PdlpParams* params);
/* ValidatePrimalDualHybridGradientParams (solvers_proto_validation.cc:171-298).
 * Returns 1 if valid; otherwise 0 and writes the reference's error text into
 * message (truncated to message_capacity). */
int32_t pdlp_b200_params_validate(const PdlpParams* params, char* message,
                                  int64_t message_capacity);

/* ---- the solve (primal_dual_hybrid_gradient.cc:3107-3152) ---------------- *
 * initial_primal/initial_dual: both NULL (start at zero) or both non-NULL with
 * the given sizes (std::optional<PrimalAndDualSolution>). interrupt_solve may
 * be NULL; it is polled like the reference's std::atomic<bool>. Callbacks run
 * synchronously on the calling thread. Returns a PDLP_B200_STATUS_* code;
 * solver-level outcomes (including invalid input) are reported in
 * result->termination_reason exactly like the reference (never thrown).     */
int32_t pdlp_b200_primal_dual_hybrid_gradient(
    const PdlpProblemView* qp, const PdlpParams* params,
    const double* initial_primal, int64_t initial_primal_size,
    const double* initial_dual, int64_t initial_dual_size,
    const volatile int32_t* interrupt_solve, PdlpMessageCallback message_callback,
    PdlpIterationStatsCallback iteration_stats_callback, void* user_data,
    PdlpResult* result);
void pdlp_b200_result_free(PdlpResult* result);

/* ---- resident solve sessions ---------------------------------------------- *
 * The same solve, split so that the problem and the iterates stay resident in
 * HBM between calls: create = PreprocessSolver::PreprocessAndSolve up to the
 * first iteration (pdhg.cc:1039-1221: upload, validation, stats, Ruiz + L2
 * rescaling, step-size / primal-weight initialisation); advance = the loop of
 * Solver::Solve (pdhg.cc:3042-3091) until `target_iterations` iterations are
 * completed or a termination criterion fires; finish = result construction
 * (pdhg.cc:1728-1818). Stopping and resuming does not change the iterates.
 * Used by callers that re-solve / warm-start and by bench.py to time PDHG
 * iterations with all inputs already on the device.                         */
typedef struct PdlpSolveSession PdlpSolveSession;
typedef struct PdlpSessionStatus {
  int32_t terminated;                  /* 1 once a termination criterion fired */
  int32_t termination_reason;
  int32_t iterations_completed;
  int32_t num_rejected_steps;
  double step_size, primal_weight;
  int64_t gpu_kernel_launches;         /* cumulative, this session             */
  double device_step_ms;               /* CUDA-event time inside the PDHG step loop (cumulative) */
  double device_total_ms;              /* CUDA-event time of all advance calls (steps + restart /
                                          termination work, host gaps included) */
  /* Sampled per-kernel CUDA-event times (pdlp_b200_session_enable_timing):
   * 0 primal step, 1 K x~ + dual epilogue, 2 K^T y' + epilogue, 3 step decision */
  double kernel_ms[4];
  int64_t kernel_samples[4];
  double kernel_algorithmic_bytes[4];  /* per launch, DESIGN.md accounting      */
} PdlpSessionStatus;
int32_t pdlp_b200_session_create(const PdlpProblemView* qp, const PdlpParams* params,
                                 const double* initial_primal, int64_t initial_primal_size,
                                 const double* initial_dual, int64_t initial_dual_size,
                                 PdlpMessageCallback message_callback,
                                 PdlpIterationStatsCallback iteration_stats_callback,
                                 void* user_data, int32_t cuda_device,
                                 PdlpSolveSession** out_session);
int32_t pdlp_b200_session_advance(PdlpSolveSession* session, int32_t target_iterations,
                                  const volatile int32_t* interrupt_solve,
                                  PdlpSessionStatus* out_status);
int32_t pdlp_b200_session_enable_timing(PdlpSolveSession* session, int32_t enable,
                                        int32_t sample_stride);
int32_t pdlp_b200_session_status(PdlpSolveSession* session, PdlpSessionStatus* out_status);
int32_t pdlp_b200_session_finish(PdlpSolveSession* session, PdlpResult* result);
void pdlp_b200_session_destroy(PdlpSolveSession* session);

/* CUDA device used by pdlp_b200_primal_dual_hybrid_gradient (default 0; one
 * process per GPU sets it to its local rank).                                */
int32_t pdlp_b200_set_default_device(int32_t cuda_device);

/* ---- multi-GPU ------------------------------------ *
 * One process per GPU. Every rank passes the SAME full problem; the library
 * keeps only its contiguous block of constraint rows on its device
 * (nnz-balanced, same mass rule as sharder.cc:51-70) plus, when the dual side
 * is the shorter one, its slice of the columns over all rows. Inside the step
 * loop the ranks exchange x~ and y' (or the K^T y' partials) through
 * peer-mapped device memory (CUDA IPC over NVLink / NVSwitch, mapped once per
 * context); NCCL carries the set-up, the small reductions outside the loop and
 * the whole exchange when peer mapping is unavailable
 * (PDLP_B200_EXCHANGE=peer-d|peer-s|nccl forces one). nccl_unique_id is the
 * 128-byte ncclUniqueId made by rank 0 (pdlp_b200_nccl_unique_id) and
 * broadcast by the caller (e.g. torch.distributed). All ranks of a context
 * must be processes on the same box.                                        */
int32_t pdlp_b200_nccl_unique_id(const char* nccl_library_path, uint8_t out_id[128]);
typedef struct PdlpDistributedContext PdlpDistributedContext;
int32_t pdlp_b200_distributed_init(const char* nccl_library_path, int32_t rank,
                                   int32_t world_size, int32_t cuda_device,
                                   const uint8_t nccl_unique_id[128],
                                   PdlpDistributedContext** out_context);
void pdlp_b200_distributed_destroy(PdlpDistributedContext* context);
/* The contiguous block [row_begin, row_end) of constraint rows that rank
 * `rank` of `world_size` keeps (host-only; no device needed): boundaries are
 * where the running (nnz + rows) count reaches rank / world_size of the total,
 * the equal-mass rule of Sharder (sharder.cc:51-70) applied to the rows of K. */
int32_t pdlp_b200_row_block(const PdlpProblemView* qp, int32_t rank, int32_t world_size,
                            int64_t* row_begin, int64_t* row_end);
/* Layout, in doubles, of the peer arena every rank of a row-sharded solve maps
 * (host-only; diagnostics / tests): out = {slice stride, padded primal length,
 * x~, K^T y' partial, y', scalar triples, barrier flags, epochs, trust-region
 * round vectors, trust-region candidates, second trust-region round vectors,
 * second trust-region candidates, total}. The step-loop exchange replaces the
 * Sharder's in-process hand-over of x and K^T y (sharder.cc:160-173).         */
int32_t pdlp_b200_peer_arena_layout(int64_t num_variables, int64_t num_constraints, int32_t world_size,
                                    int64_t out[13]);
int32_t pdlp_b200_primal_dual_hybrid_gradient_distributed(
    PdlpDistributedContext* context, const PdlpProblemView* qp, const PdlpParams* params,
    const double* initial_primal, int64_t initial_primal_size,
    const double* initial_dual, int64_t initial_dual_size,
    const volatile int32_t* interrupt_solve, PdlpMessageCallback message_callback,
    PdlpIterationStatsCallback iteration_stats_callback, void* user_data,
    PdlpResult* result);

/* Row-sharded resident session (see "resident solve sessions" above); advance /
 * status / finish / destroy are the pdlp_b200_session_* calls, made by every
 * rank with the same arguments.                                             */
int32_t pdlp_b200_session_create_distributed(PdlpDistributedContext* context,
                                             const PdlpProblemView* qp, const PdlpParams* params,
                                             PdlpSolveSession** out_session);

/* ---- device-resident problem: kernel-level entry points ------------------ *
 * These expose the individual hot-path operators of the reference so that each
 * can be checked against the CPU implementation (north_star: per-kernel parity
 * 1e-12). All vector arguments are HOST pointers unless the name says device;
 * a handle owns the device copy of the QP (both sparse orientations).        */
typedef struct PdlpDeviceProblem PdlpDeviceProblem;

/* ShardedQuadraticProgram ctor (sharded_quadratic_program.cc:79-107): uploads
 * the QP and builds the row-major and column-major device copies of K.      */
int32_t pdlp_b200_problem_create(const PdlpProblemView* qp, int32_t cuda_device,
                                 PdlpDeviceProblem** out_problem);
void pdlp_b200_problem_destroy(PdlpDeviceProblem* problem);
/* TransposedMatrixVectorProduct (sharder.cc:160-173): out[n] = K^T y.       */
int32_t pdlp_b200_transposed_matrix_vector_product(PdlpDeviceProblem* problem,
                                                   const double* y, double* out);
/* Same with the stored transpose: out[m] = K x (pdhg.cc:1912-1916).         */
int32_t pdlp_b200_matrix_vector_product(PdlpDeviceProblem* problem, const double* x,
                                        double* out);
/* ApplyRescaling (sharded_optimization_utils.cc:423-444): Ruiz + L2, rescales
 * the device QP in place, returns the scaling vectors.                      */
int32_t pdlp_b200_apply_rescaling(PdlpDeviceProblem* problem, int32_t l_inf_ruiz_iterations,
                                  int32_t l2_norm_rescaling, double* row_scaling_vec,
                                  double* col_scaling_vec);
/* One LInf / L2 scaling iteration from given vectors (sou.cc:409-421); does
 * not touch the QP. norm: 0 = LInf (Ruiz), 1 = L2.                          */
int32_t pdlp_b200_scaling_iterations(PdlpDeviceProblem* problem, int32_t norm,
                                     int32_t num_iterations, double* row_scaling_vec,
                                     double* col_scaling_vec);
/* ScaledColLInfNorm / ScaledColL2Norm (sharder.cc:288-332) of K (per column,
 * out[n]) and of K^T (per row of K, out[m]). norm: 0 = LInf, 1 = L2.        */
int32_t pdlp_b200_scaled_col_norm(PdlpDeviceProblem* problem, int32_t norm,
                                  const double* row_scaling_vec,
                                  const double* col_scaling_vec, double* out_cols);
int32_t pdlp_b200_scaled_row_norm(PdlpDeviceProblem* problem, int32_t norm,
                                  const double* row_scaling_vec,
                                  const double* col_scaling_vec, double* out_rows);
/* RescaleQuadraticProgram (sharded_quadratic_program.cc:148-181).           */
int32_t pdlp_b200_rescale_quadratic_program(PdlpDeviceProblem* problem,
                                            const double* col_scaling_vec,
                                            const double* row_scaling_vec);
/* Downloads the (possibly rescaled) QP vectors and matrix values in the CSC
 * order of the PdlpProblemView used at creation. Any pointer may be NULL.   */
int32_t pdlp_b200_problem_download(PdlpDeviceProblem* problem, double* values,
                                   double* objective_vector, double* objective_matrix_diagonal,
                                   double* constraint_lower_bounds,
                                   double* constraint_upper_bounds,
                                   double* variable_lower_bounds,
                                   double* variable_upper_bounds);
/* ComputeStats (sharded_optimization_utils.cc:270-343).                     */
int32_t pdlp_b200_compute_stats(PdlpDeviceProblem* problem, PdlpQuadraticProgramStats* out);
/* ProjectToPrimalVariableBounds / ProjectToDualVariableBounds
 * (sharded_optimization_utils.cc:725-770), in place on host vectors.        */
int32_t pdlp_b200_project_to_primal_variable_bounds(PdlpDeviceProblem* problem, double* primal,
                                                    int32_t use_feasibility_bounds);
int32_t pdlp_b200_project_to_dual_variable_bounds(PdlpDeviceProblem* problem, double* dual);
/* ComputePrimalGradient / ComputeDualGradient (sou.cc:446-527).             */
int32_t pdlp_b200_compute_primal_gradient(PdlpDeviceProblem* problem, const double* primal,
                                          const double* dual_product, double* gradient,
                                          double* value);
int32_t pdlp_b200_compute_dual_gradient(PdlpDeviceProblem* problem, const double* dual,
                                        const double* primal_product, double* gradient,
                                        double* value);
/* ComputeConvergenceInformation / ComputeInfeasibilityInformation /
 * ReducedCosts (iteration_stats.cc:383-453, 486-564, 579-593). The scaling
 * vectors may be NULL (= ones).                                             */
int32_t pdlp_b200_compute_convergence_information(
    PdlpDeviceProblem* problem, const PdlpParams* params, const double* col_scaling_vec,
    const double* row_scaling_vec, const double* scaled_primal, const double* scaled_dual,
    double componentwise_primal_residual_offset, double componentwise_dual_residual_offset,
    int32_t candidate_type, PdlpConvergenceInformation* out);
int32_t pdlp_b200_compute_infeasibility_information(
    PdlpDeviceProblem* problem, const PdlpParams* params, const double* col_scaling_vec,
    const double* row_scaling_vec, const double* scaled_primal_ray,
    const double* scaled_dual_ray, const double* primal_solution_for_residual_tests,
    int32_t candidate_type, PdlpInfeasibilityInformation* out);
int32_t pdlp_b200_reduced_costs(PdlpDeviceProblem* problem, const PdlpParams* params,
                                const double* primal, const double* dual,
                                int32_t use_zero_primal_objective, double* out);
/* ComputeLocalizedLagrangianBounds, Euclidean norm (trust_region.cc:886-1016).
 * primal_product / dual_product may be NULL (computed). out[4] =
 * {lagrangian_value, lower_bound, upper_bound, radius}.                     */
int32_t pdlp_b200_compute_localized_lagrangian_bounds(
    PdlpDeviceProblem* problem, const double* primal, const double* dual, double primal_weight,
    double radius, const double* primal_product, const double* dual_product,
    int32_t use_diagonal_qp_trust_region_solver,
    double diagonal_qp_trust_region_solver_tolerance, double out[4]);
/* The same with PrimalDualNorm::kMaxNorm (trust_region.cc:855-884): primal and
 * dual trust-region problems solved separately. Not used by the solver.      */
int32_t pdlp_b200_compute_localized_lagrangian_bounds_max_norm(
    PdlpDeviceProblem* problem, const double* primal, const double* dual, double primal_weight,
    double radius, const double* primal_product, const double* dual_product, double out[4]);
/* ---- termination.h (host-only scalar logic; no device needed) ------------- *
 * CheckSimpleTerminationCriteria / CheckIterateTerminationCriteria
 * (termination.cc:161-219): return 1 and fill *reason / *point_type when a
 * criterion fires, else 0. OptimalityCriteriaMet / ObjectiveGapMet (:26-97),
 * EffectiveOptimalityCriteria (:126-159; out = {primal abs, rel, dual abs, rel,
 * gap abs, rel}), ComputeRelativeResiduals (:239-271; out = {l_inf primal, l2
 * primal, l_inf dual, l2 dual, optimality gap}), BoundNormsFromProblemStats
 * (:221-228).                                                                */
int32_t pdlp_b200_check_simple_termination_criteria(const PdlpTerminationCriteria* criteria,
                                                    const PdlpIterationStats* stats,
                                                    const volatile int32_t* interrupt_solve,
                                                    int32_t* reason, int32_t* point_type);
int32_t pdlp_b200_check_iterate_termination_criteria(const PdlpTerminationCriteria* criteria,
                                                     const PdlpIterationStats* stats,
                                                     const PdlpBoundNorms* bound_norms,
                                                     int32_t force_numerical_termination,
                                                     int32_t* reason, int32_t* point_type);
int32_t pdlp_b200_optimality_criteria_met(const PdlpTerminationCriteria* criteria,
                                          const PdlpConvergenceInformation* stats,
                                          const PdlpBoundNorms* bound_norms,
                                          int32_t* objective_gap_met);
void pdlp_b200_effective_optimality_criteria(const PdlpTerminationCriteria* criteria, double out[6]);
void pdlp_b200_compute_relative_residuals(const PdlpTerminationCriteria* criteria,
                                          const PdlpConvergenceInformation* stats,
                                          const PdlpBoundNorms* bound_norms, double out[5]);
void pdlp_b200_bound_norms_from_problem_stats(const PdlpQuadraticProgramStats* stats,
                                              PdlpBoundNorms* out);
/* SolveTrustRegion (trust_region.h:58-64): explicit-vector problem of size
 * `size`. Writes solution[size], *step_size, *objective_value.               */
int32_t pdlp_b200_solve_trust_region(int32_t cuda_device, int64_t size,
                                     const double* objective_vector,
                                     const double* variable_lower_bounds,
                                     const double* variable_upper_bounds,
                                     const double* center_point, const double* norm_weights,
                                     double target_radius, double* solution,
                                     double* step_size, double* objective_value);
/* SolveDiagonalTrustRegion (trust_region.h:83-89).                          */
int32_t pdlp_b200_solve_diagonal_trust_region(
    int32_t cuda_device, int64_t size, const double* objective_vector,
    const double* objective_matrix_diagonal, const double* variable_lower_bounds,
    const double* variable_upper_bounds, const double* center_point,
    const double* norm_weights, double target_radius, double solve_tolerance,
    double* solution, double* step_size, double* objective_value);
/* ShardedWeightedAverage (sou.cc:43-79) over `count` datapoints of length
 * `size` stored row-major in datapoints[count*size]; out_average[size].     */
int32_t pdlp_b200_weighted_average(int32_t cuda_device, int64_t size, int64_t count,
                                   const double* datapoints, const double* weights,
                                   double* out_average, double* out_sum_weights,
                                   int32_t* out_num_terms);
/* Sharder vector ops (sharder.cc:175-286) on device, host in/out. op codes: */
enum { PDLP_VECOP_DOT = 0, PDLP_VECOP_LINF_NORM = 1, PDLP_VECOP_L1_NORM = 2,
       PDLP_VECOP_SQUARED_NORM = 3, PDLP_VECOP_NORM = 4, PDLP_VECOP_SQUARED_DISTANCE = 5,
       PDLP_VECOP_DISTANCE = 6, PDLP_VECOP_SCALED_LINF_NORM = 7,
       PDLP_VECOP_SCALED_SQUARED_NORM = 8, PDLP_VECOP_SCALED_NORM = 9 };
int32_t pdlp_b200_vector_reduce(int32_t cuda_device, int32_t op, int64_t size, const double* a,
                                const double* b, double* out);

/* ---- layout inspection (host-only; no device needed) ---------------------- *
 * The SELL-32 images of K (rows) and K^T (cols) exactly as the host builder
 * lays them out for the device (csrc/sell_builder.cc) --
 * the structure, not a product: the arrays are for checking that the image
 * stores the caller's matrix entry for entry. Slots [0, num_virtual_padded) are
 * the virtual slots of split rows (virt_pos = the row's position, -1 padding);
 * slot num_virtual_padded + (p - num_split) is the row at position p >=
 * num_split. Element j of slot s lives at slice_ptr[s / 32] + 32 j + s % 32;
 * `col` holds positions of the OTHER image's order (row_of_pos maps a position
 * back to the caller's index). Arrays are owned by the library.              */
typedef struct PdlpSellLayout {
  int64_t num_rows, num_cols, num_split, num_virtual, num_virtual_padded, num_slots, padded_nnz;
  int32_t split_len;
  int64_t* slice_ptr;    /* [num_slots / 32 + 1]  */
  int32_t* slot_len;     /* [num_slots]           */
  int32_t* col;          /* [padded_nnz]          */
  double* val;           /* [padded_nnz]          */
  int32_t* split_first;  /* [num_split + 1]       */
  int32_t* virt_pos;     /* [num_virtual_padded]  */
  int32_t* row_of_pos;   /* [num_rows]            */
  int32_t* pos_of_row;   /* [num_rows]            */
} PdlpSellLayout;
/* [row_begin, row_end): the block of constraint rows to image (0, m for all;
 * a row-sharded rank images only its block). sigma: sort window (4096).
 * natural_primal_order (row-sharded solves): the row image stores the caller's
 * column indices themselves, because primal vectors stay in the caller's order
 * on every rank; the column image keeps its own position order.              */
int32_t pdlp_b200_host_sell_layout(const PdlpProblemView* qp, int64_t row_begin, int64_t row_end,
                                   int32_t sigma, int32_t natural_primal_order,
                                   PdlpSellLayout* out_rows, PdlpSellLayout* out_cols);
void pdlp_b200_sell_layout_free(PdlpSellLayout* layout);

/* ---- misc ---------------------------------------------------------------- */
/* Number of usable CUDA devices (0 if none); never fails. */
int32_t pdlp_b200_device_count(void);
const char* pdlp_b200_version(void);
/* sizeof() of the boundary structs, for binding self-checks. index: 0
 * TerminationCriteria, 1 Params, 2 ProblemView, 3 QuadraticProgramStats,
 * 4 ConvergenceInformation, 5 InfeasibilityInformation, 6 PointMetadata,
 * 7 IterationStats, 8 BoundNorms, 9 IterationCallbackInfo, 10 Result,
 * 11 SessionStatus; -1 for an unknown index. */
int64_t pdlp_b200_sizeof(int32_t index);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* PDLP_B200_H_ */
