/* pdlp_b200_io.h -- the data formats either side of the PDLP path, natively
 * (SURVEY.md section 8f ranks 1 and 3). Host-only entry points of
 * libpdlp_b200.so: none of them needs a CUDA device except
 * pdlp_b200_solve_proto, which runs the solve.
 *
 * What each replaces in the reference:
 *   - PrimalDualHybridGradientParams as text (examples/cpp/pdlp_solve.cc:64-79,
 *     --params) or bytes (MPModelRequest.solver_specific_parameters,
 *     ortools/linear_solver/proto_solver/pdlp_proto_solver.cc:47) -> PdlpParams;
 *   - SolveLog -> bytes / text / JSON (pdlp_proto_solver.cc:127,
 *     pdlp_solve.cc:59-77 WriteSolveLog);
 *   - ReadQuadraticProgramOrDie (ortools/pdlp/quadratic_program_io.h:28-59):
 *     .mps[.gz] with the semantics of ortools/lp_data/mps_reader_template.h,
 *     MPModelProto as .pb / .textproto / .json [.gz];
 *   - QpFromMpModelProto / QpToMpModelProto (quadratic_program.cc:98-315);
 *   - WriteLinearProgramToMps / WriteQuadraticProgramToMPModelProto
 *     (quadratic_program_io.cc:80-101);
 *   - PdlpSolveProto (pdlp_proto_solver.cc:36-130): MPModelRequest bytes in,
 *     MPSolutionResponse bytes out.
 * There is no protoc / libprotobuf behind this: the proto2 wire format, text
 * format and JSON mapping of these messages are implemented in
 * csrc/proto_codec.cc from schema tables that restate the .proto files.
 *
 * Conventions: functions returning int32_t return a PDLP_B200_STATUS_* code
 * and, on PDLP_B200_STATUS_BAD_ARGUMENT, write a NUL-terminated explanation
 * into `error` (truncated to error_capacity; may be NULL). Functions that
 * produce a variable-sized blob return it in a PdlpBlob owned by the library
 * (release with pdlp_b200_blob_free).                                        */
#ifndef PDLP_B200_IO_H_
#define PDLP_B200_IO_H_

#include "pdlp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct PdlpBlob {
  uint8_t* data;   /* size bytes, followed by a NUL that is not counted     */
  int64_t size;
} PdlpBlob;
void pdlp_b200_blob_free(PdlpBlob* blob);

/* ---- PrimalDualHybridGradientParams (solvers.proto:238-497) --------------- *
 * "merge" = protobuf MergeFrom / TextFormat::Merge semantics onto *params as it
 * stands: the reference merges --params onto a message with verbosity_level
 * preset, and a field given twice keeps its last value. "parse" = defaults,
 * then the message; in text form a non-repeated field given twice, or two
 * members of a oneof, are errors (TextFormat::Parse).                        */
int32_t pdlp_b200_params_merge_text(const char* text, PdlpParams* params, char* error,
                                    int64_t error_capacity);
int32_t pdlp_b200_params_parse_text(const char* text, PdlpParams* params, char* error,
                                    int64_t error_capacity);
int32_t pdlp_b200_params_merge_bytes(const uint8_t* data, int64_t size, PdlpParams* params,
                                     char* error, int64_t error_capacity);
int32_t pdlp_b200_params_parse_bytes(const uint8_t* data, int64_t size, PdlpParams* params,
                                     char* error, int64_t error_capacity);
/* format: 0 binary, 1 text, 2 JSON. Fields equal to their proto default are
 * omitted unless the POD carries their presence (has_* / oneof case).        */
enum { PDLP_FORMAT_BINARY = 0, PDLP_FORMAT_TEXT = 1, PDLP_FORMAT_JSON = 2 };
int32_t pdlp_b200_params_serialize(const PdlpParams* params, int32_t format, PdlpBlob* out);

/* ---- SolveLog (solve_log.proto:385-459) from a finished solve -------------- */
int32_t pdlp_b200_solve_log_serialize(const PdlpResult* result, int32_t format, PdlpBlob* out);
/* WriteSolveLog of pdlp_solve.cc:59-77: the suffix picks the format
 * (.textproto, .pb, .json); anything else is PDLP_B200_STATUS_BAD_ARGUMENT.  */
int32_t pdlp_b200_write_solve_log(const PdlpResult* result, const char* path, char* error,
                                  int64_t error_capacity);

/* ---- problems -------------------------------------------------------------- *
 * A PdlpModel owns the arrays of a QuadraticProgram read from a file or a
 * proto; view() is valid until the model is freed.                           */
typedef struct PdlpModel PdlpModel;
/* ReadQuadraticProgramOrDie (returns an error instead of dying). Integrality
 * is dropped (PDLP solves the relaxation); maximisation becomes minimisation
 * with objective_scaling_factor = -1.                                        */
int32_t pdlp_b200_read_quadratic_program(const char* path, int32_t include_names,
                                         PdlpModel** out_model, char* error,
                                         int64_t error_capacity);
/* MPS text already in memory (free or fixed format). */
int32_t pdlp_b200_model_from_mps_text(const char* text, int64_t size, int32_t include_names,
                                      PdlpModel** out_model, char* error,
                                      int64_t error_capacity);
/* QpFromMpModelProto on serialized MPModelProto bytes. */
int32_t pdlp_b200_model_from_mp_model_proto(const uint8_t* data, int64_t size,
                                            int32_t relax_integer_variables,
                                            int32_t include_names, PdlpModel** out_model,
                                            char* error, int64_t error_capacity);
const PdlpProblemView* pdlp_b200_model_view(const PdlpModel* model);
/* Names: NULL if the model was read without names or the index is out of range. */
const char* pdlp_b200_model_variable_name(const PdlpModel* model, int64_t index);
const char* pdlp_b200_model_constraint_name(const PdlpModel* model, int64_t index);
void pdlp_b200_model_free(PdlpModel* model);

/* QpToMpModelProto: serialized MPModelProto of *qp. The name arrays may be
 * NULL (no names), else they have num_variables / num_constraints entries.   */
int32_t pdlp_b200_qp_to_mp_model_proto(const PdlpProblemView* qp,
                                       const char* const* variable_names,
                                       const char* const* constraint_names, PdlpBlob* out,
                                       char* error, int64_t error_capacity);
/* WriteLinearProgramToMps (free-format MPS) / WriteQuadraticProgramToMPModelProto. */
int32_t pdlp_b200_write_linear_program_to_mps(const PdlpProblemView* qp,
                                              const char* const* variable_names,
                                              const char* const* constraint_names,
                                              const char* path, char* error,
                                              int64_t error_capacity);
int32_t pdlp_b200_write_quadratic_program_to_mp_model_proto(const PdlpProblemView* qp,
                                                            const char* const* variable_names,
                                                            const char* const* constraint_names,
                                                            const char* path, char* error,
                                                            int64_t error_capacity);

/* ---- PdlpSolveProto (pdlp_proto_solver.cc:36-130) -------------------------- *
 * request: serialized MPModelRequest; response: serialized MPSolutionResponse
 * (status mapping, objective value, primal / dual values and reduced costs
 * with the sign of a maximisation restored, SolveLog bytes in
 * solver_specific_info). Invalid parameters or models are reported in the
 * response, like the reference; the return code is for device failures only
 * (PDLP_B200_STATUS_NO_DEVICE: this library has no CPU fallback).            */
int32_t pdlp_b200_solve_proto(const uint8_t* request, int64_t request_size,
                              int32_t relax_integer_variables,
                              const volatile int32_t* interrupt_solve, PdlpBlob* response);

/* ---- generic conversions between the three encodings ----------------------- *
 * message: "PrimalDualHybridGradientParams", "TerminationCriteria", "SolveLog",
 * "IterationStats", "MPModelProto", "MPModelRequest", "MPSolutionResponse".  */
int32_t pdlp_b200_proto_convert(const char* message, int32_t from_format, const uint8_t* data,
                                int64_t size, int32_t to_format, PdlpBlob* out, char* error,
                                int64_t error_capacity);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* PDLP_B200_IO_H_ */
