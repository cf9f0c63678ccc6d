// fuzz_formats.cc -- mutation fuzzer for the host-only format layer
// (csrc/proto_codec.cc, csrc/formats.cc). Built with AddressSanitizer and
// UBSan by tools/fuzz_formats.sh; the solve entry points are stubbed, so no
// CUDA is needed. Every input, however malformed, must come back as a status
// code: no crash, no out-of-bounds read, no leak.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../include/pdlp_b200_io.h"

extern "C" {
// stubs for the device side of the library (never reached with a stub device)
int32_t pdlp_b200_primal_dual_hybrid_gradient(const PdlpProblemView*, const PdlpParams*, const double*, int64_t, const double*, int64_t,
                                              const volatile int32_t*, PdlpMessageCallback, PdlpIterationStatsCallback, void*, PdlpResult*) {
  return PDLP_B200_STATUS_NO_DEVICE;
}
void pdlp_b200_result_free(PdlpResult*) {}
}
namespace pdlp_b200 { void SetDefaultParams(PdlpParams* p); }  // params.cc
extern "C" void pdlp_b200_params_set_defaults(PdlpParams* p) { pdlp_b200::SetDefaultParams(p); }  // lives in capi.cc in the library

namespace {
std::mt19937_64 rng(12345);
int Rand(int n) { return static_cast<int>(rng() % static_cast<uint64_t>(n)); }

std::string Mutate(std::string s) {
  const int edits = 1 + Rand(4);
  for (int e = 0; e < edits; ++e) {
    const int kind = Rand(5);
    if (s.empty() || kind == 0) {
      s.insert(s.begin() + (s.empty() ? 0 : Rand(static_cast<int>(s.size()) + 1)), static_cast<char>(Rand(256)));
    } else if (kind == 1) {
      s.erase(s.begin() + Rand(static_cast<int>(s.size())));
    } else if (kind == 2) {
      s[Rand(static_cast<int>(s.size()))] = static_cast<char>(Rand(256));
    } else if (kind == 3) {
      s.resize(Rand(static_cast<int>(s.size()) + 1));
    } else {
      const int a = Rand(static_cast<int>(s.size())), n = Rand(static_cast<int>(s.size()) - a + 1);
      s.insert(Rand(static_cast<int>(s.size()) + 1), s.substr(a, n));
    }
  }
  return s;
}

const char* kMessages[] = {"PrimalDualHybridGradientParams", "TerminationCriteria", "SolveLog", "IterationStats", "MPModelProto", "MPModelRequest",
                           "MPSolutionResponse"};

void Exercise(const std::string& in) {
  char err[256];
  const uint8_t* data = reinterpret_cast<const uint8_t*>(in.data());
  const int64_t size = static_cast<int64_t>(in.size());
  PdlpParams p;
  pdlp_b200_params_parse_bytes(data, size, &p, err, sizeof err);
  pdlp_b200_params_parse_text(in.c_str(), &p, err, sizeof err);
  pdlp_b200_params_set_defaults(&p);
  if (pdlp_b200_params_merge_text(in.c_str(), &p, err, sizeof err) == PDLP_B200_STATUS_OK) {
    for (int fmt = 0; fmt < 3; ++fmt) {
      PdlpBlob b{};
      if (pdlp_b200_params_serialize(&p, fmt, &b) == PDLP_B200_STATUS_OK) pdlp_b200_blob_free(&b);
    }
  }
  for (const char* m : kMessages)
    for (int from = 0; from < 3; ++from)
      for (int to = 0; to < 3; ++to) {
        PdlpBlob b{};
        if (pdlp_b200_proto_convert(m, from, data, size, to, &b, err, sizeof err) == PDLP_B200_STATUS_OK) pdlp_b200_blob_free(&b);
      }
  PdlpModel* model = nullptr;
  if (pdlp_b200_model_from_mp_model_proto(data, size, Rand(2), Rand(2), &model, err, sizeof err) == PDLP_B200_STATUS_OK) {
    PdlpBlob b{};
    if (pdlp_b200_qp_to_mp_model_proto(pdlp_b200_model_view(model), nullptr, nullptr, &b, err, sizeof err) == PDLP_B200_STATUS_OK) pdlp_b200_blob_free(&b);
    pdlp_b200_model_variable_name(model, 0);
    pdlp_b200_model_free(model);
  }
  model = nullptr;
  if (pdlp_b200_model_from_mps_text(in.data(), size, Rand(2), &model, err, sizeof err) == PDLP_B200_STATUS_OK) {
    pdlp_b200_model_constraint_name(model, 0);
    pdlp_b200_model_free(model);
  }
  PdlpBlob resp{};
  if (pdlp_b200_solve_proto(data, size, Rand(2), nullptr, &resp) == PDLP_B200_STATUS_OK) pdlp_b200_blob_free(&resp);
}

// The file layer: the input as the bytes of a .mps.bz2 / .mps.gz file (the bzip2 decoder is bound
// at run time, formats.cc ReadBzip2File). Mutated compressed streams must end as errors or models.
void ExerciseFile(const std::string& in, const char* suffix) {
  char err[256];
  const std::string path = std::string("build/fuzz_input.mps") + suffix;
  std::FILE* f = std::fopen(path.c_str(), "wb");
  if (f == nullptr) return;
  std::fwrite(in.data(), 1, in.size(), f);
  std::fclose(f);
  PdlpModel* model = nullptr;
  if (pdlp_b200_read_quadratic_program(path.c_str(), Rand(2), &model, err, sizeof err) == PDLP_B200_STATUS_OK) pdlp_b200_model_free(model);
}
}  // namespace

int main(int argc, char** argv) {
  const long iterations = argc > 1 ? std::atol(argv[1]) : 20000;
  std::vector<std::string> corpus;
  for (int i = 2; i < argc; ++i) {  // seed files
    std::FILE* f = std::fopen(argv[i], "rb");
    if (f == nullptr) continue;
    std::string s;
    char buf[4096];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
    std::fclose(f);
    corpus.push_back(s);
  }
  std::vector<std::string> packed;  // seed files that are compressed streams: fuzzed through the file layer as well
  for (const std::string& s : corpus) {
    if (s.size() > 3 && s.compare(0, 3, "BZh") == 0) packed.push_back(s);
    if (s.size() > 2 && static_cast<unsigned char>(s[0]) == 0x1f && static_cast<unsigned char>(s[1]) == 0x8b) packed.push_back(s);
  }
  corpus.push_back("");
  corpus.push_back("termination_criteria { simple_optimality_criteria { eps_optimal_absolute: 1e-4 } iteration_limit: 0x10 } random_projection_seeds: [1, 2]");
  corpus.push_back("{\"terminationCriteria\": {\"iterationLimit\": 5, \"timeSecLimit\": \"Infinity\"}, \"randomProjectionSeeds\": [1, 2]}");
  corpus.push_back("NAME x\nROWS\n N c\n E r\nCOLUMNS\n a c 1 r 1\nRHS\n rhs r 3\nRANGES\n rng r 1\nBOUNDS\n UP bnd a 4\nENDATA\n");
  for (const std::string& s : std::vector<std::string>(corpus)) Exercise(s);
  for (long it = 0; it < iterations; ++it) {
    std::string s = Mutate(corpus[static_cast<size_t>(Rand(static_cast<int>(corpus.size())))]);
    Exercise(s);
    if (s.size() < 4096 && Rand(20) == 0) corpus.push_back(s);
    if (!packed.empty() && it % 8 == 0) {
      const std::string& seed = packed[static_cast<size_t>(Rand(static_cast<int>(packed.size())))];
      ExerciseFile(Rand(8) == 0 ? seed : Mutate(seed), seed[0] == 'B' ? ".bz2" : ".gz");
    }
  }
  std::printf("fuzz ok: %ld inputs, corpus %zu\n", iterations, corpus.size());
  return 0;
}
