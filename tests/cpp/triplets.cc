// triplets.cc -- known answers of quadratic_program_test.cc:558-628 (CombineRepeatedTripletsInPlace,
// SetEigenMatrixFromTriplets) for the helpers of include/pdlp_b200.hpp. Exit code = first failing check.
#include <cstdio>
#include <string>
#include <vector>

#include "pdlp_b200.hpp"

using pdlp_b200::Triplet;

static bool Same(const std::vector<Triplet>& got, const std::vector<Triplet>& want) {
  if (got.size() != want.size()) return false;
  for (size_t k = 0; k < got.size(); ++k)
    if (got[k].row != want[k].row || got[k].col != want[k].col || got[k].value != want[k].value) return false;
  return true;
}

static std::vector<Triplet> Combined(std::vector<Triplet> t) {
  pdlp_b200::CombineRepeatedTripletsInPlace(t);
  return t;
}

int main() {
  if (!Same(Combined({}), {})) return 1;
  if (!Same(Combined({{1, 2, 3.0}}), {{1, 2, 3.0}})) return 2;
  if (!Same(Combined({{1, 2, 3.0}, {2, 1, 1.0}, {1, 1, 0.0}}), {{1, 2, 3.0}, {2, 1, 1.0}, {1, 1, 0.0}})) return 3;
  if (!Same(Combined({{1, 2, 3.0}, {1, 2, -1.0}, {1, 1, 0.0}}), {{1, 2, 2.0}, {1, 1, 0.0}})) return 4;
  if (!Same(Combined({{1, 2, 3.0}, {2, 1, 1.0}, {2, 1, 1.0}}), {{1, 2, 3.0}, {2, 1, 2.0}})) return 5;
  if (!Same(Combined({{1, 2, 3.0}, {1, 2, 1.0}, {1, 2, 2.0}}), {{1, 2, 6.0}})) return 6;
  // the matrix from triplets: an empty 2 x 2 matrix, then [[2, 0], [-1, 1]] with repeated (0, 0) entries
  pdlp_b200::QuadraticProgram qp(2, 2);
  qp.SetConstraintMatrixFromTriplets({});
  if (qp.col_starts != std::vector<int64_t>{0, 0, 0} || !qp.values.empty()) return 7;
  qp.SetConstraintMatrixFromTriplets({{0, 0, 1.0}, {1, 0, -1.0}, {0, 0, 0.0}, {1, 1, 1.0}, {0, 0, 1.0}});
  if (qp.col_starts != std::vector<int64_t>{0, 2, 3} || qp.row_indices != std::vector<int64_t>{0, 1, 1} ||
      qp.values != std::vector<double>{2.0, -1.0, 1.0})
    return 8;
  // ValidateQuadraticProgramDimensions, quadratic_program_test.cc:68-162
  {
    using pdlp_b200::QuadraticProgram;
    using pdlp_b200::ValidateQuadraticProgramDimensions;
    QuadraticProgram ok(2, 1);
    ok.objective_matrix_diagonal = std::vector<double>{4.0, 1.0};
    ok.variable_names = std::vector<std::string>{"x0", "x1"};
    ok.constraint_names = std::vector<std::string>{"c0"};
    if (!ValidateQuadraticProgramDimensions(ok).empty() || pdlp_b200::IsLinearProgram(ok)) return 9;
    int code = 10;
    auto bad = [&](void (*change)(QuadraticProgram&)) {
      QuadraticProgram q(2, 3);
      change(q);
      return ValidateQuadraticProgramDimensions(q).rfind("Inconsistent dimensions: ", 0) == 0;
    };
    if (!bad([](QuadraticProgram& q) { q.constraint_lower_bounds.resize(10); })) return code;
    ++code;
    if (!bad([](QuadraticProgram& q) { q.constraint_upper_bounds.resize(10); })) return code;
    ++code;
    if (!bad([](QuadraticProgram& q) { q.objective_vector.resize(10); })) return code;
    ++code;
    if (!bad([](QuadraticProgram& q) { q.variable_lower_bounds.resize(10); })) return code;
    ++code;
    if (!bad([](QuadraticProgram& q) { q.variable_upper_bounds.resize(10); })) return code;
    ++code;
    if (!bad([](QuadraticProgram& q) { q.col_starts.resize(11); })) return code;  // 10 columns
    ++code;
    if (!bad([](QuadraticProgram& q) { q.objective_matrix_diagonal = std::vector<double>(10, 0.0); })) return code;
    ++code;
    if (!bad([](QuadraticProgram& q) { q.variable_names = std::vector<std::string>{"x0"}; })) return code;
    ++code;
    if (!bad([](QuadraticProgram& q) { q.constraint_names = std::vector<std::string>{"c0"}; })) return code;
    if (!pdlp_b200::IsLinearProgram(QuadraticProgram(2, 3))) return 30;
  }
  std::printf("triplets ok\n");
  return 0;
}
