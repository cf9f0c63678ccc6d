// proto_codec.h -- a small proto2 codec for the messages either side of the
// PDLP path. There is no protoc / libprotobuf in
// this build, so the wire format (varint / fixed64 / length-delimited), the
// text format and the JSON mapping are implemented here, driven by schema
// tables that restate the reference's .proto files (message / field / enum
// names, tags, types): ortools/pdlp/solvers.proto, ortools/pdlp/solve_log.proto
// and the part of ortools/linear_solver/linear_solver.proto that PDLP's
// callers use. Host-only code; nothing here touches the device.
#ifndef PDLP_B200_PROTO_CODEC_H_
#define PDLP_B200_PROTO_CODEC_H_

#include <cstdint>
#include <string>
#include <string_view>
#include <vector>

namespace pdlp_b200 {
namespace proto {

// ---- wire primitives -------------------------------------------------------
enum WireType { kVarint = 0, kFixed64 = 1, kLengthDelimited = 2, kFixed32 = 5 };

class Writer {
 public:
  void Varint(uint64_t v);
  void Tag(int field, WireType t) { Varint((static_cast<uint64_t>(field) << 3) | static_cast<uint64_t>(t)); }
  void Int(int field, int64_t v) { Tag(field, kVarint); Varint(static_cast<uint64_t>(v)); }  // int32/int64/enum/bool
  void Bool(int field, bool v) { Int(field, v ? 1 : 0); }
  void Double(int field, double v);
  void Bytes(int field, std::string_view v);
  void PackedDoubles(int field, const double* v, int64_t n);
  void PackedInts(int field, const int32_t* v, int64_t n);
  void RawDouble(double v);
  std::string& out() { return out_; }
  const std::string& out() const { return out_; }

 private:
  std::string out_;
};

// One decoded field of a message: for kVarint / kFixed64 / kFixed32 `value`
// holds the raw bits; for kLengthDelimited `bytes` is the payload.
struct WireField {
  int number = 0;
  WireType type = kVarint;
  uint64_t value = 0;
  std::string_view bytes;
  double AsDouble() const;
  int64_t AsInt64() const { return static_cast<int64_t>(value); }
  int32_t AsInt32() const { return static_cast<int32_t>(static_cast<int64_t>(value)); }
  bool AsBool() const { return value != 0; }
};

class Reader {
 public:
  explicit Reader(std::string_view data) : p_(data.data()), end_(data.data() + data.size()) {}
  // Next field; false at the end of the message or on malformed input
  // (ok() tells which).
  bool Next(WireField* f);
  bool ok() const { return ok_; }
  bool Varint(uint64_t* v);

 private:
  const char* p_;
  const char* end_;
  bool ok_ = true;
};

// Repeated scalar fields may arrive packed or one by one; these append either.
bool AppendDoubles(const WireField& f, std::vector<double>* out);
bool AppendInt32s(const WireField& f, std::vector<int32_t>* out);

// ---- schema tables ---------------------------------------------------------
enum class FieldType { kDouble, kInt32, kInt64, kBool, kString, kBytes, kEnum, kMessage };
struct EnumValue { const char* name; int number; };
struct EnumDef { const char* name; std::vector<EnumValue> values; const char* NameOf(int number) const; bool NumberOf(std::string_view name, int* number) const; };
struct Schema;
struct FieldDef {
  const char* name;
  int number;
  FieldType type;
  bool repeated = false;
  bool packed = false;
  const Schema* message = nullptr;
  const EnumDef* enumeration = nullptr;
  int oneof = 0;  // members of the same oneof share a non-zero id
};
struct Schema {
  const char* name;
  std::vector<FieldDef> fields;
  const FieldDef* ByName(std::string_view name) const;
  const FieldDef* ByNumber(int number) const;
};

// solvers.proto
const Schema& TerminationCriteriaSchema();
const Schema& ParamsSchema();
// solve_log.proto
const Schema& IterationStatsSchema();
const Schema& SolveLogSchema();
// linear_solver.proto (subset: quadratic_program.cc:99-320, pdlp_proto_solver.cc:36-130)
const Schema& MPModelSchema();
const Schema& MPModelRequestSchema();
const Schema& MPSolutionResponseSchema();

// ---- text format / JSON ----------------------------------------------------
// Protobuf text format -> wire bytes (appended to *out). Unknown field names are
// errors, like google::protobuf::TextFormat. `allow_singular_overwrites` = the
// TextFormat::Merge policy (a non-repeated field may be given again, another
// member of a oneof may replace the first; the reference merges --params and
// solver_specific_parameters this way); false = the TextFormat::Parse policy
// (both are errors). Returns false and sets *error.
bool TextToWire(const Schema& schema, std::string_view text, std::string* out, std::string* error,
                bool allow_singular_overwrites = false);
// Wire bytes -> text format (fields in tag order, two-space indentation, the
// layout google::protobuf::TextFormat prints). Unknown fields are dropped.
bool WireToText(const Schema& schema, std::string_view bytes, std::string* out);
// Wire bytes -> proto3 JSON mapping (lowerCamelCase names, enums by name,
// int64 as strings, non-finite doubles as "Infinity" / "-Infinity" / "NaN").
bool WireToJson(const Schema& schema, std::string_view bytes, std::string* out);
// JSON (either lowerCamelCase or the original field names) -> wire bytes.
bool JsonToWire(const Schema& schema, std::string_view json, std::string* out, std::string* error);

// Shortest decimal text that parses back to exactly `v` (the reference's
// RoundTripDoubleFormat, util/fp_roundtrip_conv.h); "inf" / "-inf" / "nan".
std::string RoundTripDouble(double v);

}  // namespace proto
}  // namespace pdlp_b200

#endif  // PDLP_B200_PROTO_CODEC_H_
