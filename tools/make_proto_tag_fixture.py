#!/usr/bin/env python
"""Generates tests/golden/pdlp_proto_tags.json from the reference's own schema files
(/root/reference/ortools/pdlp/solvers.proto, solve_log.proto): for every message the
(field name -> [tag, label, type, default]) table and for every enum its (name -> number)
table. The fixture pins or-tools_b200/pdlp_proto.py to the reference's wire format;
run here (the reference is not present on the GPU box):  python tools/make_proto_tag_fixture.py"""
import json
import os
import re
import sys

REF = "/root/reference/ortools/pdlp"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pdlp_proto_tags.json")


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def parse(path):
    text = strip_comments(open(path).read())
    text = re.sub(r"map<[^>]*>[^;]*;", "", text)  # map fields are not used on this path
    tokens = re.findall(r"[A-Za-z_][A-Za-z0-9_.]*|0x[0-9A-Fa-f]+|-?[0-9][0-9.eE+-]*|-inf|[{}=;\[\]]|\"[^\"]*\"", text)
    messages, enums = {}, {}
    stack = []
    i = 0
    while i < len(tokens):
        t = tokens[i]
        if t in ("message", "enum", "oneof") and tokens[i + 2] == "{":
            stack.append((t, tokens[i + 1]))
            scope = ".".join(n for k, n in stack if k != "oneof")
            if t == "message":
                messages.setdefault(scope, {})
            elif t == "enum":
                enums.setdefault(scope, {})
            i += 3
            continue
        if t == "}":
            stack.pop()
            i += 1
            continue
        if stack and stack[-1][0] == "enum" and i + 2 < len(tokens) and tokens[i + 1] == "=":
            scope = ".".join(n for k, n in stack if k != "oneof")
            enums[scope][t] = int(tokens[i + 2], 0)
            i += 3
            continue
        in_msg = stack and stack[-1][0] in ("message", "oneof")
        if in_msg and t in ("optional", "repeated", "required") or (in_msg and stack[-1][0] == "oneof" and i + 3 < len(tokens) and tokens[i + 2] == "="):
            if t in ("optional", "repeated", "required"):
                label, ftype, name, number = t, tokens[i + 1], tokens[i + 2], int(tokens[i + 4])
                j = i + 5
            else:
                label, ftype, name, number = "optional", t, tokens[i + 1], int(tokens[i + 3])
                j = i + 4
            default = None
            if tokens[j] == "[":
                k = j
                while tokens[k] != "]":
                    if tokens[k] == "default":
                        default = tokens[k + 2]
                    k += 1
                j = k + 1
            scope = ".".join(n for k2, n in stack if k2 != "oneof")
            messages[scope][name] = [number, label, ftype.split(".")[-1], default]
            i = j
            continue
        i += 1
    return messages, enums


def main():
    if not os.path.isdir(REF):
        sys.exit("reference checkout not present")
    out = {"messages": {}, "enums": {}}
    for f in ("solvers.proto", "solve_log.proto"):
        m, e = parse(os.path.join(REF, f))
        out["messages"].update(m)
        out["enums"].update(e)
    # the MPSolver-side messages the PDLP proto solver reads / writes (pdlp_proto_solver.cc:36-130)
    m, e = parse(os.path.join(os.path.dirname(REF), "linear_solver", "linear_solver.proto"))
    keep = ("MPVariableProto", "MPConstraintProto", "MPQuadraticObjective", "MPModelProto", "MPModelRequest", "MPSolutionResponse")
    out["linear_solver_messages"] = {k: m[k] for k in keep}
    out["linear_solver_enums"] = {"MPSolverResponseStatus": e["MPSolverResponseStatus"], "MPModelRequest.SolverType": e["MPModelRequest.SolverType"]}
    json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, len(out["messages"]), "messages", len(out["enums"]), "enums")


if __name__ == "__main__":
    main()
