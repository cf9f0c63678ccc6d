"""Problem I/O on the host side of the path:
``ReadQuadraticProgramOrDie`` & co. of ``ortools/pdlp/quadratic_program_io.h:28-59``.

* ``read_mps_linear_program`` -- MPS (free or fixed format, optionally .gz / .bz2) to a
  ``QuadraticProgram`` with the semantics of ``ortools/lp_data/mps_reader_template.h:90-260``
  as PDLP uses it (``quadratic_program_io.cc:364-407``): rows and columns are numbered in
  order of first appearance, integrality is dropped (PDLP solves the relaxation),
  maximisation becomes minimisation with ``objective_scaling_factor = -1``, minus the
  RHS of the objective row is the objective offset.
* ``read_mp_model_proto_file`` -- MPModelProto as .pb / .textproto / .json (optionally .gz).
* ``read_quadratic_program`` -- dispatch on the file suffix.
* ``write_linear_program_to_mps`` / ``write_quadratic_program_to_mp_model_proto``.
"""
import bz2
import gzip
import io
import math

import numpy as np
import scipy.sparse as sp
from google.protobuf import json_format, text_format

from . import mp_model, pdlp

INF = float("inf")
_SECTIONS = {"NAME", "OBJSENSE", "OBJSENCE", "OBJSENSEMAX", "ROWS", "LAZYCONS", "COLUMNS", "RHS", "RANGES", "BOUNDS", "INDICATORS", "ENDATA",
             "QUADOBJ", "QMATRIX", "QSECTION", "SOS", "USERCUTS"}


class MpsError(ValueError):
    pass


def _open_text(path):
    if path.endswith(".gz"):
        return io.TextIOWrapper(gzip.open(path, "rb"), encoding="utf-8", errors="replace")
    if path.endswith(".bz2"):
        return io.TextIOWrapper(bz2.open(path, "rb"), encoding="utf-8", errors="replace")
    return open(path, "r", encoding="utf-8", errors="replace")


def _number(tok, lineno):
    try:
        v = float(tok.replace("D", "E").replace("d", "e")) if ("D" in tok or "d" in tok) and "inf" not in tok.lower() else float(tok)
    except ValueError:
        raise MpsError("line %d: cannot parse number %r" % (lineno, tok))
    if math.isnan(v):
        raise MpsError("line %d: NaN value" % lineno)
    return v


def _fixed_fields(line):
    """Fixed-format fields: columns 2-3, 5-12, 15-22, 25-36, 40-47, 50-61 (1-based)."""
    cols = [(1, 3), (4, 12), (14, 22), (24, 36), (39, 47), (49, 61)]
    out = [line[a:b].strip() for a, b in cols if len(line) > a]
    while out and out[-1] == "":
        out.pop()
    return out


def parse_mps(lines, include_names=False):
    """Parses MPS text (an iterable of lines) into a QuadraticProgram."""
    section = None
    name = None
    maximize = False
    row_index, row_type, row_names = {}, [], []
    objective_row = None
    ignored_rows = set()  # extra N rows
    col_index, col_names = {}, []
    obj = []
    lv, uv = [], []
    binary_by_default = []
    entries_r, entries_c, entries_v = [], [], []
    lc, uc = [], []
    offset = 0.0
    in_integer_block = False

    def find_col(cname, lineno, create):
        j = col_index.get(cname)
        if j is None:
            if not create:
                raise MpsError("line %d: unknown column %r" % (lineno, cname))
            j = len(col_names)
            col_index[cname] = j
            col_names.append(cname)
            obj.append(0.0)
            lv.append(0.0)
            uv.append(INF)
            binary_by_default.append(False)
        return j

    def set_rhs(rname, value, lineno):
        if rname == objective_row:
            nonlocal offset
            offset = -value  # minus the right-hand side of the objective row
            return
        if rname in ignored_rows:
            return
        i = row_index.get(rname)
        if i is None:
            raise MpsError("line %d: unknown row %r" % (lineno, rname))
        lc[i] = -INF if lc[i] == -INF else value
        uc[i] = INF if uc[i] == INF else value

    def set_range(rname, value, lineno):
        if rname == objective_row or rname in ignored_rows:
            return
        i = row_index.get(rname)
        if i is None:
            raise MpsError("line %d: unknown row %r" % (lineno, rname))
        lo, hi = lc[i], uc[i]
        if lo == hi:
            if value < 0.0:
                lo += value
            else:
                hi += value
        if lo == -INF:
            lo = hi - abs(value)
        if hi == INF:
            hi = lo + abs(value)
        lc[i], uc[i] = lo, hi

    for lineno, raw in enumerate(lines, 1):
        line = raw.rstrip("\r\n")
        if not line.strip() or line.lstrip().startswith("*"):
            continue
        if not line[0].isspace():  # section header
            parts = line.split()
            key = parts[0].upper()
            if key not in _SECTIONS:
                raise MpsError("line %d: unknown section %r" % (lineno, parts[0]))
            section = key
            if key == "NAME":
                name = " ".join(parts[1:]) if len(parts) > 1 else ""
            elif key in ("OBJSENSE", "OBJSENCE") and len(parts) > 1:
                maximize = parts[1].upper() in ("MAX", "MAXIMIZE")
            elif key == "OBJSENSEMAX":
                maximize = True
            elif key == "ENDATA":
                break
            elif key in ("QUADOBJ", "QMATRIX", "QSECTION"):
                raise MpsError("line %d: quadratic objective sections are not supported by the linear-program reader" % lineno)
            elif key in ("INDICATORS", "SOS"):
                raise MpsError("line %d: section %s is not supported" % (lineno, key))
            continue
        f = line.split()
        if section in ("OBJSENSE", "OBJSENCE"):
            maximize = f[0].upper() in ("MAX", "MAXIMIZE")
        elif section in ("ROWS", "LAZYCONS", "USERCUTS"):
            if len(f) != 2:
                f = _fixed_fields(line)
            if len(f) != 2:
                raise MpsError("line %d: expected <type> <row name>" % lineno)
            t, rname = f[0].upper(), f[1]
            if t == "N":
                if objective_row is None:
                    objective_row = rname
                else:
                    ignored_rows.add(rname)
                continue
            if t not in ("E", "L", "G"):
                raise MpsError("line %d: unknown row type %r" % (lineno, f[0]))
            if rname in row_index:
                raise MpsError("line %d: duplicate row %r" % (lineno, rname))
            row_index[rname] = len(row_names)
            row_names.append(rname)
            row_type.append(t)
            lc.append(-INF if t == "L" else 0.0)
            uc.append(INF if t == "G" else 0.0)
        elif section == "COLUMNS":
            if len(f) >= 3 and f[1].upper() == "'MARKER'":
                in_integer_block = "INTORG" in f[2].upper()
                continue
            if len(f) not in (3, 5):
                f = _fixed_fields(line)[1:]
            if len(f) not in (3, 5):
                raise MpsError("line %d: expected <column> <row> <value> [<row> <value>]" % lineno)
            new = f[0] not in col_index
            j = find_col(f[0], lineno, True)
            if new and in_integer_block:
                binary_by_default[j] = True  # integer by marker, no bound yet: [0, 1]
                uv[j] = 1.0
            for k in (1, 3):
                if k >= len(f):
                    break
                rname, value = f[k], _number(f[k + 1], lineno)
                if rname == objective_row:
                    obj[j] = value
                elif rname in ignored_rows:
                    continue
                else:
                    i = row_index.get(rname)
                    if i is None:
                        raise MpsError("line %d: unknown row %r" % (lineno, rname))
                    entries_r.append(i)
                    entries_c.append(j)
                    entries_v.append(value)
        elif section == "RHS":
            g = f if len(f) % 2 == 1 else [""] + f  # the set name may be missing
            if len(g) not in (3, 5):
                g = _fixed_fields(line)[1:]
            for k in range(1, len(g) - 1, 2):
                set_rhs(g[k], _number(g[k + 1], lineno), lineno)
        elif section == "RANGES":
            g = f if len(f) % 2 == 1 else [""] + f
            if len(g) not in (3, 5):
                g = _fixed_fields(line)[1:]
            for k in range(1, len(g) - 1, 2):
                set_range(g[k], _number(g[k + 1], lineno), lineno)
        elif section == "BOUNDS":
            kind = f[0].upper()
            needs_value = kind in ("LO", "UP", "FX", "LI", "UI", "SC")
            # ' <type> <set name> <column> <value>'; the set name may be missing
            if needs_value:
                if len(f) == 4:
                    cname, value = f[2], _number(f[3], lineno)
                elif len(f) == 3:
                    cname, value = f[1], _number(f[2], lineno)
                else:
                    g = _fixed_fields(line)
                    if len(g) < 4:
                        raise MpsError("line %d: malformed bound" % lineno)
                    cname, value = g[2], _number(g[3], lineno)
            else:
                if len(f) >= 3:
                    cname = f[2]
                elif len(f) == 2:
                    cname = f[1]
                else:
                    raise MpsError("line %d: malformed bound" % lineno)
                value = 0.0
            j = find_col(cname, lineno, True)
            lo, hi = lv[j], uv[j]
            if binary_by_default[j]:
                lo, hi = 0.0, INF
            if kind in ("LO", "LI"):
                lo = value
                if kind == "LI" and lo == 0.0:
                    hi = INF
            elif kind in ("UP", "UI"):
                hi = value
            elif kind == "FX":
                lo = hi = value
            elif kind == "FR":
                lo, hi = -INF, INF
            elif kind == "MI":
                lo = -INF
            elif kind == "PL":
                hi = INF
            elif kind == "BV":
                lo, hi = 0.0, 1.0
            elif kind == "SC":
                raise MpsError("line %d: semi-continuous variables are not supported" % lineno)
            else:
                raise MpsError("line %d: unknown bound type %r" % (lineno, f[0]))
            binary_by_default[j] = False
            lv[j], uv[j] = lo, hi
        elif section == "NAME":
            continue
        else:
            raise MpsError("line %d: data outside of a section" % lineno)

    n, m = len(col_names), len(row_names)
    qp = pdlp.QuadraticProgram(n, m)
    k = sp.csc_matrix((np.asarray(entries_v, dtype=np.float64), (np.asarray(entries_r, dtype=np.int64), np.asarray(entries_c, dtype=np.int64))), shape=(m, n))
    k.sum_duplicates()
    k.sort_indices()
    qp.constraint_matrix = k
    qp.constraint_lower_bounds = np.asarray(lc, dtype=np.float64).reshape(m)
    qp.constraint_upper_bounds = np.asarray(uc, dtype=np.float64).reshape(m)
    qp.variable_lower_bounds = np.asarray(lv, dtype=np.float64).reshape(n)
    qp.variable_upper_bounds = np.asarray(uv, dtype=np.float64).reshape(n)
    qp.objective_vector = np.asarray(obj, dtype=np.float64).reshape(n)
    qp.objective_offset = offset
    if maximize:  # quadratic_program_io.cc:259-266
        qp.objective_scaling_factor = -1.0
        qp.objective_offset *= -1
        qp.objective_vector *= -1
    if include_names:
        qp.problem_name = name
        qp.variable_names = list(col_names)
        qp.constraint_names = list(row_names)
    return qp


def read_mps_linear_program(path, include_names=False):
    with _open_text(path) as f:
        return parse_mps(f, include_names)


def read_mp_model_proto_file(path, include_names=False):
    """MPModelProto in binary (.pb), text (.textproto) or JSON (.json), optionally gzipped."""
    raw = gzip.open(path, "rb").read() if path.endswith(".gz") else open(path, "rb").read()
    base = path[:-3] if path.endswith(".gz") else path
    proto = mp_model.MPModelProto()
    if base.endswith(".textproto"):
        text_format.Parse(raw.decode("utf-8"), proto)
    elif base.endswith(".json"):
        json_format.Parse(raw.decode("utf-8"), proto)
    else:
        proto.ParseFromString(raw)
    return mp_model.qp_from_mp_model_proto(proto, relax_integer_variables=True, include_names=include_names)


def read_quadratic_program(path, include_names=False):
    """ReadQuadraticProgramOrDie, quadratic_program_io.cc:50-68 (raises instead of dying)."""
    if path.endswith((".mps", ".mps.gz", ".mps.bz2")):
        return read_mps_linear_program(path, include_names)
    if path.endswith((".pb", ".textproto", ".json", ".json.gz", ".pb.gz", ".textproto.gz")):
        return read_mp_model_proto_file(path, include_names)
    raise ValueError("Invalid filename suffix in %s. Valid suffixes are .mps, .mps.gz, .pb, .textproto, .json, and .json.gz" % path)


def _fmt(v):
    return repr(float(v))


def write_linear_program_to_mps(qp, path):
    """WriteLinearProgramToMps, quadratic_program_io.cc:80-93 (free-format MPS)."""
    if not pdlp.is_linear_program(qp):
        raise ValueError("'linear_program' has a quadratic objective")
    k = sp.csc_matrix(qp.constraint_matrix)
    k.sort_indices()
    m, n = k.shape
    s = qp.objective_scaling_factor
    rn = qp.constraint_names if qp.constraint_names else ["R%d" % i for i in range(m)]
    cn = qp.variable_names if qp.variable_names else ["C%d" % j for j in range(n)]
    lc, uc = qp.constraint_lower_bounds, qp.constraint_upper_bounds
    out = ["NAME %s" % (qp.problem_name or "")]
    if s < 0:
        out += ["OBJSENSE", "    MAX"]
    out.append("ROWS")
    out.append(" N COST")
    kinds = []
    for i in range(m):
        if lc[i] == uc[i]:
            t = "E"
        elif lc[i] == -INF and uc[i] == INF:
            raise ValueError("free constraint row %d cannot be written to MPS" % i)
        elif lc[i] == -INF:
            t = "L"
        else:
            t = "G"  # ranged rows: G with a RANGES entry
        kinds.append(t)
        out.append(" %s %s" % (t, rn[i]))
    out.append("COLUMNS")
    for j in range(n):
        wrote = False
        if qp.objective_vector[j] != 0.0:
            out.append("    %s COST %s" % (cn[j], _fmt(s * qp.objective_vector[j])))
            wrote = True
        for p in range(k.indptr[j], k.indptr[j + 1]):
            out.append("    %s %s %s" % (cn[j], rn[int(k.indices[p])], _fmt(k.data[p])))
            wrote = True
        if not wrote:
            out.append("    %s COST 0.0" % cn[j])
    out.append("RHS")
    if qp.objective_offset != 0.0:
        out.append("    RHS COST %s" % _fmt(-s * qp.objective_offset))
    for i in range(m):
        rhs = uc[i] if kinds[i] == "L" else lc[i]
        if rhs != 0.0:
            out.append("    RHS %s %s" % (rn[i], _fmt(rhs)))
    ranged = [i for i in range(m) if kinds[i] == "G" and uc[i] != INF]
    if ranged:
        out.append("RANGES")
        for i in ranged:
            out.append("    RNG %s %s" % (rn[i], _fmt(uc[i] - lc[i])))
    out.append("BOUNDS")
    lv, uv = qp.variable_lower_bounds, qp.variable_upper_bounds
    for j in range(n):
        if lv[j] == -INF and uv[j] == INF:
            out.append(" FR BND %s" % cn[j])
        elif lv[j] == uv[j]:
            out.append(" FX BND %s %s" % (cn[j], _fmt(lv[j])))
        else:
            if lv[j] == -INF:
                out.append(" MI BND %s" % cn[j])
            elif lv[j] != 0.0:
                out.append(" LO BND %s %s" % (cn[j], _fmt(lv[j])))
            if uv[j] != INF:
                out.append(" UP BND %s %s" % (cn[j], _fmt(uv[j])))
    out.append("ENDATA")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")


def write_quadratic_program_to_mp_model_proto(qp, path):
    """WriteQuadraticProgramToMPModelProto, quadratic_program_io.cc:95-101 (binary proto)."""
    with open(path, "wb") as f:
        f.write(mp_model.qp_to_mp_model_proto(qp).SerializeToString())
