"""Command-line solve, the counterpart of ``examples/cpp/pdlp_solve.cc``:

    python -m ortools_b200.pdlp_solve --input model.mps --params 'termination_criteria { ... }' \\
        --solve_log_file log.textproto --sol_file model.sol

Same flags and outputs as the reference binary: ``--input`` (.mps, .mps.gz or an
MPModelProto as .pb / .textproto / .json / .json.gz), ``--params`` (text-format
PrimalDualHybridGradientParams, verbosity_level 2 unless overridden), ``--solve_log_file``
(.textproto, .pb or .json) and ``--sol_file`` (Miplib .sol: ``=obj=`` line, then one
``name value`` line per variable). Integrality constraints are dropped on input. ^C
interrupts the solve through the interrupt flag of the C ABI. The solve runs on the GPU
through libpdlp_b200.so; there is no CPU fallback.
"""
import argparse
import ctypes
import signal
import threading
import sys

from google.protobuf import json_format, text_format

from . import mp_model, pdlp, pdlp_proto


def write_solve_log(path, log_proto):
    """WriteSolveLog, pdlp_solve.cc:64-79."""
    if path.endswith(".textproto"):
        data = text_format.MessageToString(log_proto).encode()
    elif path.endswith(".pb"):
        data = log_proto.SerializeToString()
    elif path.endswith(".json"):
        data = json_format.MessageToJson(log_proto, preserving_proto_field_name=True).encode()
    else:
        raise SystemExit("Unrecognized file extension for --solve_log_file: %s. Expected .textproto, .pb, or .json" % path)
    with open(path, "wb") as f:
        f.write(data)


def sol_string(qp, result):
    """The .sol text of pdlp_solve.cc:117-133, or None without convergence information."""
    ci = mp_model.get_convergence_information(result.solve_log.solution_stats, result.solve_log.solution_type)
    if ci is None:
        return None
    lines = ["=obj= %r" % float(ci.primal_objective)]
    for i, v in enumerate(result.primal_solution):
        name = qp.variable_names[i] if qp.variable_names is not None else "var%d" % i
        lines.append("%s %r" % (name, float(v)))
    return "\n".join(lines) + "\n"


def solve(input_path, params_text="", solve_log_file="", sol_file="", backend=None, out=sys.stderr):
    """Solve(), pdlp_solve.cc:81-135. Returns the SolverResult."""
    if not input_path:
        raise SystemExit("--input is required")
    params_msg = pdlp_proto.PrimalDualHybridGradientParamsProto()
    params_msg.verbosity_level = 2  # print iteration statistics by default
    try:
        text_format.Merge(params_text, params_msg)
    except text_format.ParseError as e:
        raise SystemExit("Error parsing --params: %s" % e)
    params = pdlp_proto.params_from_proto(params_msg)
    qp = pdlp.read_quadratic_program_or_die(input_path, include_names=True)  # the library's C++ readers; drops integrality constraints
    interrupted = ctypes.c_int32(0)
    previous = signal.getsignal(signal.SIGINT)
    try:
        signal.signal(signal.SIGINT, lambda *_: setattr(interrupted, "value", 1))
    except ValueError:  # not the main thread
        previous = None
    be = backend if backend is not None else pdlp.backend()

    def run():
        return be.primal_dual_hybrid_gradient(qp, params, interrupt_solve=interrupted,
                                              message_callback=lambda m: print(m, file=out, flush=True))
    try:
        if previous is None:
            result = run()
        else:
            # Python runs signal handlers on the main thread between bytecodes, never inside a foreign
            # call: the solve goes to a worker (ctypes drops the GIL for the call) and the main thread
            # waits in short joins, so ^C reaches interrupt_solve at any verbosity.
            box = {}

            def work():
                try:
                    box["result"] = run()
                except BaseException as e:  # re-raised on the main thread
                    box["error"] = e
            worker = threading.Thread(target=work, name="pdlp_solve", daemon=True)
            worker.start()
            while worker.is_alive():
                worker.join(0.05)
            if "error" in box:
                raise box["error"]
            result = box["result"]
    finally:
        if previous is not None:
            signal.signal(signal.SIGINT, previous)
    if solve_log_file:
        print("Writing SolveLog to '%s'." % solve_log_file, file=out)
        write_solve_log(solve_log_file, pdlp_proto.solve_log_to_proto(result.solve_log, params))
    if sol_file:
        text = sol_string(qp, result)
        if text is not None:
            print("Writing .sol solution to '%s'." % sol_file, file=out)
            with open(sol_file, "w") as f:
                f.write(text)
    return result


def main(argv=None):
    ap = argparse.ArgumentParser(description="Solve an LP / diagonal QP with PDLP on a B200.")
    ap.add_argument("--input", default="", help="REQUIRED: .mps, .mps.gz, .mps.bz2, or an MPModelProto [.pb, .textproto, .json, .json.gz]")
    ap.add_argument("--params", default="", help="PrimalDualHybridGradientParams in text format")
    ap.add_argument("--solve_log_file", default="", help="If non-empty, writes PDLP's SolveLog here (.textproto, .pb or .json)")
    ap.add_argument("--sol_file", default="", help="If non-empty, output the final primal solution in Miplib .sol format")
    a = ap.parse_args(argv)
    result = solve(a.input, a.params, a.solve_log_file, a.sol_file)
    print("termination_reason: %s" % pdlp.TerminationReason.Name(result.solve_log.termination_reason), file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
