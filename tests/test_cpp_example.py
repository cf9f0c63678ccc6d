"""The C++ face of the C ABI (include/pdlp_b200.hpp) through the counterpart of
``ortools/pdlp/samples/simple_pdlp_program.cc``: compiles and links everywhere; without a GPU it
fails loudly (no CPU fallback), on a GPU it prints the optimum of TestLp
(primal_dual_hybrid_gradient_test.cc: x = [-1, 8, 1, 2.5], y = [-2, 0, 2.375, 2/3], objective -34)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DIR = os.path.join(ROOT, "or-tools_b200", "lib")


@pytest.fixture(scope="module")
def example_binary(tmp_path_factory):
    if not os.path.exists(os.path.join(LIB_DIR, "libpdlp_b200.so")):
        import __graft_entry__
        __graft_entry__.build()
    out = str(tmp_path_factory.mktemp("cpp") / "solve_simple_lp")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "solve_simple_lp.cc"),
                           "-L" + LIB_DIR, "-lpdlp_b200", "-Wl,-rpath," + LIB_DIR, "-o", out])
    return out


def _sections(text):
    out, key = {}, None
    for line in text.splitlines():
        if line.endswith(":") and not line[0].isdigit() and not line.startswith("-"):
            key = line[:-1]
            out[key] = []
        elif key is not None and ":" not in line:
            out[key].append(float(line))
        else:
            key = None
    return out


def test_example_compiles_and_fails_loudly_without_a_gpu(example_binary):
    import ctypes
    lib = ctypes.CDLL(os.path.join(LIB_DIR, "libpdlp_b200.so"))
    if lib.pdlp_b200_device_count() > 0:
        pytest.skip("a GPU is present: covered by the gpu test")
    p = subprocess.run([example_binary], capture_output=True, text=True)
    assert p.returncode == 1
    assert "Solve not successful. Status: TERMINATION_REASON_OTHER" in p.stdout
    assert "no usable CUDA device" in p.stdout


@pytest.mark.gpu
def test_example_solves_the_sample_lp(example_binary, b200_backend):
    p = subprocess.run([example_binary], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "Solve successful" in p.stdout
    assert "Solution type: POINT_TYPE_AVERAGE_ITERATE" in p.stdout or "Solution type: POINT_TYPE_CURRENT_ITERATE" in p.stdout
    s = _sections(p.stdout)
    assert s["Primal solution"] == pytest.approx([-1, 8, 1, 2.5], abs=1e-4)
    assert s["Dual solution"] == pytest.approx([-2, 0, 2.375, 2.0 / 3.0], abs=1e-4)
    line = [l for l in p.stdout.splitlines() if l.startswith("Primal objective: ")][0]
    assert float(line.split(": ")[1]) == pytest.approx(-34.0, abs=1e-4)


def test_io_header_round_trips_formats_from_cpp(tmp_path):
    """include/pdlp_b200_io.hpp (host-only entry points) compiled with -Wall -Werror and exercised by
    tests/cpp/io_roundtrip.cc: parameter text in / out, MPS and MPModelProto files, a SolveLog in JSON."""
    import json
    if not os.path.exists(os.path.join(LIB_DIR, "libpdlp_b200.so")):
        import __graft_entry__
        __graft_entry__.build()
    exe = str(tmp_path / "io_roundtrip")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "io_roundtrip.cc"),
                           "-L" + LIB_DIR, "-lpdlp_b200", "-Wl,-rpath," + LIB_DIR, "-o", exe])
    p = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    out = p.stdout
    assert "verbosity=0 limit=100 case=9 eps=0.0001" in out          # merged onto a preset verbosity_level of 2
    text = out.split("params-text-begin\n")[1].split("params-text-end")[0]
    from google.protobuf import text_format
    from ortools_b200 import pdlp_proto
    msg = text_format.Parse(text, pdlp_proto.PrimalDualHybridGradientParamsProto())
    assert msg.termination_criteria.iteration_limit == 100 and msg.termination_criteria.simple_optimality_criteria.eps_optimal_relative == 1e-4
    assert not msg.HasField("verbosity_level")                        # back at its default: not serialised
    assert 'no field named "no_such_field"' in out
    assert "mps: n=2 m=1 nnz=2 name=tiny var1=y con0=cover c=[1, 2] lc=1 uv=4" in out
    assert "Invalid filename suffix" in out
    log = json.loads(out.split("log-json-begin\n")[1].split("log-json-end")[0])
    assert log["instanceName"] == "tiny" and log["terminationReason"] == "TERMINATION_REASON_OPTIMAL" and log["iterationCount"] == 12
    assert log["params"]["terminationCriteria"]["iterationLimit"] == 100
    assert "response bytes=" in out and "response bytes=0" not in out  # MPSOLVER_MODEL_INVALID, made without a device


def test_triplet_helpers_from_cpp(tmp_path):  # quadratic_program_test.cc:558-628
    exe = str(tmp_path / "triplets")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "triplets.cc"),
                           "-L" + LIB_DIR, "-lpdlp_b200", "-Wl,-rpath," + LIB_DIR, "-o", exe])
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0 and "triplets ok" in p.stdout, (p.returncode, p.stdout, p.stderr)
