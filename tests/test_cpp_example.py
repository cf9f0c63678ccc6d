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
