"""A CPU model of the synchronisation skeleton of ``k_peer_loop`` (tests/cpp/peer_protocol_model.cc)
under ThreadSanitizer: 3 "ranks" x 3 "blocks" as threads, the arenas and every vector as plain
memory, the grid barriers (ticket + generation, last block meets the other ranks) and the two
cross-rank barriers of an attempt exactly where the kernel has them, slices handed out by tickets.
It checks the argument of DESIGN.md 5 mechanically -- no write into an arena can race with a read of
its previous contents, for any ticket order -- and that the result equals a sequential run bit for
bit. It models the protocol; the CUDA code itself is exercised by the 2 / 4 / 8-GPU parity tests.
The rounds of the trust-region solve (one barrier per round, round vectors in alternating slots) follow
in the same program. Negative controls: without the cross-rank part of barrier A or B, or with a single
slot for the round vectors, the same program must fail."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "peer_protocol_model.cc")


@pytest.fixture(scope="module", autouse=True)
def tsan_runtime_starts(tmp_path_factory):
    """ThreadSanitizer refuses to start on kernels whose address-space layout it does not know
    ("unexpected memory mapping"); that is a property of the machine, not of the protocol."""
    d = tmp_path_factory.mktemp("tsan_probe")
    src, exe = str(d / "probe.cc"), str(d / "probe")
    open(src, "w").write("int main() { return 0; }\n")
    if subprocess.call(["g++", "-fsanitize=thread", src, "-o", exe], stderr=subprocess.DEVNULL) != 0:
        pytest.skip("g++ -fsanitize=thread is not usable here")
    p = subprocess.run([exe], capture_output=True, text=True)
    if p.returncode != 0:
        pytest.skip("the ThreadSanitizer runtime does not start on this machine: " + p.stderr.strip()[:200])


def build(tmp_path, mode, *defines):
    out = str(tmp_path / ("model_%d_%s" % (mode, "_".join(defines) or "full")))
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-pthread", "-fsanitize=thread", "-DMODE=%d" % mode,
                           *["-D" + d for d in defines], SRC, "-o", out])
    return out


@pytest.mark.parametrize("mode", [0, 1])  # 0: all-gather exchange (peer-d), 1: reduce-scatter exchange (peer-s)
def test_two_cross_rank_barriers_order_every_arena_access(tmp_path, mode):
    exe = build(tmp_path, mode)
    for _ in range(8):  # the ticket order differs from run to run
        p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
        assert "model ok" in p.stdout and "ThreadSanitizer" not in p.stderr


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("dropped", ["DROP_PEER_A", "DROP_PEER_B", "TR_SINGLE_SLOT"])
def test_the_model_fails_without_either_cross_rank_barrier(tmp_path, mode, dropped):
    exe = build(tmp_path, mode, dropped)
    for _ in range(6):  # (a race is reported when both accesses are still in the tool's history: allow a few schedules)
        p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
        if p.returncode != 0:
            assert "ThreadSanitizer: data race" in p.stderr or "MISMATCH" in p.stdout or "timed out" in p.stdout
            return
    pytest.fail("the model passed six times without " + dropped)
