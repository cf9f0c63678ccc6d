"""What the build produced, checked without a GPU: libpdlp_b200.so carries sm_100a machine code (not
PTX for a later JIT, not another architecture) for every kernel family of the hot path, and the
TMA-staged SpMV variant really is TMA (`UBLKCP` bulk copies completing on `SYNCS` mbarriers --
the mnemonics /opt/skills/guides/B200_PROFILING.md names)."""
import os
import re
import shutil
import subprocess

import pytest

from ortools_b200 import pdlp

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump is not installed")


def run(*args):
    return subprocess.run([CUOBJDUMP, *args, pdlp.library_path()], capture_output=True, text=True, timeout=600).stdout


def test_every_embedded_cubin_is_sm_100a_and_there_is_no_ptx():
    elfs = re.findall(r"ELF file\s+\d+:\s+(\S+)", run("--list-elf"))
    assert len(elfs) >= 2 and all(name.endswith(".sm_100a.cubin") for name in elfs), elfs
    assert "PTX file" not in run("--list-ptx")   # nothing is left to a driver JIT


def test_the_hot_path_kernels_are_in_the_library():
    text = run("-elf")
    kernels = set(re.findall(r"\.text\._ZN9pdlp_b2007kernels\d+(k_[a-z_0-9]+?)(?:I|E)", text))
    for name in ("k_sell", "k_sell_fixup", "k_sell_tma", "k_primal_step", "k_reduce", "k_tr_solve", "k_peer_loop", "k_peer_barrier",
                 "k_kty_finish_peer", "k_sum_push_barrier", "k_flush_average", "k_mp_primal", "k_mp_dual", "k_mp_decide"):
        assert name in kernels, (name, sorted(kernels))


def test_the_tma_staged_variant_uses_bulk_copies_and_mbarriers():
    text = run("-elf")
    symbol = re.search(r"\.text\.(_ZN9pdlp_b2007kernels10k_sell_tmaILi0ELi2ENS0_7DualEpiELi128ELi2ELi2E\w+)", text)
    assert symbol, "k_sell_tma<dot, DualEpi, 128, 2, 2> is not in the library"
    sass = run("-sass", "-fun", symbol.group(1))
    assert "UBLKCP" in sass and "SYNCS" in sass
    # the default register-staged kernel has neither: its streams are ordinary coalesced loads
    symbol = re.search(r"\.text\.(_ZN9pdlp_b2007kernels6k_sellILi0ELi2ENS0_7DualEpiELi128ELi8E\w+)", text)
    assert symbol
    sass = run("-sass", "-fun", symbol.group(1))
    assert "UBLKCP" not in sass and "LDG" in sass
