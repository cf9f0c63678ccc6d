"""Lists every TEST / TEST_P / TEST_F of the reference's ortools/pdlp/*_test.cc and flags the ones whose
line range no test under tests/ cites (``file_test.cc:LO-HI`` or a bare ``:LO-HI`` after a file was named
in the same module). A flagged test is not necessarily uncovered -- the heuristic follows citations, and a
bare range is attributed to the last file named before it -- but an unflagged one is cited somewhere.
tests/REFERENCE_TESTS.md is the hand-written table this helps to keep honest. Needs /root/reference
(so it runs in the build container, not on the GPU box).

    python tools/reference_test_map.py [--all]
"""
import collections
import glob
import os
import re
import sys

REF = "/root/reference/ortools/pdlp"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_tests():
    out = {}
    for path in sorted(glob.glob(os.path.join(REF, "*_test.cc"))):
        lines = open(path).read().split("\n")
        tests = []
        for i, line in enumerate(lines):
            m = re.match(r"^(TEST|TEST_P|TEST_F|TYPED_TEST)\((\w+),\s*(\w*)", line)
            if m:
                name = m.group(3) or re.match(r"\s*(\w+)", lines[i + 1]).group(1)
                tests.append([i + 1, m.group(2) + "." + name])
        for k, t in enumerate(tests):
            t.append(tests[k + 1][0] - 1 if k + 1 < len(tests) else len(lines))
        out[os.path.basename(path)] = tests
    return out


def citations():
    cites = collections.defaultdict(list)
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "test_*.py"))):
        current_file, current_test = None, "(module)"
        for line in open(path).read().split("\n"):
            m = re.match(r"\s*def (test_\w+)", line)
            if m:
                current_test = m.group(1)
            for m in re.finditer(r"(\w+_test\.cc)?:(\d+)(?:-(\d+))?", line):
                if m.group(1):
                    current_file = m.group(1)
                if current_file is None:
                    continue
                lo = int(m.group(2))
                hi = int(m.group(3) or lo)
                if hi >= lo:
                    cites[current_file].append((lo, hi, os.path.basename(path) + "::" + current_test))
    return cites


def main():
    if not os.path.isdir(REF):
        sys.exit("needs %s" % REF)
    show_all = "--all" in sys.argv
    cites = citations()
    uncited = 0
    for name, tests in reference_tests().items():
        print("== %s: %d tests" % (name, len(tests)))
        for lo, test, hi in tests:
            hits = sorted({c[2] for c in cites.get(name, []) if c[0] <= hi and c[1] >= lo})
            if not hits:
                uncited += 1
                print("   UNCITED %5d %s" % (lo, test))
            elif show_all:
                print("           %5d %s <- %s" % (lo, test, ", ".join(hits[:3])))
    print("uncited: %d" % uncited)


if __name__ == "__main__":
    main()
