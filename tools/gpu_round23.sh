#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_feasibility_polishing.py tests/test_problem_io.py tests/test_solver_goldens.py -m gpu -x -q > gpurun_out/pytest_gpu23.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu23.log; tail -30 gpurun_out/pytest_gpu23.log
