#!/bin/bash
# round-2 GPU call A: state of the tree before the kernel work -- full -m gpu suite (incl. the new paper-size gradient and
# activation-range tests) and a baseline bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi -L > gpurun_out/r02a_gpus.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_parity_gpu.py::test_split_modes_are_robust_to_input_scale --deselect tests/test_parity_gpu.py::test_split_modes_are_robust_to_residual_and_skip_magnitude > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "input_scale or residual_and_skip" > gpurun_out/r02a_pytest_envelope.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest_envelope.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -5 gpurun_out/r02a_pytest.log; tail -15 gpurun_out/r02a_pytest_envelope.log; head -c 600 gpurun_out/r02a_bench.json
