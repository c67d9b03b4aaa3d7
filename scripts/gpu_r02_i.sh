#!/bin/bash
# GPU run I of round 2 (1 GPU, ~3 minutes of budget left): the kernels and tests written after the last full-suite run.
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 205 python -u -m pytest -v -p no:cacheprovider tests/test_gpu_resample.py "tests/test_gpu_layers.py::test_tc_layers_match_cuda_core_kernels[c20]" \
   tests/test_gpu_configs.py --durations=0 > $O/r02i_pytest.log 2>&1
echo "exit $?" >> $O/r02i_pytest.log
