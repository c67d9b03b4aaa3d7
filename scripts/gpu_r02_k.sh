#!/bin/bash
# GPU run K of round 2 (1 GPU, the last ~2 minutes): pipelined-predict test + resample tests + the default-flag bench line.
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 60 python -u -m pytest -q -p no:cacheprovider tests/test_gpu_predict.py::test_predict_volumes_pipelined_equals_per_image_calls tests/test_gpu_resample.py > $O/r02k_pytest.log 2>&1
echo "exit $?" >> $O/r02k_pytest.log
timeout 70 python bench.py --no-cpu > $O/r02k_bench.json 2> $O/r02k_bench.err
echo "exit $?" >> $O/r02k_bench.err
