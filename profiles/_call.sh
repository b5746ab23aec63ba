set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_ragged_gpu.py -q --timeout 700 -x > gpurun_out/r2/tests_j.log 2>&1; echo "ragged rc=$?"
grep -v "Warning\|^$\|warnings.warn" gpurun_out/r2/tests_j.log | tail -25
