# validation of HEAD after the re-entry: ragged strips (fused + unfused), then the whole -m gpu suite, then the contract bench
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 800 python -m pytest tests/test_ragged_gpu.py -q --timeout 700 -x > gpurun_out/r2/tests_ragged.log 2>&1; echo "ragged rc=$?"
grep -v "Warning\|^$\|warnings.warn" gpurun_out/r2/tests_ragged.log | tail -25
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=15 --deselect tests/test_ragged_gpu.py > gpurun_out/r2/tests_gpu.log 2>&1; echo "gpu suite rc=$?"
grep -v "Warning\|^$\|warnings.warn\|tensor_from_rgb" gpurun_out/r2/tests_gpu.log | tail -30
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2/bench_cfg2_k.json 2> gpurun_out/r2/bench_cfg2_k.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/r2/bench_cfg2_k.json
