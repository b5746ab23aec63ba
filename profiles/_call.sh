set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_lovasz_gpu.py tests/test_losses_gpu.py -q --timeout 600 > gpurun_out/r2/lovasz_tests.log 2>&1; echo "lovasz rc=$?"
tail -40 gpurun_out/r2/lovasz_tests.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_lovasz_gpu.py -q --timeout 800 -k "variants or ties or argument" > gpurun_out/r2/lovasz_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -8 gpurun_out/r2/lovasz_memcheck.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_lovasz_gpu.py --deselect tests/test_losses_gpu.py > gpurun_out/r2/gpu_tests_c.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/r2/gpu_tests_c.log
TB200_FORCE_GENERIC_MERGE=1 python profiles/bench_configs.py cfg5_table_weight cfg5 cfg1 > gpurun_out/r2/cfg5_generic.json 2>&1; cat gpurun_out/r2/cfg5_generic.json
