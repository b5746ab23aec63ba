set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_ragged_gpu.py -q --timeout 700 -x > gpurun_out/r2/ragged_tests.log 2>&1; echo "ragged rc=$?"
tail -30 gpurun_out/r2/ragged_tests.log
timeout 900 python -m pytest tests/test_lovasz_gpu.py tests/test_losses_extra_gpu.py tests/test_losses_gpu.py tests/test_joint_labels_gpu.py -q --timeout 600 > gpurun_out/r2/tests_e.log 2>&1; echo "tests rc=$?"
tail -30 gpurun_out/r2/tests_e.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_losses_extra_gpu.py -q --timeout 500 > gpurun_out/r2/memcheck_e.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/r2/memcheck_e.log
for c in cfg2 cfg1 cfg5 cfg3; do timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2/bench_${c}_a.json 2> gpurun_out/r2/bench_${c}_a.err; echo "bench $c rc=$?"; cat gpurun_out/r2/bench_${c}_a.json; tail -3 gpurun_out/r2/bench_${c}_a.err; done
