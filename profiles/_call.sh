set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
TB200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_experimental_gpu.py -q --timeout 300 > gpurun_out/r2/exp_tests.log 2>&1; echo "exp rc=$?"
tail -30 gpurun_out/r2/exp_tests.log
TB200_EXPERIMENTAL=1 timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_experimental_gpu.py -q --timeout 800 -x > gpurun_out/r2/exp_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -15 gpurun_out/r2/exp_memcheck.log
timeout 600 python profiles/bench_losses.py > gpurun_out/r2/losses_a.json 2> gpurun_out/r2/losses_a.err; echo "losses rc=$?"
cat gpurun_out/r2/losses_a.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'merge_all|slice_plain' --launch-skip 4 -c 4 -o gpurun_out/r2/prof_cfg5 -f python profiles/bench_configs.py cfg5 > gpurun_out/r2/ncu_cfg5.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/r2/ncu_cfg5.log
