set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_ragged_gpu.py -q --timeout 700 -x > gpurun_out/r2/ragged_tests.log 2>&1; echo "ragged rc=$?"
tail -40 gpurun_out/r2/ragged_tests.log
timeout 900 python -m pytest tests/test_lovasz_gpu.py tests/test_merge_options_gpu.py tests/test_engine_gpu.py -q --timeout 600 > gpurun_out/r2/tests_d.log 2>&1; echo "tests rc=$?"
tail -30 gpurun_out/r2/tests_d.log
python profiles/bench_configs.py cfg5_table_weight cfg5 cfg1 > gpurun_out/r2/cfg15_sv2.json 2>&1; cat gpurun_out/r2/cfg15_sv2.json
timeout 600 python profiles/bench_lovasz.py > gpurun_out/r2/lovasz_bench.json 2>gpurun_out/r2/lovasz_bench.err; cat gpurun_out/r2/lovasz_bench.json; tail -3 gpurun_out/r2/lovasz_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'single_view' --launch-skip 2 -c 1 -o gpurun_out/r2/prof_sv_slice -f python profiles/bench_configs.py cfg5 > gpurun_out/r2/ncu_sv1.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'merge_single_view' --launch-skip 1 -c 1 -o gpurun_out/r2/prof_sv_merge -f python profiles/bench_configs.py cfg5 > gpurun_out/r2/ncu_sv2.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'merge_single_view' --launch-skip 1 -c 1 -o gpurun_out/r2/prof_sv_merge_tbl -f python profiles/bench_configs.py cfg5_table_weight > gpurun_out/r2/ncu_sv3.log 2>&1; echo "ncu rc=$?"
