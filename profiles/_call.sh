set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
python profiles/calib_bw.py > gpurun_out/r2/calib_bw.json 2>&1; cat gpurun_out/r2/calib_bw.json
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r2/gpu_tests_b.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/r2/gpu_tests_b.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'merge_all' --launch-skip 3 -c 2 -o gpurun_out/r2/prof_cfg5_merge -f python profiles/bench_configs.py cfg5_table_weight > gpurun_out/r2/ncu_cfg5m.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/r2/ncu_cfg5m.log
