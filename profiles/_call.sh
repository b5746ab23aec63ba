# last GPU call of the round: parity + timing of the shared-memory separable gaussian (cfg 5), ncu launch list and full capture of cfg 2
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2/final
mkdir -p $O
timeout 300 python -m pytest tests/test_merge_options_gpu.py tests/test_full_size_gpu.py tests/test_engine_gpu.py -q --timeout 250 -k "single_view or cfg5 or analytic or cfg1" > $O/tests_single_view.log 2>&1; echo "single-view tests rc=$?"
tail -2 $O/tests_single_view.log
timeout 200 python profiles/bench_configs.py cfg5 cfg5_table_weight cfg1 > $O/configs_smem_gauss.json 2>/dev/null; echo "rc=$?"; cat $O/configs_smem_gauss.json
TB200_SV_GAUSS=1 timeout 200 python profiles/bench_configs.py cfg5 > $O/configs_register_gauss.json 2>/dev/null; cat $O/configs_register_gauss.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_cfg2.csv python bench.py --steps 2 --warmup 1 > $O/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:'merge_all_tma|slice_augment_fast' --launch-skip 2 -c 2 -o /tmp/prof_cfg2 -f python profiles/prof_tiles.py 2 > $O/ncu_cfg2.log 2>&1; echo "ncu cfg2 rc=$?"
ncu -i /tmp/prof_cfg2.ncu-rep --page raw --csv > $O/prof_cfg2.raw.csv 2>/dev/null
timeout 200 ncu --set full --clock-control none -k regex:'merge_single_view|slice_u8c1' --launch-skip 4 -c 2 -o /tmp/prof_cfg5 -f python profiles/bench_configs.py cfg5 > $O/ncu_cfg5.log 2>&1; echo "ncu cfg5 rc=$?"
ncu -i /tmp/prof_cfg5.ncu-rep --page raw --csv > $O/prof_cfg5.raw.csv 2>/dev/null
timeout 300 python bench.py --config cfg5 --steps 10 --warmup 3 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "bench cfg5 rc=$?"; cut -c1-400 $O/bench_cfg5.json
du -sh gpurun_out
