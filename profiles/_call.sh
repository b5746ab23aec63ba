set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_ragged_gpu.py -q --timeout 700 -x > gpurun_out/r2/tests_i.log 2>&1; echo "ragged rc=$?"
grep -v "Warning\|^$\|warnings.warn" gpurun_out/r2/tests_i.log | tail -25
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_merge_options_gpu.py tests/test_tiles_gpu.py tests/test_tta_gpu.py -q --timeout 500 > gpurun_out/r2/tests_i2.log 2>&1; echo "tests rc=$?"
grep -v "Warning\|^$\|warnings.warn" gpurun_out/r2/tests_i2.log | tail -8
for v in 44 43 24 25 26; do echo "== TB200_SV_MERGE=$v"; TB200_SV_MERGE=$v python profiles/bench_configs.py cfg5_table_weight cfg5 cfg1 2>&1 | python -c "import json,sys; d=json.load(sys.stdin); print(d['cfg5_table_weight']['slice_ms'], d['cfg5_table_weight']['merge_ms'], d['cfg5_table_weight']['merge_frac'], 'sep', d['cfg5']['merge_ms'], 'cfg1', d['cfg1']['slice_ms'], d['cfg1']['merge_ms'], d['cfg5']['identity_roundtrip_ok'])"; done
python bench.py --config cfg2 --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2', d['ms_per_step'], d['roofline']['ms'], d['roofline_slice']['ms'], d['e2e']['ms_per_step'], d['e2e_labels']['ms_per_step'])"
