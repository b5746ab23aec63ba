set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_ragged_gpu.py tests/test_losses_gpu.py tests/test_tiles_gpu.py -q --timeout 700 > gpurun_out/r2/tests_h.log 2>&1; echo "tests rc=$?"
grep -v "Warning\|^$\|warnings.warn" gpurun_out/r2/tests_h.log | tail -15
for v in 0 6 7 8; do echo "== TB200_SV_SLICE=$v"; TB200_SV_SLICE=$v python profiles/bench_configs.py cfg5_table_weight 2>&1 | python -c "import json,sys; d=json.load(sys.stdin)['cfg5_table_weight']; print(d['slice_ms'], d['slice_frac'], d['merge_ms'], d['identity_roundtrip_ok'])"; done
for v in 2 3 4 5; do echo "== TB200_SV_MERGE=$v"; TB200_SV_MERGE=$v python profiles/bench_configs.py cfg5_table_weight cfg5 2>&1 | python -c "import json,sys; d=json.load(sys.stdin); print(d['cfg5_table_weight']['merge_ms'], d['cfg5_table_weight']['merge_frac'], 'sep', d['cfg5']['merge_ms'], d['cfg5']['identity_roundtrip_ok'])"; done
timeout 600 python profiles/bench_losses.py > gpurun_out/r2/losses_c.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2/losses_c.json'))
for k,v in d['cases'].items(): print(k, v.get('ours_ms'), v.get('frac_of_measured_peak'))"
