set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_ensembling_gpu.py tests/test_joint_labels_gpu.py tests/test_losses_gpu.py tests/test_merge_options_gpu.py -q --timeout 600 > gpurun_out/r2/tests_g.log 2>&1; echo "tests rc=$?"
grep -v "Warning\|^$\|warnings.warn" gpurun_out/r2/tests_g.log | tail -15
for v in 0 1 2 3 4 5; do echo "== TB200_SV_SLICE=$v"; TB200_SV_SLICE=$v python profiles/bench_configs.py cfg5_table_weight 2>&1 | python -c "import json,sys; d=json.load(sys.stdin)['cfg5_table_weight']; print(d['slice_ms'], d['slice_frac'], d['merge_ms'], d['identity_roundtrip_ok'])"; done
timeout 600 python bench.py --config cfg3 --steps 20 --warmup 3 > gpurun_out/r2/bench_cfg3_b.json 2> gpurun_out/r2/bench_cfg3_b.err; python -c "import json; d=json.load(open('gpurun_out/r2/bench_cfg3_b.json')); print('cfg3', d['ms_per_step'], d['roofline']['frac'])"
timeout 600 python profiles/bench_losses.py > gpurun_out/r2/losses_b.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2/losses_b.json'))
for k,v in d['cases'].items(): print(k, v.get('ours_ms'), v.get('frac_of_measured_peak'))"
