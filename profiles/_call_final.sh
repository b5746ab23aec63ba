# final single-GPU evidence run of the round: full -m gpu suite, every bench line, ncu launch list + full captures
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2/final
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2/final/gpu_tests.log 2>&1; echo "tests rc=$?"
grep -v "Warning\|^$\|warnings.warn\|tensor_from_rgb" gpurun_out/r2/final/gpu_tests.log | tail -12
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for c in cfg2 cfg1 cfg3 cfg4 cfg5; do timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2/final/bench_${c}.json 2> gpurun_out/r2/final/bench_${c}.err; echo "bench $c rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2/final/bench_${c}.json').read().strip().splitlines()[-1])
print('$c', round(d['value'],1), d['unit'], 'ms', round(d['ms_per_step'],4), 'roofline', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'cpu', round(d['cpu_baseline']['value'],3))"; done
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r2/final/bench_reference_cfg2.json 2> gpurun_out/r2/final/bench_reference_cfg2.err; echo "reference rc=$?"; cut -c1-400 gpurun_out/r2/final/bench_reference_cfg2.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/final/launches_bench_cfg2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2/final/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'merge_all_tma|slice_augment_fast' --launch-skip 6 -c 2 -o gpurun_out/r2/final/prof_cfg2 -f python bench.py --steps 2 --warmup 1 > gpurun_out/r2/final/ncu_cfg2.log 2>&1; echo "ncu cfg2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'single_view|u8c1' --launch-skip 6 -c 2 -o gpurun_out/r2/final/prof_cfg5 -f python bench.py --config cfg5 --steps 2 --warmup 1 > gpurun_out/r2/final/ncu_cfg5.log 2>&1; echo "ncu cfg5 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'joint_' --launch-skip 9 -c 3 -o gpurun_out/r2/final/prof_cfg3 -f python bench.py --config cfg3 --steps 10 --warmup 3 > gpurun_out/r2/final/ncu_cfg3.log 2>&1; echo "ncu cfg3 rc=$?"
timeout 600 python profiles/bench_losses.py > gpurun_out/r2/final/losses.json 2>/dev/null; timeout 600 python profiles/bench_lovasz.py > gpurun_out/r2/final/lovasz.json 2>/dev/null
ls -la gpurun_out/r2/final | head -30
