# final single-GPU evidence run of the round (trimmed to the GPU-minutes left): full -m gpu suite on the final library,
# loss-kernel numbers, fresh bench records of the configs whose kernels changed, ncu launch list + full capture of cfg 2
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2/final
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 700 > $O/gpu_tests.log 2>&1; echo "tests rc=$?"
grep -v "Warning\|^$\|warnings.warn\|tensor_from_rgb" $O/gpu_tests.log | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python profiles/bench_losses.py > $O/losses_cfg3_d.json 2> $O/losses_cfg3_d.err; echo "bench_losses rc=$?"
TB200_SFOCAL_MINB=4 timeout 300 python profiles/bench_losses.py cross_entropy_focal > $O/losses_cfg3_d_sfocal_minb4.json 2>/dev/null
python - <<'EOF'
import json
for f in ("losses_cfg3_d", "losses_cfg3_d_sfocal_minb4"):
    d = json.load(open(f"gpurun_out/r2/final/{f}.json"))
    print(f, {k: (v.get("ours_ms"), v.get("frac_of_measured_peak")) for k, v in d["cases"].items()})
EOF
for c in cfg5 cfg3 cfg1; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 > $O/bench_${c}.json 2> $O/bench_${c}.err; echo "bench $c rc=$?"; python -c "
import json; d=json.loads([l for l in open('$O/bench_${c}.json') if l.startswith('{')][-1])
print('$c', round(d['value'],1), d['unit'], 'ms', round(d['ms_per_step'],4), 'roofline', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'cpu', round(d['cpu_baseline']['value'],4))"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_cfg2.csv python bench.py --steps 2 --warmup 1 > $O/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'merge_all_tma|slice_augment_fast' --launch-skip 2 -c 2 -o $O/prof_cfg2 -f python profiles/prof_tiles.py 2 > $O/ncu_cfg2.log 2>&1; echo "ncu cfg2 rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:'softmax_focal_vec4|region_stats_mc_vec4|region_grad_kernel' --launch-skip 4 -c 4 -o $O/prof_losses_d -f python profiles/prof_losses.py > $O/ncu_losses_d.log 2>&1; echo "ncu losses rc=$?"
for r in prof_cfg2 prof_losses_d; do ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null; done
ls -la $O | head -30
