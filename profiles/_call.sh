# loss kernels: parity with the closed-form focal kernel (default) and the generic one, A/B timing, launch list, ncu --set full
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_losses_gpu.py tests/test_joint_labels_gpu.py tests/test_losses_extra_gpu.py tests/test_full_size_gpu.py -q --timeout 500 -k "loss or focal or cfg3 or joint" > gpurun_out/r2/tests_losses_fast.log 2>&1; echo "losses (fast focal) rc=$?"
tail -3 gpurun_out/r2/tests_losses_fast.log
TB200_FOCAL_GENERIC=1 timeout 600 python -m pytest tests/test_losses_gpu.py -q --timeout 500 > gpurun_out/r2/tests_losses_generic.log 2>&1; echo "losses (generic focal) rc=$?"
tail -3 gpurun_out/r2/tests_losses_generic.log
timeout 600 python profiles/bench_losses.py > gpurun_out/r2/losses_cfg3_c.json 2> gpurun_out/r2/losses_cfg3_c.err; echo "bench_losses rc=$?"
TB200_FOCAL_GENERIC=1 timeout 600 python profiles/bench_losses.py > gpurun_out/r2/losses_cfg3_c_generic_focal.json 2>/dev/null
python - <<'EOF'
import json
for f in ("losses_cfg3_c", "losses_cfg3_c_generic_focal"):
    d = json.load(open(f"gpurun_out/r2/{f}.json"))
    print(f, {k: (v.get("ours_ms"), v.get("frac_of_measured_peak")) for k, v in d["cases"].items()})
EOF
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches_losses.csv python profiles/prof_losses.py > gpurun_out/r2/launches_losses.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'softmax_focal_vec4|region_stats_mc|region_grad|focal_fast|focal_kernel|region_stats_elem_kernel' --launch-skip 6 -c 6 -o gpurun_out/r2/prof_losses_c -f python profiles/prof_losses.py > gpurun_out/r2/ncu_losses_c.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2 | tail -12
