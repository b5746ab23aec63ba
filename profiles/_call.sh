# last sanity call of the round on the final HEAD: the contract bench line (with roofline.traffic from profiles/traffic.json)
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2/final
mkdir -p $O
timeout 120 python bench.py --steps 10 --warmup 3 > $O/bench_cfg2_final.json 2> $O/bench_cfg2_final.err; echo "bench rc=$?"
python -c "
import json; d=json.loads([l for l in open('$O/bench_cfg2_final.json') if l.startswith('{')][-1])
print(round(d['value'],1), d['unit'], 'ms', round(d['ms_per_step'],4), 'roofline', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'e2e', round(d['e2e']['value'],1), 'labels', round(d['e2e_labels']['value'],1), 'cpu', round(d['cpu_baseline']['value'],3), 'eager', d.get('torch_cuda_eager_baseline',{}).get('value'), d['clocks'])"
