# usage: bash profiles/_call_multi.sh <N> [configs...]      (run under gpurun --gpus N)
set -x
N=$1; shift
CFGS=${@:-cfg2 cfg4 cfg5}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
nvidia-smi -L | head -8
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_ragged_gpu.py tests/test_full_size_gpu.py -q --timeout 500 -k "ragged or two_ranks" > gpurun_out/r2/ragged_n2.log 2>&1; echo "ragged rc=$?"
  grep -v "Warning\|^$" gpurun_out/r2/ragged_n2.log | tail -8
fi
run() {  # run <config> <tag> : one torchrun bench
  c=$1; tag=$2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config $c --steps 20 --warmup 3 > gpurun_out/r2/scale_${c}_n${N}${tag}.json 2> gpurun_out/r2/scale_${c}_n${N}${tag}.err; echo "bench $c N=$N $tag rc=$?"
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2/scale_${c}_n${N}${tag}.json') if l.startswith('{')][-1])
print('$c$tag', 'N', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'phases', d.get('phases_ms_max_over_ranks'), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline_step']['frac'],3))
" || tail -5 gpurun_out/r2/scale_${c}_n${N}${tag}.err
}
for c in $CFGS; do
  run $c _fused
  if [ "$c" != "cfg5" ] && [ "$AB" = "1" ]; then TB200_RAGGED_FUSED=0 run $c _unfused; fi
done
