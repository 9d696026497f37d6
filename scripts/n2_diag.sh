#!/bin/bash
# N=2 A/B of the gradient all-reduce variants (bench.py knobs): prints tag, pairs/s, ms/step, SM MHz, multi_gpu_check
run() { tag=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NP:-2} --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus ${NP:-2} --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-profile > gpurun_out/n2_$tag.json 2> gpurun_out/n2_$tag.err; python -c "
import json
d=json.load(open('gpurun_out/n2_$tag.json'))
c=d.get('multi_gpu_check') or {}
print('$tag', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], c.get('reduced_grads_identical_across_ranks'), c.get('reduced_vs_mean_of_local_grads_rel'))
"; }
for v in "$@"; do
  case $v in
    default) run default A=1;;
    tower) run tower OAT_LAYER_REDUCE=0;;
    noreduce) run noreduce OAT_BENCH_NO_REDUCE=1;;
    static) run static OAT_GEMM_DYNAMIC=0;;
  esac
done
