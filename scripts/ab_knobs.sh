#!/bin/bash
# Same-box A/B of runtime knobs: each argument is "tag VAR=val ..." (bench.py without the CPU / e2e legs)
one() { tag=$1; shift; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
kb=d['kernel_ms_breakdown']
print('$tag', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {k: kb[k]['ms'] for k in ('gemm','gemm_skinny','colsum','attn_bwd_0','attn_bwd_1')})
"; }
one new A=1; one noident OAT_QKV_BIAS_IDENTITY=0; one new A=1; one noident OAT_QKV_BIAS_IDENTITY=0
