#!/bin/bash
# Same-box A/B of two trees: the repo (new) and a copy of an older commit under _ab_old/ (built there), alternating.
one() { (cd $1 && python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null) | python -c "
import json,sys
d=json.loads(sys.stdin.read())
kb=d['kernel_ms_breakdown']
print('$2', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {k: kb[k]['ms'] for k in ('gemm','gemm_skinny','colsum','attn_bwd_0','attn_bwd_1','attn_fwd_1')})
"; }
one . new; one _ab_old old; one . new; one _ab_old old
