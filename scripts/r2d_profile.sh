#!/bin/bash
# Round-2 final evidence run on one B200: GPU tests, bench, attention / GEMM micro-benchmarks, ncu launch list of the
# bench command and one `ncu --set full` launch of every hot kernel. Outputs land in gpurun_out/ (r2d_*).
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2d_gputest.log
python bench.py > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench.err
python scripts/gpu_bench_attn.py > gpurun_out/r2d_attn_standalone.jsonl 2>&1
python scripts/gpu_bench_gemm.py > gpurun_out/r2d_gemm_shapes.jsonl 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16|attn_|layernorm_|colsum' \
    -o gpurun_out/r2d_kernels -f python scripts/gpu_kernels_once.py > gpurun_out/r2d_kernels.log 2>&1
ncu -i gpurun_out/r2d_kernels.ncu-rep --page raw --csv > gpurun_out/r2d_kernels_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r2d_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-profile > gpurun_out/r2d_launches_bench.log 2>&1
tail -2 gpurun_out/r2d_gputest.log
cut -c1-400 gpurun_out/r2d_bench_n1.json
