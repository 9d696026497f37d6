#!/bin/bash
# One `ncu --set full` launch of every hot kernel at the bench geometry -> gpurun_out/r2c_kernels.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16|attn_|layernorm_|colsum' \
    -o gpurun_out/r2c_kernels -f python scripts/gpu_kernels_once.py > gpurun_out/r2c_kernels.log 2>&1
tail -2 gpurun_out/r2c_kernels.log
