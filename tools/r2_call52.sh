#!/bin/bash
# Round-2 GPU call 52: issue rate of packed fp32 (FADD2 / FMUL2 / FFMA2) against the scalar instructions
set -u
mkdir -p gpurun_out
./tools/ubench/f32x2 2>&1 | tee gpurun_out/ubench_f32x2_r02.txt
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm --format=csv,noheader | tee -a gpurun_out/ubench_f32x2_r02.txt
