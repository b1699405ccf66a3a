#!/bin/bash
# build and run the pipe-rate / latency microbenchmarks (tools/ubench.cu) on the GPU box
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/ubench tools/ubench.cu && gpurun_out/ubench | tee gpurun_out/ubench.log
