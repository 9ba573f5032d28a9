#!/bin/bash
# One gpurun call: parity tests in isolated processes (a device trap must not poison later groups),
# then the micro-benchmarks.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; timeout 600 python -m pytest "$@" -q -x -p no:cacheprovider --timeout 300 > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -3 gpurun_out/$name.log; }
run k_basic tests/test_gpu_kernels.py -m gpu -k "fbank or layernorm or simt or dwconv or vq"
run k_gemm_tc tests/test_gpu_kernels.py -m gpu -k "gemm_bf16 or tensor_equals"
run k_attn tests/test_gpu_kernels.py -m gpu -k "attention_bf16"
run p_fp32 tests/test_gpu_pipeline.py -m gpu -k "fp32"
run p_rest tests/test_gpu_pipeline.py -m gpu -k "not fp32" -s
timeout 600 python tools/gpu_microbench.py > gpurun_out/microbench.log 2>&1; echo "microbench exit=$?"; tail -12 gpurun_out/microbench.log
