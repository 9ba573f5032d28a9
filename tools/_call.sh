mkdir -p gpurun_out
timeout 120 python tools/attn_sweep.py > gpurun_out/attn_sweep4.log 2>&1; echo "sweep exit=$?"; tail -16 gpurun_out/attn_sweep4.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "attention" > gpurun_out/k_attn4.log 2>&1; echo "k_attn exit=$?"; tail -3 gpurun_out/k_attn4.log
