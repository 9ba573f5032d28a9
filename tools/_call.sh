mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_acoustic.py -m gpu -x -q -p no:cacheprovider > gpurun_out/acoustic5.log 2>&1; echo "acoustic exit=$?"; tail -3 gpurun_out/acoustic5.log
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_h.json 2> gpurun_out/bench_c4_h.err; echo "bench c4 exit=$?"; cut -c1-200 gpurun_out/bench_c4_h.json; grep -o '"breakdown_ms_per_step".*' gpurun_out/bench_c4_h.json | cut -c1-400; grep -o '"e2e".\{0,120\}' gpurun_out/bench_c4_h.json
timeout 150 python tools/attn_sweep.py > gpurun_out/attn_sweep7.log 2>&1; echo "sweep exit=$?"; grep "two-pass:\|online:" gpurun_out/attn_sweep7.log | cut -c1-120
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "attention" > gpurun_out/k_attn7.log 2>&1; echo "k_attn exit=$?"; tail -3 gpurun_out/k_attn7.log
