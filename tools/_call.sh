mkdir -p gpurun_out
timeout 150 python tools/attn_sweep.py > gpurun_out/attn_sweep8.log 2>&1; echo "sweep exit=$?"; grep "two-pass:\|two-pass vs" gpurun_out/attn_sweep8.log | cut -c1-120
timeout 300 python -m pytest tests/test_gpu_acoustic.py -m gpu -x -q -p no:cacheprovider > gpurun_out/acoustic6.log 2>&1; echo "acoustic exit=$?"; tail -2 gpurun_out/acoustic6.log
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_i.json 2> gpurun_out/bench_c4_i.err; echo "bench c4 exit=$?"; cut -c1-160 gpurun_out/bench_c4_i.json; grep -o '"breakdown_ms_per_step".*' gpurun_out/bench_c4_i.json | cut -c1-400
