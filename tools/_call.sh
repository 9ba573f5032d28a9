mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_acoustic.py -m gpu -x -q -p no:cacheprovider > gpurun_out/acoustic4.log 2>&1; echo "acoustic exit=$?"; tail -3 gpurun_out/acoustic4.log
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_g.json 2> gpurun_out/bench_c4_g.err; echo "bench c4 exit=$?"; cut -c1-200 gpurun_out/bench_c4_g.json; grep -o '"breakdown_ms_per_step".*' gpurun_out/bench_c4_g.json | cut -c1-400
timeout 300 python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_e.json 2> gpurun_out/bench_c3_e.err; echo "bench c3 exit=$?"; cut -c1-200 gpurun_out/bench_c3_e.json; grep -o '"breakdown_ms_per_step".*' gpurun_out/bench_c3_e.json | cut -c1-400
