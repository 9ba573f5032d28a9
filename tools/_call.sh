mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -p no:cacheprovider -k "not fp32" -s > gpurun_out/p_rest2.log 2>&1; echo "p_rest exit=$?"; grep -i "rel\|agree\|passed\|failed\|error" gpurun_out/p_rest2.log | cut -c1-300 | tail -30
