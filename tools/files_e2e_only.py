"""Times only the files_e2e leg of bench.py (WAV files -> AudioToken.encode_batch_files -> .npy) a few times."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

ctx = bench.Ctx()
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    r = bench.run_files_e2e(ctx, 1250)
    print(json.dumps({k: r[k] for k in ('value', 'wall_s', 'windows', 'batches', 'host_seconds')}), flush=True)
