"""BASELINE config 5 alone (5M Gaussians, 3840x2160, distCUDA2 init, densify / prune events): bench.py's stress_c5 block as a
stand-alone run.  usage: python tools/run_c5.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

sys.argv = [sys.argv[0]]
print(json.dumps(bench.stress_c5_block(bench.parse(), torch.device("cuda", 0))))
