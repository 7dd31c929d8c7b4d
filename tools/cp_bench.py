#!/usr/bin/env python3
"""Frame latency of the code predictor on cuda:0 (the `code_predictor` leg of bench.py on its own)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    print(json.dumps(bench.code_predictor_leg(0, frames=int(sys.argv[1]) if len(sys.argv) > 1 else 60), indent=1))
