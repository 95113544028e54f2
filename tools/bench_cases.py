"""BASELINE configs 1-4 (the reference's own small cases) on one GPU: prints bench.small_configs() -- device us per
update and us per iteration through the batched driver -- as JSON.  bench.py carries the same object in its line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

if __name__ == "__main__":
    print(json.dumps(bench.small_configs(), indent=1))
