"""Tuning helper (GPU): bench.py's memory_free_large_tables block on its own."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
rxm = bench._load("rxm", os.path.join(bench.PKG, "rxm.py"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
print(json.dumps(bench.large_table_workloads(rxm, H, torch.device("cuda", 0)), indent=1))
