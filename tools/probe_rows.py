"""Streaming solver: rows marched per thread (grid size / wave quantisation) sweep (development aid)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from probe import probe

if __name__ == "__main__":
    cfgs = [(256, 256, 64), (2048, 2048, 1), (512, 512, 16), (1024, 1024, 4)]
    for nt, nx, c in cfgs:
        for rows in (16, 8, 4, 2):
            o = probe(nt, nx, c, 0.01, rows=rows, solver=1, reps=10, max_iter=301)
            print(json.dumps({k: o[k] for k in ("cfg", "apply_us", "apply_frac", "us_per_iter", "cg_frac")}), flush=True)
