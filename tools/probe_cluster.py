"""Cluster solver vs streaming solver on the lattices that have a cluster shape (development aid)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from probe import probe

if __name__ == "__main__":
    cfgs = [(128, 128, 33, 0.1), (128, 128, 256, 0.1), (256, 256, 8, 0.1), (256, 256, 8, 0.01), (256, 256, 64, 0.1),
            (64, 128, 148, 0.1)]
    if len(sys.argv) > 1:
        cfgs = [tuple(float(v) if "." in v else int(v) for v in a.split(",")) for a in sys.argv[1:]]
    for c in cfgs:
        for solver in (2, 1):
            print(json.dumps(probe(*c, solver=solver, reps=3)), flush=True)
