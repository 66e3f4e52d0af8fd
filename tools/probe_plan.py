"""Planned vs plain launches of the on-chip solvers (development aid): same workload with and without TB_NO_PLAN."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from probe import probe

if __name__ == "__main__":
    cfgs = [(256, 256, 8, 0.01), (256, 256, 8, 0.1), (256, 256, 12, 0.05), (128, 128, 40, 0.1), (128, 128, 64, 0.1),
            (64, 64, 256, 0.1)]
    if len(sys.argv) > 1:
        cfgs = [tuple(float(v) if "." in v else int(v) for v in a.split(",")) for a in sys.argv[1:]]
    for c in cfgs:
        for no_plan in ("1", None):
            if no_plan:
                os.environ["TB_NO_PLAN"] = no_plan
            else:
                os.environ.pop("TB_NO_PLAN", None)
            out = probe(*c, reps=2)
            out["planned"] = no_plan is None
            print(json.dumps(out), flush=True)
