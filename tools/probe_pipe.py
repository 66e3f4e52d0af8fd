"""TMA-staged vs register-marching streaming kernels (development aid): same workload with and without TB_NO_PIPE."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from probe import probe

if __name__ == "__main__":
    cfgs = [(256, 256, 64, 0.01), (2048, 2048, 1, 0.01), (128, 128, 2048, 0.01), (512, 512, 16, 0.01), (64, 64, 256, 0.01)]
    if len(sys.argv) > 1:
        cfgs = [tuple(float(v) if "." in v else int(v) for v in a.split(",")) for a in sys.argv[1:]]
    for c in cfgs:
        for no_pipe in ("1", None):
            if no_pipe:
                os.environ["TB_NO_PIPE"] = no_pipe
            else:
                os.environ.pop("TB_NO_PIPE", None)
            out = probe(*c, solver=1, reps=2, max_iter=201)
            out["staged"] = no_pipe is None
            print(json.dumps(out), flush=True)
