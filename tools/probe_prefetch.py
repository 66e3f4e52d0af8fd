"""L2 prefetch distance of the marching streaming kernels (development aid): TB_PREFETCH = 0, 2, 4, 8 rows ahead on the
many-chain configurations (the staged kernels serve the few-chain ones)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from probe import probe

if __name__ == "__main__":
    cfgs = [(256, 256, 64, 0.01), (128, 128, 2048, 0.01), (64, 64, 256, 0.01), (2048, 2048, 1, 0.01)]
    os.environ["TB_NO_PIPE"] = "1"
    for c in cfgs:
        for pf in (0, 2, 4, 8):
            os.environ["TB_PREFETCH"] = str(pf)
            out = probe(*c, solver=1, reps=2, max_iter=201)
            out["prefetch_rows"] = pf
            print(json.dumps({k: out[k] for k in ("cfg", "prefetch_rows", "apply_us", "us_per_iter", "cg_frac")}), flush=True)
