"""Streaming-kernel variants on the same workload (development aid): register-marching, TMA-staged (whole-batch tiles
or 16-chain tiles through tensor maps), and the staged two-launch iteration with the direction update folded in.

    python tools/probe_pipe.py                     # the default list
    python tools/probe_pipe.py 256,256,64,0.01     # NT,NX,chains,m ...
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from probe import probe

# name -> environment (read at context creation / at every solve)
MODES = {
    "marching": {"TB_NO_PIPE": "1"},
    "staged": {"TB_PIPE_TILED": "0", "TB_PIPE_XPAY": "0"},
    "staged+xpay": {"TB_PIPE_TILED": "0", "TB_PIPE_XPAY": "1"},
    "tiled": {"TB_PIPE_TILED": "1", "TB_PIPE_XPAY": "0"},
    "tiled+xpay": {"TB_PIPE_TILED": "1", "TB_PIPE_XPAY": "1"},
}
KEYS = ("TB_NO_PIPE", "TB_PIPE_TILED", "TB_PIPE_XPAY")

if __name__ == "__main__":
    cfgs = [(256, 256, 64, 0.01), (2048, 2048, 1, 0.01), (128, 128, 2048, 0.01), (512, 512, 16, 0.01), (512, 512, 32, 0.01)]
    if len(sys.argv) > 1:
        cfgs = [tuple(float(v) if "." in v else int(v) for v in a.split(",")) for a in sys.argv[1:]]
    for c in cfgs:
        many = c[2] > 16
        for name, env in MODES.items():
            if name.startswith("tiled") != many and name != "marching":
                continue   # batches of up to 16 chains: whole-batch tiles; larger ones: 16-chain tiles
            for k in KEYS:
                os.environ.pop(k, None)
            os.environ.update(env)
            out = probe(*c, solver=1, reps=2, max_iter=201)
            out["mode"] = name
            print(json.dumps(out), flush=True)
