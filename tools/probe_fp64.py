"""FP64 FMA rates of the device (development aid): shared-operand peak vs distinct-operand rate."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import thirring2d_b200 as tb

if __name__ == "__main__":
    with tb.Context(16, 16, 1, tb.MODE_ADJOINT, m=0.1, mu=0.0, stream=torch.cuda.current_stream().cuda_stream) as ctx:
        for kind in (0, 1, 0, 1):
            print(f"kind {kind}: {ctx.measure_fp64_rate(kind, 5):.2f} TFLOP/s", flush=True)
