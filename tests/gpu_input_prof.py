"""Timing driver for the input-pipeline kernels (B200 box): python tests/gpu_input_prof.py [--prof]
--prof: a few launches only (for ncu -k regex:augment|resample).  Test infrastructure only."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    reps = 2 if "--prof" in sys.argv else 20
    for batch in (32, 256):
        print(json.dumps(bench.input_pipeline_bench(dev, bench._peaks(), batch=batch, reps=reps)), flush=True)


if __name__ == "__main__":
    main()
