"""Phase timeline of the fused train step (B200 box): CUDA events at the phase boundaries of CustomCLIP.forward_backward,
averaged over steps.  `python tests/gpu_step_timeline.py [classes]` (125 = the per-rank shapes of an 8-GPU run).
Diagnostic only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from mudpt_b200 import synthetic as syn
    classes = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    device = torch.device("cuda:0")
    trainer = bench.build_trainer(device, False, classes)
    model = trainer.model
    B = bench.BATCH_PER_GPU
    imgs = [syn.synthetic_images(B, 224, seed=100 + i).to(device) for i in range(4)]
    labs = [syn.synthetic_labels(B, classes, seed=100 + i).to(device) for i in range(4)]

    def step(i):
        trainer.optim.zero_grad()
        model.forward_backward(imgs[i % 4], labs[i % 4])
        trainer.optim.step()

    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    acc, n = {}, 0
    for i in range(20):
        model.__dict__["_timeline"] = []
        step(i)
        end = torch.cuda.Event(enable_timing=True)
        end.record()
        torch.cuda.synchronize()
        tl = model.__dict__["_timeline"] + [("sgd_end", end)]
        t0 = tl[0][1]
        for name, ev in tl[1:]:
            acc[name] = acc.get(name, 0.0) + t0.elapsed_time(ev)
        n += 1
    model.__dict__["_timeline"] = None
    print(f"classes {classes}: ms after the start of the step (mean of {n} steps, steps issued back to back)")
    for name, v in acc.items():
        print(f"  {name:16s} {v / n:7.3f}")


if __name__ == "__main__":
    main()
