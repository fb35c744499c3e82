"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/input_pipeline.npz: the cases of tests/input_cases.py run through
the pipeline the reference's yaml selects (torchvision transforms on PIL images, bicubic; see
oracle/input_pipeline_oracle.py for the provenance).  Needs torchvision + Pillow (present in this image).

    python oracle/make_input_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_case(case):
    import torchvision.transforms as T
    import torchvision.transforms.functional as F
    from PIL import Image
    from tests import input_cases as ic
    name, seed, H, W, size, mode, box, flip = case
    pil = Image.fromarray(ic.make_image(seed, H, W))
    bicubic = T.InterpolationMode.BICUBIC
    if mode == "train":
        top, left, h, w = box
        x = F.resized_crop(pil, top, left, h, w, list(size), bicubic)  # RandomResizedCrop.forward after get_params
        if flip:
            x = F.hflip(x)                                              # RandomHorizontalFlip.forward
        tail = T.Compose([T.ToTensor(), T.Normalize(ic.MEAN, ic.STD)])
        return tail(x).numpy()
    tf = T.Compose([T.Resize(max(size), interpolation=bicubic), T.CenterCrop(size), T.ToTensor(), T.Normalize(ic.MEAN, ic.STD)])
    return tf(pil).numpy()


def main():
    import PIL
    import torchvision
    from tests import input_cases as ic
    out = {"versions": np.array(f"torchvision {torchvision.__version__} Pillow {PIL.__version__}")}
    for case in ic.CASES:
        y = run_case(case)
        assert y.dtype == np.float32 and y.shape == (3, *case[4])
        if case[0] in ic.SMALL:
            out[case[0]] = y
        else:  # 224 x 224 cases: digest + a strided sample
            out[case[0] + "_sha256"] = np.array(hashlib.sha256(np.ascontiguousarray(y).tobytes()).hexdigest())
            out[case[0] + "_sample"] = y[:, ::16, ::16].copy()
    path = os.path.join(ROOT, "tests", "golden", "input_pipeline.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
