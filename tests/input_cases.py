"""Deterministic cases for the input-pipeline tests (shared by oracle/make_input_golden.py, which runs them
through torchvision + PIL, and tests/test_input_pipeline.py).  Images are procedural (numpy PCG64), so the
fixture holds only outputs."""
import numpy as np

MEAN = (0.48145466, 0.4578275, 0.40821073)  # configs/trainers/MuDPT/vit_b16_bz4_ep10_nctx2_depth9.yaml:11-12
STD = (0.26862954, 0.26130258, 0.27577711)


def make_image(seed: int, H: int, W: int) -> np.ndarray:
    """uint8 [H, W, 3]: smooth gradients + texture + a few saturated blocks (exercises the 0 / 255 clipping of
    the bicubic overshoot)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    img = np.stack([127.5 + 127.5 * np.sin(xx / (3.0 + seed % 5) + yy / 11.0),
                    255.0 * xx / max(W - 1, 1),
                    255.0 * ((xx.astype(np.int64) // 7 + yy.astype(np.int64) // 5) % 2)], -1)
    img += rng.normal(0.0, 20.0, img.shape)
    img = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    for _ in range(3):
        y0, x0 = rng.integers(0, max(H - 4, 1)), rng.integers(0, max(W - 4, 1))
        img[y0:y0 + 9, x0:x0 + 9] = rng.choice([0, 255])
    return img


# name, seed, H, W, out size (h, w), mode, (top, left, h, w) crop box for "train", flip
CASES = [
    ("up_small", 1, 37, 53, (32, 32), "train", (3, 5, 20, 17), False),
    ("down_small", 2, 211, 173, (32, 48), "train", (10, 20, 190, 150), True),
    ("identity", 3, 64, 80, (40, 40), "train", (7, 9, 40, 40), True),       # crop == output size: both passes skipped by PIL
    ("h_only", 4, 90, 120, (40, 24), "train", (11, 3, 40, 101), False),      # vertical pass is the identity
    ("v_only", 5, 120, 70, (24, 40), "train", (2, 13, 111, 40), True),       # horizontal pass is the identity
    ("full_box", 6, 100, 75, (48, 48), "train", (0, 0, 100, 75), False),
    ("one_px_rows", 7, 60, 60, (16, 16), "train", (30, 30, 1, 2), False),   # degenerate crop
    ("eval_landscape", 8, 150, 233, (48, 48), "eval", None, False),
    ("eval_portrait", 9, 301, 170, (32, 32), "eval", None, False),
    ("eval_exact", 10, 48, 48, (48, 48), "eval", None, False),
    ("train_224", 11, 375, 500, (224, 224), "train", (40, 61, 300, 333), True),
    ("eval_224", 12, 333, 500, (224, 224), "eval", None, False),
]
SMALL = [c[0] for c in CASES if c[4][0] * c[4][1] <= 48 * 48]
