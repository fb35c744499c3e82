"""Input pipeline: crop -> bicubic resample -> window -> flip -> normalize.

CPU (`-m "not gpu"`): the oracle (oracle/input_pipeline_oracle.py) against the golden fixture made from
torchvision + PIL and, when those are importable, against them live; the host-side random draws against
torchvision under the same seed; descriptor validation through the C ABI (host-only entry point).
GPU (`-m gpu`): the CUDA kernels through the C ABI against the oracle and the fixture.
Bar: BIT-EXACT everywhere (8-bit resampling is integer work; /255 and (x - mean) / std are single IEEE fp32 ops).
"""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import input_pipeline_oracle as ipo
from tests import input_cases as ic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "input_pipeline.npz")


def _oracle_case(case):
    name, seed, H, W, size, mode, box, flip = case
    img = ic.make_image(seed, H, W)
    if mode == "train":
        return ipo.train_transform(img, *box, flip, size, ic.MEAN, ic.STD)
    return ipo.eval_transform(img, size, ic.MEAN, ic.STD)


def _check_against_golden(case, y, g):
    name = case[0]
    assert y.dtype == np.float32
    if name in ic.SMALL:
        assert np.array_equal(y, g[name]), f"{name}: max abs diff {np.abs(y - g[name]).max()}"
    else:
        assert np.array_equal(y[:, ::16, ::16], g[name + "_sample"]), name
        assert hashlib.sha256(np.ascontiguousarray(y).tobytes()).hexdigest() == str(g[name + "_sha256"]), name


# ------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("case", ic.CASES, ids=[c[0] for c in ic.CASES])
def test_oracle_matches_golden(case):
    g = np.load(GOLDEN, allow_pickle=False)
    _check_against_golden(case, _oracle_case(case), g)


def test_oracle_resample_matches_pil_live():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    for it in range(40):
        H, W = (int(v) for v in rng.integers(6, 260, 2))
        img = ic.make_image(100 + it, H, W)
        ow, oh = (int(v) for v in rng.integers(3, 200, 2))
        if it % 4 == 0:
            ow = W  # horizontal pass skipped
        if it % 5 == 0:
            oh = H
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
        assert np.array_equal(ipo.resample_u8(img, ow, oh), ref), (H, W, ow, oh)


def test_oracle_pipeline_matches_torchvision_live():
    T = pytest.importorskip("torchvision.transforms")
    Image = pytest.importorskip("PIL.Image")
    bic = T.InterpolationMode.BICUBIC
    rng = np.random.default_rng(1)
    for it in range(6):
        H, W = (int(v) for v in rng.integers(40, 300, 2))
        img = ic.make_image(200 + it, H, W)
        pil = Image.fromarray(img)
        size = (56, 56)
        # training: same seed -> same draws -> same pixels
        tf = T.Compose([T.RandomResizedCrop(size, scale=(0.08, 1.0), interpolation=bic), T.RandomHorizontalFlip(),
                        T.ToTensor(), T.Normalize(ic.MEAN, ic.STD)])
        torch.manual_seed(it)
        ref = tf(pil).numpy()
        from mudpt_b200 import input_pipeline as ip
        torch.manual_seed(it)
        bx, by, bw, bh, rw, rh, wx, wy, flip = ip.draw_geometry(H, W, size, True)
        assert (rw, rh, wx, wy) == (56, 56, 0, 0)
        got = ipo.train_transform(img, by, bx, bh, bw, bool(flip), size, ic.MEAN, ic.STD)
        assert np.array_equal(got, ref)
        ev = T.Compose([T.Resize(max(size), interpolation=bic), T.CenterCrop(size), T.ToTensor(), T.Normalize(ic.MEAN, ic.STD)])
        assert np.array_equal(ipo.eval_transform(img, size, ic.MEAN, ic.STD), ev(pil).numpy())


def test_random_draws_match_torchvision():
    T = pytest.importorskip("torchvision.transforms")
    from mudpt_b200 import input_pipeline as ip
    for seed, (H, W) in enumerate([(375, 500), (500, 333), (64, 64), (10, 400), (400, 10), (1, 1)]):
        torch.manual_seed(seed)
        ref = [T.RandomResizedCrop.get_params(torch.empty(3, H, W), [0.08, 1.0], [3 / 4, 4 / 3]) for _ in range(5)]
        ref_flip = bool(torch.rand(1) < 0.5)
        torch.manual_seed(seed)
        got = [ip.random_resized_crop_params(H, W) for _ in range(5)]
        assert got == [tuple(r) for r in ref]
        assert ip.random_flip() == ref_flip
    # evaluation geometry == torchvision's Resize(int) + center_crop arithmetic
    import torchvision.transforms.functional as F
    for H, W in [(150, 233), (301, 170), (224, 224), (225, 224), (480, 640)]:
        assert list(ip.resized_output_size(H, W, 224)) == F._compute_resized_output_size((H, W), [224])
        g = ip.draw_geometry(H, W, (224, 224), False)
        nh, nw = ip.resized_output_size(H, W, 224)
        assert g == (0, 0, W, H, nw, nh, int(round((nw - 224) / 2.0)), int(round((nh - 224) / 2.0)), 0)


def test_desc_struct_and_validation_through_abi():
    """mudpt_augment_workspace_bytes is host-only: descriptor layout and error paths are checkable without a GPU."""
    from mudpt_b200 import _lib, input_pipeline as ip
    lib = _lib.load()
    assert ip.DESC_DTYPE.itemsize == 64
    d = np.zeros(2, ip.DESC_DTYPE)
    d[0] = (0x1000, 100, 200, 600, 10, 20, 150, 60, 224, 224, 0, 0, 1, (0, 0))
    d[1] = (0x2000, 500, 400, 1200, 0, 0, 400, 500, 224, 280, 0, 28, 0, (0, 0))
    need = lib.mudpt_augment_workspace_bytes(d.ctypes.data_as(C.c_void_p), 2, 224, 224)
    kmax = max(ipo.precompute_coeffs(w, 0, w, o)[0] for w, o in [(150, 224), (60, 224), (400, 224), (500, 280)])
    assert need == 2 * 2 * 224 * (8 + 4 * kmax)
    bad = d.copy()
    bad[1]["box_w"] = 401  # outside the image
    assert lib.mudpt_augment_workspace_bytes(bad.ctypes.data_as(C.c_void_p), 2, 224, 224) < 0
    assert b"crop box" in lib.mudpt_global_last_error()
    bad = d.copy()
    bad[0]["win_x"] = 1  # window leaves the resampled image
    assert lib.mudpt_augment_workspace_bytes(bad.ctypes.data_as(C.c_void_p), 2, 224, 224) < 0
    assert b"window" in lib.mudpt_global_last_error()
    bad = d.copy()
    bad[0]["pitch"] = 599
    assert lib.mudpt_augment_workspace_bytes(bad.ctypes.data_as(C.c_void_p), 2, 224, 224) < 0
    with pytest.raises(NotImplementedError):
        ip.GpuTransform(interpolation="bilinear")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ip.GpuTransform(device="cpu")


# ------------------------------------------------------------------------------------------- GPU
def _geometry_of(case):
    from mudpt_b200 import input_pipeline as ip
    name, seed, H, W, size, mode, box, flip = case
    if mode == "train":
        top, left, h, w = box
        return (left, top, w, h, size[1], size[0], 0, 0, int(flip))
    return ip.draw_geometry(H, W, size, False)


@pytest.mark.gpu
def test_gpu_cases_bit_exact_vs_golden_and_oracle():
    from mudpt_b200 import input_pipeline as ip
    g = np.load(GOLDEN, allow_pickle=False)
    by_size = {}
    for case in ic.CASES:
        by_size.setdefault(case[4], []).append(case)
    for size, cases in by_size.items():  # one ragged batch per output size
        tf = ip.GpuTransform(size=size, is_train=True, mean=ic.MEAN, std=ic.STD)
        imgs = [torch.from_numpy(ic.make_image(c[1], c[2], c[3])).cuda() for c in cases]
        out = tf(imgs, params=[_geometry_of(c) for c in cases]).cpu().numpy()
        for c, y in zip(cases, out):
            _check_against_golden(c, y, g)
            assert np.array_equal(y, _oracle_case(c)), c[0]


@pytest.mark.gpu
def test_gpu_random_ragged_batch_bit_exact():
    """Seeded random crops at 224 x 224 (the training transform as the trainer would call it), pinned host images
    uploaded by the transform, evaluation transform on the same images, a strided (pitched) source."""
    from mudpt_b200 import input_pipeline as ip
    rng = np.random.default_rng(7)
    shapes = [(int(h), int(w)) for h, w in zip(rng.integers(30, 520, 12), rng.integers(30, 520, 12))] + [(224, 224), (500, 375)]
    host = [ic.make_image(300 + i, h, w) for i, (h, w) in enumerate(shapes)]
    tf = ip.GpuTransform(size=(224, 224), is_train=True, mean=ic.MEAN, std=ic.STD)
    torch.manual_seed(3)
    geo = [ip.draw_geometry(h, w, (224, 224), True) for h, w in shapes]
    torch.manual_seed(3)
    out = tf([torch.from_numpy(a).pin_memory() for a in host]).cpu().numpy()  # draws its own (same seed)
    for a, (bx, by, bw, bh, rw, rh, wx, wy, flip), y in zip(host, geo, out):
        assert np.array_equal(y, ipo.train_transform(a, by, bx, bh, bw, bool(flip), (224, 224), ic.MEAN, ic.STD))
    big = [a for a in host if min(a.shape[:2]) >= 100]
    ev = ip.GpuTransform(size=(224, 224), is_train=False, mean=ic.MEAN, std=ic.STD)
    out = ev([torch.from_numpy(a).cuda() for a in big]).cpu().numpy()
    for a, y in zip(big, out):
        assert np.array_equal(y, ipo.eval_transform(a, (224, 224), ic.MEAN, ic.STD))
    # pitched source: a view into a wider image
    wide = torch.from_numpy(ic.make_image(77, 200, 300)).cuda()
    view = wide[10:190, 20:260]
    y = tf([view], params=[(5, 6, 200, 150, 224, 224, 0, 0, 1)]).cpu().numpy()[0]
    ref = ipo.train_transform(view.cpu().numpy(), 6, 5, 150, 200, True, (224, 224), ic.MEAN, ic.STD)
    assert np.array_equal(y, ref)


@pytest.mark.gpu
def test_gpu_extreme_scales_bit_exact():
    """Heavy down-scaling (wide filters; the narrow-band launch when the 8-row band no longer fits shared memory)
    and heavy up-scaling."""
    from mudpt_b200 import input_pipeline as ip
    tf = ip.GpuTransform(size=(224, 224), is_train=True, mean=ic.MEAN, std=ic.STD)
    for seed, (H, W), box in [(1, (1400, 1300), (0, 0, 1300, 1400)),     # scale ~6: 8-row bands
                              (2, (6400, 240), (0, 0, 240, 6400)),       # vertical scale 28.6: 2-row bands
                              (3, (40, 40), (10, 12, 9, 7)),             # up-scaling x25 / x32
                              # staged-rows kernel, horizontal filter widths 9 / 11 / 13 / 17 taps (register-weight
                              # variants and the generic loop), short crops so that the rows fit shared memory
                              (4, (160, 420), (3, 2, 400, 150)), (5, (170, 520), (11, 5, 500, 150)),
                              (6, (160, 650), (1, 0, 640, 150)), (7, (120, 900), (50, 10, 800, 100))]:
        a = ic.make_image(seed, H, W)
        bx, by, bw, bh = box
        y = tf([torch.from_numpy(a).cuda()], params=[(bx, by, bw, bh, 224, 224, 0, 0, 0)]).cpu().numpy()[0]
        assert np.array_equal(y, ipo.train_transform(a, by, bx, bh, bw, False, (224, 224), ic.MEAN, ic.STD)), (H, W)
    with pytest.raises(RuntimeError, match="crop box"):
        tf([torch.zeros(50, 50, 3, dtype=torch.uint8, device="cuda")], params=[(0, 0, 51, 50, 224, 224, 0, 0, 0)])


@pytest.mark.gpu
def test_gpu_trainer_takes_raw_images():
    """MuDPT.forward_backward({"img_u8": [...], "label": ...}): the trainer's parse_batch_train runs the transform
    named by cfg.INPUT on the GPU; the batch it feeds the model is bit-identical to the torchvision-pinned oracle
    under the same seed, and the step equals the step on that float batch."""
    from mudpt_b200 import input_pipeline as ip
    from mudpt_b200.trainers import mudpt as M
    from tests import golden_util as gu
    case = gu.load("tiny_a")
    R = case["arch"].image_resolution
    raws = [ic.make_image(400 + i, h, w) for i, (h, w) in enumerate([(40, 61), (90, 33), (R, R)][:case["batch"]])]
    assert len(raws) == case["batch"]

    def make_trainer():
        model, cfg = gu.build_model(case, "cuda")
        cfg.INPUT["INTERPOLATION"] = "bicubic"
        cfg.INPUT["PIXEL_MEAN"], cfg.INPUT["PIXEL_STD"] = list(ic.MEAN), list(ic.STD)
        cfg.INPUT["TRANSFORMS"] = ["random_resized_crop", "random_flip", "normalize"]
        t = M.MuDPT.__new__(M.MuDPT)
        M.TrainerX.__init__(t, None, None, torch.device("cuda"))
        t.cfg, t.model = cfg, model
        t.optim = M.build_optimizer(model, cfg.OPTIM)
        t.sched = M.build_lr_scheduler(t.optim, cfg.OPTIM)
        t.register_model("MultimodalDeepPromptTuning", model, t.optim, t.sched)
        t.batch_idx, t.num_batches = 0, 10 ** 9
        return t

    torch.manual_seed(11)
    geo = [ip.draw_geometry(a.shape[0], a.shape[1], (R, R), True) for a in raws]
    ref = np.stack([ipo.train_transform(a, by, bx, bh, bw, bool(f), (R, R), ic.MEAN, ic.STD)
                    for a, (bx, by, bw, bh, _, _, _, _, f) in zip(raws, geo)])
    batch_u8 = {"img_u8": [torch.from_numpy(a) for a in raws], "label": case["labels"]}
    t1 = make_trainer()
    torch.manual_seed(11)
    image, label = t1.parse_batch_train(batch_u8)
    assert np.array_equal(image.cpu().numpy(), ref)
    torch.manual_seed(11)
    loss_u8 = t1.forward_backward(batch_u8)["loss"]
    t2 = make_trainer()
    loss_f = t2.forward_backward({"img": torch.from_numpy(ref), "label": case["labels"]})["loss"]
    assert loss_u8 == loss_f and np.isfinite(loss_u8)
    for (n1, p1), (_, p2) in zip(t1.model.named_parameters(), t2.model.named_parameters()):
        if p1.requires_grad:
            assert torch.equal(p1, p2), n1


@pytest.mark.gpu
def test_gpu_trainer_prefetch_matches_plain_steps():
    """MuDPT.prefetch(next batch) + forward_backward(batch): three optimizer steps over pinned host batches give the
    same losses and bit-identical parameters as the same steps without prefetch (the upload moves to a copy stream
    under the previous step; event-ordered slot reuse with more batches than slots)."""
    from mudpt_b200.trainers import mudpt as M
    from tests import golden_util as gu
    case = gu.load("tiny_a")

    def make_trainer():
        model, cfg = gu.build_model(case, "cuda")
        t = M.MuDPT.__new__(M.MuDPT)
        M.TrainerX.__init__(t, None, None, torch.device("cuda"))
        t.cfg, t.model = cfg, model
        t.optim = M.build_optimizer(model, cfg.OPTIM)
        t.sched = M.build_lr_scheduler(t.optim, cfg.OPTIM)
        t.register_model("MultimodalDeepPromptTuning", model, t.optim, t.sched)
        t.batch_idx, t.num_batches = 0, 10 ** 9
        return t

    g = torch.Generator().manual_seed(5)
    batches = [{"img": (case["image"] + 0.1 * i * torch.randn(case["image"].shape, generator=g)).pin_memory(),
                "label": case["labels"].roll(i).pin_memory()} for i in range(5)]
    t1, t2 = make_trainer(), make_trainer()
    l1, l2 = [], []
    t1.prefetch(batches[0])
    for i, b in enumerate(batches):
        if i + 1 < len(batches):
            t1.prefetch(batches[i + 1])
        l1.append(t1.forward_backward(b)["loss"])
        l2.append(t2.forward_backward(b)["loss"])
    assert t1.__dict__["_pf"]["ready"] == {}  # every prefetched batch was consumed by its own step
    assert l1 == l2 and all(np.isfinite(l1)), (l1, l2)
    for (n1, p1), (_, p2) in zip(t1.model.named_parameters(), t2.model.named_parameters()):
        if p1.requires_grad:
            assert torch.equal(p1, p2), n1


@pytest.mark.gpu
def test_gpu_trainer_evaluates_raw_images():
    """test() over raw 8-bit images: evaluation transform on the GPU (bit-identical to the oracle), logits from the
    cached text features, accuracy accumulated on the device."""
    from mudpt_b200.trainers import mudpt as M
    from tests import golden_util as gu
    case = gu.load("tiny_d")
    R = case["arch"].image_resolution
    model, cfg = gu.build_model(case, "cuda")
    cfg.INPUT["PIXEL_MEAN"], cfg.INPUT["PIXEL_STD"] = list(ic.MEAN), list(ic.STD)
    t = M.MuDPT.__new__(M.MuDPT)
    M.TrainerX.__init__(t, None, None, torch.device("cuda"))
    t.cfg, t.model = cfg, model
    t.register_model("MultimodalDeepPromptTuning", model, None, None)
    raws = [ic.make_image(500 + i, h, w) for i, (h, w) in enumerate([(70, 90), (R, R), (120, 64), (55, 200), (99, 98)])]
    ref = torch.from_numpy(np.stack([ipo.eval_transform(a, (R, R), ic.MEAN, ic.STD) for a in raws])).cuda()
    with torch.no_grad():
        full = model(ref)  # full forward on the oracle's batch
    labels = full.argmax(1).cpu()
    labels[2] = (labels[2] + 1) % full.shape[1]
    loader = [{"img_u8": [torch.from_numpy(a) for a in raws[i:i + 2]], "label": labels[i:i + 2]} for i in range(0, 5, 2)]
    image, _ = t.parse_batch_test(loader[0])
    assert torch.equal(image, ref[:2])
    assert t.test(loader) == pytest.approx(80.0)


def test_oracle_resample_property_vs_pil():
    """Property test of the pin: random sizes (incl. 1-pixel axes, x30 down-scaling, x40 up-scaling) and crop boxes,
    the numpy restatement equals PIL.Image.resize bit for bit."""
    Image = pytest.importorskip("PIL.Image")
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    @hyp.settings(max_examples=30, deadline=None, derandomize=True)
    @hyp.given(st.integers(1, 300), st.integers(1, 300), st.integers(1, 64), st.integers(1, 64), st.integers(0, 10 ** 6))
    def check(H, W, oh, ow, seed):
        rng = np.random.default_rng(seed)
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        x0, y0 = int(rng.integers(0, W)), int(rng.integers(0, H))
        x1, y1 = int(rng.integers(x0 + 1, W + 1)), int(rng.integers(y0 + 1, H + 1))
        crop = img[y0:y1, x0:x1]
        ref = np.asarray(Image.fromarray(crop).resize((ow, oh), Image.BICUBIC))
        assert np.array_equal(ipo.resample_u8(crop, ow, oh), ref)

    check()
