"""CPU: host-side logic of the drop-in modules (no GPU): state-dict names, prompt stacks, autograd
plumbing and the fused step, with the native engine replaced by the oracle-backed stand-in."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import mudpt_oracle as orc
from tests import fake_engine, golden_util as gu


@pytest.mark.parametrize("name", gu.TINY)
def test_state_dict_names_match_reference(name):
    c = gu.load(name)
    model, _ = gu.build_model(c, "cpu")
    assert set(model.state_dict().keys()) == set(c["sd"].keys())
    trainable = sorted(n for n, p in model.named_parameters() if p.requires_grad)
    assert trainable == sorted(orc.TRAINABLE)


def test_prompt_learner_forward_api():
    c = gu.load("tiny_a")
    model, _ = gu.build_model(c, "cpu")
    prompts, shared, deep, vis = model.mudpt_prompt_learner()
    rp, rs, rd, rv = orc.prompt_learner(c["sd"])
    for a, b in [(prompts, rp), (shared, rs), (deep, rd), (vis, rv)]:
        np.testing.assert_allclose(a.detach().numpy(), b.numpy(), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", gu.TINY)
@pytest.mark.parametrize("mode", ["fused", "autograd"])
def test_step_matches_reference_golden(name, mode):
    """forward_backward (fused) and forward + F.cross_entropy + backward (reference trainer flow,
    trainers/mudpt.py:249-251) reproduce the reference's logits, loss and 10 gradients."""
    c = gu.load(name)
    g = c["golden"]
    model, _ = gu.build_model(c, "cpu")
    fake_engine.attach(model, c)
    if mode == "fused":
        loss, logits = model.forward_backward(c["image"], c["labels"])
    else:
        logits = model(c["image"])
        loss = F.cross_entropy(logits, c["labels"])
        loss.backward()
    np.testing.assert_allclose(logits.detach().numpy(), g["logits"], atol=2e-4, rtol=0)
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=1e-5, atol=1e-5)
    params = dict(model.named_parameters())
    for k in orc.TRAINABLE:
        ref = torch.from_numpy(g["grad/" + k])
        if ref.numel() and float(ref.norm()) > 0:
            m = orc.metrics(params[k].grad, ref)
            assert m["cos"] > 0.99999 and m["rel_l2"] < 2e-3, (k, m)


def test_module_level_api_matches_reference():
    c = gu.load("tiny_d")
    g = c["golden"]
    model, _ = gu.build_model(c, "cpu")
    fake_engine.attach(model, c)
    prompts, shared, text_deep, t2v = model.mudpt_prompt_learner()
    f_img, v2t = model.image_encoder(c["image"], shared, t2v)
    f_txt = model.text_encoder(prompts, model.tokenized_prompts, text_deep + v2t)
    np.testing.assert_allclose(f_img.detach().numpy(), g["image_features"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(f_txt.detach().numpy(), g["text_features"], rtol=1e-4, atol=2e-5)
    # dense gradient path: d prompts flows back into ctx through the materialised prompts
    (f_txt.sum() + f_img.sum()).backward()
    assert model.mudpt_prompt_learner.ctx.grad is not None
    assert float(model.mudpt_prompt_learner.ctx.grad.abs().sum()) > 0


def test_inference_with_cached_text_features():
    c = gu.load("tiny_a")
    model, _ = gu.build_model(c, "cpu")
    fake_engine.attach(model, c)
    logits = model.inference(c["image"])
    np.testing.assert_allclose(logits.numpy(), c["golden"]["logits"], atol=2e-4, rtol=0)


def test_no_cpu_fallback():
    c = gu.load("tiny_c")
    model, _ = gu.build_model(c, "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(c["image"])


def test_error_conventions():
    from mudpt_b200 import clip
    cfg = gu.make_cfg(2, 0, "", 32)
    a = c = gu.load("tiny_a")["arch"]
    with pytest.raises(AssertionError):  # trainers/mudpt.py:52
        cfg2 = gu.make_cfg(2, 0, "", 32)
        m = clip.CLIP(*a.astuple(), gu.make_cfg(2, 2, "", 32))
        from mudpt_b200.trainers.mudpt import MuDPTPromptLearner
        MuDPTPromptLearner(cfg2, ["x"], m)
    with pytest.raises(AssertionError):  # :55 image size mismatch
        m = clip.CLIP(*a.astuple(), gu.make_cfg(2, 2, "", 32))
        from mudpt_b200.trainers.mudpt import MuDPTPromptLearner
        MuDPTPromptLearner(gu.make_cfg(2, 2, "", 64), ["x"], m)
    bad = gu.make_cfg(2, 2, "", 32)
    bad.TRAINER.NAME = "VPT"
    with pytest.raises(NotImplementedError):  # clip/model.py:433-434
        clip.CLIP(*a.astuple(), bad)


def test_cocoop_containers_match_reference_names():
    """BASELINE config 4: the plain (cfg=None) CLIP containers and the CoCoOp CustomCLIP expose the reference's
    parameter / buffer names (trainers/cocoop.py:166-174; freeze rule :221-225 -> 5 trainable tensors)."""
    from tests import golden_util as gu
    c = gu.load_cocoop("cocoop_tiny_b")
    model, _ = gu.build_cocoop_model(c, "cpu")
    assert set(model.state_dict()) == set(c["sd"])
    trainable = sorted(n for n, p in model.named_parameters() if p.requires_grad)
    assert trainable == sorted(["prompt_learner.ctx", "prompt_learner.meta_net.linear1.weight", "prompt_learner.meta_net.linear1.bias",
                                "prompt_learner.meta_net.linear2.weight", "prompt_learner.meta_net.linear2.bias"])
    import torch
    prompts = model.prompt_learner(torch.randn(3, c["arch"].embed_dim))
    assert tuple(prompts.shape) == (3, len(c["classnames"]), 77, c["arch"].transformer_width)
    # same prompts as the reference's per-image construct_prompts loop (trainers/cocoop.py:156-162)
    pl = model.prompt_learner
    bias = pl.meta_net(torch.zeros(1, c["arch"].embed_dim))
    ctx_i = (pl.ctx.unsqueeze(0) + bias.unsqueeze(1)).expand(pl.n_cls, -1, -1)
    ref0 = pl.construct_prompts(ctx_i, pl.token_prefix, pl.token_suffix)
    assert torch.equal(model.prompt_learner(torch.zeros(1, c["arch"].embed_dim))[0], ref0)


def test_fused_sgd_has_no_cpu_fallback():
    import pytest
    import torch
    from mudpt_b200.optim import FusedSGD
    p = torch.nn.Parameter(torch.randn(4))
    p.grad = torch.randn(4)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        FusedSGD([p], lr=0.1, momentum=0.9).step()


def test_umudpt_uumudpt_containers_match_reference_names():
    """same state-dict keys and trainable sets as the reference's UMuDPT / UUMuDPT CustomCLIP
    (key lists recorded in the fixtures by oracle/make_golden.py from the reference itself)."""
    from tests import golden_util as gu
    for name in gu.VARIANTS:
        c = gu.load_variant(name)
        model, _ = gu.build_variant_model(c, "cpu")
        assert sorted(model.state_dict().keys()) == [str(k) for k in c["golden"]["state_keys"]], name
        trainable = sorted("grad/" + n for n, p in model.named_parameters() if p.requires_grad)
        assert trainable == sorted(k for k in c["golden"] if k.startswith("grad/")), name


def test_variant_prompt_algebra_matches_reference():
    """The host-side prompt algebra of UMuDPT / UUMuDPT (LightTransformer mixing -> the two prompt stacks of the C ABI)
    fed through the fp32 oracle towers reproduces the REFERENCE logits of the fixtures to fp32 round-off: whatever
    the GPU tests then differ by is kernel numerics, not semantics."""
    import torch
    from mudpt_b200 import synthetic as syn
    from oracle import mudpt_oracle as orc
    from tests import golden_util as gu
    for name in gu.VARIANTS:
        c = gu.load_variant(name)
        model, _ = gu.build_variant_model(c, "cpu")
        with torch.no_grad():
            P_v, P_t = model.prompt_stacks()
        sd = {}
        for k, v in syn.synthetic_clip_state_dict(c["arch"], 0).items():
            if k.startswith("visual."):
                sd["image_encoder." + k[len("visual."):]] = v
            elif k.startswith("transformer.") or k in ("positional_embedding", "ln_final.weight", "ln_final.bias", "text_projection"):
                sd["text_encoder." + k] = v
            elif k == "logit_scale":
                sd[k] = v
        pl = model.mudpt_prompt_learner
        emb = torch.cat([pl.token_prefix, torch.zeros(pl.n_cls, pl.n_ctx, pl.ctx_dim), pl.token_suffix], dim=1)
        eot = c["tokenized"].argmax(-1).long()
        f_img = orc.vision_features_from_stack(sd, c["image"], P_v)
        f_txt = orc.text_features_from_stack(sd, emb, eot, P_t, 77, 0)
        logits, _ = orc.logits_and_loss(f_img, f_txt, sd["logit_scale"])
        assert float((logits - torch.from_numpy(c["golden"]["logits"])).abs().max()) <= 1e-4, name


def test_checkpoint_round_trip_dassl_layout(tmp_path):
    """a checkpoint written in Dassl's layout (model.pth.tar-N under <dir>/<model name>/) loads back
    through the reference-shaped `load_model` (trainers/mudpt.py:270-302): trained tensors restored, the fixed token
    vectors of the CHECKPOINT ignored (they belong to its class names, :294-298)."""
    import torch
    from mudpt_b200 import clip
    from mudpt_b200.trainers import mudpt as M
    from tests import golden_util as gu
    c = gu.load("tiny_c")

    def make(classnames):
        cfg = gu.make_cfg(c["n_ctx"], c["depth"], "", c["arch"].image_resolution)
        from mudpt_b200 import synthetic as syn
        clip_model = clip.CLIP(*c["arch"].astuple(), cfg).float()
        t = M.MuDPT.__new__(M.MuDPT)
        M.TrainerX.__init__(t, None, None, "cpu")
        t.cfg = cfg
        t.model = M.CustomCLIP(cfg, classnames, clip_model, tokenizer=syn.synthetic_tokenize)
        for n, p in t.model.named_parameters():
            if "prompt_learner" not in n:
                p.requires_grad_("visual_ctx" in n)
        t.optim = torch.optim.SGD([p for p in t.model.parameters() if p.requires_grad], lr=0.1, momentum=0.9)
        t.sched = None
        t.register_model("MultimodalDeepPromptTuning", t.model, t.optim, t.sched)
        return t

    a = make(["class 0", "class 1", "class 2"])
    with torch.no_grad():
        for n, p in a.model.named_parameters():
            if p.requires_grad:
                p.add_(torch.randn_like(p))
    a.save_model(4, str(tmp_path))
    path = tmp_path / "MultimodalDeepPromptTuning" / "model.pth.tar-5"
    assert path.exists()
    ck = torch.load(str(path), map_location="cpu", weights_only=False)
    assert set(ck) == {"state_dict", "epoch", "optimizer", "scheduler", "val_result"} and ck["epoch"] == 5
    b = make(["dog", "cat", "bird", "fish"])   # other class names: 4 instead of 3 classes
    before = b.model.mudpt_prompt_learner.token_prefix.clone()
    b.load_model(str(tmp_path), epoch=5)
    for (n, p), (_, q) in zip(a.model.named_parameters(), b.model.named_parameters()):
        if p.requires_grad:
            assert torch.equal(p, q), n
    assert torch.equal(b.model.mudpt_prompt_learner.token_prefix, before)
    import pytest
    with pytest.raises(FileNotFoundError):
        b.load_model(str(tmp_path), epoch=99)


def test_trainer_test_loop_uses_cached_text_features():
    """Dassl-shaped test() -> parse_batch_test -> model_inference on cached text features, top-1 accuracy
    accumulated on the device; a training step drops the cache."""
    from mudpt_b200.trainers import mudpt as M
    c = gu.load("tiny_a")
    model, cfg = gu.build_model(c, "cpu")
    fake_engine.attach(model, c)
    t = M.MuDPT.__new__(M.MuDPT)
    M.TrainerX.__init__(t, None, None, "cpu")
    t.cfg, t.model = cfg, model
    t.optim = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=0.01)
    t.sched = None
    t.register_model("MultimodalDeepPromptTuning", model, t.optim, t.sched)
    golden_pred = torch.from_numpy(c["golden"]["logits"]).argmax(1)
    labels = golden_pred.clone()
    labels[0] = (labels[0] + 1) % len(c["classnames"])  # one deliberate miss
    loader = [{"img": c["image"][i:i + 1], "label": labels[i:i + 1]} for i in range(c["batch"])]
    acc = t.test(loader)
    assert acc == pytest.approx(100.0 * (c["batch"] - 1) / c["batch"])
    assert t.last_test_result["total"] == c["batch"]
    assert model._cached_text_features is not None
    model.forward_backward(c["image"], c["labels"])
    assert model._cached_text_features is None
    with pytest.raises(ValueError):
        t.test()


@pytest.mark.parametrize("variant", ["UMuDPT", "UUMuDPT"])
@pytest.mark.parametrize("new_names", [["dog", "cat", "bird"], ["dog", "cat", "bird", "fish"]])
def test_variant_checkpoint_reload_with_other_class_names(tmp_path, variant, new_names):
    """Base -> new evaluation for the UMuDPT / UUMuDPT trainers: their checkpoints store the class-name token vectors
    under `umudpt_prompt_learner.*` / `uumudpt_prompt_learner.*` (trainers/umudpt.py:337-341, uumudpt.py:343-347);
    load_model must drop them whatever the class count (equal count: no silent overwrite; unequal: no size error)."""
    import importlib
    from mudpt_b200 import clip, synthetic as syn
    from mudpt_b200.trainers import mudpt as M
    mod = importlib.import_module("mudpt_b200.trainers." + variant.lower())
    arch = syn.ARCHS["tiny"] if "tiny" in syn.ARCHS else gu.load("tiny_c")["arch"]

    def make(classnames):
        cfg = gu.make_cfg(2, 2, "", arch.image_resolution)
        cfg.TRAINER["NAME"] = variant
        cfg.TRAINER[variant.upper()] = type(cfg)(N_CTX=2, CTX_INIT="", DEEP_PROMPT_DEPTH=2, PREC="fp32")
        clip_model = clip.CLIP(*arch.astuple(), cfg).float()
        t = getattr(mod, variant).__new__(getattr(mod, variant))
        M.TrainerX.__init__(t, None, None, "cpu")
        t.cfg = cfg
        t.model = mod.CustomCLIP(cfg, classnames, clip_model, tokenizer=syn.synthetic_tokenize)
        t.optim = torch.optim.SGD([p for p in t.model.parameters() if p.requires_grad], lr=0.1)
        t.sched = None
        t.register_model(t.MODEL_NAME, t.model, t.optim, t.sched)
        return t

    a = make(["class 0", "class 1", "class 2"])
    with torch.no_grad():
        for p in a.model.parameters():
            if p.requires_grad:
                p.add_(torch.randn_like(p))
    a.save_model(0, str(tmp_path))
    ck = torch.load(str(tmp_path / a.MODEL_NAME / "model.pth.tar-1"), map_location="cpu", weights_only=False)
    assert any(k.endswith("prompt_learner.token_prefix") and not k.startswith("mudpt_") for k in ck["state_dict"])
    b = make(new_names)
    pl = b.model.mudpt_prompt_learner
    prefix, suffix = pl.token_prefix.clone(), pl.token_suffix.clone()
    b.model._cached_text_features = torch.zeros(1)  # a stale evaluation cache from before the load
    b.load_model(str(tmp_path), epoch=1)
    assert torch.equal(pl.token_prefix, prefix) and torch.equal(pl.token_suffix, suffix)
    assert b.model._cached_text_features is None
    for (n, p), (_, q) in zip(a.model.named_parameters(), b.model.named_parameters()):
        if p.requires_grad:
            assert torch.equal(p, q), n


def test_cocoop_checkpoint_round_trip_and_freeze_rule(tmp_path):
    """BASELINE config 4 host logic: the CoCoOp trainer registers only the prompt learner (trainers/cocoop.py:232-236), saves it in
    Dassl's layout and reloads it for OTHER class names -- trained tensors restored, the checkpoint's token vectors ignored
    (:309-317), FileNotFoundError for a missing epoch (:300-301); the freeze rule leaves exactly the learner's 5 tensors trainable."""
    import pytest
    import torch
    from mudpt_b200 import clip
    from mudpt_b200 import synthetic as syn
    from mudpt_b200.trainers import cocoop as CO
    from mudpt_b200.trainers import mudpt as M
    from tests import golden_util as gu
    c = gu.load_cocoop("cocoop_tiny_b")

    def make(classnames):
        cfg = gu.make_cocoop_cfg(c["n_ctx"], "", c["arch"].image_resolution)
        t = CO.CoCoOp.__new__(CO.CoCoOp)
        M.TrainerX.__init__(t, None, None, "cpu")
        t.cfg = cfg
        t.model = CO.CustomCLIP(cfg, classnames, clip.CLIP(*c["arch"].astuple(), None).float(), tokenizer=syn.synthetic_tokenize)
        enabled = M.apply_freeze_rule(t.model, ("prompt_learner",))
        assert enabled == {"prompt_learner.ctx", "prompt_learner.meta_net.linear1.weight", "prompt_learner.meta_net.linear1.bias",
                           "prompt_learner.meta_net.linear2.weight", "prompt_learner.meta_net.linear2.bias"}
        t.optim = torch.optim.SGD(t.model.prompt_learner.parameters(), lr=0.1)
        t.register_model("prompt_learner", t.model.prompt_learner, t.optim, None)
        return t

    a = make(["class 0", "class 1", "class 2"])
    with torch.no_grad():
        for p in a.model.prompt_learner.parameters():
            p.add_(torch.randn_like(p))
    a.save_model(1, str(tmp_path))
    b = make(["dog", "cat", "bird", "fish"])
    prefix_before, suffix_before = b.model.prompt_learner.token_prefix.clone(), b.model.prompt_learner.token_suffix.clone()
    b.load_model(str(tmp_path), epoch=2)
    for (n, p), (_, q) in zip(a.model.prompt_learner.named_parameters(), b.model.prompt_learner.named_parameters()):
        assert torch.equal(p, q), n
    assert torch.equal(b.model.prompt_learner.token_prefix, prefix_before)
    assert torch.equal(b.model.prompt_learner.token_suffix, suffix_before)
    with pytest.raises(FileNotFoundError):
        b.load_model(str(tmp_path), epoch=99)
    b.load_model("")  # no directory: skipped, as in the reference
