"""GPU tests of the reference-shaped Python surface (dfd.dropin, dfd.pipeline, dfd.train_fusion) against the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _bc(head, name="tiny-hd64"):
    from dfd import dropin
    from oracle import siglip_ref as R

    c = R.CONFIGS[name]
    sd = R.init_state_dict(c, 0)
    hs = R.init_head(head, c.hidden_size, 1)
    m = dropin.BinaryClassifier(device=DEV, head=head, arch=name, max_batch=8)
    ck = {"backbone.vision_model." + k: v for k, v in sd.items()}
    ck.update(hs)
    ck["backbone.text.whatever"] = torch.zeros(4)
    m.load_state_dict(ck, strict=True)
    return m, c, sd, hs


@pytest.mark.parametrize("head,eps", [("A", 0.0), ("B", 1e-6)])
def test_binary_classifier_forward(head, eps):
    from oracle import siglip_ref as R

    m, c, sd, hs = _bc(head)
    assert m.resolution == c.image_size and set(m.classifier.state_dict()) >= {"0.weight", "2.weight", "5.weight"}
    x = R.preprocess_u8(R.synthetic_images(5, c.image_size, 3))
    z = m(x).cpu()
    pooled = R.siglip_vision_forward(sd, c, x, "fp32")["pooler_output"]
    z_ref = R.classifier_head(hs, head, pooled, eps)
    assert (z - z_ref).abs().max() < 1e-2 * max(1.0, float(z_ref.abs().max())), (z, z_ref)
    if head == "B":  # off-size input -> F.interpolate default (nearest) inside the model
        xs = R.preprocess_u8(R.synthetic_images(3, 40, 4))
        z2 = m(xs).cpu()
        p2 = R.siglip_vision_forward(sd, c, R.resize_input(xs, c.image_size, "nearest"), "fp32")["pooler_output"]
        assert (z2 - R.classifier_head(hs, "B", p2, 1e-6)).abs().max() < 2e-2


def test_fast_binary_classifier_cifake_path():
    """BASELINE config 4: 32x32 images, bilinear align_corners=False upsampling inside the model, head H-D."""
    from dfd import dropin
    from oracle import siglip_ref as R

    c = R.CONFIGS["tiny-hd64"]
    sd = R.init_state_dict(c, 0)
    for size in ("small", "large"):
        hd = R.init_head_d(size, c.hidden_size, 4)
        m = dropin.FastBinaryClassifier(size, DEV, arch="tiny-hd64", max_batch=8)
        ck = {"backbone.vision_model." + k: v for k, v in sd.items()}
        ck.update(hd)
        m.load_state_dict(ck)
        img = R.synthetic_images(7, 32, 8)
        for x in (img, R.preprocess_u8(img)):  # u8 NHWC and normalised f32 NCHW
            z = m(x).cpu()
            xr = R.resize_input(R.preprocess_u8(img), c.image_size, "bilinear")
            z_ref = R.classifier_head_d(hd, R.siglip_vision_forward(sd, c, xr, "fp32")["pooler_output"])
            assert (z - z_ref).abs().max() < 2e-2 * max(1.0, float(z_ref.abs().max())), (size, z, z_ref)


def test_run_inference_and_prototypes():
    from dfd import dropin
    from oracle import siglip_ref as R

    m, c, sd, hs = _bc("A")
    imgs = R.preprocess_u8(R.synthetic_images(12, c.image_size, 5))
    labels = torch.tensor([0, 1] * 6)
    loader = [(imgs[i:i + 4], labels[i:i + 4], [f"f{j}.png" for j in range(i, i + 4)]) for i in range(0, 12, 4)]
    lab, probs, files = dropin.run_inference(m, loader, torch.device(DEV))
    pooled = R.siglip_vision_forward(sd, c, imgs, "fp32")["pooler_output"]
    p_ref = torch.sigmoid(R.classifier_head(hs, "A", pooled, 0.0)).numpy()
    assert lab.tolist() == labels.tolist() and files[5] == "f5.png"
    assert np.abs(probs - p_ref).max() < 5e-3
    _, pinv, _ = dropin.run_inference(m, loader, torch.device(DEV), invert_logits=True)
    assert np.abs(pinv - (1 - p_ref)).max() < 5e-3
    protos = dropin.few_shot_prototype(m, loader, torch.device(DEV))
    f = R.l2_normalize(pooled)
    pr = torch.stack([f[labels == 0].mean(0), f[labels == 1].mean(0)])
    pr = pr / pr.norm(dim=-1, keepdim=True)
    assert (protos["real"].cpu() - pr[0]).abs().max() < 5e-3 and (protos["fake"].cpu() - pr[1]).abs().max() < 5e-3
    _, pp, _ = dropin.run_inference(m, loader, torch.device(DEV), prototypes=protos)
    pp_ref = R.prototype_prob(f, protos["real"].cpu(), protos["fake"].cpu()).numpy()
    assert np.abs(pp - pp_ref).max() < 5e-3


def test_hf_shaped_vision_model():
    from dfd import dropin
    from oracle import siglip_ref as R

    c = R.CONFIGS["tiny-hd72"]
    sd = R.init_state_dict(c, 0)
    m = dropin.SiglipVisionModel.from_state_dict({"vision_model." + k: v for k, v in sd.items()}, DEV, num_heads=2)
    assert m.config.hidden_size == c.hidden_size
    x = R.preprocess_u8(R.synthetic_images(2, c.image_size, 0))
    o = m(pixel_values=x)
    ref = R.siglip_vision_forward(sd, c, x, "fp32")
    assert R.cosine_report(o.pooler_output.cpu(), ref["pooler_output"])["cos_min"] >= 0.999
    assert o.last_hidden_state.shape == (2, c.tokens, c.hidden_size)
    # per-layer hidden states (SigLIP2_MTL taps, Siglip2sidafrozen.py:787-793)
    oh = m(pixel_values=x, output_hidden_states=True)
    rh = R.siglip_vision_forward(sd, c, x, "fp32", output_hidden_states=True)["hidden_states"]
    assert len(oh.hidden_states) == c.num_hidden_layers + 1 == len(rh)
    for a, b in zip(oh.hidden_states, rh):
        rep = R.cosine_report(a.cpu().reshape(-1, c.hidden_size), b.reshape(-1, c.hidden_size))
        assert rep["cos_min"] >= 0.998 and rep["rel_l2"] <= 0.03, rep
    assert torch.equal(oh.pooler_output, o.pooler_output)


def test_import_shims():
    from dfd import dropin
    from oracle import scoring_ref as S

    dropin.install_import_shims()
    import open_clip
    import pywt

    model, _, pre = open_clip.create_model_and_transforms("tiny-hd64", pretrained="webli", device=DEV)
    assert model.embed_dim == 128 and callable(pre)
    from PIL import Image

    t = pre(Image.fromarray(np.zeros((50, 70, 3), np.uint8)))
    assert t.shape == (3, 64, 64) and float(t.max()) == -1.0
    x = np.random.default_rng(0).random((16, 16)).astype(np.float32)
    cA, (cH, cV, cD) = pywt.dwt2(x, "db1")
    for a, b in zip((cA, cH, cV, cD), S.haar2(x)):
        assert np.array_equal(a, b)


def test_detection_pipeline_host_buffers_vs_oracle():
    """The public e2e call (pinned host buffers in, packed score records out) against the oracle stack."""
    from dfd import pipeline, scoring
    from oracle import scoring_ref as S
    from oracle import siglip_ref as R

    name = "tiny-hd72"
    c = R.CONFIGS[name]
    sd, hs = R.init_state_dict(c, 0), R.init_head("B", c.hidden_size, 1)
    cuts = [-1.0, -0.2, 0.3, 1.5]
    st = scoring.ScoringStack(DEV, S.init_freq_mlp_g2(2), S.init_fusion_g2(3), cuts, 1.1)
    pipe = pipeline.DetectionPipeline(name, sd, hs, st, device=0, max_batch=4)
    img = R.synthetic_images(6, c.image_size, 7)
    gray = np.stack([S.gray256_from_rgb_u8(im.numpy(), True) for im in img])
    rec = pipe.detect(img.pin_memory(), torch.from_numpy(gray).pin_memory())
    assert rec.shape == (6, len(pipeline.PACKED_FIELDS))
    col = {k: i for i, k in enumerate(pipeline.PACKED_FIELDS)}
    pooled = R.siglip_vision_forward(sd, c, R.preprocess_u8(img), "fp32")["pooler_output"]
    z_sig = R.classifier_head(hs, "B", pooled, 1e-6).numpy()
    feats = np.stack([S.extract_freq_vector(g) for g in gray])
    z_freq = S.freq_mlp_g2(S.init_freq_mlp_g2(2), feats)
    z = S.fusion_g2(S.init_fusion_g2(3), z_freq, z_sig)
    d = S.detect_scores(z, np.array(cuts, np.float32), 1.1)
    # Logit gate.  BASELINE asks for 1e-2 at the named (768/1152-wide) architectures — checked in
    # tests/test_engine_gpu.py::test_base_224_logits.  On this 144-wide toy model bf16 noise averages over 5-8x fewer
    # channels: the reference's own bf16-autocast path is already up to 8e-3 from fp32 here, so the bound is 2e-2
    # (max) with the median held at 5e-3.
    pooled_ac = R.siglip_vision_forward(sd, c, R.preprocess_u8(img), "autocast")["pooler_output"]
    z_sig_ac = R.classifier_head(hs, "B", pooled_ac, 1e-6).numpy()
    for zr in (z_sig, z_sig_ac):
        e = np.abs(rec[:, col["z_sig"]] - zr)
        assert e.max() < 2e-2 and np.median(e) < 5e-3, e
    assert np.abs(rec[:, col["z_freq"]] - z_freq).max() < 1e-3
    assert np.abs(rec[:, col["z"]] - z).max() < 2e-2
    assert np.abs(rec[:, col["p_blend"]] - d["p_blend"]).max() < 5e-3
    tp = S.coral_transition_points(np.array(cuts, np.float32))
    near = np.abs(d["z_scaled"][:, None] - tp[None]).min(1) < 1e-2
    assert np.array_equal(rec[:, col["risk_idx"]].astype(int)[~near], d["risk_idx"][~near])


def test_detect_core_multicrop_vs_oracle():
    """`detect_core(pil, multicrop=True)` for a batch of images in one call vs the reference flow restated with the
    oracle: 6 crops, weights on logits, G1 shipped heads (z-scored features), temperature, CORAL."""
    from PIL import Image

    from dfd import dropin, pipeline, scoring
    from oracle import scoring_ref as S
    from oracle import siglip_ref as R
    from tests.conftest import SIGLIP_ARTEFACTS

    name = "tiny-hd64"
    c = R.CONFIGS[name]
    sd, hs = R.init_state_dict(c, 0), R.init_head("B", c.hidden_size, 1)
    st = scoring.ScoringStack.from_dir(SIGLIP_ARTEFACTS, DEV)
    pipe = pipeline.DetectionPipeline(name, sd, hs, st, device=0, max_batch=16)
    rng = np.random.default_rng(3)
    pils = [Image.fromarray(np.clip(rng.normal(120, 50, (90 + 20 * i, 130, 3)), 0, 255).astype(np.uint8)) for i in range(2)]
    res = pipe.detect_core(pils, multicrop=True, clahe=False)
    assert len(res) == 2 and set(res[0]) >= {"z_sig", "z_freq", "z_scaled", "p_fake_raw", "p_fake_coral", "p_blend",
                                             "risk_idx", "risk_probs", "entropy", "visual_prob", "freq_prob"}
    import json
    from safetensors.torch import load_file

    fm, fu = load_file(f"{SIGLIP_ARTEFACTS}/freq_mlp.safetensors"), load_file(f"{SIGLIP_ARTEFACTS}/fusion_head.safetensors")
    cuts = S.coral_cut_logits(json.load(open(f"{SIGLIP_ARTEFACTS}/coral_cutpoints.json")))
    temp = json.load(open(f"{SIGLIP_ARTEFACTS}/coral_temp.json"))["temperature"]
    pre = dropin.make_preprocess(c.image_size, "bilinear")
    w = np.array(pipeline.DetectionPipeline.MULTICROP_WEIGHTS)
    for pil, r in zip(pils, res):
        crops = pipe.make_multicrops(pil)
        x = torch.stack([pre(v) for v in crops])
        pooled = R.siglip_vision_forward(sd, c, x, "fp32")["pooler_output"]
        z_sig = float((R.classifier_head(hs, "B", pooled, 1e-6).numpy() * w).sum())
        feats = np.stack([S.extract_freq_vector(S.gray256_from_rgb_u8(np.asarray(v), False), zscore=True) for v in crops])
        z_freq = float((S.freq_mlp_g1(fm, feats) * w).sum())
        z = S.fusion_g1(fu, np.array([z_sig]), np.array([z_freq]))
        d = S.detect_scores(z, cuts, temp)
        assert abs(r["z_sig"] - z_sig) < 2e-2 and abs(r["z_freq"] - z_freq) < 2e-3
        assert abs(r["z_scaled"] - float(d["z_scaled"][0])) < 1e-2 and abs(r["p_blend"] - float(d["p_blend"][0])) < 5e-3
        assert r["risk_idx"] == int(d["risk_idx"][0])
        assert abs(r["visual_prob"] - 1 / (1 + np.exp(-z_sig))) < 5e-3


def _reference_fit(z_freq, z_sig, labels, batch_size, epochs, seed):
    """train_fusion_head_only.py:402-447 restated with torch autograd under enable_grad."""
    import torch.nn as nn
    import torch.nn.functional as F
    from torch.utils.data import DataLoader, TensorDataset

    class Head(nn.Module):
        def __init__(self):
            super().__init__()
            self.mlp = nn.Sequential(nn.Linear(3, 32), nn.GELU(), nn.Linear(32, 2))
            self.T = nn.Parameter(torch.tensor(1.0))

        def forward(self, zf, zs):
            w = F.softmax(self.mlp(torch.stack([zf, zs, (zf - zs).abs()], -1)), -1)
            return (w[..., 0] * zf + w[..., 1] * zs) / (self.T + 1e-6)

    torch.manual_seed(seed)
    h = Head()
    loader = DataLoader(TensorDataset(torch.stack([z_freq, z_sig], 1), labels), batch_size=batch_size, shuffle=True)
    opt = torch.optim.AdamW(h.parameters(), lr=5e-4)
    with torch.enable_grad():
        for _ in range(epochs):
            for xb, yb in loader:
                opt.zero_grad()
                loss = nn.BCEWithLogitsLoss()(h(xb[:, 0], xb[:, 1]), yb)
                loss.backward()
                nn.utils.clip_grad_norm_(h.parameters(), max_norm=5.0)
                opt.step()
    return torch.cat([p.detach().reshape(-1) for p in (h.mlp[0].weight, h.mlp[0].bias, h.mlp[2].weight, h.mlp[2].bias, h.T)])


def test_fit_fusion_head_matches_reference_loop():
    from dfd import train_fusion

    rng = np.random.default_rng(0)
    n = 203  # ragged last mini-batch
    y = torch.from_numpy((rng.random(n) > 0.5).astype(np.float32))
    zs = torch.from_numpy((rng.normal(0, 2, n) + 2.0 * (y.numpy() - 0.5)).astype(np.float32))
    zf = torch.from_numpy((rng.normal(0, 2, n) + 1.0 * (y.numpy() - 0.5)).astype(np.float32))
    ref = _reference_fit(zf, zs, y, 32, 3, seed=11)
    torch.manual_seed(11)
    head, best, auc = train_fusion.fit_fusion_head(zf, zs, y, batch_size=32, epochs=3, device=DEV, verbose=False)
    got = head.flat.data.cpu()
    assert (got - ref).abs().max() < 2e-4, (got - ref).abs().max()
    assert set(best) == {"mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias", "temp.T"}
    assert tuple(best["mlp.0.weight"].shape) == (32, 3) and best["temp.T"].shape == ()
    assert auc > 0.6


def test_train_fusion_head_end_to_end(tmp_path):
    """Entry point #2 on synthetic image folders: extraction -> training -> fusion_head.safetensors in the
    reference layout, loadable by the G2 scoring stack."""
    from PIL import Image
    from safetensors.torch import load_file, save_file

    from dfd import scoring, train_fusion
    from oracle import scoring_ref as S
    from oracle import siglip_ref as R

    rng = np.random.default_rng(1)
    for cls, shift in (("real", 0), ("fake", 60)):
        os.makedirs(tmp_path / cls)
        for i in range(6):
            arr = np.clip(rng.normal(110 + shift, 40, (48 + 8 * i, 64, 3)), 0, 255).astype(np.uint8)
            Image.fromarray(arr).save(tmp_path / cls / f"{i}.png")
    c = R.CONFIGS["tiny-hd64"]
    ck = {"backbone.vision_model." + k: v.contiguous() for k, v in R.init_state_dict(c, 0).items()}
    ck.update({k: v.contiguous() for k, v in R.init_head("B", c.hidden_size, 1).items()})
    save_file(ck, str(tmp_path / "best_model.safetensors"))
    save_file({k: v.contiguous() for k, v in S.init_freq_mlp_g2(2).items()}, str(tmp_path / "freq_mlp.safetensors"))
    out = tmp_path / "fusion_head.safetensors"
    torch.manual_seed(0)
    best = train_fusion.train_fusion_head(str(tmp_path / "real"), str(tmp_path / "fake"), str(tmp_path / "best_model.safetensors"),
                                          str(tmp_path / "freq_mlp.safetensors"), str(out), batch_size=4, epochs=2,
                                          device=DEV, arch="tiny-hd64")
    sd = load_file(str(out))
    assert {k: tuple(v.shape) for k, v in sd.items()} == {"mlp.0.weight": (32, 3), "mlp.0.bias": (32,), "mlp.2.weight": (2, 32),
                                                        "mlp.2.bias": (2,), "temp.T": ()}
    assert all(torch.equal(sd[k], best[k]) for k in sd)
    st = scoring.ScoringStack(DEV, load_file(str(tmp_path / "freq_mlp.safetensors")), sd, [-1.0, -0.2, 0.3, 1.5], 1.0)
    assert st.gen == 2


def _reference_fit_freq(features, labels, batch_size, epochs, lr, seed):
    """"FreqMLP trainer.py":330-396 restated with torch autograd, dropout off (the only stochastic part)."""
    import torch.nn as nn
    import torch.nn.functional as F
    from torch.utils.data import DataLoader, TensorDataset

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.norm, self.fc1, self.fc2 = nn.LayerNorm(24), nn.Linear(24, 64), nn.Linear(64, 24)

        def forward(self, x):
            return self.fc2(F.gelu(self.fc1(self.norm(x)))) + x

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.alpha, self.beta = nn.Parameter(torch.ones(24)), nn.Parameter(torch.zeros(24))
            self.gates = nn.Parameter(torch.zeros(4))
            self.blocks = nn.ModuleList([Block(), Block()])
            self.head = nn.Linear(24, 1)
            self.T = nn.Parameter(torch.tensor(1.0))

        def forward(self, x, mean, std):
            x = torch.tanh(self.alpha * ((x - mean) / (std + 1e-6)) + self.beta)
            g = torch.sigmoid(self.gates)
            x = torch.cat([c * g[i] for i, c in enumerate(torch.split(x, 6, dim=-1))], dim=-1)
            for b in self.blocks:
                x = b(x)
            return self.head(x).squeeze(-1) / (self.T + 1e-6)

    torch.manual_seed(seed)
    m = Net()
    mean, std = features.mean(0), features.std(0) + 1e-6
    loader = DataLoader(TensorDataset(features, labels), batch_size=batch_size, shuffle=True)
    opt = torch.optim.AdamW(m.parameters(), lr=lr)
    with torch.enable_grad():
        for _ in range(epochs):
            for xb, yb in loader:
                opt.zero_grad()
                nn.BCEWithLogitsLoss()(m(xb, mean, std), yb).backward()
                nn.utils.clip_grad_norm_(m.parameters(), max_norm=5.0)
                opt.step()
    ps = [m.alpha, m.beta, m.gates]
    for b in m.blocks:
        ps += [b.norm.weight, b.norm.bias, b.fc1.weight, b.fc1.bias, b.fc2.weight, b.fc2.bias]
    ps += [m.head.weight, m.head.bias, m.T]
    return torch.cat([p.detach().reshape(-1) for p in ps])


def test_fit_freq_mlp_matches_reference_loop(tmp_path):
    from safetensors.torch import load_file, save_file

    from dfd import scoring, train_freq

    rng = np.random.default_rng(4)
    n = 101  # ragged last mini-batch
    y = torch.from_numpy((rng.random(n) > 0.5).astype(np.float32))
    feats = torch.from_numpy((rng.normal(0.2, 0.7, (n, 24)) + 0.6 * (y.numpy()[:, None] - 0.5)).astype(np.float32))
    ref = _reference_fit_freq(feats, y, 8, 3, 1e-3, seed=5)
    torch.manual_seed(5)
    model, best, auc = train_freq.fit_freq_mlp(feats, y, epochs=3, batch_size=8, lr=1e-3, device=DEV, dropout=0.0,
                                               verbose=False)
    got = model.flat.data.cpu()
    assert (got - ref).abs().max() < 5e-4, (got - ref).abs().max()
    assert auc > 0.6
    # the written file loads, strictly, into the inference-side FreqMLP (the reference's reader: train_fusion_head_only.py:390-392)
    path = str(tmp_path / "freq_mlp.safetensors")
    save_file({k: v.contiguous() for k, v in best.items()}, path)
    fm = scoring.FreqMLP()
    fm.load_state_dict(load_file(path), strict=True)
    z = fm(feats.to(DEV))
    model.load_state_dict(best)
    assert (z - model(feats.to(DEV))).abs().max() < 1e-4


def test_segformer_decoder_matches_reference_golden(golden_decoder):
    """dfd.mtl.SegFormerStrongDecoder vs the reference's own class (fp32) on its golden inputs: bf16 GEMM chain."""
    from dfd import mtl
    from oracle import decoder_ref as D

    g = golden_decoder
    C, K, E, grid, S, B = (int(v) for v in g["dims"])
    dec = mtl.SegFormerStrongDecoder(K, E, DEV).load_state_dict(D.init_decoder_state(C, K, E, 3), prefix="decoder.")
    hs = [torch.from_numpy(h).reshape(B * grid * grid, C).to(DEV, torch.bfloat16) for h in g["hidden"]]
    out = dec(hs, B, grid, S).cpu().numpy()
    ref = g["seg"]
    # five bf16 GEMMs deep; logits of magnitude ~0.35: 1 % of the range
    assert np.abs(out - ref).max() < 0.01 * max(1.0, float(np.abs(ref).max()) / 0.35), np.abs(out - ref).max()
    assert np.corrcoef(out.ravel(), ref.ravel())[0, 1] > 0.9995


def test_siglip2_mtl_end_to_end():
    """SigLIP2_MTL (encoder taps + 3-class head + decoder) vs the fp32 oracle chain at a small architecture."""
    from dfd import mtl
    from oracle import decoder_ref as D
    from oracle import siglip_ref as R

    name = "small-hd72"                      # 210 px, 15 x 15 tokens, hidden 288, 3 layers
    c = R.CONFIGS[name]
    seg_layers, E = (0, 1, -1), 64
    sd = {"encoder.vision_model." + k: v for k, v in R.init_state_dict(c, 0).items()}
    sd.update(D.init_decoder_state(c.hidden_size, len(seg_layers), E, seed=4))
    model = mtl.SigLIP2_MTL(name, 0, max_batch=2, seg_layers=seg_layers, embed_dim=E).load_state_dict(sd)
    img = R.synthetic_images(2, c.image_size, 3)
    cls, seg = model(img.to(DEV))
    o = R.siglip_vision_forward(R.init_state_dict(c, 0), c, R.preprocess_u8(img), "fp32", output_hidden_states=True)
    hs = o["hidden_states"]
    feats = [hs[i + 1 if i >= 0 else len(hs) - 1] for i in seg_layers]
    seg_ref = D.decoder_forward(sd, feats, 15, c.image_size)
    cls_ref = D.cls_head(sd, o["pooler_output"])
    torch.cuda.synchronize()
    assert tuple(seg.shape) == (2, 1, 210, 210) and tuple(cls.shape) == (2, 3)
    assert (cls.cpu() - cls_ref).abs().max() < 2e-2 * max(1.0, float(cls_ref.abs().max()))
    scale = float(seg_ref.abs().max())
    assert (seg.cpu() - seg_ref).abs().max() < 0.03 * max(scale, 0.3), ((seg.cpu() - seg_ref).abs().max(), scale)
    assert np.corrcoef(seg.cpu().numpy().ravel(), seg_ref.numpy().ravel())[0, 1] > 0.999


# ---------------------------------------------------------------------------------------------------------------------
# The reference's own flow on the GPU.  /root/reference does not exist on the GPU box, so the two model classes are
# restated here the way the reference declares them (same submodule names, same forward; the source-extracted originals
# run against the same shims in tests/test_dropin_cpu.py) and driven like its entry points drive them.
# ---------------------------------------------------------------------------------------------------------------------
def _reference_style_classifier(kind, arch_name, dim, img_size):
    """kind 'A': inference_ai_human_images.py:111-152; kind 'B': train_fusion_head_only.py:78-109."""
    import open_clip  # the dfd stand-in (install_import_shims)
    import torch.nn as nn

    class BinaryClassifierA(nn.Module):
        def __init__(self, device):
            super().__init__()
            self.resolution = img_size
            self.backbone, _, self.preprocess = open_clip.create_model_and_transforms(arch_name, pretrained="webli", device=device)
            self.classifier = nn.Sequential(nn.LayerNorm(dim), nn.Dropout(0.3), nn.Linear(dim, dim // 2), nn.GELU(),
                                            nn.Dropout(0.2), nn.Linear(dim // 2, 1))

        def forward(self, x):
            features = self.backbone.encode_image(x)
            features = features / features.norm(dim=-1, keepdim=True)
            return self.classifier(features).squeeze(-1)

    class BinaryClassifierB(nn.Module):
        def __init__(self, device):
            super().__init__()
            self.backbone, _, _ = open_clip.create_model_and_transforms(arch_name, pretrained="webli", device=device)
            self.se = nn.Sequential(nn.Linear(dim, dim // 16), nn.ReLU(), nn.Linear(dim // 16, dim), nn.Sigmoid())
            self.classifier = nn.Sequential(nn.LayerNorm(dim), nn.Dropout(0.3), nn.Linear(dim, dim // 2), nn.GELU(),
                                            nn.Dropout(0.2), nn.Linear(dim // 2, dim // 4), nn.GELU(), nn.Linear(dim // 4, 1))

        def forward(self, x):
            with torch.no_grad():
                if x.shape[-1] != img_size:
                    x = nn.functional.interpolate(x, size=(img_size, img_size))
                f = self.backbone.encode_image(x)
                f = f / (f.norm(dim=-1, keepdim=True) + 1e-6)
            return self.classifier(f * self.se(f)).squeeze(-1)

    return BinaryClassifierA if kind == "A" else BinaryClassifierB


def _timm_checkpoint(sd, hs):
    """`best_model` checkpoint as the reference's trainers write it: open_clip/timm backbone names + head + text tower."""
    from dfd import dropin
    from dfd.engine import canonicalize_state_dict

    ck = {"backbone." + k: v.clone() for k, v in dropin.timm_state_from_canonical(canonicalize_state_dict(sd)).items()}
    ck.update({k: v.clone() for k, v in hs.items()})
    ck["backbone.logit_scale"], ck["backbone.logit_bias"] = torch.tensor(2.3), torch.tensor(-10.0)
    ck["backbone.text.token_embedding.weight"] = torch.zeros(4, 8)
    return ck


@pytest.mark.parametrize("kind,eps", [("A", 0.0), ("B", 1e-6)])
def test_reference_style_module_strict_load_inference_mode_autocast(kind, eps):
    """Entry-point flow of inference_ai_human_images.py (:836-857 load, :265-309 loop) / train_fusion_head_only.py
    (:118-124, :339-347): nn.Module that owns the tower as `self.backbone`, strict load of a timm-named checkpoint,
    `inference_mode()` + `autocast()`, sigmoid, `.cpu()`.  The logits must match the oracle ON THE CHECKPOINT'S weights
    (and be far from what the tower's initial random weights give)."""
    import warnings

    from dfd import dropin
    from oracle import siglip_ref as R

    dropin.install_import_shims()
    name = "small-hd72"
    c = R.CONFIGS[name]
    sd, hs = R.init_state_dict(c, 0), R.init_head(kind, c.hidden_size, 1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = _reference_style_classifier(kind, name, c.hidden_size, c.image_size)(DEV).to(DEV)
    assert model.backbone.weights_source == "random"
    x = R.preprocess_u8(R.synthetic_images(6, c.image_size, 3))
    model.eval()
    with torch.inference_mode(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        z_random = model(x.to(DEV)).float().cpu()
    model.load_state_dict(_timm_checkpoint(sd, hs), strict=True)
    assert model.backbone.weights_source == "checkpoint"
    with torch.inference_mode():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            f = model.backbone.encode_image(x.to(DEV))
            assert f.dtype == torch.bfloat16           # the tower honours the autocast dtype like a torch module would
            z = model(x.to(DEV))
        probs = torch.sigmoid(z).float().cpu().numpy()
        assert model.backbone.encode_image(x.to(DEV)).dtype == torch.float32
    pooled = R.siglip_vision_forward(sd, c, x, "fp32")["pooler_output"]
    z_ref = R.classifier_head(hs, kind, pooled, eps)
    # the head runs as torch modules under autocast here (bf16 GEMMs, as in the reference): 3e-2 on O(1) logits
    assert (z.float().cpu() - z_ref).abs().max() < 3e-2 * max(1.0, float(z_ref.abs().max())), (z, z_ref)
    assert np.abs(probs - torch.sigmoid(z_ref).numpy()).max() < 1e-2
    assert (z_random - z_ref).abs().max() > 0.1, "the strict load must have replaced the random backbone"
    # in-place parameter updates reach the engine too (version counters), not only load_state_dict
    with torch.no_grad():
        model.backbone.visual.trunk.pos_embed.mul_(0.0)
    with torch.inference_mode():
        z2 = model(x.to(DEV)).float().cpu()
    sd2 = dict(sd)
    sd2["embeddings.position_embedding.weight"] = torch.zeros_like(sd["embeddings.position_embedding.weight"])
    z2_ref = R.classifier_head(hs, kind, R.siglip_vision_forward(sd2, c, x, "fp32")["pooler_output"], eps)
    assert (z2 - z2_ref).abs().max() < 3e-2 * max(1.0, float(z2_ref.abs().max()))


def test_reference_style_module_survives_torch_compile():
    """inference_ai_human_images.py:868-881 wraps the model in torch.compile; the tower's forward is opaque to dynamo
    (torch.compiler.disable) and keeps running on the dfd engine, the torch head around it is compiled."""
    import warnings

    from dfd import dropin
    from oracle import siglip_ref as R

    dropin.install_import_shims()
    name = "tiny-hd64"
    c = R.CONFIGS[name]
    sd, hs = R.init_state_dict(c, 0), R.init_head("A", c.hidden_size, 1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = _reference_style_classifier("A", name, c.hidden_size, c.image_size)(DEV).to(DEV).eval()
    model.load_state_dict(_timm_checkpoint(sd, hs), strict=True)
    x = R.preprocess_u8(R.synthetic_images(4, c.image_size, 3)).to(DEV)
    with torch.inference_mode():
        z_eager = model(x).float().cpu()
    try:
        compiled = torch.compile(model)
        with torch.inference_mode():
            z_c = compiled(x).float().cpu()
    except Exception as e:  # the box may lack a working inductor toolchain; the reference itself falls back to eager then
        pytest.skip(f"torch.compile unavailable here: {type(e).__name__}: {str(e)[:80]}")
    assert (z_c - z_eager).abs().max() < 1e-3, (z_c, z_eager)


def test_run_tta_inference_fused_flip_equals_two_host_passes(tmp_path):
    """run_tta_inference (inference_ai_human_images.py:321-360) with the default 2 transforms: the fused pass (mirrored view
    produced by the patch kernel from the resident pixels) equals the reference's way — a second decode + host flip —
    bit for bit, and the averaged probabilities match the oracle."""
    from PIL import Image

    from dfd import dropin
    from oracle import siglip_ref as R

    m, c, sd, hs = _bc("A")
    rng = np.random.default_rng(5)
    rows = []
    for i in range(7):
        arr = np.clip(rng.normal(128, 60, (40 + 3 * i, 50 + 2 * i, 3)), 0, 255).astype(np.uint8)
        Image.fromarray(arr).save(os.path.join(tmp_path, f"img{i}.png"))
        rows.append(f"img{i}.png,{i % 2}")
    rows.append("missing.png,1")   # rows without a file are dropped (:168-169)
    csv = os.path.join(tmp_path, "meta.csv")
    open(csv, "w").write("file_name,label\n" + "\n".join(rows) + "\n")
    tfs = dropin.create_tta_transforms(c.image_size, 2)
    assert [n for n, _ in tfs] == ["Original", "H-Flip"]
    y, p_avg, per, files = dropin.run_tta_inference(m, tmp_path, csv, tfs, batch_size=3, num_workers=0, device=torch.device(DEV))
    assert y.tolist() == [0, 1, 0, 1, 0, 1, 0] and files[3] == "img3.png" and len(per) == 2
    # the reference's way: one pass per transform, flip done by the host transform
    ref_passes = []
    for _, tf in tfs:
        ds = dropin.AIHumanDataset(tmp_path, csv, transform=tf)
        loader = torch.utils.data.DataLoader(ds, batch_size=3, shuffle=False)
        ref_passes.append(dropin.run_inference(m, loader, torch.device(DEV))[1])
    assert np.array_equal(per[0], ref_passes[0]) and np.array_equal(per[1], ref_passes[1])
    assert np.array_equal(p_avg, np.mean(ref_passes, axis=0))
    assert np.abs(per[0] - per[1]).max() > 1e-4      # the mirrored view really is a different input
    # oracle on the host-transformed tensors
    ds0 = dropin.AIHumanDataset(tmp_path, csv, transform=tfs[1][1])
    xf = torch.stack([ds0[i][0] for i in range(len(ds0))])
    pr = torch.sigmoid(R.classifier_head(hs, "A", R.siglip_vision_forward(sd, c, xf, "fp32")["pooler_output"], 0.0)).numpy()
    assert np.abs(per[1] - pr).max() < 5e-3
    # three transforms: the third (CLAHE) runs as an ordinary extra pass
    y3, p3, per3, _ = dropin.run_tta_inference(m, tmp_path, csv, dropin.create_tta_transforms(c.image_size, 3), 4, 0, torch.device(DEV))
    assert len(per3) == 3 and np.array_equal(per3[0], per[0]) and np.array_equal(per3[1], per[1])
    assert np.allclose(p3, np.mean(per3, axis=0))


@pytest.mark.parametrize("fmt", ["u8", "f32"])
@pytest.mark.parametrize("S,P,Hin", [(64, 16, 64), (60, 14, 60), (60, 14, 37)])
def test_patchify_flip_equals_flipped_input(fmt, S, P, Hin):
    """DFD_FLIP_H: mirrored read of the source image == patchify of the host-flipped image, bit for bit, on the fast
    row kernel (u8, on-grid), the generic kernel (f32) and through the in-model bilinear resample (off-grid input)."""
    from dfd import ops
    from oracle import siglip_ref as R

    img = R.synthetic_images(3, Hin, seed=S + Hin)
    src = img if fmt == "u8" else R.preprocess_u8(img)
    flipped = torch.flip(src, dims=[2] if fmt == "u8" else [3]).contiguous()
    mode = ops.RESIZE_NONE if Hin == S else ops.RESIZE_BILINEAR
    a = ops.patchify(src.to(DEV), S, P, resize_mode=mode | ops.FLIP_H)
    b = ops.patchify(flipped.to(DEV), S, P, resize_mode=mode)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    assert not torch.equal(a, ops.patchify(src.to(DEV), S, P, resize_mode=mode))
