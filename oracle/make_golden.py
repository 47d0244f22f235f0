"""Generates tests/golden/*.npz — run in the build container only (needs /root/reference and `transformers`).

    python oracle/make_golden.py

Backbone vectors come from the real `transformers.SiglipVisionModel` (the class the reference instantiates at
Siglip2sidafrozen.py:753) loaded with oracle.siglip_ref.init_state_dict weights.  Scoring vectors come from the
reference's OWN function/class bodies, extracted by `ast` from the source text under /root/reference and
exec'd (the scripts cannot be imported: open_clip / pywt / gradio are not installed, coral.py does not parse)
with a Haar shim standing in for `pywt.dwt2(x, 'db1')`.  Nothing from /root/reference is copied into the repo;
only numeric outputs are stored.
"""
from __future__ import annotations

import ast
import json
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import scoring_ref as S  # noqa: E402
from oracle import siglip_ref as R  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


# ---- reference extraction ---------------------------------------------------------------------------
def extract(path: str, names, namespace: dict, line_range=None):
    src = open(path, encoding="utf-8").read()
    if line_range is not None:  # coral.py has a SyntaxError in its module docstring: parse a slice only
        lines = src.splitlines()
        src = "\n".join(lines[line_range[0] - 1 : line_range[1]])
    tree = ast.parse(src)
    picked = []
    for node in tree.body:
        nm = None
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            nm = node.name
        elif isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            nm = node.targets[0].id
        if nm in names:
            picked.append(node)
    found = {getattr(n, "name", None) or n.targets[0].id for n in picked}
    missing = set(names) - found
    assert not missing, f"{path}: not found {missing}"
    mod = ast.Module(body=picked, type_ignores=[])
    exec(compile(mod, path, "exec"), namespace)
    return namespace


def pywt_shim():
    m = types.ModuleType("pywt")

    def dwt2(x, wavelet):
        assert wavelet == "db1"
        cA, cH, cV, cD = S.haar2(np.asarray(x, dtype=np.float32))
        return cA, (cH, cV, cD)

    m.dwt2 = dwt2
    return m


def base_namespace():
    import cv2
    from PIL import Image, ImageOps
    from typing import List, Sequence

    return {"torch": torch, "nn": nn, "F": F, "np": np, "math": math, "cv2": cv2, "Image": Image,
            "ImageOps": ImageOps, "pywt": pywt_shim(), "List": List, "Sequence": Sequence}


def test_images():
    """Four RGB u8 images with natural-ish spectra (gradients + texture + edges + noise), various sizes."""
    rng = np.random.default_rng(7)
    out = []
    for (h, w) in ((224, 224), (200, 300), (64, 64), (384, 384)):
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
        img = np.zeros((h, w, 3))
        for c in range(3):
            f1, f2 = rng.uniform(0.01, 0.2, 2)
            img[..., c] = (0.5 + 0.25 * np.sin(f1 * xx + c) * np.cos(f2 * yy) + 0.15 * ((xx // 16 + yy // 16) % 2)
                           + 0.1 * rng.standard_normal((h, w)) + 0.2 * (xx / w - 0.5))
        out.append(np.clip(img * 255.0, 0, 255).astype(np.uint8))
    return out


def make_scoring():
    from PIL import Image

    tr = extract(f"{REF}/train_fusion_head_only.py",
                 {"EPS", "SRM_K", "_pil_to_gray256_clahe", "fft_features", "srm_features", "extract_freq_vector",
                  "FeatureNormalizer", "ContrastScaler", "TemperatureScaler", "BandGating", "ResidualMLPBlock",
                  "FreqMLP", "AdaptiveFusionHead"}, base_namespace())
    app = base_namespace()
    app.update({"DETECT_USE_CLAHE": False, "CORAL_CUTS": json.load(open(f"{REF}/siglip/coral_cutpoints.json"))})
    extract(f"{REF}/deepfake-detector-v2/app.py",
            {"EPS", "SRM_K", "_pil_to_gray256", "fft_features", "srm_features", "extract_freq_vector", "SafeLayerNorm",
             "FreqMLP", "_logit", "CoralCalibrator"}, app)
    coral_ns = extract(f"{REF}/coral.py", {"fit_coral_cutpoints"}, base_namespace(), line_range=(296, 325))

    g = {}
    imgs = test_images()
    grays_clahe, grays_plain = [], []
    for im in imgs:
        pil = Image.fromarray(im, "RGB")
        grays_clahe.append(tr["_pil_to_gray256_clahe"](pil).numpy())
        grays_plain.append(app["_pil_to_gray256"](pil).numpy())
    gray = np.stack(grays_clahe + grays_plain)  # [8,256,256] float32, values k/255
    g["gray_u8"] = np.round(gray * 255.0).astype(np.uint8)
    assert np.array_equal(g["gray_u8"].astype(np.float32) / 255.0, gray)
    g["rgb_shapes"] = np.array([im.shape[:2] for im in imgs], dtype=np.int32)

    # features: trainer variant (raw) and app variant (z-scored), computed by the reference bodies on the
    # stored gray256 (their gray256 producer is swapped for a lookup; it is pinned separately above)
    raw, zs = [], []
    for i in range(gray.shape[0]):
        x = torch.from_numpy(gray[i].copy())
        tr["_pil_to_gray256_clahe"] = lambda pil, _x=x: _x.clone()
        app["_pil_to_gray256"] = lambda pil, _x=x: _x.clone()
        raw.append(tr["extract_freq_vector"](None).numpy())
        zs.append(app["extract_freq_vector"](None).numpy())
    g["feats_raw"] = np.stack(raw).astype(np.float32)
    g["feats_zscore"] = np.stack(zs).astype(np.float32)

    # FreqMLP G1 with the shipped weights (eval-time noise disabled by running in train mode: no dropout/BN)
    from safetensors.torch import load_file

    g1 = app["FreqMLP"]()
    g1.load_state_dict(load_file(f"{REF}/siglip/freq_mlp.safetensors"), strict=True)
    g1.train()
    with torch.no_grad():
        g["zfreq_g1"] = g1(torch.from_numpy(g["feats_zscore"])).numpy()
    # FreqMLP G2 with seeded weights
    g2 = tr["FreqMLP"]()
    g2.load_state_dict(S.init_freq_mlp_g2(2), strict=True)
    g2.eval()
    with torch.no_grad():
        g["zfreq_g2"] = g2(torch.from_numpy(g["feats_raw"])).numpy()

    # fusion heads
    rng = np.random.default_rng(11)
    zsig = rng.normal(0, 3, 64).astype(np.float32)
    zfreq = rng.normal(0, 3, 64).astype(np.float32)
    g["fuse_zsig"], g["fuse_zfreq"] = zsig, zfreq
    fus1 = load_file(f"{REF}/siglip/fusion_head.safetensors")
    lin = nn.Linear(2, 1)
    lin.load_state_dict({"weight": fus1["fc.weight"], "bias": fus1["fc.bias"]})
    with torch.no_grad():
        p_sig = torch.sigmoid(torch.from_numpy(zsig))
        p_freq = torch.sigmoid(torch.from_numpy(zfreq / 1.25))
        g["fuse_g1_z"] = lin(torch.stack([p_sig, p_freq], 1)).squeeze(1).numpy()
    fh = tr["AdaptiveFusionHead"]()
    fh.load_state_dict(S.init_fusion_g2(3), strict=True)
    with torch.no_grad():
        g["fuse_g2_z"] = fh(torch.from_numpy(zfreq), torch.from_numpy(zsig)).numpy()
    # reference training step = reference forward + BCEWithLogits + backward under enable_grad
    yb = (rng.random(64) > 0.5).astype(np.float32)
    g["fuse_y"] = yb
    with torch.enable_grad():
        fh.train()
        fh.zero_grad()
        loss = nn.BCEWithLogitsLoss()(fh(torch.from_numpy(zfreq), torch.from_numpy(zsig)), torch.from_numpy(yb))
        loss.backward()
        g["fuse_g2_loss"] = np.float32(loss.item())
        g["fuse_g2_grads"] = torch.cat([p.grad.reshape(-1) for p in fh.parameters()]).numpy()
        assert [n for n, _ in fh.named_parameters()] == ["mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias", "temp.T"]

    # CORAL: the reference calibrator on a sweep of z_scaled
    cal = app["CoralCalibrator"]()
    g["coral_cut_logits"] = cal.c.numpy()
    zsw = np.linspace(-8, 10, 721).astype(np.float32)
    idx, probs = [], []
    for z in zsw:
        i, p = cal.predict(torch.tensor(float(z)))
        idx.append(i)
        probs.append(p.numpy())
    g["coral_z"], g["coral_idx"], g["coral_probs"] = zsw, np.array(idx, np.int32), np.stack(probs)
    # script-style fitting on a seeded logit vector
    lg = torch.from_numpy(rng.normal(0, 2, 1001).astype(np.float32))
    g["fit_logits"] = lg.numpy()
    g["fit_cuts_script"] = np.array(coral_ns["fit_coral_cutpoints"](lg, torch.zeros(1001)), dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "scoring_golden.npz"), **g)
    print("scoring_golden.npz:", {k: v.shape for k, v in g.items()})


def make_gray():
    """gray256 through the reference's OWN functions (Pillow + OpenCV as installed here) on the seeded images of
    oracle/gray_ref.py: pins the restatement and is what the GPU kernels are compared with on the GPU box."""
    from PIL import Image

    from oracle import gray_ref as G

    tr = extract(f"{REF}/train_fusion_head_only.py", {"_pil_to_gray256_clahe"}, base_namespace())
    app = base_namespace()
    app.update({"DETECT_USE_CLAHE": False})
    extract(f"{REF}/deepfake-detector-v2/app.py", {"_pil_to_gray256"}, app)
    outs = []
    for (h, w, kind, seed) in G.GOLDEN_CASES:
        pil = Image.fromarray(G.synthetic_rgb(h, w, kind, seed), "RGB")
        for fn in (tr["_pil_to_gray256_clahe"], app["_pil_to_gray256"]):  # CLAHE on, CLAHE off
            gray = fn(pil).numpy()
            u8 = np.round(gray * 255.0).astype(np.uint8)
            assert np.array_equal(u8.astype(np.float32) / np.float32(255.0), gray)
            outs.append(u8)
    import PIL
    import cv2

    np.savez_compressed(os.path.join(OUT, "gray_golden.npz"), gray_u8=np.stack(outs),
                        cases=np.array([c[:2] + (c[3],) for c in G.GOLDEN_CASES], dtype=np.int32),
                        versions=np.array([f"Pillow {PIL.__version__}", f"opencv {cv2.__version__}"]))
    print("gray golden:", len(outs), "images")


PREPROCESS_CASES = [(97, 64, "noise", 5), (512, 384, "edges", 6), (300, 451, "waves", 4), (224, 224, "waves", 2)]  # gray_ref.synthetic_rgb
PREPROCESS_TRAIN_CASES = 2      # the training preprocess always emits 384 x 384 x 3 (442 KB each): first two cases only
PREPROCESS_TTA_SIZE = 96


def make_preprocess():
    """The pixels the backbone sees, through the reference's OWN transform objects on the seeded images of
    oracle/gray_ref.py: (1) `preprocess` of train_fusion_head_only.py:60-74 (apply_clahe -> Resize((384,384)) -> ToTensor ->
    Normalize), (2) the "Original" and "H-Flip" entries of create_tta_transforms (inference_ai_human_images.py:195-215) at 96 pixels.
    Stored as the u8 pixels behind the normalised tensor (the normalisation is inverted exactly and checked): what dfd_clahe_u8 +
    dfd_resize_u8 (+ the patch kernel's mirrored read) must reproduce bit for bit."""
    from PIL import Image
    from torchvision import transforms

    from oracle import gray_ref as G

    ns = base_namespace()
    ns["transforms"] = transforms
    tr = extract(f"{REF}/train_fusion_head_only.py", {"IMG_SIZE", "apply_clahe", "preprocess"}, ns)
    inf = base_namespace()
    inf["transforms"] = transforms
    extract(f"{REF}/inference_ai_human_images.py", {"create_tta_transforms"}, inf)
    tta = inf["create_tta_transforms"](PREPROCESS_TTA_SIZE, 2)
    assert [n for n, _ in tta] == ["Original", "H-Flip"] and tr["IMG_SIZE"] == 384

    def to_u8(t):
        u8 = torch.round((t * 0.5 + 0.5) * 255.0).to(torch.uint8)
        assert torch.equal((u8.float() / 255.0 - 0.5) / 0.5, t)          # ToTensor + Normalize of exactly these bytes
        return u8.permute(1, 2, 0).contiguous().numpy()                  # HWC

    train, orig = [], []
    for i, (h, w, kind, seed) in enumerate(PREPROCESS_CASES):
        pil = Image.fromarray(G.synthetic_rgb(h, w, kind, seed), "RGB")
        if i < PREPROCESS_TRAIN_CASES:
            train.append(to_u8(tr["preprocess"](pil)))
        orig.append(to_u8(tta[0][1](pil)))
        # the reference's H-Flip transform is the mirror image of its Original one, byte for byte: only the latter is stored
        assert np.array_equal(to_u8(tta[1][1](pil)), orig[-1][:, ::-1])
    np.savez_compressed(os.path.join(OUT, "preprocess_golden.npz"), train_u8=np.stack(train), tta_original_u8=np.stack(orig),
                        tta_size=np.int32(PREPROCESS_TTA_SIZE),
                        cases=np.array([c[:2] + (c[3],) for c in PREPROCESS_CASES], dtype=np.int32))
    print("preprocess golden:", len(train), "images")


def make_freq_train():
    """Loss and gradients of the reference's OWN FreqMLP class ("FreqMLP trainer.py":218-301) in eval mode (no dropout)
    under autograd, for seeded weights / features — pins oracle.scoring_ref.freq_mlp_g2_loss_and_grads."""
    ns = extract(f"{REF}/FreqMLP trainer.py",
                 {"FeatureNormalizer", "ContrastScaler", "TemperatureScaler", "BandGating", "ResidualMLPBlock", "FreqMLP"},
                 base_namespace())
    sd = S.init_freq_mlp_g2(7)
    m = ns["FreqMLP"]()
    m.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()}, strict=True)
    m.eval()
    g = torch.Generator().manual_seed(11)
    feats = torch.randn(37, 24, generator=g) * 0.8 + 0.3
    y = (torch.rand(37, generator=g) > 0.5).float()
    with torch.enable_grad():
        out = m(feats)
        loss = nn.BCEWithLogitsLoss()(out, y)
        loss.backward()
    grads = torch.cat([dict(m.named_parameters())[k].grad.reshape(-1) for k in S.FREQMLP_PARAM_ORDER])
    np.savez_compressed(os.path.join(OUT, "freq_train_golden.npz"), feats=feats.numpy(), y=y.numpy(),
                        loss=np.float64(loss.item()), grads=grads.numpy(), logits=out.detach().numpy())
    print("freq train golden: loss", float(loss), "|grad|", float(grads.norm()))


def make_decoder():
    """SegFormerStrongDecoder of the reference (Siglip2sidafrozen.py:698-742), instantiated from its own source text with
    seeded weights, on seeded hidden states: pins oracle/decoder_ref.py."""
    from oracle import decoder_ref as D

    ns = base_namespace()
    ns.update({"List": __import__("typing").List, "Tuple": __import__("typing").Tuple})
    extract(f"{REF}/Siglip2sidafrozen.py", {"LinearProj", "SegFormerStrongDecoder"}, ns)
    C, K, E, grid, S, B = 64, 4, 32, 5, 70, 2
    sd = D.init_decoder_state(C, K, E, seed=3)
    dec = ns["SegFormerStrongDecoder"]([C] * K, embed_dim=E, dropout_rate=0.0).eval()
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}, strict=True)
    g = torch.Generator().manual_seed(9)
    hs = [torch.randn(B, grid * grid, C, generator=g) for _ in range(K)]
    with torch.no_grad():
        out = dec(hs, (grid, grid), target_size=S)
    np.savez_compressed(os.path.join(OUT, "decoder_golden.npz"), hidden=torch.stack(hs).numpy(), seg=out.numpy(),
                        dims=np.array([C, K, E, grid, S, B], dtype=np.int32))
    print("decoder golden:", tuple(out.shape), float(out.abs().max()))


def make_heads():
    """Classifier heads as the reference defines them (inference_ai_human_images.py:131-138;
    train_fusion_head_only.py:84-99), fed seeded pooled embeddings."""
    g = {}
    for D in (128, 1152):
        pooled = torch.randn(6, D, generator=torch.Generator().manual_seed(5)) * 2.0
        g[f"pooled_{D}"] = pooled.numpy()
        a = nn.Sequential(nn.LayerNorm(D), nn.Dropout(0.3), nn.Linear(D, D // 2), nn.GELU(), nn.Dropout(0.2),
                          nn.Linear(D // 2, 1)).eval()
        sd = R.init_head("A", D, 1)
        a.load_state_dict({k.replace("classifier.", ""): v for k, v in sd.items()})
        se = nn.Sequential(nn.Linear(D, D // 16), nn.ReLU(), nn.Linear(D // 16, D), nn.Sigmoid()).eval()
        b = nn.Sequential(nn.LayerNorm(D), nn.Dropout(0.3), nn.Linear(D, D // 2), nn.GELU(), nn.Dropout(0.2),
                          nn.Linear(D // 2, D // 4), nn.GELU(), nn.Linear(D // 4, 1)).eval()
        sdb = R.init_head("B", D, 1)
        se.load_state_dict({k.replace("se.", ""): v for k, v in sdb.items() if k.startswith("se.")})
        b.load_state_dict({k.replace("classifier.", ""): v for k, v in sdb.items() if k.startswith("classifier.")})
        with torch.no_grad():
            f = pooled / pooled.norm(dim=-1, keepdim=True)
            g[f"zA_{D}"] = a(f).squeeze(-1).numpy()
            f2 = pooled / (pooled.norm(dim=-1, keepdim=True) + 1e-6)
            g[f"zB_{D}"] = b(f2 * se(f2)).squeeze(-1).numpy()
            protos = torch.randn(2, D, generator=torch.Generator().manual_seed(6))
            protos = protos / protos.norm(dim=-1, keepdim=True)
            g[f"protos_{D}"] = protos.numpy()
            dr, df = torch.cdist(f, protos[0:1]), torch.cdist(f, protos[1:2])
            g[f"pproto_{D}"] = torch.softmax(torch.cat([-dr, -df], 1), 1)[:, 1].numpy()
    # H-D (cifake_binary_classifier.py): LightweightAttention is the reference's own class (extracted); the 'large'
    # variant uses nn.MultiheadAttention; forward as in FastBinaryClassifier.forward:727-749
    ns = extract(f"{REF}/cifake_binary_classifier.py", {"LightweightAttention"}, base_namespace())
    D = 128
    pooled = torch.from_numpy(g["pooled_128"])
    for size in ("tiny", "small", "medium", "large"):
        sd = R.init_head_d(size, D, 4)
        ln = nn.LayerNorm(D)
        ln.load_state_dict({"weight": sd["layer_norm.weight"], "bias": sd["layer_norm.bias"]})
        att = None
        if size in ("tiny", "small"):
            att = ns["LightweightAttention"](D, num_heads=4)
            att.load_state_dict({k[len("attention."):]: v for k, v in sd.items() if k.startswith("attention.")})
        elif size == "large":
            att = nn.MultiheadAttention(D, min(8, D // 64), dropout=0.1, batch_first=True).eval()
            att.load_state_dict({k[len("attention."):]: v for k, v in sd.items() if k.startswith("attention.")})
        if size == "tiny":
            cls = nn.Sequential(nn.Dropout(0.05), nn.Linear(D, 1))
        elif size == "small":
            cls = nn.Sequential(nn.Linear(D, D // 4), nn.GELU(), nn.Dropout(0.1), nn.Linear(D // 4, 1))
        else:
            cls = nn.Sequential(nn.Linear(D, D // 2), nn.GELU(), nn.Dropout(0.1), nn.Linear(D // 2, D // 4), nn.GELU(),
                                nn.Dropout(0.05), nn.Linear(D // 4, 1))
        cls.eval()
        cls.load_state_dict({k[len("classifier."):]: v for k, v in sd.items() if k.startswith("classifier.")})
        with torch.no_grad():
            f = pooled / pooled.norm(dim=-1, keepdim=True)
            f = ln(f)
            if att is not None:
                f = f.unsqueeze(1)
                f = att(f) if size in ("tiny", "small") else att(f, f, f)[0]
                f = f.squeeze(1)
            g[f"zD_{size}_{D}"] = cls(f).squeeze(-1).numpy()
    np.savez_compressed(os.path.join(OUT, "heads_golden.npz"), **g)
    print("heads_golden.npz:", {k: v.shape for k, v in g.items()})


def make_backbone():
    from transformers import SiglipVisionConfig, SiglipVisionModel

    g = {}
    cases = [("tiny-hd64", 3), ("tiny-hd72", 2), ("small-hd72", 2), ("siglip2-base-patch16-224", 2),
             ("siglip2-so400m-patch14-384", 1)]
    for name, B in cases:
        c = R.CONFIGS[name]
        sd = R.init_state_dict(c, 0)
        hc = SiglipVisionConfig(hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                                num_hidden_layers=c.num_hidden_layers, num_attention_heads=c.num_attention_heads,
                                image_size=c.image_size, patch_size=c.patch_size)
        m = SiglipVisionModel(hc).eval()
        m.load_state_dict({"vision_model." + k: v for k, v in sd.items()}, strict=True)
        x = R.preprocess_u8(R.synthetic_images(B, c.image_size, 0))
        with torch.no_grad():
            o = m(pixel_values=x, output_hidden_states=True)
        g[name + "/pooled"] = o.pooler_output.numpy()
        g[name + "/last_hidden_sub"] = o.last_hidden_state[:, ::7, ::5].numpy()
        g[name + "/hidden1_sub"] = o.hidden_states[1][:, ::7, ::5].numpy()
        g[name + "/weight_checksum"] = np.float64(sum(float(v.double().sum()) for v in sd.values()))
        print(name, "pooled", o.pooler_output.shape, "|pooled|max", float(o.pooler_output.abs().max()))
    np.savez_compressed(os.path.join(OUT, "backbone_golden.npz"), **g)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    which = sys.argv[1:] or ["scoring", "heads", "backbone", "gray", "freq_train", "decoder", "preprocess"]
    if "gray" in which:
        make_gray()
    if "preprocess" in which:
        make_preprocess()
    if "freq_train" in which:
        make_freq_train()
    if "decoder" in which:
        make_decoder()
    if "scoring" in which:
        make_scoring()
    if "heads" in which:
        make_heads()
    if "backbone" in which:
        make_backbone()
