"""CPU tests (no GPU): the oracle restatements against the reference-generated golden vectors and the
reference's shipped artefacts.  These pin the oracle; the -m gpu tests then compare the CUDA path with it."""
import os

import numpy as np
import pytest
import torch

from oracle import scoring_ref as S
from oracle import siglip_ref as R


@pytest.mark.parametrize("name,B", [("tiny-hd64", 3), ("tiny-hd72", 2), ("small-hd72", 2)])
def test_backbone_oracle_matches_hf_golden(name, B, golden_backbone):
    c = R.CONFIGS[name]
    sd = R.init_state_dict(c, 0)
    assert abs(sum(float(v.double().sum()) for v in sd.values()) - float(golden_backbone[name + "/weight_checksum"])) < 1e-6
    x = R.preprocess_u8(R.synthetic_images(B, c.image_size, 0))
    o = R.siglip_vision_forward(sd, c, x, "fp32", output_hidden_states=True)
    assert np.abs(o["pooler_output"].numpy() - golden_backbone[name + "/pooled"]).max() < 5e-5
    assert np.abs(o["last_hidden_state"][:, ::7, ::5].numpy() - golden_backbone[name + "/last_hidden_sub"]).max() < 5e-5
    assert np.abs(o["hidden_states"][1][:, ::7, ::5].numpy() - golden_backbone[name + "/hidden1_sub"]).max() < 5e-5
    assert len(o["hidden_states"]) == c.num_hidden_layers + 1
    # the bf16-autocast emulation stays inside the published gate relative to fp32
    a = R.siglip_vision_forward(sd, c, x, "autocast")
    assert R.cosine_report(a["pooler_output"], o["pooler_output"])["cos_min"] > 0.9995


def test_backbone_oracle_base224_matches_hf_golden(golden_backbone):
    name = "siglip2-base-patch16-224"
    c = R.CONFIGS[name]
    sd = R.init_state_dict(c, 0)
    x = R.preprocess_u8(R.synthetic_images(2, c.image_size, 0))
    o = R.siglip_vision_forward(sd, c, x, "fp32")
    assert np.abs(o["pooler_output"].numpy() - golden_backbone[name + "/pooled"]).max() < 2e-4


def test_flops_per_image_match_survey():
    assert abs(R.flops_per_image(R.CONFIGS["siglip2-base-patch16-224"]) / 1e9 - 35.417) < 1e-3
    assert abs(R.flops_per_image(R.CONFIGS["siglip2-so400m-patch14-384"]) / 1e9 - 670.346) < 1e-3


def test_valid_conv_ignores_trailing_pixels():
    """so400m-style geometry: image 60, patch 14 -> G=4; pixels >= 56 never matter (HF padding='valid')."""
    c = R.CONFIGS["tiny-hd72"]
    sd = R.init_state_dict(c, 0)
    img = R.synthetic_images(1, 60, 0)
    img2 = img.clone()
    img2[:, 56:, :, :] = 0
    img2[:, :, 56:, :] = 0
    a = R.siglip_vision_forward(sd, c, R.preprocess_u8(img))["pooler_output"]
    b = R.siglip_vision_forward(sd, c, R.preprocess_u8(img2))["pooler_output"]
    assert torch.equal(a, b)


def test_heads_oracle_matches_golden(golden_heads):
    for D in (128, 1152):
        pooled = torch.from_numpy(golden_heads[f"pooled_{D}"])
        zA = R.classifier_head(R.init_head("A", D, 1), "A", pooled, 0.0).numpy()
        zB = R.classifier_head(R.init_head("B", D, 1), "B", pooled, 1e-6).numpy()
        assert np.abs(zA - golden_heads[f"zA_{D}"]).max() < 2e-5
        assert np.abs(zB - golden_heads[f"zB_{D}"]).max() < 2e-5
        pr = torch.from_numpy(golden_heads[f"protos_{D}"])
        p = R.prototype_prob(R.l2_normalize(pooled), pr[0], pr[1]).numpy()
        assert np.abs(p - golden_heads[f"pproto_{D}"]).max() < 1e-6


def test_head_d_oracle_matches_reference_modules(golden_heads):
    """H-D: one-token attention == proj(v(x)); checked against the reference's LightweightAttention and
    nn.MultiheadAttention outputs stored in the golden file."""
    pooled = torch.from_numpy(golden_heads["pooled_128"])
    for size in ("tiny", "small", "medium", "large"):
        z = R.classifier_head_d(R.init_head_d(size, 128, 4), pooled).numpy()
        assert np.abs(z - golden_heads[f"zD_{size}_128"]).max() < 2e-5, size


def test_freq_features_oracle_matches_reference(golden_scoring):
    gray = golden_scoring["gray_u8"].astype(np.float32) / 255.0
    for i in range(gray.shape[0]):
        v = S.extract_freq_vector(gray[i])
        ref = golden_scoring["feats_raw"][i]
        scale = np.maximum(np.abs(ref), np.maximum(S.feature_scales(gray[i]), 1e-12))
        assert (np.abs(v - ref) / scale).max() < 2e-5
        assert np.abs(S.extract_freq_vector(gray[i], zscore=True) - golden_scoring["feats_zscore"][i]).max() < 1e-5
    # SRM_K[0] is SRM_K[1] embedded in 5x5 zeros: the reference's own outputs coincide
    assert np.array_equal(golden_scoring["feats_raw"][:, 15:18], golden_scoring["feats_raw"][:, 18:21])


def test_gray256_oracle_matches_reference(golden_scoring):
    import importlib.util

    spec = importlib.util.spec_from_file_location("mk", os.path.join(os.path.dirname(__file__), "..", "oracle", "make_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    imgs = mk.test_images()
    assert np.array_equal(np.array([im.shape[:2] for im in imgs]), golden_scoring["rgb_shapes"])
    for i, im in enumerate(imgs):
        assert np.array_equal(np.round(S.gray256_from_rgb_u8(im, True) * 255).astype(np.uint8), golden_scoring["gray_u8"][i])
        assert np.array_equal(np.round(S.gray256_from_rgb_u8(im, False) * 255).astype(np.uint8), golden_scoring["gray_u8"][4 + i])


def test_grid_tables_known_answers():
    """SURVEY.md §A.3 LUT known answers."""
    band, rbin, sector = S.grid_tables()
    occ = np.bincount(rbin[rbin >= 0].astype(int), minlength=39)
    assert occ.tolist() == [0, 0, 0, 0, 0, 4, 4, 0, 12, 0, 16, 8, 24, 28, 32, 48, 64, 76, 104, 148, 180, 248, 324, 428,
                            564, 732, 956, 1248, 1660, 2176, 2828, 3716, 4848, 6352, 8304, 10832, 11246, 6140, 2185]
    assert int((rbin < 0).sum()) == 1 and rbin[128, 128] == -1
    assert int((sector < 0).sum()) == 128 and (sector[128, :128] == -1).all()
    assert band[128, 128] == 0 and band[0, 0] == 2


def test_grid_tables_are_mirror_symmetric_in_band_and_radius():
    """freq_cols_kernel folds a spectrum bin and its Hermitian mirror together and lets them share the band and the log-radius
    bin (csrc/freq.cu:fold_bin): the tables of the shifted grid — the oracle's and the product's host-built ones — must be
    symmetric under (ky, kx) -> (-ky, -kx); the sector is not (it is read per position)."""
    from dfd import scoring

    band, rbin, sector = S.grid_tables()
    ky = (np.arange(256) - 128) % 256
    sym = ((256 - ky) + 128) % 256
    assert np.array_equal(band[sym][:, sym], band) and np.array_equal(rbin[sym][:, sym], rbin)
    assert not np.array_equal(sector[sym][:, sym], sector)
    tb, tr, _ = scoring.build_freq_tables()
    assert torch.equal(tb[sym][:, sym], tb) and torch.equal(tr[sym][:, sym], tr)


def test_heads_and_fusion_oracle(golden_scoring, shipped):
    assert np.abs(S.freq_mlp_g1(shipped["freq"], golden_scoring["feats_zscore"]) - golden_scoring["zfreq_g1"]).max() < 1e-5
    assert np.abs(S.freq_mlp_g2(S.init_freq_mlp_g2(2), golden_scoring["feats_raw"]) - golden_scoring["zfreq_g2"]).max() < 1e-5
    zs, zf = golden_scoring["fuse_zsig"], golden_scoring["fuse_zfreq"]
    assert np.abs(S.fusion_g1(shipped["fusion"], zs, zf) - golden_scoring["fuse_g1_z"]).max() < 1e-6
    assert np.abs(S.fusion_g2(S.init_fusion_g2(3), zf, zs) - golden_scoring["fuse_g2_z"]).max() < 1e-5
    loss, grads, _ = S.fusion_loss_and_grads(S.init_fusion_g2(3), zf, zs, golden_scoring["fuse_y"])
    assert abs(loss - float(golden_scoring["fuse_g2_loss"])) < 1e-6
    assert np.abs(grads - golden_scoring["fuse_g2_grads"]).max() < 1e-6


def test_shipped_artefact_layout(shipped):
    """Weight-layout contract of the shipped G1 files."""
    assert {k: tuple(v.shape) for k, v in shipped["freq"].items()} == {
        "net.0.weight": (24,), "net.0.bias": (24,), "net.1.weight": (64, 24), "net.1.bias": (64,),
        "net.3.weight": (1, 64), "net.3.bias": (1,)}
    assert {k: tuple(v.shape) for k, v in shipped["fusion"].items()} == {"fc.weight": (1, 2), "fc.bias": (1,)}
    assert abs(shipped["temp"]["temperature"] - 0.9956228137016296) < 1e-12


def test_coral_known_answers(golden_scoring, shipped):
    # cutpoint fitting KAT: shipped cutpoints == quantiles/max of shipped coral_bins.npy, bit exact
    fit = S.fit_coral_shipped(shipped["bins"])
    assert fit == {k: float(shipped["cuts"][k]) for k in ("q25", "q50", "q75", "max")}
    cl = S.coral_cut_logits(shipped["cuts"])
    assert np.allclose(cl, [-1.14372, -0.25708, 0.04720, 4.00685], atol=1e-5)
    assert np.allclose(cl, golden_scoring["coral_cut_logits"], atol=1e-7)
    d = S.detect_scores(golden_scoring["coral_z"], cl, 1.0)
    assert np.array_equal(d["risk_idx"], golden_scoring["coral_idx"])
    assert np.abs(d["risk_probs"] - golden_scoring["coral_probs"]).max() < 1e-6
    tp = S.coral_transition_points(cl)
    assert np.allclose(tp, [-0.525, 3.968], atol=2e-3)
    assert S.fit_coral_script(golden_scoring["fit_logits"]) == golden_scoring["fit_cuts_script"].tolist()
    assert np.allclose(S.coral_cut_logits(None), [S.logit(v) for v in (0.32, 0.47, 0.61, 0.75)])


# ---------------------------------------------------------------------------------------------------------
# gray256 restatement (Pillow luma + resample, OpenCV CLAHE)
# ---------------------------------------------------------------------------------------------------------
def test_gray_oracle_matches_reference_golden(golden_gray):
    """oracle/gray_ref.py against the outputs of the reference's own _pil_to_gray256[_clahe] (make_golden.py)."""
    from oracle import gray_ref as G

    for i, (h, w, kind, seed) in enumerate(G.GOLDEN_CASES):
        assert tuple(golden_gray["cases"][i]) == (h, w, seed)
        rgb = G.synthetic_rgb(h, w, kind, seed)
        for j, clahe in enumerate((True, False)):
            want = golden_gray["gray_u8"][2 * i + j].astype(np.float32) / np.float32(255.0)
            assert np.array_equal(G.gray256_from_rgb_u8(rgb, clahe), want), (h, w, kind, clahe)


def _preprocess_cases():
    from oracle import gray_ref as G

    return [(97, 64, "noise", 5), (512, 384, "edges", 6), (300, 451, "waves", 4), (224, 224, "waves", 2)], G


def test_preprocess_oracle_matches_reference_golden(golden_preprocess):
    """oracle/gray_ref.py (per-channel CLAHE + Pillow bilinear Resize) against the pixels the reference's OWN transform objects
    produce: `preprocess` of train_fusion_head_only.py:60-74 and the Original / H-Flip entries of create_tta_transforms
    (inference_ai_human_images.py:195-215), run by oracle/make_golden.py."""
    cases, G = _preprocess_cases()
    g = golden_preprocess
    assert [tuple(c) for c in g["cases"]] == [(h, w, s) for h, w, _, s in cases]
    S = int(g["tta_size"])
    for i, (h, w, kind, seed) in enumerate(cases):
        rgb = G.synthetic_rgb(h, w, kind, seed)
        assert np.array_equal(G.resize_u8(rgb, S, S, "bilinear"), g["tta_original_u8"][i]), ("tta", h, w)
        if i < len(g["train_u8"]):
            cl = np.stack([G.clahe_u8(np.ascontiguousarray(rgb[..., c])) for c in range(3)], -1)
            assert np.array_equal(G.resize_u8(cl, 384, 384, "bilinear"), g["train_u8"][i]), ("train", h, w)


def test_gray_oracle_matches_installed_libraries():
    """Stage by stage against Pillow / OpenCV as installed (skipped where they are missing)."""
    PIL_Image = pytest.importorskip("PIL.Image")
    cv2 = pytest.importorskip("cv2")
    from oracle import gray_ref as G

    for (h, w, kind, seed) in [(120, 200, "noise", 11), (257, 255, "waves", 12), (64, 64, "edges", 13), (8, 9, "noise", 14)]:
        rgb = G.synthetic_rgb(h, w, kind, seed)
        L = np.array(PIL_Image.fromarray(rgb, "RGB").convert("L"))
        assert np.array_equal(L, G.luma_u8(rgb))
        assert np.array_equal(np.array(PIL_Image.fromarray(L).resize((256, 256), PIL_Image.BICUBIC)),
                              G.resize_bicubic_u8(L, 256, 256))
        assert np.array_equal(cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8)).apply(L), G.clahe_u8(L))


def test_resize_oracle_matches_pillow():
    """RGB / L, bilinear (torchvision Resize on PIL inputs) and bicubic (open_clip preprocess), up- and down-scaling."""
    PIL_Image = pytest.importorskip("PIL.Image")
    from oracle import gray_ref as G

    rng = np.random.default_rng(3)
    for (h, w, oh, ow) in [(120, 160, 96, 96), (56, 56, 96, 96), (100, 37, 64, 48), (96, 96, 96, 96), (300, 17, 20, 40)]:
        im = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for f, pf in (("bilinear", PIL_Image.BILINEAR), ("bicubic", PIL_Image.BICUBIC)):
            assert np.array_equal(np.array(PIL_Image.fromarray(im, "RGB").resize((ow, oh), pf)), G.resize_u8(im, oh, ow, f))
            assert np.array_equal(np.array(PIL_Image.fromarray(im[..., 1]).resize((ow, oh), pf)),
                                  G.resize_u8(im[..., 1], oh, ow, f))


def test_freqmlp_grad_oracle_matches_reference_class(golden_freq_train):
    """float64 autograd restatement vs loss / gradients of the reference's own FreqMLP class (make_golden.py)."""
    g = golden_freq_train
    loss, grads, out = S.freq_mlp_g2_loss_and_grads(S.init_freq_mlp_g2(7), g["feats"], g["y"])
    assert abs(loss - float(g["loss"])) < 1e-6
    assert np.abs(grads - g["grads"]).max() < 2e-6 and grads.shape == (6494,)
    assert np.abs(out - g["logits"]).max() < 1e-5


def test_decoder_oracle_matches_reference_class(golden_decoder):
    """oracle/decoder_ref.py vs the reference's own SegFormerStrongDecoder (make_golden.py)."""
    from oracle import decoder_ref as D

    g = golden_decoder
    C, K, E, grid, S, B = (int(v) for v in g["dims"])
    out = D.decoder_forward(D.init_decoder_state(C, K, E, 3), [torch.from_numpy(h) for h in g["hidden"]], grid, S)
    assert np.abs(out.numpy() - g["seg"]).max() < 1e-6
