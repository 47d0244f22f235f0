"""ORACLE (test infrastructure, not product code) — CPU restatement of the reference's scoring stack.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Restates, in numpy (float32 where the reference holds fp32 tensors, python float64 where it holds
`.item()` scalars):
  gray256            train_fusion_head_only.py:142-148, deepfake-detector-v2/app.py:736-749
  fft_features       train_fusion_head_only.py:150-209 (= "FreqMLP trainer.py":91-160, app.py:752-823)
  srm_features       train_fusion_head_only.py:211-222 (app.py:826-837)
  extract_freq_vector  train_fusion_head_only.py:224-226 (raw) / app.py:840-846 (z-scored)
  FreqMLP G1 / G2    app.py:601-628 / train_fusion_head_only.py:230-301
  fusion G1 / G2     app.py:691-696,1355-1362 / train_fusion_head_only.py:303-317
  CORAL              app.py:1265-1297,1365-1396 ; fitting coral.py:300-322 and the shipped-artefact rule
  fusion training    train_fusion_head_only.py:406-427 (BCEWithLogits mean, autograd for the gradient)

PyWavelets (`pywt.dwt2(x, 'db1')`, train_fusion_head_only.py:184-187) is an un-vendored, unpinned, un-installed
dependency: the Haar step follows the published db1 definition (orthonormal 2x2 Haar) — "parity unpinned"
against real pywt for those 8 values, pinned for everything else by tests/golden/scoring_*.npz, which
oracle/make_golden.py produced by exec'ing the reference's own function bodies.
"""
from __future__ import annotations

import math

import numpy as np
import torch  # only for the primitives whose rounding defines bin edges: logspace, bucketize, atan2

EPS = 1e-8
N = 256


# ---------------------------------------------------------------------------------------------------
# gray256 (host-side stage-1 boundary: PIL / cv2 integer arithmetic)
# ---------------------------------------------------------------------------------------------------
def gray256_from_rgb_u8(rgb_u8_hwc: np.ndarray, clahe: bool) -> np.ndarray:
    from PIL import Image

    g = Image.fromarray(np.ascontiguousarray(rgb_u8_hwc), "RGB").convert("L")
    if clahe:
        import cv2

        g = Image.fromarray(cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8)).apply(np.array(g, dtype=np.uint8)))
    g = g.resize((N, N), Image.BICUBIC)
    return np.asarray(g, dtype=np.float32) / 255.0


# ---------------------------------------------------------------------------------------------------
# constants of the 256x256 grid
# ---------------------------------------------------------------------------------------------------
def grid_tables():
    """band id (0 low, 1 mid, 2 high), log-radius bin (-1..38), sector (-1..7) per fft-SHIFTED pixel."""
    yy, xx = torch.meshgrid(torch.arange(N), torch.arange(N), indexing="ij")
    r = torch.sqrt((yy - N // 2) ** 2 + (xx - N // 2) ** 2)  # int64 -> float32, as in the reference
    rmax = float(r.max())
    r1, r2 = 0.15 * rmax, 0.45 * rmax
    band = torch.full((N, N), 2, dtype=torch.uint8)
    band[(r > r1) & (r <= r2)] = 1
    band[r <= r1] = 0
    rb = torch.logspace(math.log10(1.0), math.log10(rmax + 1.0), 40)
    ridx = (torch.bucketize(r.flatten() + 1.0, rb) - 1).reshape(N, N)
    rbin = torch.where((ridx >= 0) & (ridx < 39), ridx, torch.full_like(ridx, -1)).to(torch.int8)
    ang = torch.atan2(yy - N // 2, xx - N // 2)
    sector = torch.full((N, N), -1, dtype=torch.int8)
    for k, a0 in enumerate(np.linspace(-math.pi, math.pi, 8, endpoint=False)):
        sector[(ang >= a0) & (ang < a0 + math.pi / 4)] = k
    return band.numpy(), rbin.numpy(), sector.numpy()


_TABLES = None


def _tables():
    global _TABLES
    if _TABLES is None:
        _TABLES = grid_tables()
    return _TABLES


def haar2(x: np.ndarray):
    """One level of the orthonormal Haar (db1) 2-D DWT on an even-sized array: (cA, cH, cV, cD)."""
    a, b, c, d = x[0::2, 0::2], x[0::2, 1::2], x[1::2, 0::2], x[1::2, 1::2]
    return (a + b + c + d) * 0.5, (a + b - c - d) * 0.5, (a - b + c - d) * 0.5, (a - b - c + d) * 0.5


def fft_features(x: np.ndarray, eps: float = EPS) -> list:
    """15 spectral + wavelet features of a gray 256x256 float32 image in [0,1]."""
    band, rbin, sector = _tables()
    Fs = np.fft.fftshift(np.fft.fft2(x.astype(np.float64)))
    mag = np.abs(Fs).astype(np.float32)
    # the 4 self-conjugate bins are exactly real for a real image (+0 imaginary part)
    for (i, j) in ((128, 128), (0, 128), (128, 0), (0, 0)):
        Fs[i, j] = complex(Fs[i, j].real, 0.0)
    phase = np.angle(Fs).astype(np.float32)
    m64 = mag.astype(np.float64)
    Et = float(m64.sum()) + eps
    El, Em, Eh = (float(m64[band == k].sum()) for k in (0, 1, 2))
    mu = []
    logm = np.log(mag + np.float32(1e-6)).astype(np.float64)
    for i in range(39):
        sel = rbin == i
        mu.append(float(np.float32(logm[sel].mean())) if sel.any() else 0.0)
    xs = np.arange(39, dtype=np.float64)
    mu = np.asarray(mu)
    slope = float(((xs - xs.mean()) * (mu - mu.mean())).sum() / ((xs - xs.mean()) ** 2).sum())
    # torch.histc(bins=50, min=-pi, max=pi) on float32 phases
    lo, hi = np.float32(-math.pi), np.float32(math.pi)
    pos = ((phase - lo) / (hi - lo) * np.float32(50)).astype(np.int64)
    pos = np.minimum(pos, 49)
    hist = np.bincount(pos.ravel(), minlength=50).astype(np.float32)
    prob = hist / (hist.sum() + np.float32(eps))
    entropy = float(-(prob * np.log(prob + np.float32(eps))).sum())
    sect = []
    for k in range(8):
        sel = sector == k
        sect.append(float(np.float32(m64[sel].mean())) if sel.any() else 0.0)
    anis = float(np.var(sect))
    cA1, cH1, cV1, cD1 = haar2(x.astype(np.float32))
    cA2, cH2, cV2, cD2 = haar2(cA1)
    wave = [float(np.mean(np.abs(c.astype(np.float64)) ** 2)) for c in (cA1, cH1, cV1, cD1, cA2, cH2, cV2, cD2)]
    return [El / Et, Em / Et, Eh / Et, (Eh + eps) / (El + eps), slope, anis, entropy] + wave


_SRM = [
    np.array([[0, 0, 0, 0, 0], [0, -1, 2, -1, 0], [0, 2, -4, 2, 0], [0, -1, 2, -1, 0], [0, 0, 0, 0, 0]], np.float32),
    np.array([[-1, 2, -1], [2, -4, 2], [-1, 2, -1]], np.float32),
    np.array([[0, -1, 0], [-1, 4, -1], [0, -1, 0]], np.float32),
]


def _corr_same(x: np.ndarray, k: np.ndarray) -> np.ndarray:
    """Zero-padded 'same' cross-correlation (what F.conv2d computes)."""
    r = k.shape[0] // 2
    xp = np.pad(x.astype(np.float64), r)
    y = np.zeros_like(x, dtype=np.float64)
    for dy in range(k.shape[0]):
        for dx in range(k.shape[1]):
            if k[dy, dx] != 0:
                y += float(k[dy, dx]) * xp[dy : dy + x.shape[0], dx : dx + x.shape[1]]
    return y


def srm_features(x: np.ndarray, eps: float = EPS) -> list:
    feats = []
    for k2d in _SRM:
        k = k2d / np.float32(np.abs(k2d).sum() + np.float32(eps))
        y = _corr_same(x, k)
        m = float(np.float32(y.mean()))
        v = float(np.float32(y.var()))
        kurt = float(np.float32(((y - y.mean()) ** 4).mean())) / ((v + eps) ** 2)
        feats += [m, v, kurt]
    return feats


def extract_freq_vector(x: np.ndarray, eps: float = EPS, zscore: bool = False) -> np.ndarray:
    v = np.asarray(fft_features(x, eps) + srm_features(x, eps), dtype=np.float32)
    if zscore:
        sd = np.float32(v.std(ddof=1))
        if sd < 1e-6:
            return v * np.float32(0)
        v = (v - v.mean(dtype=np.float32)) / (sd + np.float32(1e-6))
    return v.astype(np.float32)


def feature_scales(x: np.ndarray) -> np.ndarray:
    """Natural per-feature scale for the 1e-4 relative gate: the three SRM means are sums
    of a zero-sum stencil (ill-conditioned), so their scale is rms(y)/sqrt(n), not |mean|."""
    s = np.zeros(24, dtype=np.float64)
    for i, k2d in enumerate(_SRM):
        y = _corr_same(x, k2d / np.float32(np.abs(k2d).sum()))
        s[15 + 3 * i] = math.sqrt(float((y ** 2).mean())) / math.sqrt(y.size) * 100.0
    return s


# ---------------------------------------------------------------------------------------------------
# heads
# ---------------------------------------------------------------------------------------------------
def _erf(x):
    from scipy.special import erf

    return erf(x)


def _gelu(x):
    return 0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))


def _ln(x, w, b, eps):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * w + b


def _sig(x):
    return 1.0 / (1.0 + np.exp(-x))


def _np(sd):
    return {k: (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)).astype(np.float64) for k, v in sd.items()}


def freq_mlp_g1(sd: dict, x: np.ndarray) -> np.ndarray:
    """SafeLayerNorm(24, eps 1e-5) -> Linear(24,64) -> GELU(erf) -> Linear(64,1); eval-time noise omitted."""
    p = _np(sd)
    h = _ln(x.astype(np.float64), p["net.0.weight"], p["net.0.bias"], 1e-5)
    h = _gelu(h @ p["net.1.weight"].T + p["net.1.bias"])
    return (h @ p["net.3.weight"].T + p["net.3.bias"])[..., 0]


def freq_mlp_g2(sd: dict, x: np.ndarray) -> np.ndarray:
    p = _np(sd)
    h = (x.astype(np.float64) - p["normer.mean"]) / (p["normer.std"] + 1e-6)
    h = np.tanh(p["contrast.alpha"] * h + p["contrast.beta"])
    h = h * np.repeat(_sig(p["band.gates"]), 6)
    for b in range(2):
        y = _ln(h, p[f"blocks.{b}.norm.weight"], p[f"blocks.{b}.norm.bias"], 1e-5)
        y = _gelu(y @ p[f"blocks.{b}.fc1.weight"].T + p[f"blocks.{b}.fc1.bias"])
        h = h + y @ p[f"blocks.{b}.fc2.weight"].T + p[f"blocks.{b}.fc2.bias"]
    z = (h @ p["head.weight"].T + p["head.bias"])[..., 0]
    return z / (float(p["temp.T"]) + 1e-6)


def fusion_g1(sd: dict, z_sig: np.ndarray, z_freq: np.ndarray, freq_temp: float = 1.25) -> np.ndarray:
    p = _np(sd)
    w = p["fc.weight"].reshape(-1)
    return w[0] * _sig(z_sig.astype(np.float64)) + w[1] * _sig(z_freq.astype(np.float64) / freq_temp) + p["fc.bias"].reshape(-1)[0]


def fusion_g2(sd: dict, z_freq: np.ndarray, z_sig: np.ndarray) -> np.ndarray:
    p = _np(sd)
    zf, zs = z_freq.astype(np.float64), z_sig.astype(np.float64)
    x = np.stack([zf, zs, np.abs(zf - zs)], -1)
    l = _gelu(x @ p["mlp.0.weight"].T + p["mlp.0.bias"]) @ p["mlp.2.weight"].T + p["mlp.2.bias"]
    l = l - l.max(-1, keepdims=True)
    w = np.exp(l) / np.exp(l).sum(-1, keepdims=True)
    return (w[..., 0] * zf + w[..., 1] * zs) / (float(p["temp.T"]) + 1e-6)


def init_freq_mlp_g2(seed: int = 2) -> dict:
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=1.0: torch.randn(*s, generator=g) * std
    sd = {"normer.mean": rn(24, std=0.5), "normer.std": 0.5 + torch.rand(24, generator=g),
          "contrast.alpha": 1.0 + rn(24, std=0.2), "contrast.beta": rn(24, std=0.2), "band.gates": rn(4),
          "head.weight": rn(1, 24, std=0.4), "head.bias": rn(1, std=0.1), "temp.T": torch.tensor(1.3)}
    for b in range(2):
        sd[f"blocks.{b}.norm.weight"], sd[f"blocks.{b}.norm.bias"] = 1.0 + rn(24, std=0.1), rn(24, std=0.1)
        sd[f"blocks.{b}.fc1.weight"], sd[f"blocks.{b}.fc1.bias"] = rn(64, 24, std=0.3), rn(64, std=0.1)
        sd[f"blocks.{b}.fc2.weight"], sd[f"blocks.{b}.fc2.bias"] = rn(24, 64, std=0.2), rn(24, std=0.1)
    return sd


def init_fusion_g2(seed: int = 3) -> dict:
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=1.0: torch.randn(*s, generator=g) * std
    return {"mlp.0.weight": rn(32, 3, std=0.6), "mlp.0.bias": rn(32, std=0.3), "mlp.2.weight": rn(2, 32, std=0.4),
            "mlp.2.bias": rn(2, std=0.1), "temp.T": torch.tensor(0.9)}


# ---------------------------------------------------------------------------------------------------
# CORAL
# ---------------------------------------------------------------------------------------------------
def logit(p: float) -> float:
    p = min(max(p, 1e-6), 1 - 1e-6)
    return math.log(p / (1 - p))


def coral_cut_logits(cuts: dict | None) -> np.ndarray:
    if cuts:
        return np.array([logit(cuts[k]) for k in ("q25", "q50", "q75", "max")], dtype=np.float32)
    return np.array([logit(v) for v in (0.32, 0.47, 0.61, 0.75)], dtype=np.float32)


def coral_probs(z_scaled: np.ndarray, cuts_logit: np.ndarray) -> np.ndarray:
    z = np.asarray(z_scaled, dtype=np.float32).reshape(-1, 1)
    g = (1.0 / (1.0 + np.exp(-(z - cuts_logit[None].astype(np.float32))))).astype(np.float32)
    p = np.concatenate([1.0 - g[:, :1], g[:, :-1] - g[:, 1:], g[:, -1:]], axis=1).astype(np.float32)
    return p / (p.sum(1, keepdims=True) + np.float32(1e-8))


def detect_scores(z: np.ndarray, cuts_logit: np.ndarray, coral_temp: float) -> dict:
    """app.py:1365-1396 on a batch of fused logits z."""
    z_scaled = np.asarray(z, dtype=np.float64) / max(coral_temp, 1e-3)
    p_raw = _sig(z_scaled.astype(np.float32)).astype(np.float32)
    p = coral_probs(z_scaled, cuts_logit)
    idx = p.argmax(1).astype(np.int32)
    k = np.arange(5, dtype=np.float32)
    mu = (p * k).sum(1)
    var = (p * (k[None] - mu[:, None]) ** 2).sum(1)
    p_coral = np.clip(mu / 4.0 + 0.5 * var, 0.0, 1.0)
    entropy = -(p * np.log(p + np.float32(1e-8))).sum(1)
    p_blend = np.clip(0.70 * p_raw + 0.30 * p_coral, 0.0, 1.0)
    return {"z_scaled": z_scaled.astype(np.float32), "p_raw": p_raw, "risk_probs": p, "risk_idx": idx,
            "p_coral": p_coral.astype(np.float32), "entropy": entropy.astype(np.float32),
            "p_blend": p_blend.astype(np.float32)}


def coral_transition_points(cuts_logit: np.ndarray, lo: float = -12.0, hi: float = 14.0, n: int = 260001) -> np.ndarray:
    """z_scaled values where argmax of the CORAL probabilities changes."""
    zs = np.linspace(lo, hi, n)
    idx = coral_probs(zs, cuts_logit).argmax(1)
    ch = np.nonzero(idx[1:] != idx[:-1])[0]
    return 0.5 * (zs[ch] + zs[ch + 1])


def fit_coral_shipped(probs: np.ndarray) -> dict:
    """Rule that reproduces the shipped siglip/coral_cutpoints.json from coral_bins.npy."""
    q = np.quantile(probs, [0.25, 0.5, 0.75]).astype(np.float32)  # the shipped values are float32-rounded
    return {"q25": float(q[0]), "q50": float(q[1]), "q75": float(q[2]), "max": float(np.float32(probs.max()))}


def fit_coral_script(logits: np.ndarray) -> list:
    """coral.py:300-322: sorted logits at ranks floor(q*n), q in {.15,.35,.55,.75}."""
    s = np.sort(np.asarray(logits))
    return [float(s[int(q * len(s))]) for q in (0.15, 0.35, 0.55, 0.75)]


# ---------------------------------------------------------------------------------------------------
# fusion-head training step (reference loop + enable_grad)
# ---------------------------------------------------------------------------------------------------
def fusion_loss_and_grads(sd: dict, z_freq: np.ndarray, z_sig: np.ndarray, y: np.ndarray):
    """mean BCEWithLogits(AdaptiveFusionHead(z_freq, z_sig), y) and d loss / d params (flat 195, order
    mlp.0.weight, mlp.0.bias, mlp.2.weight, mlp.2.bias, temp.T) via float64 autograd."""
    names = ("mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias", "temp.T")
    with torch.enable_grad():
        p = {k: torch.as_tensor(np.asarray(sd[k].detach().cpu() if hasattr(sd[k], "detach") else sd[k]),
                                dtype=torch.float64).clone().requires_grad_(True) for k in names}
        zf = torch.as_tensor(z_freq, dtype=torch.float64)
        zs = torch.as_tensor(z_sig, dtype=torch.float64)
        x = torch.stack([zf, zs, (zf - zs).abs()], -1)
        h = torch.nn.functional.gelu(x @ p["mlp.0.weight"].t() + p["mlp.0.bias"])
        w = torch.softmax(h @ p["mlp.2.weight"].t() + p["mlp.2.bias"], -1)
        out = (w[..., 0] * zf + w[..., 1] * zs) / (p["temp.T"] + 1e-6)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(out, torch.as_tensor(y, dtype=torch.float64))
        loss.backward()
    grads = torch.cat([p[k].grad.reshape(-1) for k in names]).numpy()
    return float(loss.detach()), grads, out.detach().numpy()


FREQMLP_PARAM_ORDER = ("contrast.alpha", "contrast.beta", "band.gates") + tuple(
    f"blocks.{b}.{n}" for b in range(2)
    for n in ("norm.weight", "norm.bias", "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")) + ("head.weight", "head.bias", "temp.T")


def freq_mlp_g2_loss_and_grads(sd: dict, feats: np.ndarray, y: np.ndarray):
    """mean BCEWithLogits(FreqMLP_G2(feats), y) and d loss / d params (flat 6494, FREQMLP_PARAM_ORDER) via float64
    autograd: "FreqMLP trainer.py":218-301 (eval-mode dropout, i.e. none) and :366-369."""
    F = torch.nn.functional
    with torch.enable_grad():
        t = lambda v: torch.as_tensor(np.asarray(v.detach().cpu() if hasattr(v, "detach") else v), dtype=torch.float64)
        p = {k: t(sd[k]).clone().requires_grad_(True) for k in FREQMLP_PARAM_ORDER}
        x = (t(feats) - t(sd["normer.mean"])) / (t(sd["normer.std"]) + 1e-6)
        x = torch.tanh(p["contrast.alpha"] * x + p["contrast.beta"])
        gates = torch.sigmoid(p["band.gates"])
        x = torch.cat([c * gates[i] for i, c in enumerate(torch.split(x, 6, dim=-1))], dim=-1)
        for b in range(2):
            r = x
            h = F.layer_norm(x, (24,), p[f"blocks.{b}.norm.weight"], p[f"blocks.{b}.norm.bias"], 1e-5)
            h = F.gelu(h @ p[f"blocks.{b}.fc1.weight"].t() + p[f"blocks.{b}.fc1.bias"])
            x = h @ p[f"blocks.{b}.fc2.weight"].t() + p[f"blocks.{b}.fc2.bias"] + r
        out = ((x @ p["head.weight"].t()).squeeze(-1) + p["head.bias"]) / (p["temp.T"] + 1e-6)
        loss = F.binary_cross_entropy_with_logits(out, t(y))
        loss.backward()
    grads = torch.cat([p[k].grad.reshape(-1) for k in FREQMLP_PARAM_ORDER]).numpy()
    return float(loss.detach()), grads, out.detach().numpy()
