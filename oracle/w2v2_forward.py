"""ORACLE (test infrastructure, never shipped or measured as the product).

CPU restatement of the Wav2Vec2ForCTC / Wav2Vec2ConformerForCTC eval-mode forward
that the reference's callbacks trigger (shap_calculation.py:42,
feasability_tests/w2v2conformer.py:36-38, lime_shap_wav2vec2_comparison.py:66-67).

The arithmetic lives in the un-vendored third-party dependency ``transformers``
(requirements.txt:3 pins only ``>=4.15.0``; the copy installed in this image is
5.5.0).  Each function cites the HF file:line it follows
(HF = site-packages/transformers/models/).  The restatement works on a plain
``state_dict`` with HF parameter names and uses only elementary torch CPU ops in
fp32 (or fp64 when ``dtype=torch.float64``), so it can be diffed op by op against
the CUDA path.

Pinning: tests/test_oracle.py checks this file against (1) the installed
``transformers`` modules with shared random-init weights (max abs logit
difference <= 2e-4 in fp32) and (2) the committed fixtures under tests/golden/
that ``oracle/make_golden.py`` generated from ``transformers`` itself.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F


def _act(name: str):
    # HF activations.py ACT2FN: "gelu" is the exact erf form, "swish" == SiLU
    if name == "gelu":
        return lambda x: F.gelu(x)
    if name in ("swish", "silu"):
        return F.silu
    raise ValueError(f"unsupported activation {name}")


def feature_encoder(sd: Dict[str, torch.Tensor], cfg: dict, x: torch.Tensor, prefix: str) -> torch.Tensor:
    """HF wav2vec2/modeling_wav2vec2.py:382-419 (+ layer classes :254-323).

    x: [B, L] -> [B, C, T'].  Layer 0 of the "group" variant is conv -> GroupNorm(C groups)
    -> GELU; the remaining layers are conv -> GELU.  The "layer" variant is
    conv(+bias) -> LayerNorm(channels) -> GELU on every layer.
    """
    h = x[:, None, :]
    n_layers = len(cfg["conv_dim"])
    for i in range(n_layers):
        p = f"{prefix}feature_extractor.conv_layers.{i}."
        bias = sd.get(p + "conv.bias") if cfg["conv_bias"] else None
        h = F.conv1d(h, sd[p + "conv.weight"], bias, stride=cfg["conv_stride"][i])
        if cfg["feat_extract_norm"] == "group":
            if i == 0:
                C = h.shape[1]
                h = F.group_norm(h, C, sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], eps=1e-5)
        else:
            h = F.layer_norm(h.transpose(1, 2), (h.shape[1],), sd[p + "layer_norm.weight"],
                             sd[p + "layer_norm.bias"], eps=1e-5).transpose(1, 2)
        h = F.gelu(h)  # feat_extract_activation == "gelu" in every named config
    return h


def feature_projection(sd, cfg, feats, prefix):
    """HF modeling_wav2vec2.py:422-434: LayerNorm(conv_dim[-1]) -> Linear."""
    p = prefix + "feature_projection."
    h = F.layer_norm(feats, (feats.shape[-1],), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"],
                     eps=cfg["layer_norm_eps"])
    return F.linear(h, sd[p + "projection.weight"], sd[p + "projection.bias"])


def pos_conv_weight(sd, prefix):
    """Fold weight_norm(dim=2) (HF modeling_wav2vec2.py:337-354): w = g * v / ||v||_{dims 0,1}."""
    p = prefix + "encoder.pos_conv_embed.conv."
    if p + "weight" in sd:
        return sd[p + "weight"]
    g = sd[p + "parametrizations.weight.original0"]
    v = sd[p + "parametrizations.weight.original1"]
    norm = v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()
    return v * (g / norm)


def pos_conv_embed(sd, cfg, h, prefix):
    """HF modeling_wav2vec2.py:326-379: grouped conv k=128 pad=64, drop last frame, GELU."""
    k = cfg["num_conv_pos_embeddings"]
    w = pos_conv_weight(sd, prefix)
    y = F.conv1d(h.transpose(1, 2), w, sd[prefix + "encoder.pos_conv_embed.conv.bias"], padding=k // 2,
                 groups=cfg["num_conv_pos_embedding_groups"])
    if k % 2 == 0:
        y = y[:, :, :-1]
    return F.gelu(y).transpose(1, 2)


def _mha(sd, p, h, n_heads):
    """HF modeling_wav2vec2.py:500-549 with eager_attention_forward :438-463 (no mask)."""
    B, T, H = h.shape
    d = H // n_heads
    q = F.linear(h, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"]).view(B, T, n_heads, d).transpose(1, 2)
    k = F.linear(h, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"]).view(B, T, n_heads, d).transpose(1, 2)
    v = F.linear(h, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"]).view(B, T, n_heads, d).transpose(1, 2)
    s = torch.matmul(q, k.transpose(2, 3)) * (d ** -0.5)
    a = torch.softmax(s, dim=-1)
    o = torch.matmul(a, v).transpose(1, 2).reshape(B, T, H)
    return F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def _ffn(sd, p, h, act):
    """HF modeling_wav2vec2.py:552-573."""
    h = F.linear(h, sd[p + "intermediate_dense.weight"], sd[p + "intermediate_dense.bias"])
    h = act(h)
    return F.linear(h, sd[p + "output_dense.weight"], sd[p + "output_dense.bias"])


def _ln(sd, p, h, eps):
    return F.layer_norm(h, (h.shape[-1],), sd[p + "weight"], sd[p + "bias"], eps=eps)


def encoder_wav2vec2(sd, cfg, h, prefix):
    """HF modeling_wav2vec2.py:658-727 (post-LN, :592-609) / :730-800 (stable-LN, :612-655)."""
    eps = cfg["layer_norm_eps"]
    act = _act(cfg["hidden_act"])
    e = prefix + "encoder."
    h = h + pos_conv_embed(sd, cfg, h, prefix)
    if not cfg["do_stable_layer_norm"]:
        h = _ln(sd, e + "layer_norm.", h, eps)
    for i in range(cfg["num_hidden_layers"]):
        p = f"{e}layers.{i}."
        if cfg["do_stable_layer_norm"]:
            r = h
            h = r + _mha(sd, p + "attention.", _ln(sd, p + "layer_norm.", h, eps), cfg["num_attention_heads"])
            h = h + _ffn(sd, p + "feed_forward.", _ln(sd, p + "final_layer_norm.", h, eps), act)
        else:
            h = _ln(sd, p + "layer_norm.", h + _mha(sd, p + "attention.", h, cfg["num_attention_heads"]), eps)
            h = _ln(sd, p + "final_layer_norm.", h + _ffn(sd, p + "feed_forward.", h, act), eps)
    if cfg["do_stable_layer_norm"]:
        h = _ln(sd, e + "layer_norm.", h, eps)
    return h


def rel_pos_embeddings(T: int, d_model: int, dtype) -> torch.Tensor:
    """HF wav2vec2_conformer/modeling_wav2vec2_conformer.py:159-205, sliced to [1, 2T-1, d].

    Row r (0..2T-2) encodes relative position (T-1-r): sin/cos interleaved.
    """
    pos = torch.arange(T - 1, -T, -1, dtype=torch.float32).unsqueeze(1)         # T-1 ... -(T-1)
    div = torch.exp(torch.arange(0, d_model, 2, dtype=torch.int64).float() * -(math.log(10000.0) / d_model))
    pe = torch.zeros(2 * T - 1, d_model)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(0).to(dtype)


def _conformer_attention(sd, p, h, cfg, pos):
    """HF modeling_wav2vec2_conformer.py:420-565 (relative: :509-565; rotary: :489-507)."""
    B, T, H = h.shape
    nh = cfg["num_attention_heads"]
    d = H // nh
    qk_in = h
    if cfg["position_embeddings_type"] == "rotary":
        cos, sin = pos
        x = h.view(B, T, nh, d)
        rot = torch.cat((-x[..., d // 2:], x[..., : d // 2]), dim=-1)
        qk_in = (x * cos[None, :, None, :] + rot * sin[None, :, None, :]).reshape(B, T, H)
    q = F.linear(qk_in, sd[p + "linear_q.weight"], sd[p + "linear_q.bias"]).view(B, T, nh, d)
    k = F.linear(qk_in, sd[p + "linear_k.weight"], sd[p + "linear_k.bias"]).view(B, T, nh, d).transpose(1, 2)
    v = F.linear(h, sd[p + "linear_v.weight"], sd[p + "linear_v.bias"]).view(B, T, nh, d).transpose(1, 2)
    if cfg["position_embeddings_type"] == "relative":
        pp = F.linear(pos, sd[p + "linear_pos.weight"]).view(1, 2 * T - 1, nh, d).permute(0, 2, 3, 1)  # [1,h,d,2T-1]
        qu = (q + sd[p + "pos_bias_u"]).transpose(1, 2)
        qv = (q + sd[p + "pos_bias_v"]).transpose(1, 2)
        ac = torch.matmul(qu, k.transpose(-2, -1))
        bd_raw = torch.matmul(qv, pp)                                            # [B,h,T,2T-1]
        # shift trick (:553-559) as explicit indexing: bd[i, j] = raw[i, T-1-i+j]
        idx = (T - 1 - torch.arange(T)[:, None] + torch.arange(T)[None, :])
        bd = torch.gather(bd_raw, 3, idx[None, None].expand(B, nh, T, T))
        s = (ac + bd) / math.sqrt(d)
    else:
        s = torch.matmul(q.transpose(1, 2), k.transpose(-2, -1)) / math.sqrt(d)
    a = torch.softmax(s, dim=-1)
    o = torch.matmul(a, v).transpose(1, 2).reshape(B, T, H)
    return F.linear(o, sd[p + "linear_out.weight"], sd[p + "linear_out.bias"])


def _conformer_conv_module(sd, p, h, cfg):
    """HF modeling_wav2vec2_conformer.py:360-417; BatchNorm in eval mode uses running stats."""
    H = h.shape[-1]
    k = cfg["conv_depthwise_kernel_size"]
    x = F.layer_norm(h, (H,), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], eps=1e-5).transpose(1, 2)
    x = F.conv1d(x, sd[p + "pointwise_conv1.weight"])
    x = F.glu(x, dim=1)
    x = F.conv1d(x, sd[p + "depthwise_conv.weight"], padding=(k - 1) // 2, groups=H)
    x = F.batch_norm(x, sd[p + "batch_norm.running_mean"], sd[p + "batch_norm.running_var"],
                     sd[p + "batch_norm.weight"], sd[p + "batch_norm.bias"], training=False, eps=1e-5)
    x = _act(cfg["hidden_act"])(x)
    x = F.conv1d(x, sd[p + "pointwise_conv2.weight"])
    return x.transpose(1, 2)


def encoder_conformer(sd, cfg, h, prefix):
    """HF modeling_wav2vec2_conformer.py:633-717 (layer :568-630).  pos_conv_embed is constructed
    there (:645) but never called in forward, so it is not applied here either."""
    eps = cfg["layer_norm_eps"]
    act = _act(cfg["hidden_act"])
    e = prefix + "encoder."
    B, T, H = h.shape
    if cfg["position_embeddings_type"] == "relative":
        pos = rel_pos_embeddings(T, H, h.dtype)
    elif cfg["position_embeddings_type"] == "rotary":
        d = H // cfg["num_attention_heads"]
        inv = 1.0 / (cfg["rotary_embedding_base"] ** (torch.arange(0, d, 2, dtype=torch.int64).float() / d))
        fr = torch.einsum("i,j->ij", torch.arange(T).float(), inv)
        emb = torch.cat((fr, fr), dim=-1)
        pos = (emb.cos().to(h.dtype), emb.sin().to(h.dtype))
    else:
        pos = None
    for i in range(cfg["num_hidden_layers"]):
        p = f"{e}layers.{i}."
        h = 0.5 * _ffn(sd, p + "ffn1.", _ln(sd, p + "ffn1_layer_norm.", h, 1e-5), act) + h
        h = _conformer_attention(sd, p + "self_attn.", _ln(sd, p + "self_attn_layer_norm.", h, 1e-5), cfg, pos) + h
        h = h + _conformer_conv_module(sd, p + "conv_module.", h, cfg)
        h = 0.5 * _ffn(sd, p + "ffn2.", _ln(sd, p + "ffn2_layer_norm.", h, 1e-5), act) + h
        h = _ln(sd, p + "final_layer_norm.", h, 1e-5)
    return _ln(sd, e + "layer_norm.", h, eps)


def ctc_logits(sd: Dict[str, torch.Tensor], cfg: dict, x: torch.Tensor) -> torch.Tensor:
    """Full eval-mode forward: waveform [B, L] -> logits [B, T', vocab].

    HF modeling_wav2vec2.py:1675-1744 (ForCTC) -> :1327-1383 (Model) -> lm_head :1708.
    The all-ones attention mask the reference passes (shap_calculation.py:39) is a no-op
    (HF :1026-1044, :679-688), so no mask is modelled.
    """
    kind = cfg.get("kind", "wav2vec2")
    prefix = "wav2vec2_conformer." if kind == "conformer" else "wav2vec2."
    dtype = x.dtype
    sd = {k: v.to(dtype) if v.is_floating_point() else v for k, v in sd.items()}
    feats = feature_encoder(sd, cfg, x, prefix).transpose(1, 2)
    h = feature_projection(sd, cfg, feats, prefix)
    if kind == "conformer":
        h = encoder_conformer(sd, cfg, h, prefix)
    else:
        h = encoder_wav2vec2(sd, cfg, h, prefix)
    return F.linear(h, sd["lm_head.weight"], sd["lm_head.bias"])


# --------------------------------------------------------------------------------------
# Construction of the real third-party model (used to pin this restatement and to make
# golden fixtures; transformers is an installed library on both the build and GPU boxes).
# --------------------------------------------------------------------------------------

def build_hf_model(cfg: dict, seed: int = 0):
    """Random-init HF model of the named architecture, eval mode (SURVEY.md 3.2 caveat)."""
    import transformers

    torch.manual_seed(seed)
    common = dict(
        conv_dim=list(cfg["conv_dim"]), conv_kernel=list(cfg["conv_kernel"]), conv_stride=list(cfg["conv_stride"]),
        conv_bias=cfg["conv_bias"], feat_extract_norm=cfg["feat_extract_norm"], hidden_size=cfg["hidden_size"],
        num_hidden_layers=cfg["num_hidden_layers"], num_attention_heads=cfg["num_attention_heads"],
        intermediate_size=cfg["intermediate_size"], num_conv_pos_embeddings=cfg["num_conv_pos_embeddings"],
        num_conv_pos_embedding_groups=cfg["num_conv_pos_embedding_groups"], vocab_size=cfg["vocab_size"],
        layer_norm_eps=cfg["layer_norm_eps"], hidden_act=cfg["hidden_act"],
        num_feat_extract_layers=len(cfg["conv_dim"]),
    )
    if cfg.get("kind", "wav2vec2") == "conformer":
        hf_cfg = transformers.Wav2Vec2ConformerConfig(
            position_embeddings_type=cfg["position_embeddings_type"],
            conv_depthwise_kernel_size=cfg["conv_depthwise_kernel_size"],
            rotary_embedding_base=cfg["rotary_embedding_base"],
            max_source_positions=cfg["max_source_positions"], **common)
        model = transformers.Wav2Vec2ConformerForCTC(hf_cfg)
    else:
        hf_cfg = transformers.Wav2Vec2Config(do_stable_layer_norm=cfg["do_stable_layer_norm"], **common)
        model = transformers.Wav2Vec2ForCTC(hf_cfg)
    return model.eval()


def randomize_affine(model, seed: int = 1):
    """Random-init leaves LayerNorm/GroupNorm at (1, 0), conformer pos_bias_u/v at 0 and BatchNorm
    running stats at (0, 1); perturb them so parity tests exercise those terms."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("layer_norm.weight") or name.endswith("batch_norm.weight"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("layer_norm.bias") or name.endswith("batch_norm.bias"):
                p.add_(0.05 * torch.randn(p.shape, generator=g))
            elif name.endswith("pos_bias_u") or name.endswith("pos_bias_v"):
                p.add_(0.05 * torch.randn(p.shape, generator=g))
        for name, b in model.named_buffers():
            if name.endswith("running_mean"):
                b.add_(0.05 * torch.randn(b.shape, generator=g))
            elif name.endswith("running_var"):
                b.mul_(1.0 + 0.2 * torch.rand(b.shape, generator=g))
    return model


def state_dict_of(model) -> Dict[str, torch.Tensor]:
    """Detached fp32 state dict with the pos-conv weight_norm folded into a plain ``weight``."""
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for pref in ("wav2vec2.", "wav2vec2_conformer."):
        k0 = pref + "encoder.pos_conv_embed.conv.parametrizations.weight.original0"
        if k0 in sd:
            sd[pref + "encoder.pos_conv_embed.conv.weight"] = pos_conv_weight(sd, pref)
    return sd
