"""Encoder configurations the reference runs with, and the algorithmic work of the hot path.

The hub `config.json` files are not reachable offline; values follow `transformers` defaults for WavLM plus what
ref:utils/create_from_pretrained.py:202-212 sets (`encoder_add_adapter=True`, `feat_proj_dropout=0`,
`final_dropout=0`, `layerdrop=0`; `mask_time_prob=0.1` is SpecAugment, host-side, switched per run).
"""
from transformers.models.wavlm.configuration_wavlm import WavLMConfig

V_LLAMA3_CTC = 128259      # Llama-3 vocabulary (128256) + <sc> + <pad> + blank (ref ...llama.py:187-193)


def wavlm_config(kind: str = "large", **over) -> WavLMConfig:
    common = dict(add_adapter=True, feat_proj_dropout=0.0, final_dropout=0.0, layerdrop=0.0, mask_time_prob=0.0,
                  hidden_dropout=0.0, attention_dropout=0.0, activation_dropout=0.0)
    if kind == "large":          # microsoft/wavlm-large
        kw = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                  feat_extract_norm="layer", do_stable_layer_norm=True, conv_bias=False)
    elif kind == "base_plus":    # microsoft/wavlm-base-plus
        kw = dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                  feat_extract_norm="group", do_stable_layer_norm=False, conv_bias=False)
    else:
        raise ValueError(kind)
    kw.update(common)
    kw.update(over)
    return WavLMConfig(**kw)


def feat_lengths(n: int, cfg: WavLMConfig):
    out = []
    for k, s in zip(cfg.conv_kernel, cfg.conv_stride):
        n = (n - k) // s + 1
        out.append(n)
    return out


def algorithmic_flops(cfg: WavLMConfig, samples: int, n_spk: int, sep_hidden: int, vocab: int, backward: bool = True,
                      adapter_backward: bool = False) -> dict:
    """2*MAC of every dense contraction on the path for ONE utterance of `samples` samples (SURVEY 8d / BASELINE.md 3).
    Backward = 2x forward for every trainable GEMM; the conv feature extractor is frozen (forward only); recompute is
    never credited.  `adapter_backward` is False for the serialized-CTC loss (it does not depend on the adapter)."""
    L = feat_lengths(samples, cfg)
    T = L[-1]
    D, F, H = cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads
    cin = [1] + list(cfg.conv_dim[:-1])
    fe = sum(2.0 * L[i] * cfg.conv_dim[i] * cfg.conv_kernel[i] * cin[i] for i in range(len(L)))
    proj = 2.0 * T * cfg.conv_dim[-1] * D
    pos = 2.0 * T * D * (D // cfg.num_conv_pos_embedding_groups) * cfg.num_conv_pos_embeddings
    layer = 2.0 * T * (4 * D * D + 2 * D * F) + 4.0 * T * T * D
    enc = cfg.num_hidden_layers * layer
    attn = cfg.num_hidden_layers * 4.0 * T * T * D          # QK^T + PV part of `enc` (runs in the fused attention kernels)
    ad, t = 0.0, T
    for _ in range(cfg.num_adapter_layers):
        t = (t + 2 * 1 - cfg.adapter_kernel_size) // cfg.adapter_stride + 1
        ad += 2.0 * t * (2 * D) * (cfg.adapter_kernel_size * D)
    sep = 2.0 * T * D * sep_hidden + 2 * (2.0 * T * (2 * sep_hidden) * (4 * sep_hidden)) \
        + n_spk * (2.0 * T * sep_hidden * sep_hidden + 2.0 * T * sep_hidden * D)
    rec = 2 * (2.0 * T * sep_hidden * (4 * sep_hidden))        # recurrent half of the LSTM (runs in the persistent kernels)
    voc = n_spk * 2.0 * T * D * vocab
    fwd = fe + proj + pos + enc + ad + sep + voc
    trainable = proj + pos + enc + sep + voc + (ad if adapter_backward else 0.0)
    total = fwd + (2.0 * trainable if backward else 0.0)
    return dict(frames=T, fwd=fwd, total=total, attention_total=attn * (3.0 if backward else 1.0),
                lstm_recurrent_total=rec * (3.0 if backward else 1.0), fe=fe, proj=proj, posconv=pos, transformer=enc, adapter=ad, separator=sep,
                vocab=voc)
