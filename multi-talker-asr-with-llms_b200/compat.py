"""transformers-5.x compatibility shim for running the REFERENCE's composite model.

The reference targets transformers ~4.47: `ref:utils/generation_utils.py:29-116` imports names that transformers 5.x removed
(`OffloadedCache`, `QuantizedCacheConfig`, `isin_mps_friendly`, `ExtensionsTrie`, the beam-search / constraint modules,
`_crop_past_key_values`, `NEED_SETUP_CACHE_CLASSES_MAPPING`, `QUANT_BACKEND_CLASSES_MAPPING`,
`HammingDiversityLogitsProcessor`, `LossKwargs`), `ref:models/llama_modules.py:117` looks up the removed
`ROPE_INIT_FUNCTIONS["default"]`, and `ref:models/llama_modules.py:316` dereferences an undefined
`ALL_ATTENTION_FUNCTIONS` unless the decoder runs with eager attention.  None of that is arithmetic of the hot path; it
only keeps `models/modeling_speech_encoder_decoder_llama.py` from importing.  This module

  * `install_transformers_shims()`  puts inert placeholders for the missing names into the transformers namespaces (classes
    that are only referenced in isinstance checks / type hints of generation modes this repo never uses, empty mappings,
    a TypedDict for `LossKwargs`) and registers the default RoPE initialiser;
  * `import_reference(root)`        imports the reference's composite module from a checkout, discovering any further missing
    name the same way (so a slightly different transformers 5.x release does not need an edit here);
  * `use_b200_classes(module)`      swaps the classes the reference's composite model instantiates -- `WavLMModel`, `Separator`,
    `CTC`, `HybridLoss`, the prefix builder, the label splitter and the greedy collapse -- for the B200-native ones of this
    package, so `SpeechEncoderDecoderModelLlama(config)` builds the hot path on the sm_100a kernels;
  * `force_eager_decoder_attention(model)`.

Pure Python; nothing of the reference is copied.  Used by oracle/gen_golden_composite.py (build container) and offered to
maintainers who run the reference on transformers 5.x.
"""
import importlib
import os
import re
import sys
import types
from typing import Optional, TypedDict

_PURGE = ("utils", "models", "modeling_llama", "modeling_wavlm", "llama_modules", "separator", "ctc", "losses", "ctc_prompt",
          "mt_ctctoken_builder", "down_sampling")


def _placeholder(name: str):
    if name == "LossKwargs":
        return TypedDict("LossKwargs", {}, total=False)
    if name.endswith("_MAPPING"):
        return {}
    if name[:1].isupper():
        return type(name, (), {"__init__": lambda self, *a, **k: None, "__doc__": "transformers-4.x name kept importable"})
    return lambda *a, **k: False


def _default_rope(config, device=None, seq_len=None, **kw):
    """transformers 4.x `_compute_default_rope_parameters`: inv_freq = base^(-2i/d), attention scaling 1."""
    import torch
    rp = getattr(config, "rope_parameters", None) or {}
    base = getattr(config, "rope_theta", None) or rp.get("rope_theta", 10000.0)
    dim = getattr(config, "head_dim", None) or config.hidden_size // config.num_attention_heads
    inv = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.int64).float().to(device) / dim))
    return inv, 1.0


KNOWN_MISSING = {
    "transformers.cache_utils": ["OffloadedCache", "QuantizedCacheConfig"],
    "transformers.pytorch_utils": ["isin_mps_friendly"],
    "transformers.tokenization_utils": ["ExtensionsTrie"],
    "transformers.generation.beam_constraints": ["DisjunctiveConstraint", "PhrasalConstraint"],
    "transformers.generation.beam_search": ["BeamScorer", "BeamSearchScorer", "ConstrainedBeamSearchScorer"],
    "transformers.generation.candidate_generator": ["_crop_past_key_values"],
    "transformers.generation.configuration_utils": ["NEED_SETUP_CACHE_CLASSES_MAPPING", "QUANT_BACKEND_CLASSES_MAPPING"],
    "transformers.generation.logits_process": ["HammingDiversityLogitsProcessor"],
    "transformers.utils": ["LossKwargs"],
}


def install_transformers_shims() -> list:
    """Idempotent.  Returns the (module, name) pairs that had to be added."""
    added = []
    for mod, names in KNOWN_MISSING.items():
        try:
            m = importlib.import_module(mod)
        except ImportError:
            m = types.ModuleType(mod)
            sys.modules[mod] = m
            parent, _, leaf = mod.rpartition(".")
            setattr(importlib.import_module(parent), leaf, m)
        for n in names:
            if not hasattr(m, n):
                setattr(m, n, _placeholder(n))
                added.append((mod, n))
    import transformers.modeling_rope_utils as rope
    if "default" not in rope.ROPE_INIT_FUNCTIONS:
        rope.ROPE_INIT_FUNCTIONS["default"] = _default_rope
        added.append(("transformers.modeling_rope_utils", "ROPE_INIT_FUNCTIONS['default']"))
    return added


def _purge_partial_imports():
    for k in list(sys.modules):
        if k in _PURGE or k.startswith(("utils.", "models.")):
            del sys.modules[k]


def import_reference(root: str, module: str = "models.modeling_speech_encoder_decoder_llama", max_fixes: int = 64):
    """Import `module` from a reference checkout at `root` under transformers 5.x."""
    for p in (os.path.join(root, "models"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    install_transformers_shims()
    for _ in range(max_fixes):
        try:
            return importlib.import_module(module)
        except ImportError as e:
            m = re.match(r"cannot import name '(\w+)' from '([\w\.]+)'", str(e))
            if m:
                setattr(importlib.import_module(m.group(2)), m.group(1), _placeholder(m.group(1)))
                _purge_partial_imports()
                continue
            m = re.match(r"No module named '(transformers[\w\.]*)'", str(e))
            if m:
                sys.modules[m.group(1)] = types.ModuleType(m.group(1))
                _purge_partial_imports()
                continue
            raise
    raise ImportError(f"could not import {module} from {root} after {max_fixes} compatibility fixes")


def use_b200_classes(module) -> None:
    """Point the names the reference's composite module instantiates at the B200-native classes (same constructors,
    forward signatures and state_dict keys)."""
    from . import ctc, greedy, losses, modeling_wavlm, mt_ctctoken_builder, separator
    module.WavLMModel = modeling_wavlm.WavLMModel
    module.Separator = separator.Separator
    module.CTC = ctc.CTC
    module.HybridLoss = losses.HybridLoss
    module.MultiSpkCTCTokenBuilder = mt_ctctoken_builder.MultiSpkCTCTokenBuilder
    module.build_multi_ctc_prefix_from_heads = greedy.build_multi_ctc_prefix_from_heads
    module.split_k_speakers_and_lengths = greedy.split_k_speakers_and_lengths
    cls = module.SpeechEncoderDecoderModelLlama
    cls.ctc_remove_duplicates_and_blank = staticmethod(
        lambda argmax_tensor, blank_id=128258, pad_id=128257, collapse_across_blanks=True:
        greedy.ctc_remove_duplicates_and_blank(argmax_tensor, blank_id, pad_id, collapse_across_blanks))


def force_eager_decoder_attention(model) -> None:
    """ref:models/llama_modules.py:307-316 only works with `_attn_implementation == "eager"` (the other branch
    dereferences an undefined name)."""
    for c in (getattr(model.config, "decoder", None), getattr(getattr(model, "decoder", None), "config", None), model.config):
        if c is not None:
            try:
                c._attn_implementation = "eager"
            except Exception:
                pass
