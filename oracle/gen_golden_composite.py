"""Golden fixture for SURVEY rows f1 / f2 from THE REFERENCE'S OWN composite model:  tests/golden/composite_tiny.npz.

Run in the build container only (needs /root/reference):
    python oracle/gen_golden_composite.py
`SpeechEncoderDecoderModelLlama` (ref:models/modeling_speech_encoder_decoder_llama.py:94-900) is imported under transformers
5.x through the compatibility shim `mtasr_b200.compat` (placeholders for removed transformers names only -- no arithmetic),
built from a tiny WavLM-Large-style encoder + a tiny LLaMA decoder with the reference's own classes, and run on CPU in fp32:
  * forward(train_mode="ctc")      -> loss, per-head CTC values            (ref :508-873, the serialized-CTC training step)
  * forward(train_mode="hybrid")   -> loss, decoder logits                 (ref :772-789, alpha * CE + (1 - alpha) * CTC)
  * forward_ctc                    -> collapsed greedy CTC token ids       (ref :873-900)
  * forward with ctc_bridge="ctcprompt" -> loss, logits, and the prefix the decoder was given  (ref :644-668)
  * teacher-forced greedy continuation through the reference forward, one call per new token with the full prefix re-fed
    (the reference's generation utilities themselves do not run under transformers 5.x: 4.8k lines against 4.47 internals),
    which is exactly the per-step recomputation row f2 removes.
Stored: inputs, labels, the full state_dict, all outputs.  Test infrastructure only.
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from mtasr_b200 import compat  # noqa: E402
from oracle.model_ref import make_composite_config, synth_batch  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "composite_tiny.npz")
V_TXT = 40           # text vocabulary 0..39, <sc> = 40, <pad> = 41  -> decoder vocab 42, CTC vocab 43 (blank = 42)


def build_config():
    cfg = make_composite_config(V_TXT)
    cfg.decoder.max_position_embeddings = 512
    return cfg


def main():
    torch.manual_seed(0)
    ref = compat.import_reference(REF)
    cfg = build_config()
    model = ref.SpeechEncoderDecoderModelLlama(cfg)
    compat.force_eager_decoder_attention(model)
    model.eval()
    with torch.no_grad():                       # make the gate / rel-pos / LayerNorm parameters non-trivial
        g = torch.Generator().manual_seed(1)
        for lyr in model.encoder.encoder.layers:
            a = lyr.attention
            a.gru_rel_pos_const.copy_(torch.rand(a.gru_rel_pos_const.shape, generator=g) + 0.5)
            a.gru_rel_pos_linear.bias.copy_(torch.randn(a.gru_rel_pos_linear.bias.shape, generator=g) * 0.5)
        w = model.encoder.encoder.layers[0].attention.rel_attn_embed.weight
        w.copy_(torch.randn(w.shape, generator=g) * 0.5)
        for head in model.serialized_ctc:       # spread the CTC logits so that the greedy path is not all-blank
            head.ctc_lo.weight.mul_(8.0)
    B, S = 3, 24000
    wav, mask, _, _ = synth_batch(B, S, 2, V_TXT + 3, seed=7, varlen=True)
    rs = np.random.RandomState(3)
    rows = []
    for b in range(B):
        s0 = rs.randint(3, V_TXT, size=rs.randint(3, 7)).tolist()
        s1 = rs.randint(3, V_TXT, size=rs.randint(2, 6)).tolist()
        rows.append(s0 + [V_TXT] + s1)
    L = max(len(r) for r in rows)
    labels = torch.tensor([r + [-100] * (L - len(r)) for r in rows])
    arrs = {"wav": wav, "mask": mask, "labels": labels}
    for k, v in model.state_dict().items():
        arrs["p:" + k] = v
    with torch.no_grad():
        model.losses.mode = "ctc"
        out = model(inputs=wav, attention_mask=mask, labels=labels.clone())
        arrs["ctc_loss"] = out.loss
        arrs["ctc_per_head"] = torch.stack(list(out.ctc_per_head))
        model.losses.mode = "hybrid"
        out = model(inputs=wav, attention_mask=mask, labels=labels.clone())
        arrs["hybrid_loss"] = out.loss
        arrs["hybrid_logits"] = out.logits
        arrs["enc_last"] = out.encoder_last_hidden_state
        arrs["forward_ctc"] = model.forward_ctc(inputs=wav, attention_mask=mask)
        # ctcprompt bridge (ref :644-668)
        model.ctc_bridge, model.ctc_bridge_type = True, "ctcprompt"
        out = model(inputs=wav, attention_mask=mask, labels=labels.clone())
        arrs["prompt_hybrid_loss"] = out.loss
        arrs["prompt_logits"] = out.logits
        arrs["prompt_enc_last"] = out.encoder_last_hidden_state          # [CTC prefix embeddings | speech embeddings]
        # greedy continuation, the reference way: every new token re-runs forward on the whole prefix (separator, the N
        # vocabulary GEMMs, collapse and prefix are recomputed inside each call; only `encoder_outputs` is reused)
        enc_out = model.encoder(wav, attention_mask=mask, return_dict=True)
        ids = torch.full((B, 1), cfg.decoder_start_token_id, dtype=torch.long)
        steps = 6
        for _ in range(steps):
            o = model(encoder_outputs=enc_out, attention_mask=mask, decoder_input_ids=ids, use_cache=False)
            nxt = o.logits[:, -1].argmax(-1, keepdim=True)
            ids = torch.cat([ids, nxt], 1)
        arrs["prompt_greedy_ids"] = ids
        arrs["prompt_last_logits"] = o.logits[:, -1]
    arrs["meta"] = np.array([V_TXT, cfg.separator_hidden, cfg.talker_numbers, steps])
    np.savez_compressed(OUT, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")
    print("ctc", float(arrs["ctc_loss"]), "hybrid", float(arrs["hybrid_loss"]), "prompt", float(arrs["prompt_hybrid_loss"]),
          "forward_ctc", tuple(arrs["forward_ctc"].shape), "greedy", arrs["prompt_greedy_ids"].tolist())


if __name__ == "__main__":
    main()
