"""`SpeechEncoderDecoderModelLlama`: the composite model that wraps the hot path.

The reference's class (ref:models/modeling_speech_encoder_decoder_llama.py:94-900) owns the B200 hot path -- encoder,
separator, N CTC heads, HybridLoss, greedy CTC prefix -- and hands its result to a LLaMA decoder.  This is the same object
surface on the B200-native classes: same constructor `(config, encoder, decoder)`, sub-module names (`encoder`, `decoder`,
`separator`, `serialized_ctc`, `enc_to_dec_proj`, `losses`: reference checkpoints load with `strict=True`), `forward`
keywords and `Seq2SeqLMOutput` (+ `.ctc_per_head`, which the PCGrad trainer reads, ref:src/trainer_seq2seq.py:1082-1087),
`forward_ctc`, and the `ctc_bridge_type="ctcprompt"` bridge (ref :644-668).

The decoder itself is outside the hot path and is the stock `transformers.LlamaForCausalLM`: for the
configurations supported here the reference's modified copy (ref:models/modeling_llama.py:170-228) only differs in how it
receives the speech embeddings -- it splices them between <bos> and the text embeddings and runs a plain causal mask over
the joint sequence (its `encoder_attention_mask` argument is never applied, ref:models/modeling_llama.py:161) -- which is
`inputs_embeds` for the stock class.  Not reproduced (LLM-side research options, they raise): `instruct` prompts, the
cross-attention adapters, `talker_ctc_refine`, `ctc_bridge_type` "raw" / "softmax".

Row f2 (ref :566, :644-668): the reference's `generate` calls `forward` once per new token; only `encoder_outputs` is cached,
so the separator (2 x 499 LSTM steps), the N vocabulary GEMMs, the collapse and the prefix embedding are recomputed for
every token.  `generate()` here computes them ONCE, then decodes against the decoder's KV cache; `generate(recompute=True)`
keeps the reference's schedule for comparison (same tokens, tests/test_composite_gpu.py; timings in bench.py --mode sot).
"""
from typing import List, Optional, Tuple

import torch
from torch import nn
from transformers import LlamaForCausalLM
from transformers.modeling_outputs import BaseModelOutput, Seq2SeqLMOutput
from transformers.modeling_utils import PreTrainedModel
from transformers.models.speech_encoder_decoder.configuration_speech_encoder_decoder import SpeechEncoderDecoderConfig

from .ctc import CTC
from .greedy import build_multi_ctc_prefix_from_heads, ctc_remove_duplicates_and_blank, split_k_speakers_and_lengths
from .losses import HybridLoss
from .modeling_wavlm import WavLMModel
from .separator import Separator


def shift_tokens_right(input_ids: torch.Tensor, pad_token_id: int, decoder_start_token_id: int) -> torch.Tensor:
    """labels -> decoder inputs: prepend the start token, drop the last one, -100 -> pad (ref :61-77)."""
    if decoder_start_token_id is None or pad_token_id is None:
        raise ValueError("decoder_start_token_id and pad_token_id have to be set in the model configuration")
    out = torch.empty_like(input_ids)
    out[:, 0] = decoder_start_token_id
    out[:, 1:] = input_ids[:, :-1]
    return out.masked_fill(out == -100, pad_token_id)


class SpeechEncoderDecoderModelLlama(PreTrainedModel):
    config_class = SpeechEncoderDecoderConfig
    base_model_prefix = "speech_encoder_decoder"
    main_input_name = "inputs"
    supports_gradient_checkpointing = True

    def __init__(self, config: Optional[SpeechEncoderDecoderConfig] = None, encoder: Optional[PreTrainedModel] = None,
                 decoder: Optional[PreTrainedModel] = None):
        if config is None:
            if encoder is None or decoder is None:
                raise ValueError("Either a configuration or an encoder and a decoder has to be provided.")
            config = SpeechEncoderDecoderConfig.from_encoder_decoder_configs(encoder.config, decoder.config)
        elif not isinstance(config, self.config_class):
            raise ValueError(f"Config: {config} has to be of type {self.config_class}")
        config.tie_word_embeddings = False
        super().__init__(config)
        g = lambda name, default=None: getattr(config, name, default)
        unsupported = {"instruct": g("instruct", False), "talker_ctc_refine": g("talker_ctc_refine", False),
                       "decoder_cross_attention": g("decoder_cross_attention", False)}
        if any(unsupported.values()) or (g("ctc_bridge", False) and g("ctc_bridge_type", "raw") != "ctcprompt"):
            raise NotImplementedError(
                f"mtasr_b200.composite: {[k for k, v in unsupported.items() if v] or ['ctc_bridge_type=' + str(g('ctc_bridge_type'))]} "
                "are LLM-side options outside the encoder + serialized-CTC hot path (SURVEY 2); supported: talker_ctc with the "
                "plain SOT decoder input or the 'ctcprompt' bridge")
        self.encoder = encoder if encoder is not None else WavLMModel(config.encoder)
        self.decoder = decoder if decoder is not None else LlamaForCausalLM._from_config(config.decoder)
        self.encoder.config = self.config.encoder
        self.decoder.config = self.config.decoder
        self.ignore_token_id = g("ignore_token_id")
        self.pad_token_id = g("pad_token_id")
        self.sc_token_id = g("sc_token_id")
        self.eos_token_id = g("eos_token_id")
        self.talker_ctc = bool(g("talker_ctc", False))
        self.talker_numbers = int(g("talker_numbers", 2))
        self.ctc_bridge = bool(g("ctc_bridge", False))
        self.ctc_bridge_type = g("ctc_bridge_type", "raw")
        self.ctc_blank_id = config.decoder.vocab_size + 1            # CTC vocabulary = decoder vocabulary + blank (last id)
        enc_out = getattr(config.encoder, "output_hidden_size", config.encoder.hidden_size)
        if self.talker_ctc:
            self.separator = Separator(in_dim=enc_out, hidden_size=config.separator_hidden, talker_numbers=self.talker_numbers)
            self.serialized_ctc = nn.ModuleList(CTC(odim=self.ctc_blank_id, encoder_output_size=enc_out)
                                                for _ in range(self.talker_numbers))
        else:
            self.serialized_ctc = []
        self.encoder_output_dim = enc_out
        if enc_out != self.decoder.config.hidden_size:
            self.enc_to_dec_proj = nn.Linear(self.encoder.config.hidden_size, self.decoder.config.hidden_size)
        if self.encoder.get_output_embeddings() is not None:
            raise ValueError(f"The encoder {self.encoder} should not have a LM Head.")
        self.losses = HybridLoss(alpha=config.ctc_alpha, mode=config.train_mode, blank_id=self.ctc_blank_id - 1,
                                 enable_blank_check=True, log_every_steps=100)

    # ------------------------------------------------------------------------------------------ reference accessors
    def get_encoder(self):
        return self.encoder

    def get_decoder(self):
        return self.decoder

    def get_input_embeddings(self):
        return self.decoder.get_input_embeddings()

    def get_output_embeddings(self):
        return self.decoder.get_output_embeddings()

    def freeze_feature_encoder(self):
        self.encoder.freeze_feature_encoder()

    def ctc_remove_duplicates_and_blank(self, argmax_tensor, blank_id: int = 128258, pad_id: int = 128257,
                                        collapse_across_blanks: bool = True):
        return ctc_remove_duplicates_and_blank(argmax_tensor, blank_id, pad_id, collapse_across_blanks)

    # ------------------------------------------------------------------------------------------ hot-path pieces
    def _encode(self, inputs, attention_mask, encoder_outputs, **kw):
        if encoder_outputs is None:
            if inputs is None:
                raise ValueError("You have to specify either input_values or input_features")
            encoder_outputs = self.encoder(inputs, attention_mask=attention_mask, return_dict=True, **kw)
        elif isinstance(encoder_outputs, tuple):
            encoder_outputs = BaseModelOutput(*encoder_outputs)
        return encoder_outputs

    def _greedy_transcripts(self, sep_hidden_states) -> List[torch.Tensor]:
        """Per head: fused vocabulary-GEMM argmax -> device-side collapse (ref :644-652, :886-896)."""
        return [self.ctc_remove_duplicates_and_blank(head.argmax(x), blank_id=self.ctc_blank_id - 1, pad_id=self.pad_token_id)[0]
                for head, x in zip(self.serialized_ctc, sep_hidden_states)]

    def speech_context(self, encoder_outputs, attention_mask):
        """Everything the decoder receives from the audio, computed once: (speech embeddings incl. the CTC prompt prefix when
        the 'ctcprompt' bridge is on, separator streams, frame mask of the CTC heads)."""
        enc_h, frames = encoder_outputs[0], encoder_outputs[1]
        sep = self.separator(frames) if self.talker_ctc else None
        if self.encoder_output_dim != self.decoder.config.hidden_size:
            enc_h = self.enc_to_dec_proj(enc_h)
        ctc_mask = None
        if attention_mask is not None and self.talker_ctc:
            ctc_mask = self.encoder._get_feature_vector_attention_mask_x0(frames.shape[1], attention_mask)
        if self.ctc_bridge and self.ctc_bridge_type == "ctcprompt":
            prefix_embeds, _, _ = build_multi_ctc_prefix_from_heads(self._greedy_transcripts(sep), self.decoder, self.pad_token_id, None)
            enc_h = torch.cat([prefix_embeds.to(enc_h.dtype), enc_h], dim=1)
        return enc_h, sep, ctc_mask

    def _decoder_inputs(self, decoder_input_ids, speech):
        """<bos> | speech embeddings | remaining text embeddings (ref:models/modeling_llama.py:221-228)."""
        emb = self.decoder.get_input_embeddings()(decoder_input_ids)
        return torch.cat([emb[:, :1], speech.to(emb.dtype), emb[:, 1:]], dim=1)

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, inputs=None, attention_mask=None, prompt_ids=None, decoder_input_ids=None, decoder_attention_mask=None,
                encoder_outputs=None, past_key_values=None, decoder_inputs_embeds=None, labels=None, use_cache=None,
                output_attentions=None, output_hidden_states=None, input_values=None, input_features=None, return_dict=None,
                **kwargs):
        if decoder_attention_mask is not None or decoder_inputs_embeds is not None:
            raise NotImplementedError("mtasr_b200.composite: decoder_attention_mask / decoder_inputs_embeds are not supported "
                                      "(the reference runs a plain causal mask over <bos> | speech | text)")
        if inputs is None:
            if input_values is not None and input_features is not None:
                raise ValueError("You cannot specify both input_values and input_features at the same time")
            inputs = input_values if input_values is not None else input_features
        enc_kw = {k: v for k, v in kwargs.items() if not k.startswith("decoder_") and k != "num_items_in_batch"}
        encoder_outputs = self._encode(inputs, attention_mask, encoder_outputs, **enc_kw)
        speech, sep, ctc_mask = self.speech_context(encoder_outputs, attention_mask)
        label_spks = label_lens = None
        if labels is not None and decoder_input_ids is None:
            decoder_input_ids = shift_tokens_right(labels, self.config.pad_token_id, self.config.decoder_start_token_id)
            label_spks, label_lens = split_k_speakers_and_lengths(
                decoder_input_ids[:, 1:], self.talker_numbers, self.sc_token_id, self.config.pad_token_id, ignore_id=-100,
                end_token_id=self.config.pad_token_id, allow_empty_segment=False)
            B = labels.shape[0]
            # one more decoder position for <eos>: the input gets a pad, the labels get <eos> at their first ignored slot
            decoder_input_ids = torch.cat([decoder_input_ids, decoder_input_ids.new_full((B, 1), self.pad_token_id)], 1)
            labels = torch.cat([labels, labels.new_full((B, 1), self.ignore_token_id)], 1)
            eos = self.config.eos_token_id[0] if isinstance(self.config.eos_token_id, (list, tuple)) else self.config.eos_token_id
            first_ignored = (labels == self.ignore_token_id).float().argmax(dim=1)
            labels[torch.arange(B, device=labels.device), first_ignored] = eos
            labels = torch.cat([labels.new_full((B, speech.shape[1]), self.ignore_token_id), labels], 1)   # no loss on the speech span
        if decoder_input_ids is None:
            raise ValueError("decoder_input_ids or labels have to be given")
        dec = self.decoder(inputs_embeds=self._decoder_inputs(decoder_input_ids, speech), past_key_values=past_key_values,
                           use_cache=use_cache, output_attentions=output_attentions, output_hidden_states=output_hidden_states,
                           return_dict=True)
        loss = None
        if labels is not None:
            loss = self.losses(decoder_outputs=dec, labels=labels, decoder_vocab_size=self.decoder.config.vocab_size,
                               talker_ctc=self.serialized_ctc, sep_hidden_states=sep, encoder_attention_mask_ctc=ctc_mask,
                               label_spks=label_spks, label_spks_lengths=label_lens, talker_numbers=self.talker_numbers,
                               return_dict=True)
        out = Seq2SeqLMOutput(loss=loss, logits=dec.logits, past_key_values=dec.past_key_values,
                              decoder_hidden_states=dec.hidden_states, decoder_attentions=dec.attentions,
                              encoder_last_hidden_state=speech, encoder_hidden_states=encoder_outputs.hidden_states,
                              encoder_attentions=encoder_outputs.attentions)
        if self.losses.last_ctc_per_head is not None:
            out.ctc_per_head = self.losses.last_ctc_per_head
        return out

    @torch.no_grad()
    def forward_ctc(self, inputs=None, attention_mask=None, encoder_outputs=None, input_values=None, **kwargs) -> torch.Tensor:
        """CTC-only greedy decode (ref :873-900): collapsed ids of every head, concatenated along time."""
        if inputs is None:
            inputs = input_values
        encoder_outputs = self._encode(inputs, attention_mask, encoder_outputs)
        return torch.cat(self._greedy_transcripts(self.separator(encoder_outputs[1])), dim=1)

    # ------------------------------------------------------------------------------------------ greedy generation (f2)
    @torch.no_grad()
    def generate(self, inputs=None, attention_mask=None, max_new_tokens: int = 32, recompute: bool = False,
                 encoder_outputs=None) -> torch.Tensor:
        """Greedy SOT decoding, (B, 1 + max_new_tokens) ids starting with the decoder start token.

        recompute=False: separator, CTC heads, collapse and prompt prefix run ONCE; every new token is one decoder call on a
        single position against the KV cache.  recompute=True: the reference's schedule (ref :566, :644-668) -- each new token
        re-runs `forward` on the whole prefix with only `encoder_outputs` reused -- for parity and timing."""
        encoder_outputs = self._encode(inputs, attention_mask, encoder_outputs)
        B = encoder_outputs[0].shape[0]
        ids = torch.full((B, 1), self.config.decoder_start_token_id, dtype=torch.long, device=encoder_outputs[0].device)
        if recompute:
            for _ in range(max_new_tokens):
                logits = self.forward(encoder_outputs=encoder_outputs, attention_mask=attention_mask, decoder_input_ids=ids,
                                      use_cache=False).logits
                ids = torch.cat([ids, logits[:, -1].argmax(-1, keepdim=True)], 1)
            return ids
        speech, _, _ = self.speech_context(encoder_outputs, attention_mask)
        out = self.decoder(inputs_embeds=self._decoder_inputs(ids, speech), use_cache=True, return_dict=True)
        cache = out.past_key_values
        for step in range(max_new_tokens):
            nxt = out.logits[:, -1].argmax(-1, keepdim=True)
            ids = torch.cat([ids, nxt], 1)
            if step + 1 < max_new_tokens:
                out = self.decoder(input_ids=nxt, past_key_values=cache, use_cache=True, return_dict=True)
                cache = out.past_key_values
        return ids
