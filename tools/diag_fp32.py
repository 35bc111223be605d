"""Diagnostic: fp32 parity mode (3 / 6 split terms) and the oracle's own fp32 run, both against the oracle in fp64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import copy
import torch
from _util import build_oracle, build_ours, perturb_, rel, oracle_run
from oracle.model_ref import make_config, synth_batch
from mtasr_b200 import precise

cuda = torch.device("cuda:0")
kind, S, B, layers = sys.argv[1], int(sys.argv[2]), 2, int(sys.argv[3])
torch.manual_seed(5)
cfg = make_config(kind, num_hidden_layers=layers)
enc, sep, heads, _ = build_ours(cfg, 2, 896, 515)
perturb_(enc, 3)
o = build_oracle(cfg, 2, 896, 515, ours=(enc, sep, heads))
wav, mask, labels, lens = synth_batch(B, S, 2, 515, seed=9, varlen=True)
wav, mask = wav.to(cuda), mask.to(cuda)
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
with torch.no_grad():
    last32, enc32, down32, feats32 = o[0](wav, mask)
    o64 = copy.deepcopy(o[0]).double()
    last64, enc64, down64, feats64 = o64(wav.double(), mask)
    fm = o[0].frame_mask_x0(enc32.shape[1], mask)
    m8 = enc._get_feature_vector_attention_mask(last32.shape[1], mask)
    m4 = enc._get_feature_vector_attention_mask_x4(down32.shape[1], mask)
    def errs(out, ref):
        return dict(feats=rel(out[3], ref[3], fm), enc=rel(out[1], ref[1], fm), last=rel(out[0], ref[0], m8), down=rel(out[2], ref[2], m4))
    ref64 = (last64, enc64, down64, feats64)
    print("oracle fp32 vs fp64      ", errs((last32, enc32, down32, feats32), ref64))
    for terms in (3, 6):
        precise.R = terms
        with precise.precision("fp32"):
            out = enc(wav, attention_mask=mask)
        print(f"ours terms={terms} vs fp64    ", errs(out, ref64))
        print(f"ours terms={terms} vs oracle32", errs(out, (last32, enc32, down32, feats32)))
