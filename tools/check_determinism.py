"""Run the encoder forward repeatedly in one process and report the first stage whose output changes between calls."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from _util import build_ours, load_model_golden
from oracle.model_ref import make_config
cuda = torch.device("cuda:0")
g, params, _ = load_model_golden("tiny_large")
enc, sep, heads, _ = build_ours(make_config("tiny_large"), int(g["n_spk"]), int(g["hidden_sep"]), int(g["vocab"]), params)
wav = torch.from_numpy(g["wav"]).to(cuda); mask = torch.from_numpy(g["mask"]).to(cuda)
runs = []
for it in range(4):
    if it == 2:
        junk = torch.full((64 * 1024 * 1024,), float("nan"), device=cuda); del junk
    with torch.no_grad():
        o = enc(wav, attention_mask=mask, output_hidden_states=True)
    runs.append([o[3].clone()] + [h.clone() for h in o.hidden_states] + [o[0].clone(), o[2].clone()])
names = ["extract_features"] + [f"hidden[{i}]" for i in range(len(runs[0]) - 3)] + ["last(x8)", "down(x4)"]
for it in range(1, 4):
    d = [float((a.float() - b.float()).abs().max()) for a, b in zip(runs[it], runs[0])]
    first = next((n for n, x in zip(names, d) if x != 0.0), None)
    print(f"run {it} vs run 0: first differing stage = {first}; max abs diffs = {[round(x, 6) for x in d]}")
