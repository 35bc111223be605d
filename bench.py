#!/usr/bin/env python
"""Headline benchmark: WavLM-Large encoder + separator + serialized-CTC forward+backward throughput in audio-seconds/s.

    python bench.py [--gpus N] [--steps K] [--warmup W]              (N > 1: launched under torch.distributed.run)
    python bench.py --impl reference ...                             (the reference algorithm on the host CPU cores)

Workload = BASELINE.json configs[1] ("cfg2"): WavLM-Large (24 layers, D=1024) + Separator(896) + 2 CTC heads with the
Llama-3 vocabulary (V=128259) on synthetic LibriMix-shaped 2-speaker 10 s 16 kHz mixtures, batch 32 per GPU, bf16
operands / fp32 accumulation, feature encoder frozen (as in every reference run).  One step = one forward + backward of
the serialized-CTC loss through the repo's public modules (every gradient of encoder, separator and CTC heads is
produced; for N > 1 the NCCL gradient all-reduce is inside the step).  Prints ONE JSON line (see README / DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "encoder+serialized-CTC fwd+bwd audio-sec/s"
UNIT = "audio-s/s"


# BASELINE.json configs: cfg2 = the headline; cfg3 = 3-mix 15 s training step; cfg5 = 30 s forward + greedy collapse (B = 64
# over 8 GPUs = 8 per GPU)
PRESETS = {"cfg2": dict(speakers=2, seconds=10.0, batch=32, mode="train"),
           "cfg3": dict(speakers=3, seconds=15.0, batch=32, mode="train"),
           "cfg4": dict(speakers=2, seconds=10.0, batch=16, mode="sot"),
           "cfg5": dict(speakers=3, seconds=30.0, batch=8, mode="infer")}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="utterances per GPU")
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--speakers", type=int, default=2)
    ap.add_argument("--layers", type=int, default=0, help="debug only: override the number of encoder layers")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--mode", default="train", choices=["train", "infer", "sot"],
                    help="train: fwd+bwd of the serialized-CTC loss (the headline metric); infer: encoder + separator + greedy "
                         "CTC argmax + collapse (BASELINE config 5 shape); sot: hybrid fwd+bwd of the composite model with the "
                         "'ctcprompt' bridge and a LLaMA-3.2-1B-shaped random decoder (BASELINE config 4); neither is the headline")
    ap.add_argument("--profile-run", action="store_true", help="ncu helper: 1 warm-up + 1 step, no e2e/roofline/cpu legs")
    ap.add_argument("--config", default=None, choices=sorted(PRESETS),
                    help="BASELINE.json configuration preset (sets --speakers/--seconds/--batch/--mode); default = cfg2, the headline")
    ap.add_argument("--dropout", type=float, default=0.0,
                    help="encoder hidden / activation / attention dropout rate; > 0 runs the step in train() mode the way the "
                         "reference trains (0.1 each, separator LSTM dropout 0.2); the headline is quoted at 0 (eval mode, both arms)")
    ap.add_argument("--no-stock-gpu", action="store_true", help="skip the stock-PyTorch-on-GPU leg (oracle modules on the same GPU)")
    a = ap.parse_args()
    if a.config:
        for k, v in PRESETS[a.config].items():
            setattr(a, k, v)
    return a


def synth_batch(B, S, n_spk, vocab, seed):
    """LibriMix-shaped synthetic batch: sum of low-passed noise 'speakers' with random gains, per-utterance
    zero-mean / unit-variance (what Wav2Vec2FeatureExtractor(do_normalize=True) yields), fixed length; per-speaker targets
    of U(2,6) tokens per second with ids U[0, vocab-3), pad = vocab-2, blank = vocab-1, one forced repeat."""
    g = torch.Generator().manual_seed(seed)
    wav = torch.zeros(B, S)
    k = 0.85
    w = (k ** torch.arange(64, dtype=torch.float32)).flip(0)[None, None] * (1 - k)
    for _ in range(n_spk):
        n = torch.randn(B, S, generator=g)
        lp = torch.nn.functional.conv1d(torch.nn.functional.pad(n[:, None], (63, 0)), w)[:, 0]
        wav += lp * (0.5 + 0.5 * torch.rand(B, 1, generator=g))
    wav = (wav - wav.mean(1, keepdim=True)) / torch.sqrt(wav.var(1, keepdim=True, unbiased=False) + 1e-7)
    mask = torch.ones(B, S, dtype=torch.int32)
    sec = S / 16000.0
    lo, hi = max(1, int(2 * sec)), max(2, int(6 * sec))
    pad_id = vocab - 2
    labels, lens = [], []
    for _ in range(n_spk):
        L = torch.randint(lo, hi + 1, (B,), generator=g)
        y = torch.full((B, int(L.max())), pad_id, dtype=torch.long)
        for b in range(B):
            y[b, : L[b]] = torch.randint(0, vocab - 3, (int(L[b]),), generator=g)
            if L[b] >= 3:
                y[b, 2] = y[b, 1]
        labels.append(y)
        lens.append(L)
    return wav, mask, labels, lens


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed regions run."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)) or sorted(sm)
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def source_sha16():
    """Hash of the sources that determine the kernels and the launch sequence: a committed ncu traffic figure is only quoted
    for the build it was measured on."""
    import hashlib
    h = hashlib.sha256()
    pkg = os.path.join(ROOT, "multi-talker-asr-with-llms_b200")
    names = sorted(os.listdir(os.path.join(pkg, "csrc")))
    for f in [os.path.join(pkg, "csrc", n) for n in names if n.endswith((".cu", ".cuh"))] + \
            [os.path.join(pkg, n) for n in ("ops.py", "kernels.py", "modeling_wavlm.py", "separator.py", "ctc.py")]:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    except Exception:
        return 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"


# --------------------------------------------------------------------------------------------------- reference arm
def run_cpu_reference(args, steps, warmup, budget_s):
    """The reference algorithm on the host cores: the oracle's restatement of the reference modules (torch CPU fp32, the
    same library kernels the reference itself runs), Large encoder + separator + N CTC heads + serialized-CTC loss,
    fwd+bwd with the feature encoder frozen, on a bounded sample (few utterances per step) of the cfg2 workload."""
    from oracle.model_ref import RefCTC, RefSeparator, RefWavLMModel, ref_hybrid_ctc
    from mtasr_b200.configs import V_LLAMA3_CTC, wavlm_config
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    cfg = wavlm_config("large", **({"num_hidden_layers": args.layers} if args.layers else {}))
    S = int(args.seconds * 16000)
    enc = RefWavLMModel(cfg).eval()
    for p in enc.feature_extractor.parameters():
        p.requires_grad_(False)
    for p in enc.adapter.parameters():
        p.requires_grad_(False)
    sep = RefSeparator(cfg.hidden_size, 896, args.speakers).eval()
    heads = torch.nn.ModuleList(RefCTC(V_LLAMA3_CTC, cfg.hidden_size) for _ in range(args.speakers))
    params = [p for m in (enc, sep, heads) for p in m.parameters() if p.requires_grad]

    def step(Bc, seed):
        wav, mask, labels, lens = synth_batch(Bc, S, args.speakers, V_LLAMA3_CTC, seed)
        t0 = time.perf_counter()
        if args.mode == "infer":                          # ref ...llama.py:873-900: argmax per head + python collapse
            from oracle import host_ref
            with torch.no_grad():
                _, h, _, _ = enc(wav, mask.long())
                for hd, x in zip(heads, sep(h)):
                    host_ref.collapse(hd.argmax(x).tolist(), V_LLAMA3_CTC - 1, V_LLAMA3_CTC - 2)
            return time.perf_counter() - t0
        _, h, _, _ = enc(wav, mask.long())
        sp = sep(h)
        fm = enc.frame_mask_x0(h.shape[1], mask.long())
        loss, _ = ref_hybrid_ctc(heads, sp, fm, labels, lens)
        torch.autograd.grad(loss, params, allow_unused=True)
        return time.perf_counter() - t0

    t1 = step(1, 100)                                   # warm-up / calibration on one utterance
    n_steps = max(1, steps) + max(0, warmup - 1)
    Bc = 1
    for cand in (4, 2):
        if t1 * cand * n_steps <= budget_s:
            Bc = cand
            break
    for i in range(max(0, warmup - 1)):
        step(Bc, 200 + i)
    ts = [step(Bc, 300 + i) for i in range(max(1, steps))]
    dt = sum(ts) / len(ts)
    return dict(value=Bc * args.seconds / dt, ms_per_step=dt * 1e3, cores=cores, batch=Bc, steps=len(ts),
                sample=f"{Bc} utterance(s) x {args.seconds:g} s per step of the {workload_name(args).split(':')[0]} workload "
                       f"(WavLM-Large, {args.speakers} heads, V={V_LLAMA3_CTC}, mode {args.mode}), {len(ts)} timed step(s), "
                       f"torch CPU fp32, {cores} threads")


def run_stock_torch_gpu(args, dev):
    """SURVEY 8d "what you get today": the oracle's restatement of the reference modules (the transformers WavLM classes
    the reference instantiates + its separator / CTC / loss wiring, i.e. stock PyTorch / cuBLAS / cuDNN kernels) on the SAME
    GPU and workload, in fp32 (TF32 off: the reference's default precision, ref:slurm/template.slurm:99) and under
    torch.autocast(bf16) (its --bf16 option), with the largest batch <= the benched one that fits.  A yardstick printed
    beside `cpu_baseline`; nothing of the product path runs here."""
    from oracle.model_ref import RefCTC, RefSeparator, RefWavLMModel, ref_hybrid_ctc
    from mtasr_b200.configs import V_LLAMA3_CTC, wavlm_config
    torch.manual_seed(0)
    cfg = wavlm_config("large", **({"num_hidden_layers": args.layers} if args.layers else {}))
    S = int(args.seconds * 16000)
    enc = RefWavLMModel(cfg).to(dev).eval()
    for p in list(enc.feature_extractor.parameters()) + list(enc.adapter.parameters()):
        p.requires_grad_(False)
    sep = RefSeparator(cfg.hidden_size, 896, args.speakers).to(dev).eval()
    heads = torch.nn.ModuleList(RefCTC(V_LLAMA3_CTC, cfg.hidden_size) for _ in range(args.speakers)).to(dev)
    params = [p for m in (enc, sep, heads) for p in m.parameters() if p.requires_grad]
    prev = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    out = {}

    def step(batch, autocast):
        wav, mask, labels, lens = batch
        if args.mode == "infer":
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                _, h, _, _ = enc(wav, mask)
                return sum(hd.argmax(x.float()).sum() for hd, x in zip(heads, sep(h)))
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            _, h, _, _ = enc(wav, mask)
            sp = sep(h)
        fm = enc.frame_mask_x0(h.shape[1], mask)
        loss, _ = ref_hybrid_ctc(heads, [x.float() for x in sp], fm, labels, lens)
        torch.autograd.grad(loss, params, allow_unused=True)
        return loss

    try:
        for name, autocast in (("fp32", False), ("bf16_autocast", True)):
            Bc = args.batch
            while Bc >= 1:
                try:
                    wav, mask, labels, lens = synth_batch(Bc, S, args.speakers, V_LLAMA3_CTC, 77)
                    batch = (wav.to(dev), mask.to(dev).long(), [y.to(dev) for y in labels], [l.to(dev) for l in lens])
                    step(batch, autocast)                  # warm-up (cuDNN autotune, allocator)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    n = 2
                    e0.record()
                    for _ in range(n):
                        step(batch, autocast)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / n
                    out[name] = {"value": Bc * args.seconds / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "batch": Bc, "steps": n}
                    break
                except torch.OutOfMemoryError:
                    del batch
                    torch.cuda.empty_cache()
                    Bc //= 2
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev
    out["what"] = ("oracle modules (transformers WavLM + python-loop LSTM separator + nn.Linear/log_softmax/CTCLoss heads) with stock "
                   "PyTorch kernels on this GPU, same workload; fp32 = TF32 off")
    return out


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = run_cpu_reference(args, args.steps, args.warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, r["batch"], 1),
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(args):
    """Truthful description of what is being timed, derived from the arguments (not a fixed string)."""
    if args.layers:
        return f"DEBUG {args.layers}-layer encoder"
    shape = (args.speakers, args.seconds, args.mode)
    tag = {(2, 10.0, "train"): "cfg2", (3, 15.0, "train"): "cfg3", (3, 30.0, "infer"): "cfg5", (2, 10.0, "sot"): "cfg4"}.get(shape, "custom")
    what = {"train": "fwd+bwd, feature encoder frozen", "infer": "forward + greedy CTC argmax + collapse (no backward)",
            "sot": "composite model (ctcprompt bridge, LLaMA-3.2-1B-shaped random decoder from stock transformers under bf16 "
                   "autocast), hybrid loss fwd+bwd, feature encoder frozen"}[args.mode]
    return (f"{tag}: WavLM-Large + Separator(896) + serialized CTC ({args.speakers}mix), {args.seconds:g} s 16 kHz, {what}")


def workload_config(args, batch, world):
    return {"workload": workload_name(args),
            "encoder": "wavlm-large-shaped (24L, D=1024, H=16, F=4096), random init", "speakers": args.speakers,
            "seconds": args.seconds, "batch_per_gpu": batch, "global_batch": batch * world, "vocab": 128259,
            "separator_hidden": 896, "launch": getattr(args, "graph_note", None), "parallelism": f"dp{world}", "spec_augment": "off", "dropout": getattr(args, "dropout", 0.0),
            "module_mode": "train()" if getattr(args, "dropout", 0.0) > 0 and args.mode == "train" else "eval() (dropout inactive in both arms)",
            "l2": "working set per step (>1 GB of weights, >20 GB of activations) exceeds the 126 MB L2; no explicit flush"}


# --------------------------------------------------------------------------------------------------------- our arm
def main_ours(args):
    import torch.distributed as dist
    from mtasr_b200 import kernels as K
    from mtasr_b200.configs import V_LLAMA3_CTC, algorithmic_flops, wavlm_config
    from mtasr_b200.pipeline import SerializedCTCPath

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # SMs left to the NCCL all-reduce kernels during the backward (0 = none reserved).  The persistent GEMM / attention
    # kernels own whole SMs, so without a reservation the collective only gets SMs in the gaps between kernels and the
    # statically scheduled tiles of the kernel it displaces finish late.  Measured at 8 GPUs (cfg2, ms per step): no
    # gradient exchange 102.0, in-place reducer without reservation 115.5, with 8 reserved SMs + NCCL_MAX_CTAS=8 106.6.
    comm_sms = int(os.environ.get("MTASR_COMM_SMS", "8")) if world > 1 else 0
    total_sms = torch.cuda.get_device_properties(dev).multi_processor_count
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if comm_sms:
            os.environ.setdefault("NCCL_MAX_CTAS", str(comm_sms))
        # the reducer waits on every NCCL work itself (GradGroupReducer.finish): no record_stream bookkeeping on the
        # gradient tensors, whose deferred frees otherwise churn the caching allocator when the host runs ahead
        os.environ.setdefault("TORCH_NCCL_AVOID_RECORD_STREAMS", "1")
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(1234)
    over = {"num_hidden_layers": args.layers} if args.layers else {}
    if args.dropout > 0:
        over.update(hidden_dropout=args.dropout, activation_dropout=args.dropout, attention_dropout=args.dropout)
    cfg = wavlm_config("large", **over)
    S = int(args.seconds * 16000)
    B = args.batch
    if args.mode == "sot":
        # BASELINE configs[3]: the full SOT pipeline on the new encoder / CTC path (mtasr_b200.composite)
        from transformers import LlamaConfig
        from transformers.models.speech_encoder_decoder.configuration_speech_encoder_decoder import SpeechEncoderDecoderConfig
        from mtasr_b200.composite import SpeechEncoderDecoderModelLlama
        vdec = V_LLAMA3_CTC - 1                        # 128256 text ids + <sc> + <pad>
        dcfg = LlamaConfig(vocab_size=vdec, hidden_size=2048, intermediate_size=8192, num_hidden_layers=16, num_attention_heads=32,
                           num_key_value_heads=8, max_position_embeddings=4096, pad_token_id=vdec - 1, bos_token_id=128000,
                           eos_token_id=128001)
        ccfg = SpeechEncoderDecoderConfig.from_encoder_decoder_configs(cfg, dcfg)
        for k, v in dict(talker_ctc=True, talker_numbers=args.speakers, separator_hidden=896, ctc_alpha=0.7, train_mode="hybrid",
                         pad_token_id=vdec - 1, sc_token_id=vdec - 2, ignore_token_id=-100, eos_token_id=128001,
                         decoder_start_token_id=128000, instruct=False, ctc_bridge=True, ctc_bridge_type="ctcprompt").items():
            setattr(ccfg, k, v)
        with torch.device(dev):
            model = SpeechEncoderDecoderModelLlama(ccfg)
        model.release_graph = lambda: setattr(model.losses, "last_ctc_per_head", None)
    else:
        model = SerializedCTCPath(cfg, talker_numbers=args.speakers, separator_hidden=896, vocab_size=V_LLAMA3_CTC - 1).to(dev)
    model.encoder.freeze_feature_encoder()
    if args.mode != "sot":
        for p in model.encoder.adapter.parameters():   # the serialized-CTC loss does not depend on the adapter branch
            p.requires_grad_(False)
    model.eval()      # headline: dropout 0 / SpecAugment off in BOTH arms (eval-mode step; gradients still flow)
    if args.dropout > 0 and args.mode == "train":
        model.train()  # fused Philox-free counter-based dropout in the GEMM epilogues / attention kernels, LSTM dropout 0.2
    n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
    net = model
    # Data parallel: one process per GPU, the only collective is the gradient all-reduce (mean).  Default: the in-place
    # mtasr_b200.dp.GradGroupReducer (coalesced NCCL all-reduce of the gradient tensors where the backward kernels wrote
    # them, launched per group as the backward produces them).  MTASR_DP=ddp selects torch DistributedDataParallel
    # (flat buckets + pre-division: +3.4 ms of copies per step at cfg2), MTASR_DP=none no reduction at all (debug).
    dp_mode = os.environ.get("MTASR_DP", "group") if world > 1 else "single"
    reducer = None
    if dp_mode == "ddp":
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True,
                                                        bucket_cap_mb=int(os.environ.get("MTASR_DDP_BUCKET_MB", "100")),
                                                        broadcast_buffers=False)
    elif dp_mode == "group":
        from mtasr_b200.dp import GradGroupReducer
        for p_ in model.parameters():                       # same start on every rank (what DDP's constructor does)
            dist.broadcast(p_.data, src=0)
        reducer = GradGroupReducer(model.parameters(), group_bytes=int(os.environ.get("MTASR_DP_GROUP_MB", "64")) << 20)

    wav, mask, labels, lens = synth_batch(B, S, args.speakers, V_LLAMA3_CTC, seed=1234 + rank)
    host = [wav.pin_memory(), mask.pin_memory()] + [y.pin_memory() for y in labels] + [l.pin_memory() for l in lens]
    sot_host = [[labels[k][b, :int(lens[k][b])].tolist() for b in range(B)] for k in range(args.speakers)]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)
    ns = args.speakers

    def to_dev():
        d = [t.to(dev, non_blocking=True) for t in host]
        return d[0], d[1], d[2:2 + ns], d[2 + ns:2 + 2 * ns]

    graphed = [None]        # mtasr_b200.graphs.GraphedTrainStep once captured (MTASR_GRAPH=0: eager launches)

    def step(w, m, ys, yl):
        if args.mode == "infer":
            with torch.no_grad():
                ids = model.forward_ctc(w, attention_mask=m)
            return ids.sum().float()
        if graphed[0] is not None:
            return graphed[0](w, m, *ys, *yl)
        if args.mode == "sot":
            for p in model.parameters():
                p.grad = None
            # SOT labels: speaker 1 tokens, <sc>, speaker 2 tokens ..., -100 padded (ref:src/data_collator.py:47-50)
            rows = []
            for b in range(w.shape[0]):
                r = []
                for k in range(ns):
                    r += sot_host[k][b] + ([model.sc_token_id] if k + 1 < ns else [])
                rows.append(r)
            L = max(len(r) for r in rows)
            lab = torch.tensor([r + [-100] * (L - len(r)) for r in rows], device=w.device)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = model(inputs=w, attention_mask=m, labels=lab).loss
            loss.backward()
            return loss
        for p in model.parameters():
            p.grad = None
        if reducer is not None:
            reducer.begin()
        loss = net(w, attention_mask=m, label_spks=ys, label_spks_lengths=yl)
        if comm_sms:
            K.set_sm_budget(total_sms - comm_sms)           # the backward's persistent kernels leave `comm_sms` SMs to NCCL
        try:
            loss.backward()
        finally:
            if comm_sms:
                K.set_sm_budget(0)
        if reducer is not None:
            reducer.finish()                                # the current stream waits for the gradient all-reduce
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_issue = {}

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        host_issue["ms_per_step"] = (time.perf_counter() - t0) * 1e3 / steps   # time the host needs to ENQUEUE a step
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    dw, dm, dys, dyl = to_dev()
    graph_note = "eager launches"
    if args.mode == "train" and dp_mode != "ddp" and not args.profile_run and os.environ.get("MTASR_GRAPH", "1") != "0":
        # the whole step (forward + backward + gradient all-reduce) as one CUDA graph: the eager step is bound by the host
        # (see mtasr_b200/graphs.py); `eager_step` below still measures the per-kernel timings of the roofline leg
        from mtasr_b200.graphs import GraphedTrainStep
        try:
            graphed[0] = GraphedTrainStep(
                lambda w, m, *rest: net(w, attention_mask=m, label_spks=list(rest[:ns]), label_spks_lengths=list(rest[ns:])),
                [dw, dm, *dys, *dyl], model.parameters(), reducer=reducer,
                backward_sm_budget=(total_sms - comm_sms) if comm_sms else 0, release=model.release_graph)
            graph_note = "one CUDA graph per step (forward + backward" + (" + gradient all-reduce)" if reducer is not None else ")")
        except Exception as ex:                              # capture unsupported in this environment: stay eager, say so
            graphed[0] = None
            graph_note = f"eager launches (graph capture failed: {type(ex).__name__}: {str(ex)[:160]})"
            torch.cuda.synchronize()
    args.graph_note = graph_note
    if args.profile_run:
        step(dw, dm, dys, dyl)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("timed_step")
        step(dw, dm, dys, dyl)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        return
    for _ in range(max(3, args.warmup)):
        loss = step(dw, dm, dys, dyl)
    torch.cuda.synchronize()
    loss0 = loss.item()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = K.launch_count()
    ms = timed(lambda: step(dw, dm, dys, dyl), args.steps)
    launches = K.launch_count() - l0
    if graphed[0] is not None:                              # replays do not pass through the C ABI: kernels per captured step
        launches = graphed[0].launches_per_replay * args.steps

    e2e = None
    if not args.no_e2e:
        # end to end through the public API with HOST inputs: every step's batch is copied from pinned host memory and
        # every step's loss is read back on the host, both inside the timed region.  The copies are pipelined the way a
        # pinned-memory data loader does it (mtasr_b200.io): batch i+1 travels on a copy stream while step i computes, and
        # the loss of step i is read when step i+1 has been launched.
        from mtasr_b200.io import HostPrefetcher, ScalarReader
        pre, reader = HostPrefetcher(dev), ScalarReader()
        e2e_losses = []

        def e2e_loop(steps):
            staged = pre.stage(host)                   # step 0's inputs: copied inside the timed region as well
            pending = None
            for i in range(steps):
                d = pre.take(staged)
                if i + 1 < steps:
                    staged = pre.stage(host)
                loss = step(d[0], d[1], d[2:2 + ns], d[2 + ns:2 + 2 * ns])
                nxt = reader.submit(loss)
                if pending is not None:
                    e2e_losses.append(pending.result())
                pending = nxt
            e2e_losses.append(pending.result())

        e2e_loop(2)
        e2e_losses.clear()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_loop(args.steps)
        e1.record()
        barrier()
        t_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        ms_e2e = t_e2e.item()
        assert len(e2e_losses) == args.steps
        e2e = {"value": world * B * args.seconds * args.steps / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps}
    clocks = sampler.stop() if sampler else None

    # roofline of the dominant kernel (gemm_bf16_kernel): per-launch CUDA events on its stream over one more step
    g_keep, graphed[0] = graphed[0], None                   # eager launches for this leg: per-launch events on the GEMM
    K.profile_begin()
    step(dw, dm, dys, dyl)
    torch.cuda.synchronize()
    gemm_ms, exec_flops, gemm_launches = K.profile_end()
    graphed[0] = g_keep
    fl = algorithmic_flops(cfg, S, args.speakers, 896, V_LLAMA3_CTC, backward=args.mode != "infer",
                           adapter_backward=args.mode == "sot")   # (the decoder's own FLOPs run in library kernels: not counted)
    alg_step = fl["total"] * B
    # the QK^T / PV contractions run in the fused attention kernels and the recurrent half of the LSTM in the persistent
    # LSTM kernels, not in the GEMM kernel: not credited to it
    alg_gemm = (fl["total"] - fl["attention_total"] - fl["lstm_recurrent_total"]) * B
    peak, peak_src = peaks()
    ms_step = ms / args.steps
    # DRAM traffic of the kernel's launches of ONE step (sum over the launches, like `achieved`), from the committed ncu
    # metrics pass of this workload (profiles/gemm_traffic_r1.json); only quoted for the configuration it was captured on
    traffic = None
    traffic_src = "null: no ncu DRAM-bytes pass of THIS build / workload is committed (profiles/gemm_traffic_r2.json)"
    tpath = os.path.join(ROOT, "profiles", "gemm_traffic_r2.json")
    if os.path.exists(tpath) and not args.layers:
        with open(tpath) as f:
            tj = json.load(f)
        # only quoted when the capture was taken on this very build (hash of the kernel + op sources) and workload
        if (tj.get("source_sha16") == source_sha16() and tj.get("workload") == workload_name(args) and tj.get("batch") == B
                and tj.get("launches_per_step") == gemm_launches):
            traffic = tj["dram_bytes"]
            traffic_src = ("DRAM bytes (read + write) summed over the kernel's launches of one step, ncu dram__bytes_{read,write}.sum, "
                           "profiles/gemm_traffic_r2.json (same source hash, workload and launch count as this run)")
    roof = {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05/TMEM/TMA)", "achieved": alg_gemm / (gemm_ms / 1e3) / 1e12,
            "peak": peak, "unit": "TFLOP/s", "frac": alg_gemm / (gemm_ms / 1e3) / 1e12 / peak, "traffic": traffic,
            "traffic_unit": traffic_src,
            "algorithmic_tflop_per_step_in_kernel": alg_gemm / 1e12,
            "peak_source": peak_src, "launches_per_step": gemm_launches, "kernel_ms_per_step": gemm_ms,
            "kernel_share_of_step": gemm_ms / ms_step, "algorithmic_tflop_per_step": alg_step / 1e12,
            "executed_tflop_per_step": exec_flops / 1e12,
            "step_achieved_tflops": alg_step / (ms_step / 1e3) / 1e12, "step_frac": alg_step / (ms_step / 1e3) / 1e12 / peak}

    cpu = None
    stock = None
    if rank == 0 and world == 1 and not args.no_stock_gpu:
        del net, model
        net = model = None
        torch.cuda.empty_cache()
        try:
            stock = run_stock_torch_gpu(args, dev)
        except Exception as ex:                             # a yardstick must never take the bench line down
            stock = {"error": f"{type(ex).__name__}: {ex}"[:300]}
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        del net, model
        torch.cuda.empty_cache()
        r = run_cpu_reference(args, steps=2, warmup=1, budget_s=25.0)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    if rank == 0:
        line = {"metric": {"train": METRIC, "infer": "encoder+greedy-CTC forward audio-sec/s",
                           "sot": "full SOT (ctcprompt) hybrid fwd+bwd audio-sec/s"}[args.mode], "value": world * B * args.seconds * args.steps / (ms / 1e3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(args, B, world), "roofline": roof, "cpu_baseline": cpu, "stock_torch_gpu": stock, "e2e": e2e,
                "gpu_launches": int(launches), "host_issue_ms_per_step": host_issue.get("ms_per_step"), "clocks": clocks, "loss": loss0, "trainable_params": n_train,
                "gflop_per_audio_s": fl["total"] / args.seconds / 1e9}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Teardown.  A captured step holds NCCL kernels inside a CUDA graph; destroying the communicator while the graph
        # object is alive blocked for minutes in the 2-GPU run (the JSON line had already been printed).  Release the graph
        # first, and never let a stuck teardown turn a finished measurement into a time-out: the process leaves after 20 s.
        graphed[0] = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        killer = threading.Timer(20.0, lambda: os._exit(0))
        killer.daemon = True
        killer.start()
        try:
            dist.barrier()
            dist.destroy_process_group()
        finally:
            killer.cancel()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
