#!/usr/bin/env python
"""KD-step throughput of the B200-native path (BASELINE.json metric: KD-step utterances/s, distil-large-v3 student <-
large-v3 teacher, batch 32 x (80 x 3000 mel, 128 tokens) per GPU, bf16).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference's own CPU implementation (HF modules)

One JSON line on rank 0.  `value` = utterances/s with the batch already in HBM; `e2e` = the same step with the batch in
pinned HOST memory (H2D copy and the D2H read of the loss inside the timed region).  `roofline` is measured live with
CUDA events around every tcgen05 GEMM launch of the timed region.  Timing: barrier + synchronize on both sides, CUDA
events, max over ranks.  L2: the step's working set (4.6 GB of weights + activations) is far larger than L2, no flush.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STUDENT = dict(vocab_size=51866, num_mel_bins=80, d_model=1280, encoder_layers=32, encoder_attention_heads=20, encoder_ffn_dim=5120,
               decoder_layers=2, decoder_attention_heads=20, decoder_ffn_dim=5120, max_source_positions=1500,
               max_target_positions=448, pad_token_id=50256, decoder_start_token_id=50258)
TEACHER = dict(STUDENT, decoder_layers=32)
BATCH, N_TOK = 32, 128
# algorithmic FLOPs per utterance of the frozen-encoder recipe (variant B of BASELINE.md section 2): encoder fwd once,
# teacher decoder + LM head fwd, student decoder + LM head fwd + bwd (2x)
TF_PER_UTT_B = (2272.7 + 553.6 + 3 * 50.5) / 1e3
# variant A: trainable student encoder (fwd + bwd = 3x) and a separate teacher encoder forward
TF_PER_UTT_A = (3 * (2272.7 + 50.5) + 2272.7 + 553.6) / 1e3


def synthetic_batch(batch, n_tok, seed, dims, device="cpu"):
    """mel-like features in [-1, 1.5]; ragged label rows (length ~U[n_tok/2, n_tok]) whose tail is
    pad / -100; decoder_input_ids = labels shifted right with SOT (ref:training/run_distillation.py:460-476)."""
    g = torch.Generator().manual_seed(seed)
    feats = (0.5 * torch.randn((batch, dims["num_mel_bins"], 2 * dims["max_source_positions"]), generator=g)).clamp_(-1.0, 1.5)
    hi = min(dims["pad_token_id"], dims["decoder_start_token_id"]) - 1
    ids = torch.randint(0, hi, (batch, n_tok + 1), generator=g)
    ids[:, 0] = dims["decoder_start_token_id"]
    lens = torch.randint(n_tok // 2, n_tok + 2, (batch,), generator=g)
    lens[-1] = n_tok + 1
    pos = torch.arange(n_tok + 1)[None, :]
    valid = pos < lens[:, None]
    ids = torch.where(valid, ids, torch.full_like(ids, dims["pad_token_id"]))
    dec_in = ids[:, :-1].contiguous()
    labels = torch.where(valid[:, 1:], ids[:, 1:], torch.full_like(ids[:, 1:], -100)).contiguous()
    return {"input_features": feats.to(device), "decoder_input_ids": dec_in.to(device), "labels": labels.to(device)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (the recipe's clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        # the recipe's period (B200_PROFILING.md: -lms 200); DWB_BENCH_SMI_MS overrides for A/B runs (0 = no sampling)
        self.period_ms = int(os.environ.get("DWB_BENCH_SMI_MS", "200"))

    def start(self):
        try:
            if self.period_ms <= 0:
                return
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = max((float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()), default=None)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    except Exception:  # noqa: BLE001
        return 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"


def run_reference(args):
    """The reference's own CPU path (HF modules + train_step restatement) on a bounded sample of the workload.  Under
    torchrun (N > 1) rank 0 alone runs it; torchrun exports OMP_NUM_THREADS=1, so the thread count is set explicitly."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    from oracle.reference_step import ReferenceKDStep
    from oracle.whisper_oracle import WhisperDims
    b = args.cpu_batch
    ref = ReferenceKDStep(WhisperDims(**STUDENT), WhisperDims(**TEACHER), freeze_encoder=True)
    batch = synthetic_batch(b, N_TOK, 1234, STUDENT)
    sec, loss = ref.time_steps(batch, args.steps, args.warmup)
    val = b / sec
    cb = {"value": val, "unit": "utterances/s", "cores": ref.cores, "kind": ref.kind,
          "sample": f"{b} utterances per step (same models, 80x3000 mel, {N_TOK} tokens, fp32 on host cores)"}
    print(json.dumps({
        "impl": "reference", "metric": "kd_step_utterances_per_s", "value": val, "unit": "utterances/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "distil-large-v3 student + large-v3 teacher KD step, frozen+shared encoder (README recipe), "
                               f"{b}x(80x3000 mel, {N_TOK} tok) per step on CPU", "global_batch": b, "seq_len": N_TOK, "parallelism": "cpu"},
        "cpu_baseline": cb, "e2e": {"value": val, "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "loss": loss}))


def time_hf_gpu(variant, steps, warmup):
    """The reference's own GPU configuration (HF modules, fp32 student under bf16 autocast, bf16 teacher, sdpa, torch AdamW,
    ref:training/run_distillation.py:798-813,985-1004) on this box: the comparator SURVEY.md 8d calls the number to beat."""
    from oracle.reference_step import ReferenceKDStep
    from oracle.whisper_oracle import WhisperDims
    ref = ReferenceKDStep(WhisperDims(**STUDENT), WhisperDims(**TEACHER), freeze_encoder=variant == "B", device="cuda")
    batch = {k: v.cuda() for k, v in synthetic_batch(BATCH, N_TOK, 1234, STUDENT).items()}
    sec, loss = ref.time_steps(batch, steps, warmup)
    out = {"impl": "hf_gpu", "value": BATCH / sec, "unit": "utterances/s", "ms_per_step": sec * 1e3, "steps": steps, "warmup": warmup,
           "dtype": "bf16 autocast (fp32 master student, bf16 teacher)", "kind": ref.kind, "loss": loss,
           "workload": f"HF transformers Whisper modules + sdpa + torch AdamW on one B200, variant {variant}, {BATCH}x(80x3000, {N_TOK} tok), "
                       "device-resident batch, wall clock with synchronize on both sides"}
    del ref, batch
    return out


def run_hf_gpu(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    out = time_hf_gpu(args.variant, args.steps, max(args.warmup, 5))
    out.update({"metric": "kd_step_utterances_per_s", "n_gpus": 1})
    print(json.dumps(out))


def build_models(device, variant="B"):
    from distil_whisper_b200.modeling import DistilWhisperB200ForConditionalGeneration
    torch.manual_seed(0)
    with torch.device(device):
        student = DistilWhisperB200ForConditionalGeneration(STUDENT)
        teacher = DistilWhisperB200ForConditionalGeneration(TEACHER)
    teacher = teacher.to(torch.bfloat16)                                   # ref :985-992 teacher_dtype bf16
    if variant == "B":
        for p in student.model.encoder.parameters():                      # ref :1023-1026 --freeze_encoder
            p.requires_grad = False
    return student, teacher


def _timed(fn, n, world, finish=None):
    """Seconds per call: barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks.  `finish`
    joins work the calls left on other streams (the pipelined optimiser tail) before the end event is recorded."""
    from distil_whisper_b200 import ddp
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    if finish is not None:
        finish()
    e.record()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    return ddp.max_over_ranks(s.elapsed_time(e) / 1e3 / n)


def measure_kd_step(variant, steps, warmup, rank, local, world, use_graph=True, full=True):
    """One variant of the KD step on this rank's GPU.  full=True adds the eager replay with per-GEMM CUDA events (live
    roofline + launch count), the host-buffer e2e arm, the clock sampler and (N > 1) the gradient-parity check."""
    from distil_whisper_b200 import _abi, ops
    from distil_whisper_b200.kd import DistillationStep, PipelinedTrainer
    from distil_whisper_b200.optim import FusedAdamW
    from distil_whisper_b200 import ddp
    dev = torch.device("cuda", local)
    student, teacher = build_models(dev, variant)
    ddp.broadcast_parameters(student)
    ddp.broadcast_parameters(teacher)
    step = DistillationStep(student, teacher, kl_weight=1.0)
    opt = FusedAdamW.for_model(student, lr=1e-4, weight_decay=0.0, max_grad_norm=1.0)
    host_batch = {k: v.pin_memory() for k, v in synthetic_batch(BATCH, N_TOK, 1234 + rank, STUDENT).items()}
    dev_batch = {k: v.to(dev) for k, v in host_batch.items()}

    def eager_step(batch):
        loss, metrics = step.train_step(batch, temperature=2.0)
        loss.backward()
        opt.all_reduce_gradients()          # THE multi-GPU collective: one NCCL all-reduce of the flat student gradients
        opt.step()
        opt.zero_grad()
        return loss

    for _ in range(2):
        eager_step(dev_batch)
    trainer = PipelinedTrainer(step, opt, dev_batch, temperature=2.0) if use_graph else None

    def one_step(batch):
        if trainer is None:
            return eager_step(batch)
        return trainer.step(batch)          # graph replay(s) + all-reduce + clip + AdamW (overlapped with the next step's encoder)

    for _ in range(warmup):
        one_step(dev_batch)
    res = {"variant": variant}
    sampler = ClockSampler(local) if (full and rank == 0) else None
    if sampler:
        sampler.start()
    finish = trainer.flush if trainer is not None else None
    sec = _timed(lambda: one_step(dev_batch if trainer is None else None), steps, world, finish)
    res["clocks"] = sampler.stop() if sampler else None
    res["sec"] = sec
    res["launch_mode"] = "eager" if trainer is None else trainer.describe()
    if trainer is not None and full:
        # exposed part of the tail (all-reduce + clip + AdamW): CUDA events inside the trainer -- how long the main stream sat between
        # the end of graph E and the start of graph D waiting for the previous step's tail; the tail's own duration on its stream
        n2 = max(4, steps // 2)
        trainer.profile = []
        _timed(lambda: trainer.step(None), n2, world, finish)
        prof, trainer.profile = trainer.profile[1:], None
        if prof:
            res["exposed_comm_ms"] = {"value": sum(p[1].elapsed_time(p[5]) for p in prof) / len(prof),
                                      "tail_ms_on_side_stream": sum(p[3].elapsed_time(p[4]) for p in prof) / len(prof),
                                      "graph_E_ms": sum(p[0].elapsed_time(p[1]) for p in prof) / len(prof),
                                      "graph_D_ms": sum(p[5].elapsed_time(p[2]) for p in prof) / len(prof),
                                      "what": "main-stream wait between the end of graph E and the start of graph D (CUDA events), i.e. the part of "
                                              "all-reduce + clip + AdamW of the previous step that graph E did not cover; max over ranks not taken"}
    if full:
        # live GEMM roofline: the same K steps launched eagerly with a CUDA-event pair around every tcgen05 GEMM launch
        # (kernels inside a replayed graph cannot be bracketed by events); the library counts its own kernel launches
        _abi.launch_count(reset=True)
        ops.GEMM_PROFILE = []
        res["sec_eager"] = _timed(lambda: eager_step(dev_batch), steps, world)
        res["launches"] = _abi.launch_count(reset=True)
        prof, ops.GEMM_PROFILE = ops.GEMM_PROFILE, None
        res["gemm_ms"] = sum(s.elapsed_time(e) for s, e, _ in prof)
        res["gemm_flops"] = sum(f for _, _, f in prof)
        res["gemm_launches"] = len(prof)

        def e2e_step():
            b = host_batch if trainer is not None else {k: v.to(dev, non_blocking=True) for k, v in host_batch.items()}
            loss = one_step(b)                  # pinned host -> static device buffers (H2D) inside the step
            return float(loss.item())           # D2H read of the step's result
        e2e_step()
        res["sec_e2e"] = _timed(e2e_step, steps, world, finish)
        res["h2d_bytes"] = sum(v.numel() * v.element_size() for v in host_batch.values())
        if world > 1:
            res["ddp_parity"] = ddp_gradient_parity(step, opt, rank, world, dev)
    res["loss"] = float(one_step(dev_batch).item())
    if trainer is not None:
        trainer.flush()
    torch.cuda.synchronize()
    del trainer, step, opt, student, teacher, dev_batch
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return res


def ddp_gradient_parity(step, opt, rank, world, dev):
    """BASELINE.json configs[2] on hardware, outside the timed region: the NCCL all-reduced student gradient / N against the
    mean of the per-shard gradients recomputed serially on rank 0 (every rank's shard is regenerated from its seed).  The
    yardstick is the run-to-run noise of ONE shard's gradient on ONE GPU: dQ tiles are summed by fp32 TMA reduce-add in
    arrival order, and with a random-init model (near-uniform attention) dQ = sum_j dS_ij K_j is a cancelling sum, which
    turns fp32 ordering round-off into ~1e-3 of the gradient norm (profiles/r02_gradient_noise.md); every other kernel of
    the step is bitwise reproducible."""
    flat = opt.flat

    def shard_grad(r):
        flat.grad.zero_()
        b = {k: v.to(dev) for k, v in synthetic_batch(BATCH, N_TOK, 1234 + r, STUDENT).items()}
        step.forward_backward(b, 2.0)
        return flat.grad.clone()
    shard_grad(rank)
    opt.all_reduce_gradients()
    reduced = flat.grad.clone().mul_(1.0 / world)
    out = None
    if rank == 0:
        shards = [shard_grad(r) for r in range(world)]
        serial = torch.stack(shards).sum(0).mul_(1.0 / world)
        again = shard_grad(0)
        den = serial.double().norm()
        out = {"rel_err": float((reduced.double() - serial.double()).norm() / den),
               "run_to_run_noise": float((again.double() - shards[0].double()).norm() / shards[0].double().norm()),
               "grad_norm": float(den), "shards": world,
               "what": "||allreduce(grad)/N - mean_r grad_r|| / ||mean_r grad_r|| over the flat fp32 student gradient (shards recomputed "
                       "serially on rank 0); run_to_run_noise = the same shard's gradient computed twice on rank 0 (fp32 reduce-add order "
                       "of the attention dQ tiles) -- parity holds when rel_err is of the order of that noise"}
    flat.grad.zero_()
    torch.distributed.barrier()
    return out


def measure_logmel(n_clips=1024, n_mels=80, iters=5):
    """BASELINE.json configs[3]: 1024 x 480000-sample fp32 waveforms -> [1024, 80, 3000] fp32; algorithmic bytes 2.88 MB per clip."""
    from distil_whisper_b200.feature_extraction import WhisperFeatureExtractorB200
    fe = WhisperFeatureExtractorB200(n_mels)
    g = torch.Generator(device="cuda").manual_seed(7)
    wav = torch.randn((n_clips, 480000), device="cuda", generator=g) * 0.1
    out = torch.empty((n_clips, n_mels, 3000), device="cuda")
    for _ in range(3):
        fe.extract_device(wav, out=out)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fe.extract_device(wav, out=out)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    byts = n_clips * (480000 * 4 + n_mels * 3000 * 4)
    hbm = 6571.6
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm = float(json.load(f)["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    del wav, out
    torch.cuda.empty_cache()
    return {"workload": f"configs[3]: {n_clips} x 480000-sample fp32 clips -> [{n_clips}, {n_mels}, 3000] (2.95 GB in+out > L2)", "ms": ms,
            "gbs": byts / ms / 1e6, "hbm_peak_gbs": hbm, "frac_of_hbm": byts / ms / 1e6 / hbm, "clips_per_s": n_clips / ms * 1e3}


def measure_config5(batch=64, iters=3):
    """BASELINE.json configs[4]: distil-medium.en encoder-only forward + backward, batch 64, S = 1500 (3 x 1138.1 GF per clip)."""
    from distil_whisper_b200 import engine
    from distil_whisper_b200.modeling import DistilWhisperB200ForConditionalGeneration
    cfg = dict(vocab_size=51864, num_mel_bins=80, d_model=1024, encoder_layers=24, encoder_attention_heads=16, encoder_ffn_dim=4096,
               decoder_layers=2, decoder_attention_heads=16, decoder_ffn_dim=4096, max_source_positions=1500, max_target_positions=448,
               pad_token_id=50256, decoder_start_token_id=50257)
    with torch.device("cuda"):
        model = DistilWhisperB200ForConditionalGeneration(cfg)
    enc = model.model.encoder
    st = engine.state_of(enc)
    feats = (0.5 * torch.randn((batch, 80, 3000), device="cuda")).clamp_(-1, 1.5)
    denc = torch.randn((batch * 1500, 1024), device="cuda").bfloat16() * 1e-3

    def step():
        _, ctx = engine.encoder_forward(st, feats, save=True)
        engine.encoder_backward(st, ctx, denc)
        for p in enc.parameters():
            p.grad = None
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    tf = 3 * 1138.1 * batch / 1e3
    del model, enc, st, feats, denc
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    peak, _ = peaks()
    return {"workload": f"configs[4]: distil-medium.en encoder fwd+bwd, batch {batch}, S=1500", "ms": ms, "tflop_per_step": tf,
            "tflops": tf / ms * 1e3, "frac_of_peak": tf / ms * 1e3 / peak}


def measure_decode(batch=32, n_new=124, reps=3):
    """KV-cached greedy `generate()` -- the eval loop's call on the student and the pseudo-labelling loop's call
    on the large-v3-shaped teacher (ref:training/run_distillation.py:1526, run_pseudo_labelling.py:903) -- tokens/s at `batch` clips,
    128 tokens per row (4 initial + 124 generated, EOS never emitted), encoder included; best of `reps` calls after one that builds the
    decode session and its CUDA graph."""
    import time
    from distil_whisper_b200.modeling import DistilWhisperB200ForConditionalGeneration
    out = {"workload": f"greedy generate, {batch} x (80 x 3000 mel) -> 128 tokens per row, one CUDA-graph replay per token", "batch": batch}
    for name, dims, dtype in (("teacher_large_v3_bf16", TEACHER, torch.bfloat16), ("student_distil_large_v3_fp32", STUDENT, None)):
        torch.manual_seed(0)
        with torch.device("cuda"):
            m = DistilWhisperB200ForConditionalGeneration(dims)
        if dtype is not None:
            m = m.to(dtype)
        feats = synthetic_batch(batch, 8, 3, dims, device="cuda")["input_features"]
        kw = dict(max_new_tokens=n_new, eos_token_id=10 ** 6,
                  decoder_input_ids=torch.tensor([[50258, 50259, 50360, 50364]], device="cuda").expand(batch, -1))
        m.generate(feats, **kw)
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ids = m.generate(feats, **kw)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        steps = ids.shape[1] - 1
        out[name] = {"s_per_call": best, "tokens_per_s": batch * steps / best, "utterances_per_s": batch / best, "ms_per_token_step_incl_encoder": best / steps * 1e3}
        del m, feats
        import gc
        gc.collect()
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "hf_gpu"],
                    help="reference: the reference's CPU path (driver arm).  hf_gpu: HF modules on the GPU (bf16 autocast + sdpa) alone")
    ap.add_argument("--cpu-batch", type=int, default=2, help="utterances per CPU reference step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the sub-objects measured after the headline at N=1 (variant A, gpu_reference, logmel, config5)")
    ap.add_argument("--variant", default="B", choices=["A", "B"],
                    help="B: frozen + shared encoder (reference README / paper recipe, default).  A: trainable student encoder, "
                         "separate teacher encoder (BASELINE.md section 2)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "hf_gpu":
        return run_hf_gpu(args)
    if args.warmup < 3:
        args.warmup = 3

    from distil_whisper_b200 import _abi, ddp
    rank, local, world = ddp.init_from_env()
    torch.cuda.set_device(local)
    _abi.call("dwb_check_device")
    tf_per_utt = {"A": TF_PER_UTT_A, "B": TF_PER_UTT_B}
    r = measure_kd_step(args.variant, args.steps, args.warmup, rank, local, world, use_graph=not args.no_graph, full=True)
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    peak, peak_src = peaks()
    utt = BATCH * world
    sec, sec_eager, sec_e2e = r["sec"], r["sec_eager"], r["sec_e2e"]
    achieved = r["gemm_flops"] / (r["gemm_ms"] * 1e-3) / 1e12 if r["gemm_ms"] > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "gemm_dram_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    vname = "B (--freeze_encoder)" if args.variant == "B" else "A (trainable encoder)"
    out = {
        "metric": "kd_step_utterances_per_s", "value": utt / sec, "unit": "utterances/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "configs[1]: distil-large-v3 student + large-v3 teacher KD step (fwd student+teacher, fused CE+KL, "
                               "bwd, grad all-reduce, clip, AdamW), " + ("frozen+shared encoder = README/paper recipe (BASELINE.md variant B), " if args.variant == "B"
                               else "trainable student encoder + separate teacher encoder (BASELINE.md variant A), ") +
                               f"{BATCH}x(80x3000 mel, {N_TOK} tok) per GPU", "global_batch": utt, "seq_len": N_TOK,
                   "parallelism": f"dp{world}", "l2": "working set >> 126 MB L2, no flush", "variant": vname,
                   "algorithmic_tflop_per_utt": tf_per_utt[args.variant]},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": traffic,
                     "traffic_source": "profiles/gemm_dram_traffic.json: dram__bytes_read+write of ONE ncu --set full launch of the fused-QKV GEMM "
                                       "(M 48000, N 3840, K 1280; algorithmic 501 MB) -- a committed capture, not measured in this run",
                     "kernel": "gemm_bf16_2cta_kernel + gemm_bf16_tcgen05_kernel (every GEMM launch of the timed region)",
                     "peak_source": peak_src, "gemm_share_of_step": r["gemm_ms"] * 1e-3 / args.steps / sec_eager if sec_eager > 0 else None,
                     "measured_in": "eager replay of the same steps (ms_per_step_eager below)",
                     "gemm_launches": r["gemm_launches"]},
        "step_roofline": {"achieved_tflops_per_gpu": utt / sec * tf_per_utt[args.variant] / world,
                          "frac_of_peak": utt / sec * tf_per_utt[args.variant] / world / peak},
        "e2e": {"value": utt / sec_e2e, "unit": "utterances/s", "h2d_bytes_per_step": r["h2d_bytes"], "d2h_bytes_per_step": 4},
        "gpu_launches": r["launches"], "gpu_launches_counted": "inside libdwb.so (dwb_launch_count) over the eager replay of the timed steps",
        "clocks": r["clocks"], "loss": r["loss"], "ms_per_step_eager": sec_eager * 1e3, "launch_mode": r["launch_mode"],
        "exposed_comm_ms": r.get("exposed_comm_ms"),
    }
    if world > 1:
        out["ddp_parity"] = r.get("ddp_parity")
    if world == 1 and not args.no_extras:
        def guarded(name, fn):
            try:
                out[name] = fn()
            except Exception as ex:  # noqa: BLE001
                out[name] = {"failed": f"{type(ex).__name__}: {ex}"[:300]}
            torch.cuda.empty_cache()
        other = "A" if args.variant == "B" else "B"

        def other_variant():
            ro = measure_kd_step(other, max(3, args.steps // 2), 3, rank, local, world, use_graph=not args.no_graph, full=False)
            return {other: {"value": BATCH / ro["sec"], "unit": "utterances/s", "ms_per_step": ro["sec"] * 1e3,
                            "algorithmic_tflop_per_utt": tf_per_utt[other], "frac_of_peak": BATCH / ro["sec"] * tf_per_utt[other] / peak,
                            "loss": ro["loss"], "launch_mode": ro["launch_mode"]},
                    args.variant: {"value": utt / sec, "unit": "utterances/s", "ms_per_step": sec * 1e3,
                                   "algorithmic_tflop_per_utt": tf_per_utt[args.variant],
                                   "frac_of_peak": utt / sec * tf_per_utt[args.variant] / peak}}
        guarded("variants", other_variant)
        guarded("logmel", measure_logmel)
        guarded("config5", measure_config5)
        guarded("decode", measure_decode)
        guarded("gpu_reference", lambda: time_hf_gpu(args.variant, 20, 5))
        if isinstance(out.get("gpu_reference"), dict) and out["gpu_reference"].get("value"):
            out["gpu_reference"]["speedup_of_this_repo"] = (utt / sec_e2e) / out["gpu_reference"]["value"]
    if not args.no_cpu_baseline and world == 1:
        try:
            from oracle.reference_step import ReferenceKDStep
            from oracle.whisper_oracle import WhisperDims
            ref = ReferenceKDStep(WhisperDims(**STUDENT), WhisperDims(**TEACHER), freeze_encoder=True)
            b = args.cpu_batch
            s_cpu, _ = ref.time_steps(synthetic_batch(b, N_TOK, 1234, STUDENT), 2, 1)
            out["cpu_baseline"] = {"value": b / s_cpu, "unit": "utterances/s", "cores": ref.cores, "kind": ref.kind,
                                   "sample": f"2 timed steps of {b} utterances (same models and shapes, fp32, HF modules on host cores)"}
        except Exception as ex:  # noqa: BLE001
            out["cpu_baseline"] = {"value": None, "unit": "utterances/s", "cores": os.cpu_count(), "kind": "reference",
                                   "sample": f"failed: {type(ex).__name__}: {ex}"}
    print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
