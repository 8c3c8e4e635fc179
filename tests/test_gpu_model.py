"""End-to-end parity of the CUDA KD step (through the reference-shaped Python surface, every kernel behind the C ABI)
against the oracle and the committed HF 5.5.0 golden vectors.

Tolerances: the kernels compute in bf16 with fp32 accumulation (BASELINE.json config 2 is bf16); bf16 has an 8-bit
mantissa (eps 7.8e-3), so element-wise agreement with the fp32 oracle is bounded by ~1e-2 relative L2, not by the
1e-3 the north star quotes for fp16.  Scalars that average over many elements (loss, ce, kl) are held to 2e-3."""
import os

import numpy as np
import pytest
import torch

from oracle import whisper_oracle as wo

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

LOGITS_REL = 1.5e-2
GRAD_REL = 4e-2
LOSS_REL = 2e-3


def _rel(x, y):
    x, y = torch.as_tensor(x).double().cpu(), torch.as_tensor(y).double().cpu()
    return float((x - y).norm() / (y.norm() + 1e-30))


def _build(dims, sd, dtype=torch.float32, freeze_encoder=False):
    from distil_whisper_b200.modeling import DistilWhisperB200ForConditionalGeneration
    m = DistilWhisperB200ForConditionalGeneration(dims.to_dict())
    m.load_hf_state_dict({k: v.clone() for k, v in sd.items()})
    m = m.to("cuda", dtype)
    if freeze_encoder:
        for p in m.model.encoder.parameters():
            p.requires_grad = False
    return m


def _cuda(batch):
    return {k: v.cuda() for k, v in batch.items()}


def test_state_dict_names_match_hf_layout():
    sc = wo.PRESETS["tiny-student"]
    m = _build(sc, wo.init_state_dict(sc, 1))
    ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    expected = {k: tuple(v) for k, v in wo.param_shapes(sc).items()}
    expected["proj_out.weight"] = expected["model.decoder.embed_tokens.weight"]
    assert ours == expected
    assert m.proj_out.weight.data_ptr() == m.model.decoder.embed_tokens.weight.data_ptr()
    names = [n for n, _ in m.named_parameters()]
    assert "proj_out.weight" not in names and "model.decoder.embed_tokens.weight" in names   # tied weight listed once


def test_tiny_kd_step_matches_hf_golden(golden_dir):
    """Branch B (frozen + shared encoder, the reference recipe): loss / logits / grads vs HF golden."""
    from distil_whisper_b200.kd import DistillationStep
    g = np.load(os.path.join(golden_dir, "kd_tiny.npz"))
    s_seed, t_seed, b_seed = [int(x) for x in g["seeds"]]
    sc, tc = wo.PRESETS["tiny-student"], wo.PRESETS["tiny-teacher"]
    ssd, tsd = wo.init_state_dict(sc, s_seed), wo.init_state_dict(tc, t_seed)
    student = _build(sc, ssd, freeze_encoder=True)
    teacher = _build(tc, tsd, dtype=torch.bfloat16)
    step = DistillationStep(student, teacher, kl_weight=1.0, keep_logits=True)
    assert step.share_hidden_states and teacher.model.encoder is student.model.encoder
    batch = _cuda(wo.synthetic_batch(sc, batch=3, n_tok=12, seed=b_seed))
    loss, metrics = step.train_step(batch, temperature=2.0)
    loss.backward()
    assert abs(loss.item() - float(g["B_loss"])) / float(g["B_loss"]) < LOSS_REL
    assert abs(metrics["ce_loss"].item() - float(g["B_ce_loss"])) / float(g["B_ce_loss"]) < LOSS_REL
    # the teacher runs in bf16 weights (reference: teacher_dtype bf16) -> KL itself is a small number; 5% of it
    assert abs(metrics["kl_loss"].item() - float(g["B_kl_loss"])) / float(g["B_kl_loss"]) < 5e-2
    assert _rel(step.last_student_logits, g["B_student_logits"]) < LOGITS_REL
    assert _rel(step.last_teacher_logits, g["B_teacher_logits"]) < 2 * LOGITS_REL
    enc = step.last_encoder_states.float().view(3, -1, sc.d_model)
    assert _rel(enc, g["B_encoder_last_hidden_state"]) < LOGITS_REL
    # token-id argmax: exact except where the fp32 top-2 margin is inside the bf16 tolerance
    ref_logits = torch.from_numpy(g["B_student_logits"])
    top2 = ref_logits.topk(2, dim=-1).values
    safe = (top2[..., 0] - top2[..., 1]) > 4 * LOGITS_REL * ref_logits.abs().max()
    ours = step.last_student_logits.float().cpu().argmax(-1)
    assert (ours[safe] == ref_logits.argmax(-1)[safe]).all()
    names = [str(n) for n in g["B_grad_names"]]
    got = {n for n, p in student.named_parameters() if p.grad is not None and p.requires_grad}
    assert got == set(names), (sorted(got ^ set(names)))
    params = dict(student.named_parameters())
    worst = 0.0
    for n, norm in zip(names, g["B_grad_norms"]):
        gr = params[n].grad
        assert abs(float(gr.norm()) - norm) / (norm + 1e-12) < GRAD_REL, (n, float(gr.norm()), norm)
        key = f"B_grad::{n}"
        if key in g.files:
            r = _rel(gr, g[key])
            worst = max(worst, r)
            assert r < GRAD_REL, (n, r)
    print("worst grad rel err", worst)
    assert float(student.model.decoder.embed_tokens.weight.grad[sc.pad_token_id].abs().sum()) > 0


def test_tiny_kd_step_variant_a_trainable_encoder_matches_hf_golden(golden_dir):
    """Branch A: trainable student encoder + the teacher's own encoder (teacher_model(**batch), ref :1481)."""
    from distil_whisper_b200.kd import DistillationStep
    g = np.load(os.path.join(golden_dir, "kd_tiny.npz"))
    s_seed, t_seed, b_seed = [int(x) for x in g["seeds"]]
    sc, tc = wo.PRESETS["tiny-student"], wo.PRESETS["tiny-teacher"]
    student = _build(sc, wo.init_state_dict(sc, s_seed))
    teacher = _build(tc, wo.init_state_dict(tc, t_seed), dtype=torch.bfloat16)
    step = DistillationStep(student, teacher, kl_weight=1.0, keep_logits=True)
    assert not step.share_hidden_states and teacher.model.encoder is not student.model.encoder
    batch = _cuda(wo.synthetic_batch(sc, batch=3, n_tok=12, seed=b_seed))
    loss, metrics = step.train_step(batch, temperature=2.0)
    loss.backward()
    assert abs(loss.item() - float(g["A_loss"])) / float(g["A_loss"]) < LOSS_REL
    assert abs(metrics["kl_loss"].item() - float(g["A_kl_loss"])) / float(g["A_kl_loss"]) < 5e-2
    assert _rel(step.last_student_logits, g["A_student_logits"]) < LOGITS_REL
    assert _rel(step.last_teacher_logits, g["A_teacher_logits"]) < 2 * LOGITS_REL
    names = [str(n) for n in g["A_grad_names"]]
    params = dict(student.named_parameters())
    got = {n for n, p in params.items() if p.grad is not None and p.requires_grad}
    assert got == set(names), sorted(got ^ set(names))
    bad = []
    for n, norm in zip(names, g["A_grad_norms"]):
        gr = params[n].grad
        tol = 2 * GRAD_REL if (".q_proj." in n or ".k_proj." in n) else GRAD_REL
        if abs(float(gr.norm()) - norm) / (norm + 1e-12) > tol:
            bad.append((n, "norm", float(gr.norm()), norm))
        key = f"A_grad::{n}"
        if key in g.files and _rel(gr, g[key]) > tol:
            bad.append((n, "rel", _rel(gr, g[key])))
    assert not bad, bad


def test_generic_forward_backward_path_matches_oracle():
    """model(**batch).loss.backward() -- the un-fused HF-shaped path (CE only) -- against the fp32 oracle."""
    sc = wo.PRESETS["tiny-student"]
    ssd = wo.init_state_dict(sc, 3)
    student = _build(sc, ssd, freeze_encoder=True)
    batch = wo.synthetic_batch(sc, batch=2, n_tok=9, seed=4)
    out = student(**_cuda(batch))
    assert out.logits.shape == (2, 9, sc.vocab_size) and out.encoder_last_hidden_state.shape == (2, 50, sc.d_model)
    out.loss.backward()
    osd = {k: v.clone().requires_grad_(k.startswith("model.decoder")) for k, v in ssd.items()}
    ref = wo.model_forward(osd, sc, **batch)
    ref["loss"].backward()
    assert abs(out.loss.item() - ref["loss"].item()) / ref["loss"].item() < LOSS_REL
    assert _rel(out.logits, ref["logits"].detach()) < LOGITS_REL
    for n, p in student.named_parameters():
        if p.requires_grad:
            assert _rel(p.grad, osd[n].grad) < GRAD_REL, n
    # labels only -> decoder inputs built by shift_tokens_right (teacher call shape, ref :1477-1478)
    with torch.no_grad():
        out2 = student(encoder_outputs=(out.encoder_last_hidden_state,), labels=batch["labels"].cuda())
        ref2 = wo.model_forward(ssd, sc, labels=batch["labels"], encoder_hidden_states=ref["encoder_last_hidden_state"].detach())
    assert _rel(out2.logits, ref2["logits"]) < LOGITS_REL
    with pytest.raises(ValueError):
        student(input_features=torch.zeros(1, sc.num_mel_bins, 98, device="cuda"), decoder_input_ids=batch["decoder_input_ids"][:1].cuda())


def test_plumbing_config_small_en_vs_oracle():
    """BASELINE.json configs[0]: distil-small.en-shaped student (12/4/768) + 12/12 teacher, 2 x (80 x 3000), 32 labels."""
    from distil_whisper_b200.kd import DistillationStep
    sc, tc = wo.PRESETS["distil-small.en"], wo.PRESETS["small.en-teacher"]
    ssd, tsd = wo.init_state_dict(sc, 5), wo.init_state_dict(tc, 6)
    for k in list(tsd):
        if k.startswith("model.encoder."):
            tsd[k] = ssd[k]
    batch = wo.synthetic_batch(sc, batch=2, n_tok=32, seed=7)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for k, v in ssd.items():
        v.requires_grad_(k.startswith("model.decoder"))
    loss_ref, m_ref, so, to = wo.kd_train_step(ssd, sc, tsd, tc, batch, 2.0, 1.0, share_hidden_states=True)
    loss_ref.backward()
    student = _build(sc, {k: v.detach() for k, v in ssd.items()}, freeze_encoder=True)
    teacher = _build(tc, tsd, dtype=torch.bfloat16)
    step = DistillationStep(student, teacher, keep_logits=True)
    loss, metrics = step.train_step(_cuda(batch), temperature=2.0)
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) / loss_ref.item() < LOSS_REL, (loss.item(), loss_ref.item())
    assert _rel(step.last_student_logits, so["logits"].detach()) < LOGITS_REL
    bad = []
    for n, p in student.named_parameters():
        if p.requires_grad:
            r = _rel(p.grad, ssd[n].grad)
            # softmax-logit gradients (q / k projections) go through the P * (dP - delta) cancellation: 2x looser in bf16
            tol = 2 * GRAD_REL if (".q_proj." in n or ".k_proj." in n) else GRAD_REL
            if r > tol:
                bad.append((n, r))
    assert not bad, bad


def test_fused_adamw_and_second_step_changes_loss():
    from distil_whisper_b200.kd import DistillationStep
    from distil_whisper_b200.optim import FusedAdamW
    sc, tc = wo.PRESETS["tiny-student"], wo.PRESETS["tiny-teacher"]
    student = _build(sc, wo.init_state_dict(sc, 11), freeze_encoder=True)
    teacher = _build(tc, wo.init_state_dict(tc, 23), dtype=torch.bfloat16)
    step = DistillationStep(student, teacher)
    opt = FusedAdamW.for_model(student, lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
    assert len(opt.param_groups) == 2
    n_decay = sum(p.numel() for p in opt.param_groups[0]["params"])
    n_nodecay = sum(p.numel() for p in opt.param_groups[1]["params"])
    assert n_decay > n_nodecay > 0
    batch = _cuda(wo.synthetic_batch(sc, batch=3, n_tok=12, seed=5))
    losses = []
    for _ in range(4):
        loss, _ = step.train_step(batch, 2.0)
        loss.backward()
        opt.all_reduce_gradients()
        opt.step()
        opt.zero_grad()
        losses.append(loss.item())
    assert losses[-1] < losses[0], losses
    assert float(opt.flat.grad.abs().sum()) == 0.0


def test_cuda_graph_replay_matches_eager():
    """GraphedDistillationStep (fwd + loss + bwd as one CUDA graph) == train_step + loss.backward(), across weight updates."""
    from distil_whisper_b200.kd import DistillationStep, GraphedDistillationStep
    from distil_whisper_b200.optim import FusedAdamW
    sc, tc = wo.PRESETS["tiny-student"], wo.PRESETS["tiny-teacher"]

    def make():
        student = _build(sc, wo.init_state_dict(sc, 11), freeze_encoder=True)
        teacher = _build(tc, wo.init_state_dict(tc, 23), dtype=torch.bfloat16)
        step = DistillationStep(student, teacher)
        return step, FusedAdamW.for_model(student, lr=1e-3, max_grad_norm=1.0)
    b1 = _cuda(wo.synthetic_batch(sc, batch=3, n_tok=12, seed=5))
    b2 = _cuda(wo.synthetic_batch(sc, batch=3, n_tok=12, seed=6))
    step_e, opt_e = make()
    step_g, opt_g = make()
    graphed = GraphedDistillationStep(step_g, b1, temperature=2.0)
    assert float(opt_g.flat.grad.abs().sum()) == 0.0          # capture and warm-up left no gradient behind
    for it, batch in enumerate((b1, b2, b1)):
        le, _ = step_e.train_step(batch, 2.0)
        le.backward()
        lg, _ = graphed(batch)
        assert abs(le.item() - lg.item()) < (2e-5 if it == 0 else 1e-3) * abs(le.item())      # weights drift apart by atomics-order round-off
        # identical weights on the first pass: only the fp32 atomics order differs (LN / bias column sums, embedding scatter);
        # afterwards AdamW's m / sqrt(v) has amplified that noise into the weights (measured up to 4e-4 on the gradients)
        assert _rel(opt_g.flat.grad, opt_e.flat.grad) < (2e-4 if it == 0 else 5e-3)
        opt_e.step()
        opt_g.step()
    assert _rel(opt_g.flat.data, opt_e.flat.data) < 1e-3          # a skipped / doubled optimiser step would show as ~1e-1


def test_greedy_generate_follows_the_oracle_argmax():
    """generate() (ref:training/run_distillation.py:1526 eval path, greedy): every token it emits must be the oracle's
    arg-max on the same prefix, up to the bf16 logit tolerance (random-init logits have small top-2 margins, so the check
    is 'the oracle's logit of our token is within tolerance of the oracle's maximum', position by position); rows stop at
    EOS and are padded; the prompt is preserved; unsupported decoding modes raise."""
    sc = wo.PRESETS["tiny-student"]
    sd = wo.init_state_dict(sc, 5)
    m = _build(sc, sd)
    batch = wo.synthetic_batch(sc, batch=3, n_tok=4, seed=17)
    feats = batch["input_features"]
    prompt = torch.tensor([[sc.decoder_start_token_id, 7], [sc.decoder_start_token_id, 11], [sc.decoder_start_token_id, 3]])
    out = m.generate(feats.cuda(), decoder_input_ids=prompt.cuda(), max_new_tokens=10, eos_token_id=10 ** 6)   # eos never hit
    assert out.shape == (3, 12) and out.dtype == torch.long
    assert torch.equal(out[:, :2].cpu(), prompt)
    with torch.no_grad():
        ref = wo.model_forward(sd, sc, input_features=feats, decoder_input_ids=out[:, :-1].cpu())["logits"]
    tol = LOGITS_REL * float(ref.abs().max())
    for t in range(1, out.shape[1] - 1):
        chosen = ref[torch.arange(3), t, out[:, t + 1].cpu()]
        assert (chosen >= ref[:, t].max(-1).values - tol).all(), t
    # EOS handling: declare the first generated token of row 0 to be EOS -> the rest of that row is padding
    eos = int(out[0, 2])
    out2 = m.generate(feats.cuda(), decoder_input_ids=prompt.cuda(), max_new_tokens=6, eos_token_id=eos, pad_token_id=0)
    assert int(out2[0, 2]) == eos and (out2[0, 3:] == 0).all()
    # default prompt = decoder_start_token_id, max_length from the config; training mode is restored
    m.train()
    out3 = m.generate(feats.cuda(), max_length=5)
    assert out3.shape[1] <= 5 and (out3[:, 0] == sc.decoder_start_token_id).all() and m.training
    with pytest.raises(NotImplementedError):
        m.generate(feats.cuda(), num_beams=4)
    with pytest.raises(ValueError):
        m.generate(feats.cuda(), return_timestamps=True)          # no no_timestamps_token_id in the generation config


def _tiny_pair(freeze_encoder=True, lr=1e-3):
    from distil_whisper_b200.kd import DistillationStep
    from distil_whisper_b200.optim import FusedAdamW
    sc, tc = wo.PRESETS["tiny-student"], wo.PRESETS["tiny-teacher"]
    student = _build(sc, wo.init_state_dict(sc, 11), freeze_encoder=freeze_encoder)
    teacher = _build(tc, wo.init_state_dict(tc, 23), dtype=torch.bfloat16)
    step = DistillationStep(student, teacher)
    return sc, step, FusedAdamW.for_model(student, lr=lr, weight_decay=0.01, max_grad_norm=1.0)


@pytest.mark.parametrize("freeze_encoder", [True, False])
def test_pipelined_trainer_matches_the_eager_loop_from_pinned_host_batches(freeze_encoder):
    """PipelinedTrainer (graph E | graph D | all-reduce + clip + AdamW on a side stream under the next E) == the reference loop
    body (ref :1606-1614) issued eagerly, with batches arriving in pinned HOST memory (the bench's e2e path), and with
    gradient_accumulation_steps = 2 (1/2 folded into the loss kernel, one optimiser step per two micro-batches)."""
    from distil_whisper_b200.kd import PipelinedTrainer
    sc, step_e, opt_e = _tiny_pair(freeze_encoder)
    _, step_p, opt_p = _tiny_pair(freeze_encoder)
    host = [{k: v.pin_memory() for k, v in wo.synthetic_batch(sc, batch=3, n_tok=12, seed=s).items()} for s in (5, 6, 7, 8, 9, 10)]
    trainer = PipelinedTrainer(step_p, opt_p, _cuda(host[0]), temperature=2.0, gradient_accumulation_steps=2)
    assert float(opt_p.flat.grad.abs().sum()) == 0.0
    assert ("student" in trainer.pre) == freeze_encoder and ("teacher" in trainer.pre) == (not freeze_encoder)
    for i, hb in enumerate(host):
        le, _ = step_e.train_step(_cuda(hb), 2.0, loss_scale=0.5)
        le.backward()
        lp = trainer.step(hb)
        assert abs(le.item() - lp.item()) < (5e-5 if i < 2 else 1e-3) * abs(le.item()), (i, le.item(), lp.item())
        if i % 2 == 1:
            opt_e.all_reduce_gradients()
            opt_e.step()
            opt_e.zero_grad()
    trainer.flush()
    torch.cuda.synchronize()
    assert opt_p.step_count == opt_e.step_count == 3
    # measured 1e-5 .. 2e-4: the fp32-atomics ordering noise of the gradients (dQ, LayerNorm / bias column sums) after three
    # AdamW steps, whose m / sqrt(v) normalisation amplifies it on near-zero gradients
    assert _rel(opt_p.flat.data, opt_e.flat.data) < 1e-3          # a skipped / doubled optimiser step would show as ~1e-1
    assert float(opt_p.flat.grad.abs().sum()) == 0.0


def test_upstream_gradient_of_backward_is_honoured():
    """accelerator.backward(loss) divides the loss by gradient_accumulation_steps before .backward() (ref :1607-1609): the
    gradient handed to _KDStepFn.backward must scale d loss / d logits, not be dropped."""
    sc, step, opt = _tiny_pair()
    batch = _cuda(wo.synthetic_batch(sc, batch=3, n_tok=12, seed=5))
    loss, _ = step.train_step(batch, 2.0)
    loss.backward()
    g1 = opt.flat.grad.clone()
    opt.flat.grad.zero_()
    loss, _ = step.train_step(batch, 2.0)
    (loss / 4).backward()
    assert _rel(opt.flat.grad, g1 / 4) < 1e-2              # dlogits are re-rounded to bf16 after the scale
    opt.flat.grad.zero_()
    loss, _ = step.train_step(batch, 2.0, loss_scale=0.25)
    loss.backward()
    assert _rel(opt.flat.grad, g1 / 4) < 2e-3


def test_fused_adamw_state_dict_round_trip_resumes_identically():
    """accelerator.save_state / load_state (ref :1559, :1640) go through optimizer.state_dict(): save -> load into a fresh
    optimiser -> the next steps match the uninterrupted run; the layout is torch.optim.AdamW's."""
    import io
    sc, step_a, opt_a = _tiny_pair()
    _, step_b, opt_b = _tiny_pair()
    batches = [_cuda(wo.synthetic_batch(sc, batch=3, n_tok=12, seed=s)) for s in (5, 6, 7)]

    def run(step, opt, b):
        loss, _ = step.train_step(b, 2.0)
        loss.backward()
        opt.step()
        opt.zero_grad()
    run(step_a, opt_a, batches[0])
    run(step_a, opt_a, batches[1])
    buf = io.BytesIO()
    torch.save({"opt": opt_a.state_dict(), "model": step_a.student.state_dict()}, buf)
    buf.seek(0)
    ck = torch.load(buf)
    st = ck["opt"]["state"]
    assert all(set(v) == {"step", "exp_avg", "exp_avg_sq"} for v in st.values()) and len(st) == len(opt_a.flat.layout)
    step_b.student.load_state_dict(ck["model"])
    opt_b.load_state_dict(ck["opt"])
    assert opt_b.step_count == 2
    run(step_a, opt_a, batches[2])
    run(step_b, opt_b, batches[2])
    # (not bit-identical: dQ / LayerNorm / bias gradients are accumulated with fp32 atomics whose order varies run to run)
    assert _rel(opt_b.exp_avg, opt_a.exp_avg) < 2e-3 and _rel(opt_b.exp_avg_sq, opt_a.exp_avg_sq) < 2e-3
    assert _rel(opt_b.flat.data, opt_a.flat.data) < 1e-3          # zeroed moments / a reset step count would show as ~1e-1
    # a torch.optim.AdamW over the same parameters accepts the checkpoint (same per-parameter layout)
    ta = torch.optim.AdamW([{"params": g["params"]} for g in opt_a.param_groups], lr=1e-3)
    ta.load_state_dict({k: v for k, v in opt_a.state_dict().items() if k != "dwb"})


def test_out_of_range_label_or_token_id_poisons_the_loss_instead_of_reading_out_of_bounds():
    sc, step, opt = _tiny_pair()
    batch = _cuda(wo.synthetic_batch(sc, batch=3, n_tok=12, seed=5))
    bad = {k: v.clone() for k, v in batch.items()}
    bad["labels"][1, 2] = sc.vocab_size + 5
    loss, _ = step.train_step(bad, 2.0)
    assert torch.isnan(loss).item()
    bad = {k: v.clone() for k, v in batch.items()}
    bad["decoder_input_ids"][0, 1] = sc.vocab_size
    loss, _ = step.train_step(bad, 2.0)
    assert torch.isnan(loss).item()
    loss, _ = step.train_step(batch, 2.0)
    assert torch.isfinite(loss).item()


GEN_MULTI = dict(decoder_start_token_id=501, eos_token_id=502, pad_token_id=500, bos_token_id=502,
                 suppress_tokens=[1, 2, 7, 8, 9, 220], begin_suppress_tokens=[220, 502], is_multilingual=True,
                 lang_to_id={"<|en|>": 503, "<|fr|>": 504, "<|de|>": 505}, task_to_id={"transcribe": 506, "translate": 507},
                 no_timestamps_token_id=508)
GEN_EN = dict(decoder_start_token_id=501, eos_token_id=502, pad_token_id=500, bos_token_id=502, is_multilingual=False,
              no_timestamps_token_id=508, suppress_tokens=None, begin_suppress_tokens=None)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_kv_cached_generate_reproduces_hf_token_ids(golden_dir, dtype):
    """model.generate(features, language=, task=, max_length= / max_new_tokens=) -- the reference's gen_kwargs (ref :1434-1445) --
    against token ids minted from HF's initial-token logic + suppress processors + greedy search (tests/golden/generate_tiny.npz):
    exact ids, row for row, through the KV-cached single-token CUDA graph.  fp32 weights = the student in the eval loop,
    bf16 weights = the teacher in the pseudo-labelling loop."""
    g = np.load(os.path.join(golden_dir, "generate_tiny.npz"))
    sc = wo.PRESETS["tiny-student"]
    m = _build(sc, wo.init_state_dict(sc, int(g["model_seed"]), std=float(g["model_std"])), dtype=dtype)
    cases = {"A": (GEN_MULTI, dict(language="fr", task="transcribe", max_new_tokens=10, return_timestamps=False)),
             "B": (GEN_MULTI, dict(max_length=12, return_timestamps=False)),
             "C": (GEN_EN, dict(max_new_tokens=6, return_timestamps=False)),
             # the reference's recommended pseudo-labelling mode: timestamp rules inside the token pick
             "D": (dict(GEN_MULTI, max_initial_timestamp_index=1), dict(language="en", task="transcribe", max_new_tokens=8, return_timestamps=True))}
    for name, (cfg, kw) in cases.items():
        eos = int(g[f"{name}_eos"])
        gen = dict(cfg, eos_token_id=eos)
        if cfg.get("begin_suppress_tokens"):
            gen["begin_suppress_tokens"] = [int(t) for t in g[f"{name}_begin_suppress"]]
        m.generation_config = gen
        feats = torch.from_numpy(g[f"{name}_feats"]).cuda()
        out = m.generate(feats, num_beams=1, **kw)
        ref = torch.from_numpy(g[f"{name}_seq"])
        assert out.dtype == torch.long and tuple(out.shape) == tuple(ref.shape), (name, out.shape, ref.shape)
        assert torch.equal(out.cpu(), ref), (name, out.cpu(), ref)
        # a second call on the same session (graph reuse, weights unchanged) and a sub-batch give the same rows
        assert torch.equal(m.generate(feats, num_beams=1, **kw).cpu(), ref)
        sub = m.generate(feats[1:3], num_beams=1, **kw).cpu()
        for r in range(2):
            row, want = sub[r].tolist(), ref[1 + r].tolist()
            n = min(len(row), len(want))
            assert row[:n] == want[:n] and all(t == cfg["pad_token_id"] for t in want[n:] + row[n:])
    # the arguments the reference passes for multilingual models are honoured or refused -- never swallowed
    m.generation_config = GEN_EN
    with pytest.raises(ValueError):
        m.generate(feats, language="fr", task="transcribe")
    with pytest.raises(NotImplementedError):
        m.generate(feats, some_unknown_flag=True)


def test_generate_sees_weight_updates_between_calls():
    """The eval loop calls generate() between optimiser steps: the captured decode graph must read the re-cast bf16 shadows."""
    sc, step, opt = _tiny_pair(lr=5e-2)
    student = step.student
    feats = wo.synthetic_batch(sc, batch=3, n_tok=4, seed=17)["input_features"].cuda()
    a = student.generate(feats, max_new_tokens=8, eos_token_id=10 ** 6)
    assert student.training
    batch = _cuda(wo.synthetic_batch(sc, batch=3, n_tok=12, seed=5))
    for _ in range(3):
        loss, _ = step.train_step(batch, 2.0)
        loss.backward()
        opt.step()
        opt.zero_grad()
    b = student.generate(feats, max_new_tokens=8, eos_token_id=10 ** 6)
    # same session object, new weights: compare with a fresh (uncached) decode of the updated model through the full-prefix path
    from distil_whisper_b200 import engine
    with torch.no_grad():
        enc, S, _ = engine.run_encoder(student, feats, None)
        st = engine.state_of(student.model.decoder)
        hf, _ = engine.decoder_forward(st, b[:, :-1].contiguous(), enc, 3, S, save=False)
        logits = engine.lm_head(st, hf)[:, :sc.vocab_size].view(3, -1, sc.vocab_size)
    tol = LOGITS_REL * float(logits.abs().max())
    for t in range(b.shape[1] - 1):
        chosen = logits[torch.arange(3), t, b[:, t + 1]]
        assert (chosen >= logits[:, t].max(-1).values - tol).all(), t
    assert not torch.equal(a, b) or True          # (tokens usually change after 3 large steps; equality is not an error)


def test_integration_stub_loop_matches_the_reference_loop():
    """The INTEGRATION.md stub, executed: models built with from_hf, the reference's own weight-decay grouping
    (ref:training/run_distillation.py:1386-1400) fed to FusedAdamW, HF get_scheduler, gradient accumulation over 2 micro-batches
    the way accelerator.backward does it ((loss / k).backward(), ref :1607-1609), optimiser step only on the sync step,
    checkpoint (model.save_pretrained + optimizer.state_dict) -> resume.  Reference arm: the HF modules + torch.optim.AdamW +
    clip_grad_norm_ + the same scheduler in fp32 on the GPU.  After 2 optimiser steps the parameters agree to bf16-gradient noise."""
    import tempfile
    from torch import nn
    from transformers import get_scheduler
    from transformers.modeling_outputs import BaseModelOutput
    from distil_whisper_b200.kd import DistillationStep
    from distil_whisper_b200.modeling import DistilWhisperB200ForConditionalGeneration
    from distil_whisper_b200.optim import FusedAdamW, get_parameter_names
    from oracle.gen_golden import hf_model
    sc, tc = wo.PRESETS["tiny-student"], wo.PRESETS["tiny-teacher"]
    hf_s, _ = hf_model(sc, wo.init_state_dict(sc, 11))
    hf_t, _ = hf_model(tc, wo.init_state_dict(tc, 23))
    hf_s, hf_t = hf_s.cuda(), hf_t.cuda()
    student = DistilWhisperB200ForConditionalGeneration.from_hf(hf_s).cuda()
    teacher = DistilWhisperB200ForConditionalGeneration.from_hf(hf_t, dtype=torch.bfloat16).cuda()
    for model in (student, hf_s):                                   # --freeze_encoder (ref :1023-1026)
        for p in model.model.encoder.parameters():
            p.requires_grad = False
    hf_t.model.encoder = hf_s.model.encoder
    lr, wd, accum, clip = 1e-3, 0.1, 2, 1.0

    def grouped(model):                                             # ref :1386-1400, verbatim rule
        decay = [n for n in get_parameter_names(model, [nn.LayerNorm]) if "bias" not in n]
        return [{"params": [p for n, p in model.named_parameters() if n in decay and p.requires_grad], "weight_decay": wd},
                {"params": [p for n, p in model.named_parameters() if n not in decay and p.requires_grad], "weight_decay": 0.0}]
    opt = FusedAdamW(grouped(student), lr=lr, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=clip)
    sched = get_scheduler("linear", optimizer=opt, num_warmup_steps=1, num_training_steps=4)
    kd = DistillationStep(student, teacher, kl_weight=1.0, share_hidden_states=True)
    ref_opt = torch.optim.AdamW(grouped(hf_s), lr=lr, betas=(0.9, 0.999), eps=1e-8)
    ref_sched = get_scheduler("linear", optimizer=ref_opt, num_warmup_steps=1, num_training_steps=4)
    batches = [_cuda(wo.synthetic_batch(sc, batch=3, n_tok=12, seed=s)) for s in (5, 6, 7, 8)]
    init = {n: p.detach().clone() for n, p in hf_s.named_parameters() if p.requires_grad}

    def ref_micro(batch):
        hf_s.train()
        hf_t.eval()
        so = hf_s(**batch)
        with torch.no_grad():
            to = hf_t(encoder_outputs=BaseModelOutput(so.encoder_last_hidden_state), labels=batch["labels"])
        kl = wo.kl_divergence(torch.softmax(to.logits / 2.0, -1), torch.log_softmax(so.logits / 2.0, -1), batch["labels"]) * 4.0
        loss = 0.8 * so.loss + kl
        (loss / accum).backward()
        return loss

    def our_micro(batch):
        loss, _ = kd.train_step(batch, temperature=2.0)
        (loss / accum).backward()                                   # accelerator.backward(loss) under accumulate()
        return loss
    for i, b in enumerate(batches):
        lo_, lr_ = our_micro(b), ref_micro(b)
        assert abs(lo_.item() - lr_.item()) / lr_.item() < 5e-3
        if (i + 1) % accum == 0:                                    # accelerator.sync_gradients
            opt.all_reduce_gradients()
            opt.step(); sched.step(); opt.zero_grad()               # noqa: E702
            torch.nn.utils.clip_grad_norm_([p for p in hf_s.parameters() if p.requires_grad], clip)
            ref_opt.step(); ref_sched.step(); ref_opt.zero_grad()   # noqa: E702
        if i == 1:                                                  # checkpoint + resume in the middle of training
            with tempfile.TemporaryDirectory() as d:
                student.save_pretrained(d)
                osd = opt.state_dict()
                student2 = DistilWhisperB200ForConditionalGeneration.from_pretrained(d).cuda()
            for p in student2.model.encoder.parameters():
                p.requires_grad = False
            # the order accelerate's load_state uses: objects are built first, then optimizer state (which carries the
            # current lr of every group), then scheduler state
            opt2 = FusedAdamW(grouped(student2), lr=lr, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=clip)
            sched2 = get_scheduler("linear", optimizer=opt2, num_warmup_steps=1, num_training_steps=4)
            opt2.load_state_dict(osd)
            sched2.load_state_dict(sched.state_dict())
            assert opt2.param_groups[0]["lr"] == opt.param_groups[0]["lr"] > 0
            student, opt, sched = student2, opt2, sched2
            kd = DistillationStep(student, teacher, kl_weight=1.0, share_hidden_states=True)
    assert opt.step_count == 2 and abs(opt.param_groups[0]["lr"] - ref_opt.param_groups[0]["lr"]) < 1e-12
    hp = dict(hf_s.named_parameters())
    # AdamW's first steps move every element by ~lr * sign(gradient): where |gradient| is below the bf16 noise of the CUDA
    # path the sign is a coin toss for BOTH implementations, so single small tensors (biases) may differ by O(lr) per element.
    # The checks are therefore on the whole parameter vector and on the direction of the total update.
    ours = torch.cat([p.detach().reshape(-1) for n, p in student.named_parameters() if p.requires_grad]).double()
    ref = torch.cat([hp[n].detach().reshape(-1) for n, p in student.named_parameters() if p.requires_grad]).double()
    start = torch.cat([init[n].reshape(-1) for n, p in student.named_parameters() if p.requires_grad]).double()
    assert float((ours - ref).norm() / ref.norm()) < 2e-2
    du, dr = ours - start, ref - start
    n_du, n_dr = float(du.norm()), float(dr.norm())
    assert n_du > 0 and n_dr > 0, (n_du, n_dr, float(ours.norm()), float(ref.norm()), float(start.norm()))
    cos = float((du * dr).sum()) / (n_du * n_dr)
    assert cos > 0.9, (cos, n_du, n_dr)
    assert abs(n_du / n_dr - 1.0) < 0.1, (n_du, n_dr)
