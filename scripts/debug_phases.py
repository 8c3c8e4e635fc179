"""Where does a pipelined step spend its time?  CUDA events around graph E, graph D and the optimiser tail, with and without overlap."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
from distil_whisper_b200.kd import DistillationStep, PipelinedTrainer
from distil_whisper_b200.optim import FusedAdamW
dev = torch.device("cuda", 0)
student, teacher = bench.build_models(dev, "B")
step = DistillationStep(student, teacher, kl_weight=1.0)
opt = FusedAdamW.for_model(student, lr=1e-4, weight_decay=0.0, max_grad_norm=1.0)
batch = {k: v.to(dev) for k, v in bench.synthetic_batch(32, 128, 1234, bench.STUDENT).items()}
for _ in range(2):
    loss, _ = step.train_step(batch, 2.0); loss.backward(); opt.step()
for overlap in ("1", "0"):
    os.environ["DWB_TAIL_OVERLAP"] = overlap
    tr = PipelinedTrainer(step, opt, batch, temperature=2.0)
    for _ in range(4):
        tr.step(None)
    tr.flush(); torch.cuda.synchronize()
    tr.profile = []
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        tr.step(None)
    tr.flush(); e.record(); torch.cuda.synchronize()
    P = tr.profile
    avg = lambda f: sum(f(p) for p in P[1:]) / (len(P) - 1)
    print(f"overlap={overlap}: step {s.elapsed_time(e) / 10:.2f} ms | E {avg(lambda p: p[0].elapsed_time(p[1])):.2f} | wait {avg(lambda p: p[1].elapsed_time(p[5])):.3f} | D {avg(lambda p: p[5].elapsed_time(p[2])):.2f} | "
          f"tail {avg(lambda p: p[3].elapsed_time(p[4])):.2f} | D end -> tail start {avg(lambda p: p[2].elapsed_time(p[3])):.3f}")
    tr.profile = None
    s.record()
    for _ in range(10):
        tr.step(None, tail=False)
    e.record(); torch.cuda.synchronize()
    opt.flat.grad.zero_()
    print(f"   no tail: step {s.elapsed_time(e) / 10:.2f} ms")
    del tr
