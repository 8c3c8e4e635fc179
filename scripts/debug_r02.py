"""(1) run-to-run noise of the full-size student gradient on one GPU (same weights, same batch, twice), per parameter;
(2) the integration-stub test body with prints."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench  # noqa: E402
from distil_whisper_b200.kd import DistillationStep  # noqa: E402
from distil_whisper_b200.optim import FusedAdamW  # noqa: E402

dev = torch.device("cuda", 0)
student, teacher = bench.build_models(dev, "B")
step = DistillationStep(student, teacher, kl_weight=1.0)
opt = FusedAdamW.for_model(student, lr=1e-4, weight_decay=0.0, max_grad_norm=1.0)
batch = {k: v.to(dev) for k, v in bench.synthetic_batch(32, 128, 1234, bench.STUDENT).items()}
grads = []
for rep in range(3):
    opt.flat.grad.zero_()
    step.forward_backward(batch, 2.0)
    torch.cuda.synchronize()
    grads.append(opt.flat.grad.clone())
for a in (1, 2):
    d = (grads[a].double() - grads[0].double())
    print(f"rep{a} vs rep0: rel {float(d.norm() / grads[0].double().norm()):.3e}  max abs {float(d.abs().max()):.3e}")
names = {id(p): n for n, p in student.named_parameters()}
rows = []
for p, off in opt.flat.layout:
    n = p.numel()
    a, b = grads[0][off:off + n].double(), grads[1][off:off + n].double()
    rows.append((float((a - b).norm() / (a.norm() + 1e-30)), float((a - b).norm()), float(a.norm()), names[id(p)]))
for r in sorted(rows, reverse=True)[:12]:
    print("  %.3e  |d| %.3e  |g| %.3e  %s" % r)
print("by absolute contribution:")
for r in sorted(rows, key=lambda r: -r[1])[:8]:
    print("  %.3e  |d| %.3e  |g| %.3e  %s" % r)
