"""One KD step between cudaProfilerStart/Stop (for `ncu --profile-from-start off` launch lists)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench  # noqa: E402
from distil_whisper_b200.kd import DistillationStep  # noqa: E402
from distil_whisper_b200.optim import FusedAdamW  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
student, teacher = bench.build_models(dev, os.environ.get("DWB_VARIANT", "B"))
step = DistillationStep(student, teacher, kl_weight=1.0)
opt = FusedAdamW.for_model(student, lr=1e-4, max_grad_norm=1.0)
batch = {k: v.to(dev) for k, v in bench.synthetic_batch(bench.BATCH, bench.N_TOK, 1234, bench.STUDENT).items()}


def one():
    loss, _ = step.train_step(batch, 2.0)
    loss.backward()
    opt.all_reduce_gradients()
    opt.step()
    opt.zero_grad()


for _ in range(2):
    one()
torch.cuda.synchronize()
torch.cuda.profiler.start()
one()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
