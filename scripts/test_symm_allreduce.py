"""N-rank check of the peer-memory gradient all-reduce against NCCL (torchrun, one rank per GPU)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from distil_whisper_b200 import ddp  # noqa: E402
from distil_whisper_b200.optim import FlatBuffers  # noqa: E402

rank, local, world = ddp.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
torch.manual_seed(0)
params = [torch.nn.Parameter(torch.randn(n, device=dev)) for n in (1280 * 1280, 51866 * 1280, 1280, 5120 * 1280, 7)]
flat = FlatBuffers([params])
print(f"rank {rank}: symm {'on' if flat.symm is not None else 'OFF'}"
      + (f", multicast_ptr {hex(int(getattr(flat.symm, 'multicast_ptr', 0) or 0))}" if flat.symm is not None else ""), flush=True)
for trial, no_mc in ((0, "0"), (1, "1"), (2, "0")):
    os.environ["DWB_SYMM_NO_MC"] = no_mc
    g = torch.Generator(device=dev).manual_seed(100 * trial + rank)
    flat.grad.copy_(torch.randn(flat.grad.numel(), device=dev, generator=g))
    ref = flat.grad.clone()
    dist.all_reduce(ref)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    flat.all_reduce()
    e.record()
    torch.cuda.synchronize()
    err = float((flat.grad.double() - ref.double()).norm() / ref.double().norm())
    print(f"rank {rank} trial {trial} (no_mc={no_mc}): rel err vs NCCL {err:.2e}, {s.elapsed_time(e):.3f} ms for {flat.grad.numel() * 4 / 1e6:.0f} MB", flush=True)
    assert err < 1e-6, err
# timing vs NCCL
os.environ["DWB_SYMM_NO_MC"] = "0"
for name, fn in (("symm", flat.all_reduce), ("nccl", lambda: dist.all_reduce(flat.grad))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        fn()
    e.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"{name}: {s.elapsed_time(e) / 10:.3f} ms per all-reduce", flush=True)
dist.barrier()
dist.destroy_process_group()
