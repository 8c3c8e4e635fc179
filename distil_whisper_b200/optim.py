"""Optimiser tail of the KD step: flat fp32 parameter / gradient / moment buffers, ONE gradient all-reduce, grad-norm
clip + AdamW in one pass (dwb_grad_sumsq, dwb_adamw_step).

Mirrors ref:training/run_distillation.py:1386-1407 (two AdamW groups: weight decay on everything that is not a
LayerNorm parameter or a bias) and :1610-1614 (clip_grad_norm_(max_grad_norm) -> optimizer.step() -> zero_grad()).
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops


def _symmetric_grad_buffer(total: int, dev):
    """The flat gradient buffer in symmetric memory (every rank maps every peer's buffer), so that the gradient all-reduce can be
    the small-footprint peer-memory kernel dwb_allreduce_symm instead of NCCL.  Collective: every rank must call it at the same
    point (FusedAdamW construction).  Returns (tensor, handle) or (None, None) when not applicable / not available -- then the
    buffer is an ordinary tensor and the all-reduce goes through NCCL.

    Opt-in (DWB_SYMM_ALLREDUCE=1).  Measured at N = 2 (profiles/r02_tail_overlap.md): results are bitwise equal to NCCL's and the
    kernel takes 0.85 ms standalone for 298 MB (NCCL 0.58 ms); underneath the encoder forward NCCL's all-reduce + the optimiser take
    3.7 ms on the side stream and the main stream never waits (0.003 ms), while this kernel's 32 CTAs are starved by the persistent
    GEMM CTAs (11 ms) and slow the encoder graph by 1.4 ms -- so NCCL stays the default."""
    import os
    import torch.distributed as dist
    if os.environ.get("DWB_SYMM_ALLREDUCE", "0") != "1" or dev.type != "cuda":
        return None, None
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and dist.get_backend() == "nccl"):
        return None, None
    try:
        import torch.distributed._symmetric_memory as symm_mem
        group = dist.group.WORLD
        if hasattr(symm_mem, "enable_symm_mem_for_group"):
            try:
                symm_mem.enable_symm_mem_for_group(group.group_name)
            except Exception:  # noqa: BLE001  (newer versions enable lazily)
                pass
        buf = symm_mem.empty(total, dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(buf, group)
        buf.zero_()
        torch.cuda.synchronize(dev)
        dist.barrier()
        return buf, hdl
    except Exception as ex:  # noqa: BLE001
        import warnings
        warnings.warn(f"symmetric-memory gradient buffer unavailable ({type(ex).__name__}: {ex}); the all-reduce uses NCCL")
        return None, None


def get_parameter_names(model, forbidden_layer_types, forbidden_module=None):
    """Same contract as ref:training/run_distillation.py:760-778 (names outside forbidden layer types / modules)."""
    result = []
    for name, child in model.named_children():
        if isinstance(child, tuple(forbidden_layer_types)) or (forbidden_module is not None and child in tuple(forbidden_module)):
            continue
        result += [f"{name}.{n}" for n in get_parameter_names(child, forbidden_layer_types, forbidden_module)]
    result += list(model._parameters.keys())
    return result


def decay_split(model):
    """(decay params, no-decay params) of the trainable parameters, by the reference's rule (ref :1392-1400)."""
    decay_names = [n for n in get_parameter_names(model, [nn.LayerNorm]) if "bias" not in n]
    decay, no_decay = [], []
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (decay if n in decay_names else no_decay).append(p)
    return decay, no_decay


class FlatBuffers:
    """Re-homes parameters (and their .grad) into one contiguous fp32 buffer each, so that the gradient exchange is a
    single collective and the optimiser a single launch per group.  Device-agnostic (CPU tensors work: used by the gloo
    tests); 16-byte alignment per parameter keeps every view usable as a TMA / vector operand."""

    def __init__(self, groups):
        self.groups = []            # (offset, numel, params)
        params = [p for g in groups for p in g]
        if not params:
            raise ValueError("no trainable parameters")
        dev = params[0].device
        total = 0
        layout = []
        for g in groups:
            start = total
            for p in g:
                if p.dtype != torch.float32:
                    raise TypeError("trainable parameters must be fp32 master weights")
                layout.append((p, total))
                total += (p.numel() + 3) // 4 * 4
            self.groups.append((start, total - start, list(g)))
        self.data = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad, self.symm = _symmetric_grad_buffer(total, dev)
        if self.grad is None:
            self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, off in layout:
                view = self.data[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                gview = self.grad[off:off + p.numel()].view(p.shape)
                if p.grad is not None:
                    gview.copy_(p.grad)
                p.grad = gview
        self.layout = layout

    def rebind_grads(self):
        """Point .grad back at the flat views (after someone set them to None)."""
        for p, off in self.layout:
            if p.grad is None or p.grad.data_ptr() != self.grad[off:].data_ptr():
                p.grad = self.grad[off:off + p.numel()].view(p.shape)

    def all_reduce(self, group=None):
        """THE multi-GPU step of the KD path: one sum all-reduce of the student gradients (averaging is folded into
        the optimiser's grad_scale).  ref: implicit DDP all-reduce inside accelerator.backward, :1609."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if self.symm is not None and group is None:
                self._all_reduce_symm()
            else:
                dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
            return dist.get_world_size(group)
        return 1

    def _all_reduce_symm(self):
        """Peer-memory all-reduce (csrc/collective.cu): barrier (every rank's gradient is complete and visible), one kernel in which
        rank r sums slice r of the N buffers -- inside the NVSwitch when the buffer has a multicast mapping -- and writes the sum
        into all of them, barrier (every slice is written).  Runs on the current stream; a few no-smem CTAs."""
        import ctypes as C
        import os
        from . import _abi
        h, n = self.symm, self.grad.numel()
        world, rank = h.world_size, h.rank
        off = int(getattr(h, "offset", 0) or 0)
        mc = 0 if os.environ.get("DWB_SYMM_NO_MC", "0") == "1" else int(getattr(h, "multicast_ptr", 0) or 0)
        if not hasattr(self, "_symm_peers"):
            peers = [h.get_buffer(r, (n,), torch.float32) for r in range(world)]
            self._symm_peer_tensors = peers                     # keep the mappings alive
            self._symm_peers = (C.c_void_p * world)(*[C.c_void_p(t.data_ptr()) for t in peers])
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        h.barrier(channel=0)
        _abi.call("dwb_allreduce_symm", C.c_void_p(mc + off) if mc else None, self._symm_peers, rank, world, n,
                  int(os.environ.get("DWB_SYMM_CTAS", "32")), stream)
        h.barrier(channel=1)


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics on flat buffers.  `max_grad_norm` > 0 applies clip_grad_norm_ inside the same pass."""

    def __init__(self, params, lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.flat = FlatBuffers([[p for p in g["params"] if p.requires_grad] for g in self.param_groups])
        self.exp_avg = torch.zeros_like(self.flat.data)
        self.exp_avg_sq = torch.zeros_like(self.flat.data)
        self.max_grad_norm = float(max_grad_norm)
        self.step_count = 0
        self.grad_scale = 1.0
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=self.flat.data.device)
        self._bind_state()

    # ---- checkpointing: torch.optim.AdamW's per-parameter layout ------------------------------------------------
    def _bind_state(self):
        """Expose the flat moments as `state[p] = {step, exp_avg, exp_avg_sq}` (views), the layout torch.optim.AdamW
        checkpoints, so `accelerator.save_state` / `load_state` (ref:training/run_distillation.py:1559, :1640) round-trip."""
        for p, off in self.flat.layout:
            n = p.numel()
            self.state[p] = {"step": torch.tensor(float(self.step_count)),
                             "exp_avg": self.exp_avg[off:off + n].view(p.shape),
                             "exp_avg_sq": self.exp_avg_sq[off:off + n].view(p.shape)}

    def state_dict(self):
        for st in self.state.values():
            st["step"] = torch.tensor(float(self.step_count))
        sd = super().state_dict()
        sd["dwb"] = {"step_count": self.step_count, "grad_scale": self.grad_scale, "max_grad_norm": self.max_grad_norm}
        return sd

    def load_state_dict(self, state_dict):
        extra = state_dict.get("dwb")
        super().load_state_dict({k: v for k, v in state_dict.items() if k != "dwb"})
        steps = set()
        with torch.no_grad():
            for p, off in self.flat.layout:
                st = self.state.get(p)
                if not st:                       # a checkpoint taken before the first step: moments stay zero
                    continue
                n = p.numel()
                self.exp_avg[off:off + n].view(p.shape).copy_(st["exp_avg"])
                self.exp_avg_sq[off:off + n].view(p.shape).copy_(st["exp_avg_sq"])
                steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"FusedAdamW keeps one step count for all parameters; the checkpoint has {sorted(steps)}")
        self.step_count = steps.pop() if steps else 0
        if extra:
            self.step_count = int(extra.get("step_count", self.step_count))
            self.grad_scale = float(extra.get("grad_scale", self.grad_scale))
        self._bind_state()

    @classmethod
    def for_model(cls, model, lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=1.0):
        decay, no_decay = decay_split(model)
        groups = [g for g in ({"params": decay, "weight_decay": weight_decay}, {"params": no_decay, "weight_decay": 0.0}) if g["params"]]
        return cls(groups, lr=lr, betas=betas, eps=eps, max_grad_norm=max_grad_norm)

    def all_reduce_gradients(self, group=None):
        world = self.flat.all_reduce(group)
        self.grad_scale = 1.0 / world
        return world

    def grad_norm(self):
        """Global L2 norm of the (scaled) gradients as a 0-d device tensor, no host sync."""
        self._sumsq.zero_()
        ops.grad_sumsq(self.flat.grad, self._sumsq)
        return self._sumsq.sqrt() * self.grad_scale

    @torch.no_grad()
    def step(self, closure=None):
        self.flat.rebind_grads()
        self.step_count += 1
        clip = self.max_grad_norm > 0
        if clip:
            self._sumsq.zero_()
            ops.grad_sumsq(self.flat.grad, self._sumsq)
        for g, (off, n, _) in zip(self.param_groups, self.flat.groups):
            if n == 0:
                continue
            sl = slice(off, off + n)
            ops.adamw_step(self.flat.data[sl], self.flat.grad[sl], self.exp_avg[sl], self.exp_avg_sq[sl], None, g["lr"],
                           g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], self.step_count,
                           self._sumsq if clip else None, self.max_grad_norm, self.grad_scale, zero_grad=True)
        # parameters were rewritten through raw pointers (no autograd version bump): tell the engine's shadow cache
        from . import engine
        engine.bump_param_epoch()

    def zero_grad(self, set_to_none=False):
        # gradients are zeroed by the fused kernel right after they are consumed; keep the flat views bound
        self.flat.rebind_grads()
