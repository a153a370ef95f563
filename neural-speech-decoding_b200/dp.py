"""Data-parallel training of the decoder: one process per GPU, batch sharded across ranks, ONE
exchange step per optimizer step -- a single flat fp32 gradient bucket all-reduced (SUM) over NCCL
(NVLink 5 / NVSwitch), then an identical optimizer step on every rank.  SURVEY 8(e).

The model has 31,764 parameters (127 KB of gradients): the all-reduce is latency-bound, so there is
nothing to overlap or bucket -- every parameter's ``.grad`` is a VIEW into one contiguous buffer (the scalar loss rides in one
extra slot behind the gradients, so a step issues a single collective), the
backward kernels' results are accumulated straight into it and the collective runs on that buffer
in place (no flatten / unflatten copies).

Inference needs no communication at all: shard the batch and call the model (see bench.py).
"""
from __future__ import annotations

from typing import Iterable, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class FlatGradBucket:
    """All gradients of a module as views into one contiguous fp32 buffer."""

    def __init__(self, params: Sequence[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        # one extra slot behind the gradients carries the scalar loss through the SAME collective (one launch instead of two)
        self.store = torch.zeros(n + 1, dtype=torch.float32, device=dev)
        self.flat = self.store[:n]
        self.loss_slot = self.store[n:]
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self) -> None:
        self.store.zero_()

    def check_views(self) -> None:
        """Autograd accumulates in place into an existing .grad; if someone replaced it
        (e.g. optimizer.zero_grad(set_to_none=True)) re-attach the views."""
        off = 0
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                if p.grad is not None:
                    v.copy_(p.grad)
                p.grad = v
            off += p.numel()


class DataParallelTrainer:
    """``step(micro_batches, global_batch)``: forward/backward over this rank's micro-batches with the
    loss ``sum(CE) / global_batch`` (so that the SUM all-reduce yields the gradient of the mean loss
    over the GLOBAL batch, exactly what a single process on the concatenated batch computes), one
    flat all-reduce, optimizer step.  Returns the global mean loss (a device scalar)."""

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, world_size: Optional[int] = None,
                 process_group=None):
        self.model, self.optimizer, self.group = model, optimizer, process_group
        if world_size is None:
            world_size = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.world_size = world_size
        self.bucket = FlatGradBucket(list(model.parameters()))
        if hasattr(optimizer, "attach"):            # optim.FusedAdam updates in place: it must drop the model's packed-weight caches
            optimizer.attach(model)

    def backward_only(self, micro_batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], global_batch: int) -> torch.Tensor:
        self.bucket.check_views()
        self.bucket.zero_()
        total = None
        for x, y in micro_batches:
            logits = self.model(x)
            loss = torch.nn.functional.cross_entropy(logits.float(), y, reduction="sum") / float(global_batch)
            loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
        if self.world_size > 1:
            # one exchange step: the gradients and, in the slot behind them, the scalar loss (for logging) -- ONE collective
            self.bucket.loss_slot.copy_(total.reshape(1))
            dist.all_reduce(self.bucket.store, op=dist.ReduceOp.SUM, group=self.group)
            total = self.bucket.loss_slot[0].clone()
        return total

    def step(self, micro_batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], global_batch: int) -> torch.Tensor:
        loss = self.backward_only(micro_batches, global_batch)
        self.optimizer.step()
        return loss


def bind_to_gpu_numa(device_index: int) -> Optional[list]:
    """Pin this process to the CPUs that are local to GPU ``device_index`` (NVML's affinity mask, intersected with the
    CPUs the container may use), so that pinned host buffers are first-touched on the GPU's NUMA node and the H2D
    copies of N ranks do not all cross the socket interconnect.  Call it before allocating pinned memory.
    Returns the CPU list, or None if NVML / the affinity call is unavailable (nothing changes then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        prop = torch.cuda.get_device_properties(device_index)
        handle = None
        if hasattr(prop, "pci_bus_id"):              # NVML enumerates every GPU of the box; CUDA only the visible ones
            try:
                handle = pynvml.nvmlDeviceGetHandleByPciBusId(
                    f"{int(getattr(prop, 'pci_domain_id', 0)):08x}:{int(prop.pci_bus_id):02x}:{int(getattr(prop, 'pci_device_id', 0)):02x}.0")
            except Exception:
                handle = None
        if handle is None:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        local = {i for i in range(ncpu) if (int(mask[i // 64]) >> (i % 64)) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = sorted(local & allowed)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def shard_batch(n: int, rank: int, world_size: int) -> slice:
    """Contiguous B/N slices (keeps all trials of a session on one GPU)."""
    per = (n + world_size - 1) // world_size
    return slice(min(n, rank * per), min(n, (rank + 1) * per))
