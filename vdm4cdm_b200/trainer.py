"""Data-parallel training loop for ``LightVDM`` / ``LightSFM``: what ``lightning.Trainer(devices=...,
gradient_clip_val=0.5, max_epochs=...).fit(model, datamodule)`` does for the reference
(trainVDM3D128_c_c_from_field_name_thick_lowbatch.py:134-160, trainSFM3D160_...:129-150), without Lightning.

One process per GPU (``torch.distributed``, NCCL on GPUs / gloo on CPU for the host-logic tests):

  * every parameter is a view into ONE flat fp32 bucket, every ``.grad`` a view into a second one, so
    the gradient exchange is a single ``all_reduce(SUM)`` over NVLink and the optimizer is a single
    kernel (``vdm_adamw_step``: grad averaging, global-norm clipping, AdamW) over the bucket;
  * GroupNorm is per sample, so there is nothing else to synchronise;
  * replicas start from rank 0's weights (broadcast) and stay bit-identical because every rank applies
    the same update to the same bucket.

  * after a few eager steps the WHOLE step (zero grads, weight re-packing, forward, backward, all-reduce,
    gradient norm, clip + AdamW) is captured in one CUDA graph and replayed: a 128^3 step is ~1500 kernel
    launches and torch ops, which took the host as long to issue as the GPU needs to run them (r01m).  The
    batch is copied into static buffers; dropout masks and Adam's bias correction advance through device
    counters, torch's own RNG is graph-safe.

``Trainer`` needs a CUDA device (the optimizer kernel has no CPU path, like every compute entry point of this
package); ``FlatBuckets`` / ``allreduce_gradients`` / ``shard_indices`` are device agnostic and are what the
gloo tests exercise.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Iterable, Optional

import torch
import torch.distributed as dist

from . import ops


class FlatBuckets:
    """Re-homes the parameters of ``module`` into one flat fp32 buffer (and their grads into another)."""

    def __init__(self, module: torch.nn.Module):
        params = [p for p in module.parameters() if p.requires_grad]
        assert all(p.dtype == torch.float32 for p in params), "fp32 master parameters expected"
        self.params = params
        sizes = [(p.numel() + 3) // 4 * 4 for p in params]            # 16-byte aligned slots
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + s)
        n = self.offsets[-1]
        dev = params[0].device
        self.flat_param = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(params, self.offsets):
                view = self.flat_param[o:o + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)

    @property
    def numel(self) -> int:
        return self.flat_param.numel()

    def zero_grad(self, detach_small: bool = False) -> None:
        """Zero the flat gradient bucket and (re-)attach the ``.grad`` views.  With ``detach_small`` the parameters that are
        not conv filters (biases, GroupNorm affines, the embedding MLPs, the noise schedule: ~115 tensors) get ``.grad =
        None`` instead: autograd then STORES their gradients instead of launching one in-place add per tensor, and
        ``collect_small_grads`` moves them into the bucket with a few multi-tensor copies."""
        self.flat_grad.zero_()
        for p, o in zip(self.params, self.offsets):         # re-attach views a caller may have dropped
            if detach_small and p.dim() <= 2:
                p.grad = None
            elif p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)

    def collect_small_grads(self) -> None:
        """After backward of a ``zero_grad(detach_small=True)`` step: copy the stored gradients of the small parameters into
        their bucket slots (``torch._foreach_copy_``: a handful of launches for all of them) and re-attach the views."""
        dsts, srcs = [], []
        for p, o in zip(self.params, self.offsets):
            if p.dim() > 2:
                continue
            view = self.flat_grad[o:o + p.numel()].view_as(p)
            if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
                dsts.append(view)
                srcs.append(p.grad.detach())
            p.grad = view
        if dsts:
            with torch.no_grad():
                torch._foreach_copy_(dsts, srcs)


def allreduce_gradients(buckets: FlatBuckets) -> None:
    """Sum the flat gradient bucket over all ranks (no-op without an initialised process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buckets.flat_grad, op=dist.ReduceOp.SUM)


def broadcast_parameters(buckets: FlatBuckets, src: int = 0) -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(buckets.flat_param, src=src)


def shard_indices(n_items: int, rank: int, world: int) -> range:
    """Item ids of this rank: i -> rank i mod world (the sampling shard rule of SURVEY.md section 8e)."""
    return range(rank, n_items, world)


def _clone_tree(obj):
    if torch.is_tensor(obj):
        return obj.clone()
    if isinstance(obj, dict):
        return {k: _clone_tree(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_clone_tree(v) for v in obj)
    return obj


def _same_layout(a, b) -> bool:
    if torch.is_tensor(a):
        return torch.is_tensor(b) and a.shape == b.shape and a.dtype == b.dtype
    if isinstance(a, dict):
        return isinstance(b, dict) and a.keys() == b.keys() and all(_same_layout(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)):
        return isinstance(b, (list, tuple)) and len(a) == len(b) and all(_same_layout(x, y) for x, y in zip(a, b))
    return a == b


def _copy_tree(dst, src) -> None:
    if torch.is_tensor(dst):
        if dst.data_ptr() != src.data_ptr():
            dst.copy_(src, non_blocking=True)
    elif isinstance(dst, dict):
        for k in dst:
            _copy_tree(dst[k], src[k])
    elif isinstance(dst, (list, tuple)):
        for x, y in zip(dst, src):
            _copy_tree(x, y)


class Trainer:
    """``Trainer(model, ...).training_step(batch)`` = forward + backward + all-reduce + clip + AdamW.

    ``model`` is a ``LightVDM`` / ``LightSFM`` (anything with ``training_step(batch) -> loss`` and
    ``learning_rate``)."""

    def __init__(self, model: torch.nn.Module, gradient_clip_val: float = 0.5, weight_decay: float = 0.01,
                 betas=(0.9, 0.999), eps: float = 1e-8, learning_rate: Optional[float] = None,
                 use_cuda_graph: bool = True, graph_warmup_steps: int = 3, seed: int = 42):
        self.model = model
        self.use_cuda_graph, self.graph_warmup_steps = use_cuda_graph, graph_warmup_steps
        self._graph, self._static_batch, self._static_loss, self._eager_steps = None, None, None, 0
        self.lr = float(model.learning_rate if learning_rate is None else learning_rate)
        self.clip, self.wd, self.betas, self.eps = float(gradient_clip_val), float(weight_decay), betas, float(eps)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.buckets = FlatBuckets(model)
        broadcast_parameters(self.buckets)
        # dropout masks are keyed by (base seed + step counter, layer, element): fold the rank into the base seed so
        # that the replicas of a data-parallel run draw DIFFERENT masks for their micro-batches
        for m in model.modules():
            if hasattr(m, "drop_counter"):
                m.dropout_base_seed = int(seed) * self.world + self.rank
        self._nets_changed()
        dev = self.buckets.flat_param.device
        if dev.type != "cuda":
            raise RuntimeError("vdm4cdm_b200.Trainer needs a CUDA device (there is no CPU optimizer path)")
        self.exp_avg = torch.zeros_like(self.buckets.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.buckets.flat_param)
        self.grad_sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)      # optimizer step number, on the device
        self.step_count = 0
        # warm-up steps and the capture run on ONE side stream, so that autograd's AccumulateGrad nodes (created by
        # the first backward) live on the stream the graph is captured on
        self._stream = torch.cuda.Stream(device=dev) if use_cuda_graph else None

    def _nets_changed(self):
        for m in self.model.modules():
            if hasattr(m, "invalidate_packed"):
                m.invalidate_packed()

    def training_step(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """One optimizer step on this rank's micro-batch; returns the (local) loss tensor, not synchronised."""
        self.model.train()
        if not self.use_cuda_graph:
            return self._eager_step(batch)
        if self._static_batch is None or not _same_layout(self._static_batch, batch):
            self._static_batch, self._graph, self._eager_steps = _clone_tree(batch), None, 0
        _copy_tree(self._static_batch, batch)
        if self._graph is not None:
            self._graph.replay()
            self.step_count += 1
            return self._static_loss.clone()
        if self._eager_steps < self.graph_warmup_steps:
            self._eager_steps += 1
            cur = torch.cuda.current_stream()
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream):
                loss = self._eager_step(self._static_batch)
            cur.wait_stream(self._stream)
            return loss
        # A failed capture RAISES: there is no silent eager fallback (pass use_cuda_graph=False to ask for eager launches)
        self._capture()
        self._graph.replay()
        self.step_count += 1
        return self._static_loss.clone()

    def _eager_step(self, batch) -> torch.Tensor:
        nvtx = torch.cuda.nvtx.range          # no-ops without a profiler attached
        with nvtx("vdm.train_step"):
            self.buckets.zero_grad(detach_small=True)
            with nvtx("vdm.forward_loss"):
                loss = self.model.training_step(batch)
            with nvtx("vdm.backward"):
                loss.backward()
                self.buckets.collect_small_grads()
            with nvtx("vdm.allreduce_clip_adamw"):
                self.optimizer_step()
        return loss.detach()

    def _capture(self) -> None:
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        self._nets_changed()                          # the weight re-packing must be part of the graph
        with torch.cuda.graph(graph, stream=self._stream):
            self.buckets.zero_grad(detach_small=True)
            loss = self.model.training_step(self._static_batch)
            loss.backward()
            self.buckets.collect_small_grads()
            self._optimizer_kernels()
            self._static_loss = loss.detach()
        self._graph = graph

    def _optimizer_kernels(self) -> None:
        allreduce_gradients(self.buckets)
        ops.increment(self.step_dev)
        self.grad_sumsq.zero_()
        ops.sumsq(self.buckets.flat_grad, self.grad_sumsq)
        ops.adamw_step(self.buckets.flat_param, self.buckets.flat_grad, self.exp_avg, self.exp_avg_sq, lr=self.lr,
                       step=0, step_ptr=self.step_dev, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps,
                       weight_decay=self.wd, grad_sumsq=self.grad_sumsq, max_norm=self.clip,
                       grad_scale=1.0 / self.world)
        self._nets_changed()

    def optimizer_step(self) -> None:
        """all-reduce + gradient norm + clip + AdamW on the flat buckets (gradients must be in place)."""
        self._optimizer_kernels()
        self.step_count += 1

    def sync_inference_weights(self) -> None:
        """Call before using the model for inference between training steps (validation, sampling): the captured step
        re-packs the bf16 conv filters at its START, so after a replay they are one optimizer update behind the fp32
        parameters; this marks them stale and the next forward re-packs them (in place, pointers stay valid)."""
        self._nets_changed()

    def grad_norm(self) -> float:
        """Global gradient norm of the last step (after averaging over ranks, before clipping)."""
        return math.sqrt(self.grad_sumsq.item()) / self.world

    def fit(self, batches: Iterable[Dict[str, torch.Tensor]], max_steps: Optional[int] = None,
            log: Optional[Callable[[int, float], None]] = None) -> None:
        for i, batch in enumerate(batches):
            if max_steps is not None and i >= max_steps:
                break
            loss = self.training_step(batch)
            if log is not None:
                log(i, loss.item())

    def _dropout_nets(self):
        return [(n, m) for n, m in self.model.named_modules() if hasattr(m, "drop_counter") and torch.is_tensor(m.drop_counter)]

    def state_dict(self) -> dict:
        """Everything a run needs to continue: ``state_dict`` (the key ``utils.get_model`` and Lightning checkpoints use,
        src/utils.py:467), ``global_step``, the optimizer (AdamW moments + step number), the dropout-mask counters and
        the RNG states the loss draws its times and noise from."""
        dev = self.buckets.flat_param.device
        return {"state_dict": self.model.state_dict(), "global_step": self.step_count,
                "optimizer": {"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                              "step": int(self.step_dev.item()), "lr": self.lr, "betas": tuple(self.betas), "eps": self.eps,
                              "weight_decay": self.wd, "gradient_clip_val": self.clip},
                "drop_counters": {n: int(m.drop_counter.item()) for n, m in self._dropout_nets()},
                "rng": {"cpu": torch.get_rng_state(), "cuda": torch.cuda.get_rng_state(dev)}}

    def load_state_dict(self, state: dict, restore_rng: bool = True) -> None:
        self.model.load_state_dict(state["state_dict"])         # copies into the flat views in place
        opt = state.get("optimizer")
        if opt is None and "exp_avg" in state:                  # round-1 layout
            opt = {"exp_avg": state["exp_avg"], "exp_avg_sq": state["exp_avg_sq"], "step": state["step"]}
        if opt is not None:
            self.exp_avg.copy_(opt["exp_avg"])
            self.exp_avg_sq.copy_(opt["exp_avg_sq"])
            self.step_dev.fill_(int(opt["step"]))
        self.step_count = int(state.get("global_step", 0 if opt is None else opt["step"]))
        nets = dict(self._dropout_nets())
        for n, v in state.get("drop_counters", {}).items():
            if n in nets:
                nets[n].drop_counter.fill_(int(v))
        if restore_rng and "rng" in state:
            torch.set_rng_state(state["rng"]["cpu"].cpu())
            torch.cuda.set_rng_state(state["rng"]["cuda"].cpu(), self.buckets.flat_param.device)
        self._graph, self._eager_steps = None, 0                # pointers are unchanged, but re-capture from clean state
        self._nets_changed()

    def save_checkpoint(self, path: str, extra: Optional[dict] = None) -> None:
        """``torch.save`` of ``state_dict()`` (+ ``extra``, e.g. the data loader's position) -- written to a temporary
        name first: a killed job never leaves a torn file."""
        import os
        state = self.state_dict()
        if extra:
            state.update(extra)
        tmp = path + ".tmp"
        torch.save(state, tmp)
        os.replace(tmp, path)

    def load_checkpoint(self, path: str, restore_rng: bool = True) -> dict:
        """Restores weights, optimizer, step and RNG; returns the whole checkpoint dict (``global_step``, extras)."""
        state = torch.load(path, map_location=self.buckets.flat_param.device, weights_only=False)
        self.load_state_dict(state, restore_rng=restore_rng)
        return state
