"""CUDA-graph replay of one forward + backward pass of the hot path.

A training step of the path is ~2000 short launches (36 ConvNeXt blocks x ~25 kernels each way, 16
decoders); issued eagerly from Python the host falls behind the GPU (55 ms of launch work for 40 ms of
kernels at 16 images).  The shapes of a training run are static, so the step is captured once and
replayed: every kernel of libdgtd_ops.so takes its stream from the caller and allocates nothing, so the
capture needs no special casing.

Gradients live in ONE flat fp32 buffer (`.grad` of every parameter is a view into it).  With data parallel
training the buffer is reduced in ~25 MiB buckets whose NCCL all-reduces are launched from gradient hooks INSIDE the
captured step (twig/buckets.py), i.e. as a forked branch of the graph that overlaps the rest of backward -- the
mechanism of the reference's `MMDistributedDataParallel` (cod.py:8,238) without its bucket copies.
`reduce="after"` keeps the round-1 behaviour (one all-reduce of the whole buffer after the replay), `reduce="none"`
leaves the gradients local (used to measure the exposed reduction time).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.nn as nn

from . import flat
from .buckets import GradBucketer
from .model import texture_diffuser as TD


import weakref

_LIVE_GRAPHS: "weakref.WeakSet" = weakref.WeakSet()     # Graphed* objects whose captured graphs are still alive


def _mark_captured(owner, params) -> list:
    """Pointers a captured graph reads parameters from; `_check_captured` raises when one moved (an optimizer that
    re-homes `p.data`, `module.to()`, `load_state_dict(assign=True)`): the replay would read freed storage."""
    owner._param_ids = {id(p) for p in params}
    _LIVE_GRAPHS.add(owner)
    return [p.data_ptr() for p in params]


def captured_by_live_graph(params) -> bool:
    """True when a still-alive GraphedTrainStep / GraphedPredict captured any of `params` (twig/optim.py refuses to
    re-home those)."""
    ids = {id(p) for p in params}
    return any(ids & g._param_ids for g in list(_LIVE_GRAPHS))


def _check_captured(params, ptrs, what: str) -> None:
    for p, q in zip(params, ptrs):
        if p.data_ptr() != q:
            raise RuntimeError(f"{what}: a parameter's storage moved after capture (was an optimizer / .to() / "
                               "load_state_dict(assign=True) applied afterwards?); re-capture the step")


def default_loss(emb1, emb3, tokens) -> torch.Tensor:
    """Scalar stand-in for the downstream loss: every prompt tensor and embedding3 contribute."""
    return sum(t.float().mean() for row in tokens for t in row) + emb3.float().mean()


class GraphedTrainStep:
    """step = GraphedTrainStep(enc, dec, image, depth); loss = step(image, depth) replays fwd + bwd.

    After the call `p.grad` of every hot-path parameter holds this step's gradient (averaged over
    `process_group` when given).  Inputs must keep the shape / dtype of the example tensors."""

    def __init__(self, enc: nn.Module, dec: nn.Module, image: torch.Tensor, depth: torch.Tensor,
                 loss_fn: Callable = default_loss, precision: Optional[str] = None, warmup: int = 3,
                 process_group=None, flat_grad: Optional[torch.Tensor] = None, reduce: str = "bucketed",
                 bucket_bytes: int = 25 << 20):
        self.enc, self.dec, self.loss_fn, self.precision = enc, dec, loss_fn, precision
        self.image, self.depth = image.detach().clone(), depth.detach().clone()
        self._capture([p for p in list(enc.parameters()) + list(dec.parameters()) if p.requires_grad], warmup,
                      process_group, flat_grad, reduce, bucket_bytes)

    def _capture(self, params, warmup, process_group, flat_grad, reduce, bucket_bytes) -> None:
        """Flat gradient buffer, bucketed reduction, warm-up on a side stream, capture of `self._fwd_bwd`."""
        assert self.image.is_cuda and self.depth.is_cuda, "a captured training step needs CUDA tensors"
        assert reduce in ("bucketed", "after", "none")
        self.group = process_group
        import torch.distributed as dist
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.reduce = reduce if self.world > 1 else "none"
        self.bucketer = None
        self.params = params
        self.offsets, n = flat.flat_offsets(self.params)
        if flat_grad is None:
            self.flat_grad = torch.zeros(n, device=self.image.device, dtype=torch.float32)
            flat.bind_views(self.params, self.flat_grad, "grad")
        else:   # an optimizer (twig/optim.py::FusedAdamW) already owns the flat buffer and bound the views
            assert flat_grad.dtype == torch.float32 and flat.views_match(self.params, flat_grad), \
                "flat_grad does not follow the twig/flat.py layout of these parameters"
            self.flat_grad = flat_grad
        if self.reduce == "bucketed":
            self.bucketer = GradBucketer(self.params, self.flat_grad, process_group, bucket_bytes)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):              # warm-up off the capture stream (allocator, lazy inits, NCCL)
            for i in range(max(2, warmup)):
                self._fwd_bwd()
                if i == 0 and self.bucketer is not None:
                    self.bucketer.calibrate()      # which parameters receive gradients at all (adaptor.* do not)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.loss = self._fwd_bwd()
        self._ptrs = _mark_captured(self, self.params)

    def _fwd_bwd(self) -> torch.Tensor:
        self.flat_grad.zero_()
        if self.bucketer is not None:
            self.bucketer.begin()
        out = TD.texture_prompts_train(self.enc, self.dec, self.image, self.depth, precision=self.precision)
        loss = self.loss_fn(*out)
        loss.backward()                            # accumulates in place into the flat buffer's views; the hooks of
        if self.bucketer is not None:              # twig/buckets.py launch each bucket's all-reduce as it completes
            self.bucketer.finish()
        return loss.detach()

    def close(self) -> None:
        """Remove the gradient hooks (the captured graph keeps working; eager backward of the same parameters outside
        this object no longer triggers reductions).  Also called when the object is garbage-collected."""
        if self.bucketer is not None:
            self.bucketer.enabled = False
            self.bucketer.remove()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __call__(self, image: Optional[torch.Tensor] = None, depth: Optional[torch.Tensor] = None) -> torch.Tensor:
        if image is not None:
            self.image.copy_(image, non_blocking=True)
        if depth is not None:
            self.depth.copy_(depth, non_blocking=True)
        _check_captured(self.params, self._ptrs, "GraphedTrainStep")
        self.graph.replay()
        if self.reduce == "after":
            import torch.distributed as dist
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.AVG, group=self.group)
        return self.loss


class GraphedModelTrainStep(GraphedTrainStep):
    """The training step of the WHOLE model captured once and replayed: `cod.forward(raw, input, label, depth,
    mode='loss')` (cod.py:118-146: backbone with the texture prompts, Hitnet decoder with train-mode BatchNorm, deep
    supervision + SSIM constant) followed by backward into the flat gradient buffer.

        opt  = FusedAdamW(model.named_parameters(), ...)            # first: it re-homes the parameters
        step = GraphedModelTrainStep(model, image, depth, label, flat_grad=opt.flat_grad)
        loss = step(next_image, next_depth, next_label); opt.step()

    ~2900 launches per step at 384^2; the BatchNorm running statistics and `num_batches_tracked` are updated by kernels
    inside the graph, DropPath draws from torch's graph-safe CUDA generator.  Bucketed NCCL reduction as in the base class
    (per-GPU BatchNorm statistics, like the reference's plain nn.BatchNorm2d under DDP)."""

    def __init__(self, model: nn.Module, image: torch.Tensor, depth: torch.Tensor, label: torch.Tensor,
                 precision: Optional[str] = None, warmup: int = 3, process_group=None,
                 flat_grad: Optional[torch.Tensor] = None, reduce: str = "bucketed", bucket_bytes: int = 25 << 20):
        self.model, self.precision = model, precision
        if precision is not None:
            TD.set_precision(model, precision)
        self.image, self.depth = image.detach().clone(), depth.detach().clone()
        self.label = label.detach().clone().float()
        self._capture([p for p in model.parameters() if p.requires_grad], warmup, process_group, flat_grad, reduce,
                      bucket_bytes)

    def _fwd_bwd(self) -> torch.Tensor:
        self.flat_grad.zero_()
        if self.bucketer is not None:
            self.bucketer.begin()
        loss = self.model(None, self.image, self.label, self.depth, mode="loss")["loss"]
        loss.backward()
        if self.bucketer is not None:
            self.bucketer.finish()
        return loss.detach()

    def __call__(self, image: Optional[torch.Tensor] = None, depth: Optional[torch.Tensor] = None,
                 label: Optional[torch.Tensor] = None) -> torch.Tensor:
        if label is not None:
            self.label.copy_(label, non_blocking=True)
        return super().__call__(image, depth)


class GraphedPredict:
    """Low-latency serving of the whole model: `cod.forward(mode='tensor')` (cod.py:147-149) captured once for a
    fixed batch shape and replayed.

        run = GraphedPredict(model, image, depth)        # model = twig.model.hitnet.cod(...).eval()
        logits = run(next_image, next_depth)             # (B,1,H,W) fp32; valid until the next call

    One predict step is ~630 short launches; at batch 1 the GPU work is shorter than the time Python needs to issue
    them, so the eager path is launch-bound.  The replay removes that: inputs are copied into the captured
    buffers, the result is a captured buffer too (clone it if it must outlive the next call)."""

    def __init__(self, model: nn.Module, image: torch.Tensor, depth: torch.Tensor, warmup: int = 2):
        assert image.is_cuda and depth.is_cuda, "GraphedPredict needs CUDA tensors"
        assert not model.training, "GraphedPredict captures the eval-mode forward"
        self.model = model
        self.image, self.depth = image.detach().clone(), depth.detach().clone()
        self.size = tuple(image.shape[-2:])
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                    # warm-up off the capture stream (weight shadows, scratch)
            for _ in range(max(1, warmup)):
                self._forward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.logits = self._forward()
        self.params = list(model.parameters()) + list(model.buffers())
        self._ptrs = _mark_captured(self, self.params)
        self._versions = [p._version for p in self.params]

    @torch.no_grad()
    def _forward(self) -> torch.Tensor:
        _, out = self.model.hitnet.predict_logits(self.image, self.depth, self.size)
        return out

    def __call__(self, image: Optional[torch.Tensor] = None, depth: Optional[torch.Tensor] = None) -> torch.Tensor:
        if image is not None:
            self.image.copy_(image, non_blocking=True)
        if depth is not None:
            self.depth.copy_(depth, non_blocking=True)
        _check_captured(self.params, self._ptrs, "GraphedPredict")
        versions = [p._version for p in self.params]
        if versions != self._versions:
            # weights changed in place (load_state_dict / optimizer step): the captured kernels read re-packed /
            # down-cast / BN-folded SHADOWS of the parameters (texture_diffuser._Packed).  One eager forward refreshes
            # every shadow in place (same storage the graph reads), then the replay sees the new weights.
            self._forward()
            self._versions = versions
        self.graph.replay()
        return self.logits
