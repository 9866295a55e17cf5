"""Bucketed, overlapped gradient all-reduce over the flat gradient buffer (BASELINE configs[2]: data-parallel SOD
training; the reference wraps the model in `MMDistributedDataParallel`, cod.py:8,238, i.e. DDP's bucketed
all-reduce overlapped with backward, `find_unused_parameters` semantics for the 5 grad-less parameters).

The gradients of the hot path already live in ONE flat fp32 buffer (twig/flat.py), so a bucket is a contiguous slice
of it -- no gather / scatter copies as in DDP.  Buckets are cut in REVERSE registration order (the decoders and the
trunk's last stages receive their gradients first) at `bucket_bytes` (DDP default 25 MiB); every parameter gets a
post-accumulate-grad hook, and the hook of the last parameter of a bucket launches `all_reduce(AVG)` on that slice
asynchronously: NCCL runs it on its own stream after the producing kernels, while backward continues on the compute
stream.  `finish()` joins the streams.  Inside a CUDA-graph capture (twig/graphs.py) the same calls become a forked
branch of the captured graph, so a replay overlaps the reduction with backward with zero host work.

Parameters that never receive a gradient (`prompt_encoder.adaptor.*`) would keep their bucket open for ever:
`calibrate()` runs once after a first backward and drops them from the pending counts (the reference needs
`find_unused_parameters=True` for the same reason)."""
from __future__ import annotations

import weakref
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import flat

__all__ = ["GradBucketer"]


class GradBucketer:
    def __init__(self, params: Sequence[torch.nn.Parameter], flat_grad: torch.Tensor, process_group=None,
                 bucket_bytes: int = 25 << 20):
        self.params = list(params)
        self.flat_grad = flat_grad
        self.group = process_group
        offs, total = flat.flat_offsets(self.params)
        assert flat_grad.numel() == total
        # contiguous [lo, hi) element ranges, cut from the END of the buffer towards the start
        self.buckets: List[dict] = []
        cap = max(1, bucket_bytes // 4)
        hi = total
        members: List[int] = []
        for i in range(len(self.params) - 1, -1, -1):
            members.append(i)
            lo = offs[i]
            if hi - lo >= cap or i == 0:
                self.buckets.append({"lo": lo, "hi": hi, "members": members, "pending": 0, "need": len(members)})
                hi, members = lo, []
        self._bucket_of = {}
        for b, bk in enumerate(self.buckets):
            for i in bk["members"]:
                self._bucket_of[i] = b
        self._fired = [False] * len(self.params)
        self._used: Optional[List[bool]] = None       # set by calibrate()
        self._works: list = []
        self._handles = []
        self.enabled = True
        self.launch_order: List[int] = []             # buckets in the order their reductions were launched (last pass)
        for i, p in enumerate(self.params):
            self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))

    # ---------------------------------------------------------------------------------------------
    def _make_hook(self, i: int):
        ref = weakref.ref(self)      # the hook lives on the parameter: it must neither keep the bucketer (and its flat
                                     # buffer) alive nor fire once the owner is gone
        def hook(_param):
            me = ref()
            if me is None or not me.enabled:
                return
            me._fired[i] = True
            if me._used is None:                      # calibration pass: only record who fires
                return
            bk = me.buckets[me._bucket_of[i]]
            bk["pending"] += 1
            if bk["pending"] == bk["need"]:
                me._launch(me._bucket_of[i])
        return hook

    def __del__(self):
        try:
            self.remove()
        except Exception:
            pass

    def _launch(self, b: int) -> None:
        bk = self.buckets[b]
        view = self.flat_grad[bk["lo"]:bk["hi"]]
        self.launch_order.append(b)
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        if view.is_cuda:
            self._works.append(dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        else:                                         # gloo (CPU tests): no AVG
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
            view.div_(dist.get_world_size(self.group))

    # ---------------------------------------------------------------------------------------------
    def begin(self) -> None:
        """Call before every backward."""
        for bk in self.buckets:
            bk["pending"] = 0
        self._fired = [False] * len(self.params)
        self._works = []
        self.launch_order = []

    def calibrate(self) -> None:
        """After ONE backward run with the hooks recording only: parameters that did not fire are treated as unused
        (their slice of the flat buffer stays zero and is reduced with its bucket)."""
        self._used = list(self._fired)
        for bk in self.buckets:
            bk["need"] = sum(1 for i in bk["members"] if self._used[i])

    def finish(self) -> None:
        """Call after backward: launches buckets without any used parameter (all-zero slices stay consistent across
        ranks anyway, so they are skipped) and joins the reductions into the current stream."""
        if self._used is None:
            return
        for b, bk in enumerate(self.buckets):
            if bk["need"] > 0 and bk["pending"] != bk["need"]:
                raise RuntimeError(f"gradient bucket {b} incomplete after backward ({bk['pending']}/{bk['need']}): a "
                                   "parameter that fired during calibration did not receive a gradient")
        for w in self._works:
            w.wait()                                  # stream-level join (no host block on CUDA)
        self._works = []

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []

    def describe(self) -> str:
        sizes = [(bk["hi"] - bk["lo"]) * 4 / 2 ** 20 for bk in self.buckets]
        return f"{len(self.buckets)} buckets of the flat gradient buffer in reverse registration order, " \
               f"{min(sizes):.1f}-{max(sizes):.1f} MiB, all_reduce(AVG) launched from the last gradient hook of each bucket"
