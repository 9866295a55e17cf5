"""Host-to-host inference pipeline of the hot path: pinned host batches in, prompt tokens out.

Three streams: the host->device copy of batch i+1 runs on a copy stream while batch i is on the
compute stream, and the device->host read of a result runs on a third stream, so a step costs
max(compute, copies) instead of their sum (151 MB of RGB-D in, 18.9 MB of tokens out per 64 images).
Two device input buffers are recycled; events order every reuse.
"""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, Optional, Tuple

import torch
import torch.nn as nn

from .model import texture_diffuser as TD


class HostPipeline:
    """pipe = HostPipeline(enc, dec); for res in pipe.run(batches, select, out_bufs): ...

    `batches` yields (image_host, depth_host) pinned CPU tensors of one fixed shape; `select(emb1, emb3,
    tokens)` picks the tensor to bring back; `out_bufs` is a list of >= 2 pinned host tensors written
    round-robin.  Yields (index, host_tensor, done_event); the host tensor is valid after
    done_event.synchronize()."""

    def __init__(self, enc: Optional[nn.Module], dec: Optional[nn.Module], precision: Optional[str] = None,
                 want_embedding3: bool = False, device: Optional[torch.device] = None,
                 forward: Optional[Callable] = None):
        """`forward(*device_tensors) -> tuple` replaces the hot-path call (e.g. the whole model in predict mode
        followed by the metric kernels); a batch may then be any tuple of pinned host tensors."""
        self.enc, self.dec, self.precision, self.want_e3 = enc, dec, precision, want_embedding3
        self.forward = forward
        self.device = device or next(enc.parameters()).device
        self.copy_stream = torch.cuda.Stream(self.device)
        self.out_stream = torch.cuda.Stream(self.device)
        self._bufs = None

    def _buffers(self, *host: torch.Tensor):
        key = tuple((tuple(t.shape), t.dtype) for t in host)
        if self._bufs is None or self._bufs[0] != key:
            mk = lambda t: torch.empty(t.shape, device=self.device, dtype=t.dtype)
            self._bufs = (key, [tuple(mk(t) for t in host) for _ in range(2)],
                          [torch.cuda.Event() for _ in range(2)],    # H2D of slot done
                          [torch.cuda.Event() for _ in range(2)])    # compute finished reading slot
        return self._bufs[1], self._bufs[2], self._bufs[3]

    def run(self, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], select: Callable,
            out_bufs) -> Iterator[Tuple[int, torch.Tensor, torch.cuda.Event]]:
        compute = torch.cuda.current_stream(self.device)
        it = iter(batches)
        nxt = next(it, None)
        if nxt is None:
            return
        bufs, ready, free = self._buffers(*nxt)
        used = [False, False]

        def submit(slot, pair):
            with torch.cuda.stream(self.copy_stream):
                if used[slot]:
                    self.copy_stream.wait_event(free[slot])
                for dst, src in zip(bufs[slot], pair):
                    dst.copy_(src, non_blocking=True)
                ready[slot].record(self.copy_stream)
            used[slot] = True

        submit(0, nxt)
        i = 0
        while nxt is not None:
            slot = i & 1
            nxt = next(it, None)
            if nxt is not None:
                submit(slot ^ 1, nxt)                      # overlaps with the compute of batch i
            compute.wait_event(ready[slot])
            if self.forward is not None:
                out = self.forward(*bufs[slot])
            else:
                out = TD.texture_prompts(self.enc, self.dec, bufs[slot][0], bufs[slot][1], precision=self.precision,
                                         want_embedding3=self.want_e3)
            free[slot].record(compute)
            res = select(*out)
            computed = torch.cuda.Event()
            computed.record(compute)
            host = out_bufs[i % len(out_bufs)]
            done = torch.cuda.Event()
            with torch.cuda.stream(self.out_stream):
                self.out_stream.wait_event(computed)
                host.copy_(res, non_blocking=True)
                done.record(self.out_stream)
            res.record_stream(self.out_stream)
            yield i, host, done
            i += 1
