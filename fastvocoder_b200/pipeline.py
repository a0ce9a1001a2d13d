"""Host <-> device pipelining for serving loops: a stream of host-resident mel batches in, host-resident waveforms out.

The generator forward of one batch takes ~10-20 ms on a B200 while its PCIe traffic (80 x T floats in, 240 x T floats out per
utterance) takes ~1-2 ms; issued back to back on one stream (copy in, compute, copy out) the copies add to the step time.
`HostPipeline` runs the three phases on three CUDA streams with double-buffered device inputs, so that the H2D copy of batch
i+1 and the D2H copy of batch i-1 overlap the compute of batch i.  Every batch still makes both trips over PCIe; nothing is
cached or skipped.  (The reference has no counterpart: bin/test.py:123-131 loops over utterances on the CPU.)

    pipe = HostPipeline(model)                      # or HostPipeline(model, fwd=lambda x: model(x, synthesize=True)[1])
    for mel_host, out_host in batches:              # pinned host tensors
        pipe.submit(mel_host, out_host)             # returns immediately
    pipe.finish()                                   # all outputs have landed in their host buffers
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["HostPipeline"]


class HostPipeline:
    def __init__(self, model, fwd=None, depth: int = 2):
        if depth < 2:
            raise ValueError("depth must be >= 2 (double buffering)")
        self.model = model
        self.device = model.device
        if self.device.type != "cuda":
            raise _lib.FvError("HostPipeline needs the model on a CUDA device (no CPU path)")
        self.fwd = fwd if fwd is not None else self._default_fwd
        self.depth = depth
        with torch.cuda.device(self.device):
            self.s_in, self.s_compute, self.s_out = (torch.cuda.Stream(device=self.device) for _ in range(3))
        self._x = [None] * depth              # device input buffers
        self._y = [None] * depth              # device outputs kept alive until their D2H copy was issued
        self._ev_in = [torch.cuda.Event() for _ in range(depth)]
        self._ev_compute = [None] * depth     # compute of the step that last used slot s
        self._ev_out = [None] * depth
        self._i = 0
        # everything enqueued so far on the caller's stream (e.g. weight uploads) happens before the pipeline's first step
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_compute, self.s_out):
            s.wait_stream(cur)

    def _default_fwd(self, x):
        y = self.model(x)
        return y[0] if isinstance(y, tuple) else y

    def submit(self, mel_host: torch.Tensor, out_host: torch.Tensor):
        """Enqueue one batch: mel_host [B, C, T] (pinned for a truly asynchronous copy) -> out_host (pinned, the forward's shape)."""
        s = self._i % self.depth
        self._i += 1
        if self._x[s] is None or self._x[s].shape != mel_host.shape:
            self._x[s] = torch.empty(mel_host.shape, dtype=torch.float32, device=self.device)
        with torch.cuda.stream(self.s_in):
            if self._ev_compute[s] is not None:
                self.s_in.wait_event(self._ev_compute[s])      # the compute that read this input buffer has finished
            self._x[s].copy_(mel_host, non_blocking=True)
            self._ev_in[s].record(self.s_in)
        with torch.cuda.stream(self.s_compute), torch.no_grad():
            self.s_compute.wait_event(self._ev_in[s])
            y = self.fwd(self._x[s])
            ev = torch.cuda.Event()
            ev.record(self.s_compute)
            self._ev_compute[s] = ev
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(ev)
            out_host.copy_(y, non_blocking=True)
            y.record_stream(self.s_out)                         # allocated on the compute stream, read on the copy stream
            evo = torch.cuda.Event()
            evo.record(self.s_out)
            self._ev_out[s] = evo
        self._y[s] = y
        return evo

    def finish(self):
        """Block until every submitted batch has landed in its host buffer."""
        self.s_in.synchronize()
        self.s_compute.synchronize()
        self.s_out.synchronize()
        self._y = [None] * self.depth
