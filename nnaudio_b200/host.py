"""Host-buffer entry point: spectrograms of waveforms that live in (pinned) host
memory, with the PCIe copies overlapped with the kernels.

The batch is cut into chunks of clips; chunk i+1 is copied host->device on a
copy stream while chunk i is transformed on the compute stream and chunk i-1 is
copied device->host on a third stream (the two DMA directions are independent
engines), so the call costs ~max(H2D, compute, D2H) instead of their sum.
"""
from __future__ import annotations

from typing import Optional

import torch


class HostPipeline:
    def __init__(self, module: torch.nn.Module, chunk_clips: int = 8, copy_streams: int = 1,
                 **forward_kwargs):
        """``copy_streams`` > 1 splits each chunk's host->device copy over several streams
        (several DMA engines can then drive the PCIe link together)."""
        self.module = module
        self.chunk = int(chunk_clips)
        self.n_in = max(1, int(copy_streams))
        self.kw = forward_kwargs
        self._dev = None
        self._in = None
        self._streams = None

    def _setup(self, device, L):
        if self._dev != (device, L):
            self._in = [torch.empty((self.chunk, L), dtype=torch.float32, device=device) for _ in range(2)]
            self._streams = (torch.cuda.Stream(device), torch.cuda.Stream(device))
            self._extra_in = [torch.cuda.Stream(device) for _ in range(self.n_in - 1)]
            self._dev = (device, L)

    @torch.no_grad()
    def __call__(self, x_host: torch.Tensor, out_host: Optional[torch.Tensor] = None,
                 device: Optional[torch.device] = None) -> torch.Tensor:
        """x_host: (B, L) fp32 pinned CPU tensor -> spectrograms in a pinned CPU
        tensor (allocated on first use if ``out_host`` is None).  Asynchronous with
        respect to the host: the CURRENT stream waits for the last copy, so
        ``torch.cuda.current_stream().synchronize()`` (or an event recorded on it)
        marks completion."""
        if x_host.is_cuda or x_host.dtype != torch.float32 or x_host.dim() != 2:
            raise ValueError("x_host must be a (B, L) float32 CPU tensor")
        if not x_host.is_pinned():
            raise ValueError("x_host must be pinned (torch.Tensor.pin_memory()) for asynchronous copies")
        device = device or next(self.module.buffers()).device
        B, L = x_host.shape
        self._setup(device, L)
        s_in, s_out = self._streams
        cur = torch.cuda.current_stream(device)
        s_in.wait_stream(cur)
        s_out.wait_stream(cur)
        for se in self._extra_in:
            se.wait_stream(cur)
        done_compute = [None, None]
        n_chunks = (B + self.chunk - 1) // self.chunk
        for c in range(n_chunks):
            lo, hi = c * self.chunk, min(B, (c + 1) * self.chunk)
            buf = self._in[c & 1][: hi - lo]
            in_streams = [s_in] + self._extra_in
            n_rows = hi - lo
            parts = min(len(in_streams), n_rows)
            for pi in range(parts):
                r0, r1 = (n_rows * pi) // parts, (n_rows * (pi + 1)) // parts
                st = in_streams[pi]
                with torch.cuda.stream(st):
                    if done_compute[c & 1] is not None:
                        st.wait_event(done_compute[c & 1])  # buffer still read by chunk c-2
                    buf[r0:r1].copy_(x_host[lo + r0: lo + r1], non_blocking=True)
                    ev_in = torch.cuda.Event()
                    ev_in.record(st)
                cur.wait_event(ev_in)
            y = self.module(buf, **self.kw)
            ev_c = torch.cuda.Event()
            ev_c.record(cur)
            done_compute[c & 1] = ev_c
            if out_host is None:
                out_host = torch.empty((B,) + tuple(y.shape[1:]), dtype=y.dtype).pin_memory()
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_c)
                out_host[lo:hi].copy_(y, non_blocking=True)
            y.record_stream(s_out)
        cur.wait_stream(s_out)
        return out_host
