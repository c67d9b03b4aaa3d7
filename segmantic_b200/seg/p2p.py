"""NVLink peer-memory seam exchange of the multi-GPU driver (one process per GPU).

The reference predicts on a single device (``/root/reference/src/segmantic/seg/utils.py:4-12`` keeps ``gpu_ids[0]``);
BASELINE.json's north_star adds a data-parallel driver.  With window OWNERSHIP (``sliding_window.window_partition``)
the only data-path exchange is one push per seam: the importance-weighted logits of the last windows of rank ``r`` also
cover the first output planes of rank ``r + 1``.  ``SeamLink`` moves them WITHOUT a collective library on the data
path: every rank exports its receive buffer and a small flag array as CUDA IPC handles (``sgm_p2p_export``), the
previous rank maps them once (``sgm_p2p_open``) and, per volume, pushes its tail with the copy engines over NVLink
(``sgm_p2p_put`` on a side stream: no SM is taken from the persistent conv kernels) and raises the peer's DATA counter
(``sgm_p2p_signal``); the receiver waits for the counter on its compute stream right before the blend
(``sgm_p2p_wait``) and raises the sender's ACK counter after it, so the next volume's push cannot overtake the blend.
``torch.distributed`` carries the handles (once) and gathers the label slabs -- nothing else.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

from .. import _lib

DATA, ACK, TIMEOUT = 0, 1, 8   # slots of the per-rank int32 flag array


def _export(lib, t: torch.Tensor):
    handle = C.create_string_buffer(64)
    off = C.c_int64(0)
    _lib.check(lib.sgm_p2p_export(t.data_ptr(), handle, C.byref(off)), "sgm_p2p_export")
    return bytes(handle.raw), int(off.value)


class SeamLink:
    """Peer mappings and counters of one rank for a fixed receive buffer (built collectively, once per geometry)."""

    def __init__(self, recv_buf: torch.Tensor, rank: int, world_size: int, group=None, timeout_s: float = 20.0):
        self.lib = _lib.load()
        self.rank, self.world = int(rank), int(world_size)
        self.device = recv_buf.device
        self.recv_buf = recv_buf
        self.timeout_s = float(timeout_s)
        with torch.cuda.device(self.device):
            self.flags = torch.zeros(16, dtype=torch.int32, device=self.device)
            torch.cuda.synchronize(self.device)
            mine = dict(buf=_export(self.lib, recv_buf), flags=_export(self.lib, self.flags), ptr=recv_buf.data_ptr())
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self._opened = {}
        self.next_buf = self.next_flags = self.prev_flags = None
        with torch.cuda.device(self.device):
            if self.rank + 1 < self.world:
                nxt = everyone[self.rank + 1]
                self.next_buf = self._open(*nxt["buf"])
                self.next_flags = self._open(*nxt["flags"])
            if self.rank > 0:
                self.prev_flags = self._open(*everyone[self.rank - 1]["flags"])
        self.step = 0
        self.copy_stream = torch.cuda.Stream(self.device)
        self.tail_done = torch.cuda.Event()
        self.put_done: Optional[torch.cuda.Event] = None
        dist.barrier(group=group)  # every mapping exists before the first push

    def _open(self, handle: bytes, offset: int) -> int:
        base = self._opened.get(handle)
        if base is None:
            out = C.c_void_p()
            _lib.check(self.lib.sgm_p2p_open(handle, C.byref(out)), "sgm_p2p_open")
            base = self._opened[handle] = int(out.value)
        return base + offset

    def close(self):
        for base in self._opened.values():
            try:
                self.lib.sgm_p2p_close(base)
            except Exception:  # noqa: BLE001
                pass
        self._opened = {}

    # ---- per volume --------------------------------------------------------------------------------------------
    def begin(self):
        """Start of a volume on the compute stream: the previous volume's push has left the send region."""
        self.step += 1
        if self.put_done is not None:
            torch.cuda.current_stream(self.device).wait_event(self.put_done)

    def push(self, send_view: torch.Tensor, dst_offset_elems: int = 0):
        """Queue the push of ``send_view`` (float32, just computed on the current stream) into the next rank's
        receive buffer at element ``dst_offset_elems`` -- copy engines on the side stream, then DATA = step."""
        if self.next_buf is None or send_view.numel() == 0:
            return
        cur = torch.cuda.current_stream(self.device)
        self.tail_done.record(cur)
        with torch.cuda.device(self.device):
            t = self.copy_stream
            t.wait_event(self.tail_done)
            st = int(t.cuda_stream)
            flag_me = self.flags.data_ptr()
            # the next rank has blended the previous volume (its blend read the region this push overwrites)
            _lib.check(self.lib.sgm_p2p_wait(flag_me + 4 * ACK, (self.step - 1) & 0xFFFFFFFF, flag_me + 4 * TIMEOUT,
                                             self.timeout_s, st), "sgm_p2p_wait")
            _lib.check(self.lib.sgm_p2p_put(self.next_buf + 4 * int(dst_offset_elems), send_view.data_ptr(),
                                            send_view.numel() * 4, st), "sgm_p2p_put")
            _lib.check(self.lib.sgm_p2p_signal(self.next_flags + 4 * DATA, self.step & 0xFFFFFFFF, st), "sgm_p2p_signal")
            self.put_done = torch.cuda.Event()
            self.put_done.record(t)

    def wait_data(self):
        """On the compute stream, before the blend: the previous rank's windows of this volume have landed."""
        if self.prev_flags is None:
            return
        with torch.cuda.device(self.device):
            flag_me = self.flags.data_ptr()
            _lib.check(self.lib.sgm_p2p_wait(flag_me + 4 * DATA, self.step & 0xFFFFFFFF, flag_me + 4 * TIMEOUT,
                                             self.timeout_s, int(torch.cuda.current_stream(self.device).cuda_stream)),
                       "sgm_p2p_wait")

    def ack(self):
        """On the compute stream, after the blend: tell the previous rank its next push may overwrite the buffer."""
        if self.prev_flags is None:
            return
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sgm_p2p_signal(self.prev_flags + 4 * ACK, self.step & 0xFFFFFFFF,
                                               int(torch.cuda.current_stream(self.device).cuda_stream)), "sgm_p2p_signal")

    def check(self):
        """Synchronise and raise if a wait gave up (a peer died or the protocol is broken)."""
        torch.cuda.synchronize(self.device)
        if int(self.flags[TIMEOUT].item()) != 0:
            self.flags[TIMEOUT] = 0
            raise RuntimeError(f"rank {self.rank}: NVLink seam exchange timed out after {self.timeout_s} s")
