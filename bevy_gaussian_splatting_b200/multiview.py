"""Multi-GPU view batch: one process per GPU, the cloud replicated on every GPU,
rank r renders view r, and ONE collective per frame gathers the finished frames on `root`.

torch.distributed is plumbing only (rendezvous + broadcasting the 128-byte NCCL unique id); the
frame gather itself is `bgs_gather_frames` (ncclSend/ncclRecv on the render stream, libbgs).  With
backend="gloo" (CPU tests of the host logic, no GPU) the same rank/view/ordering logic runs over
host tensors; that path never renders anything.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .camera import View, orbit_view


def view_for_rank(rank: int, world: int, width: int = 1920, height: int = 1080) -> View:
    """Config C5: `world` cameras on a circle of radius 5 around (0, 1.5, 0); rank r gets camera r."""
    return orbit_view(rank, world, width, height)


def broadcast_unique_id(make_id, rank: int, root: int = 0) -> bytes:
    """Root creates the 128-byte id (make_id() -> bytes), everyone receives it via torch.distributed."""
    import torch
    import torch.distributed as dist

    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == root:
        raw = make_id()
        assert len(raw) == 128
        buf = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        t = buf.to(dev)
        dist.broadcast(t, src=root)
        buf = t.cpu()
    else:
        dist.broadcast(buf, src=root)
    return bytes(buf.numpy().tobytes())


class MultiViewSession:
    def __init__(self, rank: int, world: int, root: int = 0, plugin=None, share_comm_of: "MultiViewSession | None" = None):
        """`share_comm_of`: reuse another session's NCCL communicator (the contexts of one rank -- frames in flight --
        share ONE communicator; every rank issues its gathers in the same order, which is all NCCL asks for)."""
        self.rank, self.world, self.root, self.plugin = rank, world, root, plugin
        self._comm = C.c_void_p()
        self._lib = None
        self._owns_comm = True
        if plugin is not None and share_comm_of is not None:
            self._lib = plugin._lib
            self._comm = share_comm_of._comm
            self._owns_comm = False
        elif plugin is not None:
            self._lib = plugin._lib

            def make_id():
                raw = (C.c_ubyte * 128)()
                st = self._lib.bgs_nccl_unique_id(raw)
                if st != abi.BGS_OK:
                    raise abi.BgsError(st, "bgs_nccl_unique_id failed (libnccl.so.2 not loadable?)")
                return bytes(raw)

            ident = broadcast_unique_id(make_id, rank, root)
            idbuf = (C.c_ubyte * 128).from_buffer_copy(ident)
            st = self._lib.bgs_nccl_comm_init(plugin._ctx, world, rank, idbuf, C.byref(self._comm))
            if st != abi.BGS_OK:
                raise abi.BgsError(st, "bgs_nccl_comm_init failed")

    def view(self, width: int = 1920, height: int = 1080) -> View:
        return view_for_rank(self.rank, self.world, width, height)

    def gather_device(self, local_ptr: int, all_ptr: int, nbytes: int) -> None:
        """Enqueue the gather on the context's stream (device pointers)."""
        st = self._lib.bgs_gather_frames(self.plugin._ctx, self._comm, self.root, C.c_void_p(local_ptr),
                                         C.c_void_p(all_ptr) if all_ptr else None, nbytes)
        if st != abi.BGS_OK:
            raise abi.BgsError(st, "bgs_gather_frames failed")

    # ---- copy-engine variant (CUDA IPC + peer-to-peer copies; NCCL stays the default)
    def setup_peer_frames(self, device: int, nbytes_per_frame: int) -> int:
        """Root: create the exportable frame array (world x nbytes) and broadcast its IPC handle; others: open it.
        Returns the device pointer every rank pushes into (the root's own array on the root)."""
        import torch
        import torch.distributed as dist

        handle = (C.c_ubyte * 64)()
        self._peer_ptr, self._peer_opened = C.c_void_p(), 0
        # layout of the exported allocation: world frames, then (256 B-aligned) one 32-bit completion word per slot
        self._peer_flags_off = (self.world * nbytes_per_frame + 255) & ~255
        self._peer_seq = 0
        if self.rank == self.root:
            st = self._lib.bgs_peer_buffer_create(device, self._peer_flags_off + 4 * self.world, C.byref(self._peer_ptr), handle)
            if st != abi.BGS_OK:
                raise abi.BgsError(st, "bgs_peer_buffer_create failed")
        t = torch.frombuffer(bytearray(bytes(handle)), dtype=torch.uint8).clone().to(torch.device("cuda", torch.cuda.current_device()))
        dist.broadcast(t, src=self.root)
        if self.rank != self.root:
            raw = (C.c_ubyte * 64).from_buffer_copy(bytes(t.cpu().numpy().tobytes()))
            st = self._lib.bgs_peer_buffer_open(device, raw, C.byref(self._peer_ptr))
            if st != abi.BGS_OK:
                raise abi.BgsError(st, "bgs_peer_buffer_open failed (no peer access between the GPUs?)")
            self._peer_opened = 1
        return int(self._peer_ptr.value)

    def push_device(self, local_ptr: int, nbytes: int, signal: bool = False) -> int:
        """Queue the push of this rank's finished frame into its slot of the root's array.  signal=True also stores
        this push's sequence number (1, 2, ... per session) into the slot's completion word, ordered after the copy on
        the same stream; the root pairs it with wait_frames(stream, sequence).  Returns the sequence."""
        self._peer_seq += 1
        if signal:
            st = self._lib.bgs_push_frame_signal(self.plugin._ctx, C.c_void_p(local_ptr), self._peer_ptr, self.rank, nbytes,
                                                 C.c_void_p(self._peer_ptr.value + self._peer_flags_off), self._peer_seq & 0xFFFFFFFF)
        else:
            st = self._lib.bgs_push_frame(self.plugin._ctx, C.c_void_p(local_ptr), self._peer_ptr, self.rank, nbytes)
        if st != abi.BGS_OK:
            raise abi.BgsError(st, "bgs_push_frame failed")
        return self._peer_seq

    def peer_slot_ptr(self, nbytes: int) -> int:
        """Device address of this rank's slot in the root's frame stack: rendering into it (`render_view_to_device`)
        makes the blend kernel write the frame across NVLink itself; follow it with push_device(slot, nbytes, signal=True),
        which then only stores the completion word."""
        return int(self._peer_ptr.value) + self.rank * nbytes

    def wait_frames(self, stream_ptr: int, sequence: int) -> None:
        """Root: make `stream_ptr` wait (on the device, no host round-trip) until every rank's push number `sequence`
        has landed.  Every rank must queue that push, or the stream never resumes."""
        st = self._lib.bgs_wait_frames(C.c_void_p(stream_ptr), C.c_void_p(self._peer_ptr.value + self._peer_flags_off),
                                       self.world, sequence & 0xFFFFFFFF)
        if st != abi.BGS_OK:
            raise abi.BgsError(st, "bgs_wait_frames failed")

    def peer_flags_ptr(self) -> int:
        return int(self._peer_ptr.value) + self._peer_flags_off

    def release_peer_frames(self) -> None:
        if getattr(self, "_peer_ptr", None) is not None and self._peer_ptr:
            self._lib.bgs_peer_buffer_release(self._peer_ptr, self._peer_opened)
            self._peer_ptr = C.c_void_p()

    def gather_host(self, frame: np.ndarray):
        """gloo stand-in used by the CPU tests: same ordering contract (frames[r] = rank r's frame)."""
        import torch
        import torch.distributed as dist

        t = torch.from_numpy(np.ascontiguousarray(frame))
        out = [torch.empty_like(t) for _ in range(self.world)] if self.rank == self.root else None
        dist.gather(t, out, dst=self.root)
        return None if out is None else np.stack([o.numpy() for o in out])

    def destroy(self):
        if not self._owns_comm:
            self._comm = C.c_void_p()
            return
        if self._comm and self._lib is not None:
            self._lib.bgs_nccl_comm_destroy(self._comm)
            self._comm = C.c_void_p()
