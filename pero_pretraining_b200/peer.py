"""Peer-memory exchange buffers for the sharded modes of the path (include/pero_b200.h
"peer-memory collectives").

A ``PeerBuffer`` is one symmetric allocation per rank, mapped by every rank of the group over NVLink 5 /
NVSwitch (``torch.distributed._symmetric_memory`` does the allocation and the handle exchange — plumbing only).
Kernels that produce an exchanged quantity (EMA sums|counts, d_W|d_b|loss, packed winners) write straight into
a range of the buffer; ``all_reduce_sum_`` / ``all_reduce_min_`` then run ONE libpero_b200.so kernel per rank
that reduces the range in place through the switch (multimem.ld_reduce + multimem.st) or, without a multicast
object, with peer loads/stores.  Nothing here calls NCCL on the data path.
"""
import torch
import torch.distributed as dist

from . import _lib
from ._lib import check

HEADER_BYTES = 16384          # PERO_PEER_HEADER_BYTES
TIMEOUT_OFFSET = 12288        # PERO_PEER_TIMEOUT_OFFSET: u32 milliseconds, 0 = default (600 s, NCCL's watchdog)
ERROR_OFFSET = 12292          # PERO_PEER_ERROR_OFFSET: non-zero after a barrier timed out
DEFAULT_BLOCKS = 24


def _round_up(a, b):
    return (a + b - 1) // b * b


class PeerBuffer:
    """[header flags | payload] symmetric buffer.  ``carve(nbytes)`` hands out 256-byte aligned payload ranges
    (same sequence of carves on every rank -> same offsets everywhere)."""

    def __init__(self, payload_bytes, device, group=None, n_blocks=DEFAULT_BLOCKS, use_multicast=None, timeout_s=None):
        """timeout_s: how long an exchange kernel waits for a late rank (default 600 s, like NCCL's watchdog).  A kernel
        that gives up does not trap: it flags the buffer (see check()) and leaves the range unreduced.
        use_multicast: None = switch-side reduction whenever the group has a multicast object, True / False to
        force.  Per-direction link traffic for a payload S: multimem 1.5 S at 2 ranks, 1.125 S at 8; peer pointers
        1.0 S at 2 ranks, 1.75 S at 8 — but the multimem kernel needs far fewer threads to keep the links busy
        (it is the switch that fans out), which matters beside the GEMMs it overlaps with."""
        import torch.distributed._symmetric_memory as symm_mem
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group if group is not None else dist.group.WORLD
        self.device = torch.device(device)
        self.total = HEADER_BYTES + _round_up(int(payload_bytes), 256)
        self.storage = symm_mem.empty(self.total, dtype=torch.uint8, device=self.device)
        self.handle = symm_mem.rendezvous(self.storage, self.group)
        self.storage.zero_()
        if timeout_s is not None:
            self.storage[TIMEOUT_OFFSET:TIMEOUT_OFFSET + 4].view(torch.int32).fill_(max(1, int(timeout_s * 1000)))
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)       # every rank's flag words are zero before any kernel touches them
        self.rank, self.world = int(self.handle.rank), int(self.handle.world_size)
        self.ptrs_dev = int(self.handle.buffer_ptrs_dev)
        if use_multicast is None:
            use_multicast = True
        mc = int(self.handle.multicast_ptr) if use_multicast else 0
        self.multicast = mc if mc != 0 else None
        self.n_blocks = int(n_blocks)
        self._next = HEADER_BYTES

    def check(self):
        """Raise if an exchange kernel of this rank ever gave up waiting for a peer (synchronises the device)."""
        code = int(self.storage[ERROR_OFFSET:ERROR_OFFSET + 4].view(torch.int32).item())
        if code != 0:
            raise _lib.PeroError(f"peer exchange timed out waiting for rank {code & 0xff} (block {(code >> 8) & 0xff}); "
                                 "the last exchanged ranges are not reduced")

    @property
    def transport(self):
        return "nvswitch-multimem" if self.multicast else "nvlink-p2p"

    def carve(self, nbytes):
        """Offset (bytes from the buffer base) of a fresh payload range."""
        off = self._next
        end = off + _round_up(int(nbytes), 256)
        if end > self.total:
            raise ValueError(f"PeerBuffer exhausted: need {end} bytes, have {self.total}")
        self._next = end
        return off

    def view(self, offset_bytes, numel, dtype):
        """Local tensor view of a payload range (what the producing kernel writes into)."""
        nbytes = int(numel) * torch.empty((), dtype=dtype).element_size()
        return self.storage[offset_bytes:offset_bytes + nbytes].view(dtype)

    def _call(self, fn, name, offset_bytes, numel, n_blocks):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(fn(self.ptrs_dev, self.multicast, self.rank, self.world, int(offset_bytes), int(numel),
                 int(n_blocks or self.n_blocks), stream), name)

    def all_reduce_sum_(self, offset_bytes, numel, n_blocks=None):
        """In-place fp32 SUM of [offset, offset + 4*numel) over all ranks, on the current stream."""
        self._call(_lib.lib().pero_peer_allreduce_sum_f32, "pero_peer_allreduce_sum_f32", offset_bytes, numel, n_blocks)

    def all_reduce_min_(self, offset_bytes, numel, n_blocks=None):
        """In-place int64 MIN of [offset, offset + 8*numel) over all ranks, on the current stream."""
        self._call(_lib.lib().pero_peer_allreduce_min_i64, "pero_peer_allreduce_min_i64", offset_bytes, numel, n_blocks)


class PeerRange:
    """A typed payload range of a PeerBuffer: `.tensor` is the local view, `.all_reduce_*_()` exchange it."""

    def __init__(self, buf, numel, dtype):
        self.buf = buf
        self.numel = int(numel)
        pad = 4 if dtype == torch.float32 else 2
        self.padded = _round_up(self.numel, pad)
        self.offset = buf.carve(self.padded * torch.empty((), dtype=dtype).element_size())
        self.full = buf.view(self.offset, self.padded, dtype)
        self.tensor = self.full[:self.numel]
        if self.padded != self.numel:
            self.full[self.numel:].zero_()

    def all_reduce_sum_(self, n_blocks=None):
        self.buf.all_reduce_sum_(self.offset, self.padded, n_blocks)
        return self.tensor

    def all_reduce_min_(self, n_blocks=None):
        self.buf.all_reduce_min_(self.offset, self.padded, n_blocks)
        return self.tensor


def emulate_all_reduce(buffers, op, offset_bytes, numel, n_blocks=4):
    """Single-GPU test of the exchange protocol: `buffers` are `world` equal-size uint8 tensors on ONE device
    (header zeroed); the ranks are played by blockIdx.y of one cooperative launch."""
    dev = buffers[0].device
    ptrs = torch.tensor([b.data_ptr() for b in buffers], dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    check(_lib.lib().pero_peer_allreduce_emulate(ptrs.data_ptr(), len(buffers), {"sum": 0, "min": 1}[op], int(offset_bytes),
                                                 int(numel), int(n_blocks), stream), "pero_peer_allreduce_emulate")
    torch.cuda.synchronize(dev)     # `ptrs` must outlive the kernel
