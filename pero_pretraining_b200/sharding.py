"""Multi-GPU partitioning of the path over one NVLink/NVSwitch box.  The reference is
single-process; both modes target "equals the single-process run on the concatenated batch".

  batch-sharded   frames split by whole lines across ranks, codebook / head replicated.  The only exchange
                  is one SUM all-reduce of the [K, D+1] EMA sums|counts buffer (VectorQuantizer.
                  enable_data_parallel) and, for the masked CE, of (loss_sum, M) and (d_W, d_b)
                  (MaskedTransformerEncoder.enable_data_parallel / LinearHead.masked_loss(dp_group=...)).
  codebook-sharded every rank sees all frames and K/g contiguous codewords; local winners are packed as the
                  signed word (order_key(distance) << 32 | global index) and merged with ONE int64 MIN
                  all-reduce, which picks the nearest codeword and the lowest index on exact ties.

On GPUs the exchanges run in libpero_b200's own peer-memory kernels (peer.PeerBuffer: NVSwitch multimem
reduction, or NVLink peer loads/stores); torch.distributed collectives (NCCL, or gloo in the CPU tests of the
semantics) remain as the peer=False path.  Either way the exchange is enqueued on the current stream directly
after the producing kernel.
"""
import torch
import torch.distributed as dist

from . import ops

INT64_MAX = 0x7FFFFFFFFFFFFFFF


def shard_bounds(total, world_size, rank):
    """Contiguous [lo, hi) of `total` items owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(int(total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_dist_index_reference(dmin, idx):
    """Pure-torch statement of the packing done on the device (csrc/ptx.cuh pack_dist_index); used by the
    CPU tests of the merge rule and to merge results that were produced unpacked."""
    b = dmin.contiguous().view(torch.int32).to(torch.int64)
    key = b ^ ((b >> 31) & 0x7FFFFFFF)
    return (key << 32) | (idx.to(torch.int64) & 0xFFFFFFFF)


def unpack_index_reference(packed):
    return packed & 0xFFFFFFFF


def merge_packed(packed, group=None):
    """In-place MIN all-reduce of the packed (distance, index) words across codebook shards."""
    dist.all_reduce(packed, op=dist.ReduceOp.MIN, group=group)
    return packed


class ShardedCodebook:
    """Rank-local slice [k_lo, k_hi) of a [K, D] codebook for codebook-sharded assignment."""

    def __init__(self, full_weight_or_shard, K_total, rank, world_size, group=None, is_shard=False, peer_frames=0):
        """peer_frames > 0: packed winners of up to that many frames are merged by libpero_b200's own int64 MIN
        all-reduce in a peer-mapped buffer (NVLink/NVSwitch) instead of torch.distributed."""
        self.rank, self.world_size, self.group = rank, world_size, group
        self._peer = None
        self._peer_frames = int(peer_frames)
        self.K_total = int(K_total)
        self.k_lo, self.k_hi = shard_bounds(K_total, world_size, rank)
        w = full_weight_or_shard if is_shard else full_weight_or_shard[self.k_lo:self.k_hi]
        self.weight = w.detach().float().contiguous()
        assert self.weight.shape[0] == self.k_hi - self.k_lo
        self.codebook = ops.PreparedCodebook(self.weight.shape[0], self.weight.shape[1], self.weight.device)
        self.codebook.prepare(self.weight)
        if self._peer_frames > 0:
            from .peer import PeerBuffer
            self._peer = PeerBuffer(8 * (self._peer_frames + 2), self.weight.device, group)
            self._peer_off = self._peer.carve(8 * (self._peer_frames + 2))

    def assign(self, x, n_lines, frames_per_line, channels_first, want_dmin=False):
        """Global nearest-codeword index for every frame (all ranks return the same tensor)."""
        N = int(n_lines) * int(frames_per_line)
        if self._peer is not None and N <= self._peer_frames:
            n_pad = N + (N & 1)
            packed = self._peer.view(self._peer_off, n_pad, torch.int64)
            ops.vq_packed_init(n_pad, x.device, out=packed)
            ops.vq_assign(x, self.codebook, n_lines, frames_per_line, channels_first, index_offset=self.k_lo,
                          packed=packed[:N])
            self._peer.all_reduce_min_(self._peer_off, n_pad)
            return ops.vq_unpack(packed[:N], want_dmin)
        packed = ops.vq_packed_init(N, x.device)
        ops.vq_assign(x, self.codebook, n_lines, frames_per_line, channels_first, index_offset=self.k_lo, packed=packed)
        if self.world_size > 1:
            merge_packed(packed, self.group)
        return ops.vq_unpack(packed, want_dmin)
