"""Observation + legal-mask producer for AlphaZero-style inference (BASELINE config C5).

The reference hands a network its input one state at a time: `ObservationTensor` fills a caller-owned
span with 12*n*(n-2) floats (twixt.cc:101-132) and `LegalActions` (twixt.h:86-90) gives the policy mask.
Here one kernel launch (`twixt_observation_and_mask`) writes both for a whole batch of envs straight into
device tensors -- `[B, 12, n, n-2]` float32 and `[B, n*n]` uint8 -- reading each env's record from HBM once.
The outputs are ordinary torch CUDA tensors and are also offered through DLPack, so a consumer in another
framework takes them without a copy; a consumer may equally LEND its own buffers through DLPack.

torch is plumbing here (device memory, streams, the DLPack capsule); the product is the kernel.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .batch import TwixTBatch


class ObservationMaskProducer:
    def __init__(self, batch: TwixTBatch, max_count: Optional[int] = None, use_torch_stream: bool = True):
        self.batch = batch
        self.max_count = int(max_count if max_count is not None else batch.num_envs)
        self.device = torch.device("cuda", batch.device)
        n = batch.board_size
        # the producer's own output buffers (reused by every produce() call that brings none)
        self.obs = torch.empty((self.max_count,) + tuple(batch.obs_shape), dtype=torch.float32, device=self.device)
        self.mask = torch.empty((self.max_count, n * n), dtype=torch.uint8, device=self.device)
        if use_torch_stream:
            batch.use_torch_stream()  # the kernel is ordered with the consumer's torch work, no sync needed

    @staticmethod
    def _adopt(x):
        """A torch tensor, or anything that speaks DLPack (`__dlpack__`), as a torch view of the same memory."""
        if x is None or isinstance(x, torch.Tensor):
            return x
        return torch.from_dlpack(x)

    def produce(self, first: int = 0, count: Optional[int] = None, out_obs=None,
                out_mask=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """obs [count,12,n,n-2] f32 and mask [count,n*n] u8 of envs [first, first+count), on the device.

        out_obs / out_mask: optional destination buffers (torch tensors or DLPack exporters) of at least that
        size; without them views of the producer's own buffers are returned (valid until the next call)."""
        if count is None:
            count = min(self.batch.num_envs - first, self.max_count)
        obs = self._adopt(out_obs)
        mask = self._adopt(out_mask)
        if obs is None:
            if count > self.max_count:
                raise ValueError("count %d > max_count %d" % (count, self.max_count))
            obs = self.obs[:count]
        if mask is None:
            if count > self.max_count:
                raise ValueError("count %d > max_count %d" % (count, self.max_count))
            mask = self.mask[:count]
        for t, dt in ((obs, torch.float32), (mask, torch.uint8)):
            if not t.is_cuda or t.device.index != self.batch.device or t.dtype != dt or not t.is_contiguous():
                raise ValueError("output must be a contiguous %s tensor on cuda:%d" % (dt, self.batch.device))
        self.batch.observation_and_mask(first, count, out_obs=obs, out_mask=mask)
        n = self.batch.board_size
        return obs.view((count,) + tuple(self.batch.obs_shape)), mask.view(count, n * n)

    def produce_dlpack(self, first: int = 0, count: Optional[int] = None):
        """The same two arrays as DLPack capsules (zero-copy hand-over to another framework).

        The capsules are made on the stream the kernel ran on, so a consumer that honours DLPack's stream
        protocol needs no extra synchronisation."""
        obs, mask = self.produce(first, count)
        return torch.utils.dlpack.to_dlpack(obs), torch.utils.dlpack.to_dlpack(mask)
