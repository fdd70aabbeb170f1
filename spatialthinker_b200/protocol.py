"""A minimal stand-in for the slice of ``verl.protocol.DataProto`` the actor loop touches.

The reference's container (protocol.py:166-689) wraps a TensorDict plus numpy side data and meta info; the actor only
uses ``select`` (:326), ``split`` (:521) / ``chunk`` (:488), ``.batch[...]`` and ``.meta_info``. Any object with those
members (the real DataProto included) can be handed to :class:`spatialthinker_b200.dp_actor.DataParallelPPOActor`;
this class exists so the path runs without tensordict / ray.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch


class TensorBatch:
    def __init__(self, batch: Dict[str, torch.Tensor], non_tensor_batch: Optional[Dict[str, np.ndarray]] = None,
                 meta_info: Optional[Dict[str, Any]] = None):
        sizes = {v.shape[0] for v in batch.values()} | {len(v) for v in (non_tensor_batch or {}).values()}
        if len(sizes) > 1:
            raise ValueError(f"inconsistent batch sizes: {sorted(sizes)}")
        self.batch = dict(batch)
        self.non_tensor_batch = dict(non_tensor_batch or {})
        self.meta_info = dict(meta_info or {})

    def __len__(self) -> int:
        for v in self.batch.values():
            return v.shape[0]
        for v in self.non_tensor_batch.values():
            return len(v)
        return 0

    def select(self, batch_keys: Optional[Sequence[str]] = None, non_tensor_batch_keys: Optional[Sequence[str]] = None,
               meta_info_keys: Optional[Sequence[str]] = None) -> "TensorBatch":
        b = self.batch if batch_keys is None else {k: self.batch[k] for k in batch_keys}
        n = self.non_tensor_batch if non_tensor_batch_keys is None else {k: self.non_tensor_batch[k] for k in non_tensor_batch_keys}
        m = self.meta_info if meta_info_keys is None else {k: self.meta_info[k] for k in meta_info_keys}
        return TensorBatch(b, n, m)

    def chunk(self, chunks: int) -> List["TensorBatch"]:
        n = len(self)
        assert n % chunks == 0, f"only support equal chunk. Got size of DataProto {n} and chunk {chunks}."
        step = n // chunks
        out = []
        for i in range(chunks):
            sl = slice(i * step, (i + 1) * step)
            out.append(TensorBatch({k: v[sl] for k, v in self.batch.items()},
                                   {k: v[sl] for k, v in self.non_tensor_batch.items()}, self.meta_info))
        return out

    def split(self, split_size: int) -> List["TensorBatch"]:
        return self.chunk(len(self) // split_size)

    def take(self, index: Sequence[int]) -> "TensorBatch":
        """Rows ``index`` in that order (what ``torch.cat([batch[i:i+1] for i in partition])`` builds in the reference's
        ``rearrange_micro_batches``, seqlen_balancing.py:245-251)."""
        idx = torch.as_tensor(list(index), dtype=torch.long)
        batch = {k: v.index_select(0, idx.to(v.device)) for k, v in self.batch.items()}
        return TensorBatch(batch, {k: v[idx.numpy()] for k, v in self.non_tensor_batch.items()}, self.meta_info)

    def to(self, device) -> "TensorBatch":
        self.batch = {k: v.to(device, non_blocking=True) for k, v in self.batch.items()}
        return self
