"""ActionTokenizer (reference: vla/action_tokenizer.py) with the binning on the GPU — integer bins bit-exact.

Same constructor / attributes (`bins`, `bin_centers`, `action_token_begin_idx`, `vocab_size`).  `__call__` keeps the
reference behaviour (decode to text through the wrapped tokenizer); `encode_ids` / `decode_token_ids_to_actions`
expose the integer path used by training batches (`labels`), computed by `mla_action_digitize` /
`mla_action_decode` against the same float64 `np.linspace` edge table the reference builds."""
from __future__ import annotations

import ctypes as C
from typing import List, Union

import numpy as np
import torch

from . import _lib, ops
from ._lib import check


class ActionTokenizer:
    def __init__(self, tokenizer, bins: int = 256, min_action: int = -1, max_action: int = 1) -> None:
        self.tokenizer, self.n_bins, self.min_action, self.max_action = tokenizer, bins, min_action, max_action
        self.bins = np.linspace(min_action, max_action, self.n_bins)
        self.bin_centers = (self.bins[:-1] + self.bins[1:]) / 2.0
        self.action_token_begin_idx: int = int(self.tokenizer.vocab_size - (self.n_bins + 1))
        self._dev = {}

    def _tables(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (torch.from_numpy(self.bins).to(device), torch.from_numpy(self.bin_centers).to(device))
        return self._dev[key]

    def encode_ids(self, action: Union[np.ndarray, torch.Tensor], device="cuda") -> torch.Tensor:
        """Token ids (int64, same shape as `action`): vocab_size - np.digitize(np.clip(action, lo, hi), bins)."""
        a = torch.as_tensor(action)
        if a.dtype not in (torch.float32, torch.float64):
            a = a.to(torch.float64)
        a = a.to(device).contiguous()
        edges, _ = self._tables(a.device)
        ids = torch.empty(a.shape, dtype=torch.int64, device=a.device)
        check(_lib.lib().mla_action_digitize(ops._p(a), C.c_int32(int(a.dtype == torch.float64)), C.c_int64(a.numel()),
                                             ops._p(edges), C.c_int32(self.n_bins), C.c_double(float(self.min_action)),
                                             C.c_double(float(self.max_action)), C.c_int64(self.tokenizer.vocab_size),
                                             ops._p(ids), ops._stream()))
        return ids

    def __call__(self, action: np.ndarray) -> Union[str, List[str]]:
        ids = self.encode_ids(action).cpu().numpy()
        if ids.ndim == 1:
            return self.tokenizer.decode(list(ids))
        return self.tokenizer.batch_decode(ids.tolist())

    def decode_token_ids_to_actions(self, action_token_ids: Union[np.ndarray, torch.Tensor], device="cuda") -> np.ndarray:
        ids = torch.as_tensor(action_token_ids).to(torch.int64).to(device).contiguous()
        _, centers = self._tables(ids.device)
        out = torch.empty(ids.shape, dtype=torch.float64, device=ids.device)
        check(_lib.lib().mla_action_decode(ops._p(ids), C.c_int64(ids.numel()), ops._p(centers),
                                           C.c_int32(self.bin_centers.shape[0]), C.c_int64(self.tokenizer.vocab_size),
                                           ops._p(out), ops._stream()))
        return out.cpu().numpy()

    @property
    def vocab_size(self) -> int:
        return self.n_bins
