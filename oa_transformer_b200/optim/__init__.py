"""Fused AdamW (one liboat launch for every parameter tensor) with the semantics of `transformers.AdamW`, the optimizer
the reference instantiates through config.initialize('optimizer', transformers, ...) (OATrans/train_dist_multi.py:66;
`transformers.AdamW` no longer exists in transformers 5.x). Same constructor keywords: lr, betas, eps, weight_decay,
correct_bias; `state_dict()` / `load_state_dict()` keep torch.optim's layout (state[i] = {step, exp_avg, exp_avg_sq})."""
import ctypes

import torch

from .. import ops
from .._lib import check, lib, ptr, stream_ptr

_CHUNK = 1024


class AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True):
        if lr < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or eps < 0.0:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias))
        self._tables = {}

    def _table(self, gi, plist):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr()) for p in plist)
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == key:
            return cached[1:]
        dev = plist[0].device
        rows, sizes, prefix = [], [], [0]
        for p in plist:
            st = self.state[p]
            rows.append([p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()])
            sizes.append(p.numel())
            prefix.append(prefix[-1] + (p.numel() + _CHUNK - 1) // _CHUNK)
        table = torch.tensor(rows, dtype=torch.int64, device=dev)
        sizes_t = torch.tensor(sizes, dtype=torch.int64, device=dev)
        prefix_t = torch.tensor(prefix, dtype=torch.int64, device=dev)
        self._tables[gi] = (key, table, prefix_t, sizes_t, len(plist), prefix[-1])
        return self._tables[gi][1:]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            plist = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                assert p.is_cuda and p.dtype == torch.float32 and p.is_contiguous(), "fused AdamW: fp32 CUDA parameters"
                assert p.grad.dtype == torch.float32 and p.grad.is_contiguous()
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                plist.append(p)
            if not plist:
                continue
            steps = {int(self.state[p]["step"]) for p in plist}
            assert len(steps) == 1, "fused AdamW expects every parameter of a group to have taken the same number of steps"
            table, prefix, sizes, n, total = self._table(gi, plist)
            ops._count(1)
            check(lib().oat_adamw_multi(ptr(table), ptr(prefix), ptr(sizes), ctypes.c_int32(n), ctypes.c_int64(total),
                                        ctypes.c_float(group["lr"]), ctypes.c_float(group["betas"][0]),
                                        ctypes.c_float(group["betas"][1]), ctypes.c_float(group["eps"]),
                                        ctypes.c_float(group["weight_decay"]), ctypes.c_int32(steps.pop()),
                                        ctypes.c_int32(1 if group["correct_bias"] else 0), stream_ptr()),
                  "oat_adamw_multi")
        return loss
