"""AdamW with HuggingFace `transformers.AdamW` semantics (the optimizer type the reference's configs name;
removed from transformers 5.x) on one fused multi-tensor CUDA kernel."""
import ctypes as C
import math

import torch

from ._lib import call, lib


class _TensorDesc(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_longlong)]


class AdamW(torch.optim.Optimizer):
    """transformers.AdamW(params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True):
        if lr < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or eps < 0.0:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias))
        self._tables = {}

    def _table(self, gi, plist):
        # the device-side pointer table is rebuilt whenever a tensor moved (new .grad after zero_grad(set_to_none),
        # optimizer state replaced by load_state_dict, ...)
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                     self.state[p]["exp_avg_sq"].data_ptr()) for p in plist)
        ent = self._tables.get(gi)
        if ent is not None and ent[0] == key:
            return ent[1:]
        chunk = lib().egovlp_adamw_chunk_elems()
        descs = (_TensorDesc * len(plist))()
        ct, co = [], []
        for i, p in enumerate(plist):
            st = self.state[p]
            descs[i] = _TensorDesc(p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(),
                                   st["exp_avg_sq"].data_ptr(), p.numel())
            n = (p.numel() + chunk - 1) // chunk
            ct += [i] * n
            co += list(range(n))
        dev = plist[0].device
        raw = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(dev)
        ct_t = torch.tensor(ct, dtype=torch.int32, device=dev)
        co_t = torch.tensor(co, dtype=torch.int32, device=dev)
        self._tables[gi] = (key, raw, ct_t, co_t)
        return raw, ct_t, co_t

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            for p in plist:
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                assert p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()
                st["step"] += 1
            t = self.state[plist[0]]["step"]
            b1, b2 = group["betas"]
            step_size = group["lr"]
            if group["correct_bias"]:
                step_size = step_size * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
            raw, ct, co = self._table(gi, plist)
            call("egovlp_adamw_multi", C.c_void_p(raw.data_ptr()), C.c_void_p(ct.data_ptr()), C.c_void_p(co.data_ptr()),
                 ct.numel(), C.c_float(group["lr"]), C.c_float(b1), C.c_float(b2), C.c_float(group["eps"]),
                 C.c_float(group["weight_decay"]), C.c_float(step_size), C.c_void_p(0),
                 C.c_void_p(torch.cuda.current_stream().cuda_stream))
            for p in plist:
                torch.autograd.graph.increment_version(p)   # the kernel wrote p.data: let bf16 caches / autograd see it
        return loss
