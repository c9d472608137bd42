"""`FluteLinear` and `prepare_model_flute` -- the module-level caller of the hot path.

Mirrors `flute/integrations/base.py:45-326`: same constructor arguments, buffer names (`weight`, `scales`, `tables`,
`tables2`), extra state (`num_bits`, `group_size`, `template_id`) and forward (`flute.qgemm` + in-place bias add), so a
state dict written by the reference loads into this module and vice versa.  What is NOT carried over: the CUTLASS-era
auto-tuner (packing is the canonical tile_P = 32 layout, `flute_b200.tune.tune_and_pack`), bitsandbytes conversion and
accelerate hook juggling (both optional third-party packages that are absent here; a module that carries accelerate
hooks is refused rather than silently stripped).
"""
from __future__ import annotations

import warnings
from typing import Dict, Optional

import torch

from .. import nf_utils, tune, utils
from ..templates import default_template_id


class FluteLinear(torch.nn.Module):
    __constants__ = ["in_features", "out_features", "num_bits", "group_size", "template_id", "num_sms", "workspace_lazy_init"]

    def __init__(self, in_features: int, out_features: int, num_bits: int, group_size: int, template_id: int,
                 workspace_lazy_init: bool = False, bias: bool = False, device: Optional[torch.device] = None,
                 dtype: Optional[torch.dtype] = None) -> None:
        if dtype not in (torch.float16, torch.bfloat16):
            raise NotImplementedError("FluteLinear: fp16 / bf16 only")
        if not isinstance(device, torch.device):
            raise NotImplementedError("FluteLinear: pass a torch.device")
        super().__init__()
        K, N = in_features, out_features
        P, G = int(N / 16 * num_bits), int(K / group_size)
        tables = torch.arange(2 ** num_bits, dtype=dtype, device=device)
        self.in_features, self.out_features = in_features, out_features
        self.num_bits, self.group_size, self.template_id = num_bits, group_size, template_id
        self.workspace_lazy_init = workspace_lazy_init
        if workspace_lazy_init or device.type != "cuda":     # (meta / cpu construction, e.g. when loading a checkpoint)
            self.num_sms, self.workspace = None, None
        else:
            self.num_sms = utils.get_device_num_sms(device)
            self.workspace = utils.get_workspace_streamk(device)
        self.register_buffer("weight", torch.empty((P, K), dtype=torch.int16, device=device))
        self.register_buffer("scales", torch.ones((N, G), dtype=dtype, device=device))
        self.register_buffer("tables", tables)
        self.register_buffer("tables2", utils.make_qmap2_from_qmap(tables))
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_features, device=device, dtype=dtype))
        else:
            self.register_parameter("bias", None)

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        import flute_b200 as flute
        if self.workspace is None:
            num_sms = utils.get_device_num_sms(inputs.device)
            workspace = utils.get_workspace_streamk(inputs.device)
        else:
            num_sms, workspace = self.num_sms, self.workspace
        output = flute.qgemm(inputs, self.weight, self.scales, self.tables, self.tables2, workspace, self.num_bits,
                             self.group_size, self.template_id, num_sms)
        if self.bias is not None:
            output.add_(self.bias)
        return output

    def extra_repr(self) -> str:
        return (f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}, "
                f"num_bits={self.num_bits}, group_size={self.group_size}")

    def get_extra_state(self) -> Dict:
        return {"num_bits": self.num_bits, "group_size": self.group_size, "template_id": self.template_id}

    def set_extra_state(self, state: Dict) -> None:
        if self.num_bits != state["num_bits"] or self.group_size != state["group_size"]:
            raise ValueError("FluteLinear: checkpoint quantisation config differs from the module's")
        if self.template_id is None:
            self.template_id = state["template_id"]
        if self.template_id != state["template_id"]:
            raise ValueError("FluteLinear: checkpoint template_id differs from the module's")


@torch.no_grad()
def prepare_model_flute(name: str, module: torch.nn.Module, num_bits: int, group_size: int, example_batch_size: int = 1,
                        fake: bool = False, custom_scales_dict: Optional[Dict[str, torch.Tensor]] = None,
                        check_correctness: bool = False) -> None:
    """Replace every `nn.Linear` below `module` (fp16 / bf16 weights) by a `FluteLinear` holding its NF-quantised,
    packed weight -- or, with `fake=True`, overwrite the weight by the dense tensor the kernel's arithmetic yields
    (`nf_quantize_2`), which is how quantised models are checked against the kernel (reference tests/vllm.py:57-82).
    Quantisation runs on the weight's own device; packing is the canonical tile_P = 32 layout."""

    def _replace(prefix: str, parent: torch.nn.Module) -> None:
        for child_name, child in parent.named_children():
            full = f"{prefix}.{child_name}"
            if not isinstance(child, torch.nn.Linear):
                _replace(full, child)
                continue
            if child.weight.dtype not in (torch.float16, torch.bfloat16):
                raise NotImplementedError(f"{full}: fp16 / bf16 weights only")
            if getattr(child, "_hf_hook", None) is not None or hasattr(child, "_old_forward"):
                raise ValueError(f"`{full}` carries accelerate hooks; remove them before quantising")
            if fake:
                child.weight = torch.nn.Parameter(nf_utils.nf_quantize_2(child.weight, num_bits, group_size, child.weight.dtype),
                                                  requires_grad=False)
                continue
            dtype, device = child.weight.dtype, child.weight.device
            custom = None if custom_scales_dict is None else custom_scales_dict[full]
            # (on the weight in its OWN dtype, like the reference, integrations/base.py:133-137: code indices then agree
            # with nf_quantize_2's, whose fp16 / bf16 arithmetic decides values next to a pivot differently from fp32)
            _, idx, scales, qmap = nf_utils.nf_quantize(child.weight, num_bits, group_size, custom_scales=custom)
            if int(idx.max()) >= 2 ** num_bits:
                raise ValueError(f"{full}: code index out of range")
            W_idx = idx.to(torch.uint8).T.contiguous()                    # [K, N] code indices
            if device.type == "cuda":
                example = torch.randn(example_batch_size, child.in_features, dtype=dtype, device=device)
                Q, meta = tune.tune_and_pack(inputs=example, weight=W_idx, num_bits=num_bits, group_size=group_size,
                                             check_correctness=check_correctness)
                template_id = meta.template_id
            else:       # packing is pure index arithmetic: a model can be quantised on the host and moved later
                template_id = default_template_id(num_bits)
                Q = utils.pack_tile_p(W_idx, num_bits, 32)
            new = FluteLinear(child.in_features, child.out_features, num_bits, group_size, template_id,
                              bias=child.bias is not None, device=device, dtype=dtype)
            new.weight.copy_(Q.to(device))
            new.scales.copy_(scales.view(new.scales.shape).to(dtype))
            new.tables.copy_(qmap.to(dtype))
            new.tables2.copy_(utils.make_qmap2_from_qmap(qmap.to(dtype)))
            if new.bias is not None:
                new.bias.copy_(child.bias)
            setattr(parent, child_name, new)

    if any(True for _ in module.parameters()) and next(module.parameters()).device.type != "cuda" and not fake:
        warnings.warn("prepare_model_flute: module is not on a CUDA device; FluteLinear.forward needs one")
    _replace(name, module)
