"""WeightOnlyQuantLinear / WeightOnlyQuantRowLinear -- mirror of T/tensorrt_llm/quantization/layer.py:268-382 and
of weight_only_quantize()'s module swap (T/tensorrt_llm/models/quantized/quant.py:51-97), as eager torch modules
holding the preprocessed int8 weight, the per-channel fp16 scales and an optional bias."""
import torch

from .. import ops
from .functional import weight_only_quant_matmul
from .mode import QuantMode


class WeightOnlyQuantLinear(torch.nn.Module):

    def __init__(self, in_features, out_features, bias=True, dtype=torch.float16, tp_group=None, tp_size=1,
                 gather_output=True, quant_mode=QuantMode.use_weight_only()):
        super().__init__()
        if not quant_mode.is_int8_weight_only():
            raise ValueError("only int8 weight-only quantization is supported")
        if tp_size != 1:
            raise ValueError("tensor parallelism is out of scope: Whisper shards by utterance (SURVEY.md 8e)")
        self.weight_only_quant_mode = 1
        self.in_features = in_features
        self.out_features = out_features
        # same trick as the reference: the parameter is declared float32 [K, N/4]; it holds int8 bytes
        self.register_buffer("weight", torch.zeros((in_features, out_features // 4), dtype=torch.float32))
        self.register_buffer("per_channel_scale", torch.zeros((out_features,), dtype=dtype))
        if bias:
            self.register_buffer("bias", torch.zeros((out_features,), dtype=dtype))
        else:
            self.bias = None

    @torch.no_grad()
    def load_from_linear_weight(self, weight_out_in, bias=None):
        """weight_out_in: torch Linear weight [out, in]; quantized like examples/whisper/weight.py:76-80."""
        w_kn = weight_out_in.detach().to(torch.float16).t().contiguous()
        proc, scales = ops.symmetric_quantize_last_axis_of_batched_matrix(w_kn.to(self.weight.device), torch.int8)
        self.weight.copy_(proc.view(torch.float32))
        self.per_channel_scale.copy_(scales)
        if bias is not None and self.bias is not None:
            self.bias.copy_(bias.detach().to(self.bias.dtype))

    def forward(self, x):
        # reference: matmul plugin, then a separate `x + bias` layer; here the add is fused into the epilogue with the
        # same rounding (fp16 after the matmul, fp16 after the add)
        return weight_only_quant_matmul(x, self.weight, self.per_channel_scale, self.weight_only_quant_mode,
                                        bias=self.bias)


WeightOnlyQuantColumnLinear = WeightOnlyQuantLinear


class WeightOnlyQuantRowLinear(WeightOnlyQuantLinear):

    def __init__(self, in_features, out_features, bias=True, dtype=torch.float16, tp_group=None, tp_size=1,
                 quant_mode=QuantMode.use_weight_only()):
        super().__init__(in_features, out_features, bias=bias, dtype=dtype, tp_group=tp_group, tp_size=tp_size,
                         quant_mode=quant_mode)


def weight_only_quantize(model, quant_mode, exclude_modules=None):
    """Swaps every torch.nn.Linear for a WeightOnlyQuantLinear (quant.py:51-97 swaps ColumnLinear/RowLinear),
    skipping names in `exclude_modules` (default: 'lm_head', as the reference)."""
    assert quant_mode.is_weight_only()
    exclude_modules = ['lm_head'] if exclude_modules is None else exclude_modules
    for name, module in list(model.named_children()):
        if name in exclude_modules:
            continue
        if isinstance(module, torch.nn.Linear):
            q = WeightOnlyQuantLinear(module.in_features, module.out_features, bias=module.bias is not None,
                                      quant_mode=quant_mode).to(module.weight.device)
            q.load_from_linear_weight(module.weight, module.bias)
            setattr(model, name, q)
        else:
            weight_only_quantize(module, quant_mode, exclude_modules)
    setattr(model, 'quant_mode', quant_mode)
    return model
