"""LayerNorm under the reference's name (T/tensorrt_llm/layers/normalization.py), fp16 parameters, fp32 statistics
(b200_layernorm_fp16)."""
import torch

from .. import functional


class LayerNorm(torch.nn.Module):

    def __init__(self, normalized_shape, eps=1e-05, elementwise_affine=True, dtype=torch.float16):
        super().__init__()
        if isinstance(normalized_shape, int):
            normalized_shape = (normalized_shape,)
        self.normalized_shape = tuple(normalized_shape)
        self.eps = eps
        self.register_buffer("weight", torch.ones(self.normalized_shape, dtype=dtype))
        self.register_buffer("bias", torch.zeros(self.normalized_shape, dtype=dtype))

    def forward(self, x):
        return functional.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
