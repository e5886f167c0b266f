"""Conv1d under the reference's name and constructor (T/tensorrt_llm/layers/conv.py:52-94, added by the hackathon
entry for the Whisper encoder stem).  The reference lowers it to TensorRT's IConvolutionLayer on a [B, C, T, 1] view
with a weight of shape (out, in / groups, k, 1) (conv.py:83-85, examples/whisper/weight.py:52-55); here forward() runs
the tcgen05 implicit-GEMM kernel (b200_conv1d_fp16_tc) through functional.conv1d with the same weight tensor."""
import torch

from .. import functional


class Conv1d(torch.nn.Module):

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 padding_mode='zeros', dtype=torch.float16):
        super().__init__()
        if groups != 1 or dilation != 1 or padding_mode != 'zeros':
            raise ValueError("the Whisper stem uses groups = 1, dilation = 1, zero padding (model.py:135-136)")
        if in_channels % groups != 0 or out_channels % groups != 0:
            raise ValueError("in_channels and out_channels must be divisible by groups")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = (kernel_size, 1)  # the reference's 2-D view of a 1-D kernel
        self.stride, self.padding, self.dilation, self.groups = stride, padding, dilation, groups
        self.padding_mode = padding_mode
        self.register_buffer("weight", torch.zeros((out_channels, in_channels // groups, kernel_size, 1), dtype=dtype))
        if bias:
            self.register_buffer("bias", torch.zeros((out_channels,), dtype=dtype))
        else:
            self.bias = None

    def forward(self, input, activation=None):
        """input [B, C_in, T] fp16 -> [B, C_out, T_out]; `activation` ('gelu') is fused into the kernel's epilogue
        (the reference applies gelu as a separate layer, model.py:154-157)."""
        return functional.conv1d(input, self.weight.squeeze(-1), self.bias, stride=self.stride, padding=self.padding,
                                 activation=activation)
