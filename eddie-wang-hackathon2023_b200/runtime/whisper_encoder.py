"""WhisperEncoder -- the audio encoder of the quantized Whisper model on one B200.

Counterpart of T/tensorrt_llm/models/whisper/model.py:124-172 (WhisperEncoder: conv stem, positional embedding,
n_audio_layer pre-LN blocks with bidirectional attention, ln_post) with int8 weight-only Linear layers as
T/examples/whisper/weight.py quantizes them.  Every op is a kernel of libb200_whisper.so:

    conv1 + GELU, conv2 (stride 2) + GELU      b200_conv1d_fp16_tc       tcgen05 implicit GEMM
    permute + positional embedding             b200_transpose_add_pos_fp16
    LayerNorm                                  b200_layernorm_fp16
    q|k|v, out, fc1 (+GELU), fc2 (+residual)   b200_woq_int8_gemm_fused  tcgen05, M = B * 1500 rows
    attention over the 1500 frames             b200_attention_bidirectional_fp16

The output feeds WhisperDecoding.set_encoder_output(), which builds the int8 cross-KV caches."""
import torch

from .. import _lib, ops
from .whisper_decoding import _QLinear


class WhisperEncoder:

    def __init__(self, dims, state_dict, device="cuda"):
        self.lib = _lib.load()
        _lib.check(self.lib.b200_init(), "b200_init")
        self.dims = dims
        self.device = torch.device(device)
        self.d, self.H, self.L = dims.n_audio_state, dims.n_audio_head, dims.n_audio_layer
        self.T = dims.n_audio_ctx
        dev, sd = self.device, state_dict
        f16 = lambda t: t.detach().to(device=dev, dtype=torch.float16).contiguous()  # noqa: E731
        self.conv1 = (f16(sd["encoder.conv1.weight"]), f16(sd["encoder.conv1.bias"]))
        self.conv2 = (f16(sd["encoder.conv2.weight"]), f16(sd["encoder.conv2.bias"]))
        self.pos = f16(sd["encoder.positional_embedding"])
        self.ln_post = (f16(sd["encoder.ln_post.weight"]), f16(sd["encoder.ln_post.bias"]))
        self.layers = []
        for i in range(self.L):
            p = f"encoder.blocks.{i}"
            w = torch.cat([sd[p + ".attn.query.weight"], sd[p + ".attn.key.weight"], sd[p + ".attn.value.weight"]], dim=0)
            qb = sd[p + ".attn.query.bias"]
            b = torch.cat([qb, torch.zeros_like(qb), sd[p + ".attn.value.bias"]], dim=0)  # key has no bias
            self.layers.append({
                "attn_ln": (f16(sd[p + ".attn_ln.weight"]), f16(sd[p + ".attn_ln.bias"])),
                "qkv": _QLinear(w, b, dev),
                "attn_out": _QLinear(sd[p + ".attn.out.weight"], sd[p + ".attn.out.bias"], dev),
                "mlp_ln": (f16(sd[p + ".mlp_ln.weight"]), f16(sd[p + ".mlp_ln.bias"])),
                "fc1": _QLinear(sd[p + ".mlp.0.weight"], sd[p + ".mlp.0.bias"], dev),
                "fc2": _QLinear(sd[p + ".mlp.2.weight"], sd[p + ".mlp.2.bias"], dev),
            })
        self._bufs = {}

    def _st(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _buf(self, name, shape, dtype=torch.float16):
        key = (name, tuple(shape), dtype)
        if key not in self._bufs:
            self._bufs[key] = torch.empty(shape, dtype=dtype, device=self.device)
        return self._bufs[key]

    def _conv(self, x, wb, stride, out):
        B, cin, t_in = x.shape
        w, b = wb
        cout, _, k = w.shape
        ws = self._buf("conv_ws", (self.lib.b200_conv1d_workspace_bytes(B, cin, cout, t_in, k),), torch.uint8)
        rc = self.lib.b200_conv1d_fp16_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), B, cin, cout, t_in, k,
                                          stride, 1, _lib.ACT_GELU_ERF, ws.data_ptr(), ws.numel(), self._st())
        _lib.check(rc, "conv1d")

    def _gemm(self, x, rows, lin, out, act=_lib.ACT_NONE, residual=None):
        ws = self._buf("gemm_ws", (max(self.lib.b200_woq_workspace_bytes(rows, 4 * self.d, 4 * self.d), 1 << 20),),
                       torch.uint8)
        rc = self.lib.b200_woq_int8_gemm_fused(
            x.data_ptr(), rows, lin.k, lin.weight.data_ptr(), lin.scales.data_ptr(), lin.n,
            lin.bias.data_ptr() if lin.bias is not None else None, act,
            residual.data_ptr() if residual is not None else None, out.data_ptr(), ws.data_ptr(), ws.numel(), self._st())
        _lib.check(rc, "woq gemm")

    def _ln(self, x, wb, out, rows):
        _lib.check(self.lib.b200_layernorm_fp16(x.data_ptr(), wb[0].data_ptr(), wb[1].data_ptr(), out.data_ptr(), rows,
                                                self.d, 1e-5, self._st()), "layernorm")

    def forward(self, mel):
        """mel [B, n_mels, 2 * n_audio_ctx] (log-mel) -> encoder output [B, n_audio_ctx, d] fp16."""
        B, n_mels, t_in = mel.shape
        assert n_mels == self.dims.n_mels and t_in == 2 * self.T
        d, H, T = self.d, self.H, self.T
        x0 = mel.to(device=self.device, dtype=torch.float16).contiguous()
        c1 = self._buf("c1", (B, d, t_in))
        c2 = self._buf("c2", (B, d, T))
        self._conv(x0, self.conv1, 1, c1)
        self._conv(c1, self.conv2, 2, c2)
        rows = B * T
        x = self._buf("x", (rows, d))
        _lib.check(self.lib.b200_transpose_add_pos_fp16(c2.data_ptr(), self.pos.data_ptr(), x.data_ptr(), B, d, T,
                                                        self._st()), "transpose_add_pos")
        h = self._buf("h", (rows, d))
        qkv = self._buf("qkv", (rows, 3 * d))
        ctx = self._buf("ctx", (rows, d))
        u = self._buf("u", (rows, 4 * d))
        for lay in self.layers:
            self._ln(x, lay["attn_ln"], h, rows)
            self._gemm(h, rows, lay["qkv"], qkv)
            _lib.check(self.lib.b200_attention_bidirectional_fp16(qkv.data_ptr(), ctx.data_ptr(), B, T, H, d // H,
                                                                  self._st()), "bidirectional_attention")
            self._gemm(ctx, rows, lay["attn_out"], x, residual=x)
            self._ln(x, lay["mlp_ln"], h, rows)
            self._gemm(h, rows, lay["fc1"], u, act=_lib.ACT_GELU_ERF)
            self._gemm(u, rows, lay["fc2"], x, residual=x)
        out = torch.empty((rows, d), dtype=torch.float16, device=self.device)
        self._ln(x, self.ln_post, out, rows)
        return out.view(B, T, d)

    __call__ = forward
