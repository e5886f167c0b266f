"""build.py -- the reference's `examples/whisper/build.py` entry point (T/examples/whisper/build.py:42-143 flags,
:145-330 the three builders) over this library: checkpoint `{dims, model_state_dict}` + `quantize_dir` calibration
files in, an ENGINE DIRECTORY out.

The reference serializes three TensorRT engines (whisper_encoder / whisper_decoder / whsiper_crossattn [sic]) plus a
JSON config per builder.  TensorRT is not part of this library: an "engine" here is the weight container the B200
runtime consumes -- per model one file under the reference's engine name holding the int8 weights ALREADY in the
preprocessed layout (quantized on the GPU by b200_whisper.ops.symmetric_quantize..., bit-identical with the reference's
preprocessor), the fp16 scales / biases / LayerNorm vectors, and the KV-cache scales; the config JSON files carry the
builder settings under the reference's keys.  run.py loads the directory without touching the checkpoint again.

Same flags, same defaults.  Only the configuration on the hot path builds: `--use_weight_only --weight_only_precision
int8` (int4, fp16 Linear layers and tensor parallelism raise), `--int8_kv_cache` optional, `--use_gpt_attention_plugin`
accepted with the reference's meaning (the attention always runs through the GPTAttention-plugin kernels here)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODEL_ENCODER_NAME = "whisper_encoder"
MODEL_DECODER_NAME = "whisper_decoder"
MODEL_CROSSATTN_NAME = "whsiper_crossattn"  # the reference's spelling (build.py:28); kept so directories interchange


def get_engine_name(model, dtype, tp_size, rank):
    return '{}_{}_tp{}_rank{}.engine'.format(model, dtype, tp_size, rank)


def parse_arguments(args=None):
    from b200_whisper.quantization import QuantMode
    parser = argparse.ArgumentParser()
    parser.add_argument('--world_size', type=int, default=1, help='world size, only support tensor parallelism now')
    parser.add_argument('--model_dir', type=str, default="large-v2.pt")
    parser.add_argument('--quantize_dir', type=str, default="quantize/1-gpu")
    parser.add_argument('--dtype', type=str, default='float16', choices=['float16', 'float32', 'bfloat16'])
    parser.add_argument('--log_level', type=str, default='info')
    parser.add_argument('--max_batch_size', type=int, default=256)
    parser.add_argument('--max_input_len', type=int, default=200)
    parser.add_argument('--max_output_len', type=int, default=200)
    parser.add_argument('--max_beam_width', type=int, default=1)
    parser.add_argument('--use_gpt_attention_plugin', nargs='?', const=None, type=str, default=False, choices=['float16'],
                        help="Activates attention plugin. You can specify the plugin dtype or leave blank to use the model dtype.")
    parser.add_argument('--use_gemm_plugin', nargs='?', const=None, type=str, default=False,
                        choices=['float16', 'float32', 'bfloat16'],
                        help="Activates GEMM plugin. You can specify the plugin dtype or leave blank to use the model dtype.")
    parser.add_argument('--use_layernorm_plugin', nargs='?', const=None, type=str, default=False,
                        choices=['float16', 'float32', 'bfloat16'],
                        help="Activates layernorm plugin. You can specify the plugin dtype or leave blank to use the model dtype.")
    parser.add_argument('--output_dir', type=str, default='whisper_outputs',
                        help='The path to save the serialized engine files, timing cache file and model configs')
    parser.add_argument('--use_weight_only', default=False, action="store_true",
                        help='Quantize weights for the various GEMMs to INT4/INT8.'
                             'See --weight_only_precision to set the precision')
    parser.add_argument('--weight_only_precision', const='int8', type=str, nargs='?', default='int8', choices=['int8', 'int4'],
                        help='Define the precision for the weights when using weight-only quantization.'
                             'You must also use --use_weight_only for that argument to have an impact.')
    parser.add_argument('--int8_kv_cache', default=False, action="store_true",
                        help='By default, we use dtype for KV cache. int8_kv_cache chooses int8 quantization for KV')
    args = parser.parse_args(args)

    for plugin_arg in ('use_gemm_plugin', 'use_layernorm_plugin', 'use_gpt_attention_plugin'):
        if getattr(args, plugin_arg) is None:  # flag given without a value: the model dtype (build.py:123-130)
            setattr(args, plugin_arg, args.dtype)
    if args.use_weight_only:
        args.quant_mode = QuantMode.use_weight_only(args.weight_only_precision == 'int4')
    else:
        args.quant_mode = QuantMode(0)
    if args.int8_kv_cache:
        args.quant_mode = args.quant_mode.set_int8_kv_cache()
    return args


def check_supported(args):
    if args.world_size != 1:
        raise ValueError("tensor parallelism is not on the Whisper hot path (20 heads do not divide 8 GPUs): "
                         "utterances are sharded across GPUs instead, one engine directory serves every rank")
    if not args.quant_mode.is_int8_weight_only():
        raise ValueError("only --use_weight_only --weight_only_precision int8 builds (no fp16 / int4 Linear kernels here)")
    if args.dtype != 'float16':
        raise ValueError("activations are fp16 on the hot path")
    if args.max_beam_width != 1:
        raise ValueError("the reference's run flow decodes greedily (beam width 1)")


def _module_tensors(module):
    """parameters + buffers of an eager module, on the CPU, under their attribute paths (the names weight.py assigns)"""
    return {k: v.detach().cpu() for k, v in module.state_dict().items()}


def _config(name, dims_dict, args, **extra):
    # the keys Builder.save_config writes (T/tensorrt_llm/builder.py) that run-time code reads back
    cfg = {"builder_config": {"name": name, "precision": args.dtype, "tensor_parallel": 1,
                              "max_batch_size": args.max_batch_size, "max_input_len": args.max_input_len,
                              "max_output_len": args.max_output_len, "max_beam_width": args.max_beam_width,
                              "int8": bool(args.quant_mode.has_act_and_weight_quant() or args.quant_mode.has_int8_kv_cache()),
                              "quant_mode": int(args.quant_mode)},
           "plugin_config": {"gpt_attention_plugin": args.use_gpt_attention_plugin,
                             "gemm_plugin": args.use_gemm_plugin, "layernorm_plugin": args.use_layernorm_plugin,
                             "weight_only_quant_matmul_plugin": args.dtype if args.use_weight_only else False},
           "dims": dims_dict}
    cfg["builder_config"].update(extra)
    return cfg


def _save(output_dir, engine_name, config_name, tensors, config):
    import torch
    os.makedirs(output_dir, exist_ok=True)
    torch.save(tensors, os.path.join(output_dir, engine_name))
    with open(os.path.join(output_dir, config_name), "w") as f:
        json.dump(config, f, indent=1)


def build_encoder(model, args, device="cuda"):
    """build.py:145-196: WhisperEncoder -> `whisper_encoder_float16_tp1_rank0.engine` + encoder_config.json"""
    from b200_whisper.models import WhisperEncoder
    d, sd = model['dims'], model['model_state_dict']
    enc = WhisperEncoder(d['n_mels'], d['n_audio_ctx'], d['n_audio_state'], d['n_audio_head'], d['n_audio_layer']).to(device)
    enc.load_from_state_dict(sd)
    _save(args.output_dir, get_engine_name(MODEL_ENCODER_NAME, 'float16', 1, 0), 'encoder_config.json', _module_tensors(enc),
          _config(MODEL_ENCODER_NAME, d, args, num_layers=d['n_audio_layer'], num_heads=d['n_audio_head'],
                  hidden_size=d['n_audio_state']))
    return enc


def build_decoder(model, args, device="cuda"):
    """build.py:198-305: WhisperDecoder (+ the self-attention KV scales read from quantize_dir, weight.py:236-243)"""
    from b200_whisper.models import WhisperDecoder
    from b200_whisper.runtime.checkpoint import read_kv_scales
    d, sd = model['dims'], model['model_state_dict']
    dec = WhisperDecoder(d['n_vocab'], d['n_text_ctx'], d['n_text_state'], d['n_text_head'], d['n_text_layer'],
                         quant_mode=args.quant_mode).to(device)
    kv = ckv = None
    if args.quant_mode.has_int8_kv_cache():
        kv = read_kv_scales(args.quantize_dir, d['n_text_layer'])
        ckv = read_kv_scales(args.quantize_dir, d['n_text_layer'], cross=True)
    dec.load_from_state_dict(sd, kv_scales=kv, cross_kv_scales=ckv)
    _save(args.output_dir, get_engine_name(MODEL_DECODER_NAME, args.dtype, 1, 0), 'decoder_config.json', _module_tensors(dec),
          _config(MODEL_DECODER_NAME, d, args, num_layers=d['n_text_layer'], num_heads=d['n_text_head'],
                  hidden_size=d['n_text_state'], vocab_size=d['n_vocab'], max_position_embeddings=d['n_text_ctx']))
    return dec


def build_crossattn_kv_linear(model, args, device="cuda"):
    """build.py:307-365: CrossAttn_KV, the `cross_kv_cache_warping` model"""
    from b200_whisper.models import CrossAttn_KV
    from b200_whisper.runtime.checkpoint import read_kv_scales
    d, sd = model['dims'], model['model_state_dict']
    ckv_model = CrossAttn_KV(d['n_text_state'], d['n_text_head'], d['n_text_layer'], quant_mode=args.quant_mode).to(device)
    ckv = read_kv_scales(args.quantize_dir, d['n_text_layer'], cross=True) if args.quant_mode.has_int8_kv_cache() else None
    ckv_model.load_from_state_dict(sd, cross_kv_scales=ckv)
    _save(args.output_dir, get_engine_name(MODEL_CROSSATTN_NAME, args.dtype, 1, 0), 'crossattn_config.json',
          _module_tensors(ckv_model),
          _config(MODEL_CROSSATTN_NAME, d, args, num_layers=d['n_text_layer'], num_heads=d['n_text_head'],
                  hidden_size=d['n_text_state']))
    return ckv_model


def run_build(args=None):
    import torch
    args = parse_arguments(args)
    check_supported(args)
    if not torch.cuda.is_available():
        raise RuntimeError("build.py quantizes the weights with the library's GPU quantizer: a B200 is required "
                           "(there is no CPU path)")
    tik = time.time()
    model = torch.load(args.model_dir, map_location="cpu", weights_only=False)
    dims = model['dims']
    model = {'dims': dims if isinstance(dims, dict) else dict(vars(dims)), 'model_state_dict': model['model_state_dict']}
    build_encoder(model, args)
    build_decoder(model, args)
    build_crossattn_kv_linear(model, args)
    print(f"Total time of building all engines: {time.strftime('%H:%M:%S', time.gmtime(time.time() - tik))}")
    return args


if __name__ == '__main__':
    run_build()
