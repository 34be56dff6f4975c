"""Parameter layout of the SAiD inference path.

The drop-in contract is that ``SAID_UNet1D`` is an ``nn.Module`` whose
``state_dict`` carries exactly the reference's tensor names and shapes, so that
``load_state_dict(torch.load("SAiD.pth"))`` works unchanged (reference caller:
``script/inference.py:152-159``).  The arithmetic itself lives in CUDA kernels, so the
modules here are *parameter containers only*: they have no ``forward``.

Names and shapes follow
  - denoiser:      ``said/model/ldm/openaimodel.py:463-479, 154-194, 665-669`` and
                   ``said/model/ldm/attention.py:69-85, 143-158, 210-221`` as instantiated by
                   ``said/model/unet_1d_condition.py:36-49`` (model_channels=192, channel_mult=(1,),
                   one ResBlock + one SpatialTransformer per level, 6 heads x 32);
  - audio encoder: Hugging Face ``Wav2Vec2Model`` with the default ``Wav2Vec2Config()`` (base: 7 conv
                   layers, 12 post-LN transformer layers, hidden 768), checkpoint names as written by
                   transformers 4.30.2 (``pos_conv_embed.conv.weight_g / weight_v``); the
                   ``parametrizations.weight.original0/1`` names that transformers >= 4.35 writes are
                   accepted on load and mapped onto them.
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Tuple

import torch
from torch import nn

Spec = List[Tuple[str, Tuple[int, ...]]]

MODEL_CH = 192          # UNetModel.model_channels          (unet_1d_condition.py:40)
TIME_CH = 4 * MODEL_CH  # time_embed_dim                    (openaimodel.py:462)
HEADS = 6               # 192 // num_head_channels(32)      (openaimodel.py:498-510)
HEAD_DIM = 32
FF_INNER = 4 * MODEL_CH  # FeedForward mult=4               (attention.py:36-38)
GN_GROUPS = 32

# ResBlocks / SpatialTransformers in execution order (openaimodel.py:697-704)
RESBLOCK_PATHS = (
    ("input_blocks.1.0", MODEL_CH),
    ("middle_block.0", MODEL_CH),
    ("middle_block.2", MODEL_CH),
    ("output_blocks.0.0", 2 * MODEL_CH),
    ("output_blocks.1.0", 2 * MODEL_CH),
)
TRANSFORMER_PATHS = (
    "input_blocks.1.1",
    "middle_block.1",
    "output_blocks.0.1",
    "output_blocks.1.1",
)


def _resblock_spec(path: str, cin: int) -> Spec:
    c = MODEL_CH
    spec: Spec = [
        (f"{path}.in_layers.0.weight", (cin,)),
        (f"{path}.in_layers.0.bias", (cin,)),
        (f"{path}.in_layers.2.weight", (c, cin, 3)),
        (f"{path}.in_layers.2.bias", (c,)),
        (f"{path}.emb_layers.1.weight", (c, TIME_CH)),
        (f"{path}.emb_layers.1.bias", (c,)),
        (f"{path}.out_layers.0.weight", (c,)),
        (f"{path}.out_layers.0.bias", (c,)),
        (f"{path}.out_layers.3.weight", (c, c, 3)),
        (f"{path}.out_layers.3.bias", (c,)),
    ]
    if cin != c:
        spec += [
            (f"{path}.skip_connection.weight", (c, cin, 1)),
            (f"{path}.skip_connection.bias", (c,)),
        ]
    return spec


def _transformer_spec(path: str, ctx_dim: int) -> Spec:
    c = MODEL_CH
    tb = f"{path}.transformer_blocks.0"
    return [
        (f"{path}.norm.weight", (c,)),
        (f"{path}.norm.bias", (c,)),
        (f"{tb}.attn1.to_q.weight", (c, c)),
        (f"{tb}.attn1.to_k.weight", (c, c)),
        (f"{tb}.attn1.to_v.weight", (c, c)),
        (f"{tb}.attn1.to_out.0.weight", (c, c)),
        (f"{tb}.attn1.to_out.0.bias", (c,)),
        (f"{tb}.ff.net.0.proj.weight", (2 * FF_INNER, c)),
        (f"{tb}.ff.net.0.proj.bias", (2 * FF_INNER,)),
        (f"{tb}.ff.net.2.weight", (c, FF_INNER)),
        (f"{tb}.ff.net.2.bias", (c,)),
        (f"{tb}.attn2.to_q.weight", (c, c)),
        (f"{tb}.attn2.to_k.weight", (c, ctx_dim)),
        (f"{tb}.attn2.to_v.weight", (c, ctx_dim)),
        (f"{tb}.attn2.to_out.0.weight", (c, c)),
        (f"{tb}.attn2.to_out.0.bias", (c,)),
        (f"{tb}.norm1.weight", (c,)),
        (f"{tb}.norm1.bias", (c,)),
        (f"{tb}.norm2.weight", (c,)),
        (f"{tb}.norm2.bias", (c,)),
        (f"{tb}.norm3.weight", (c,)),
        (f"{tb}.norm3.bias", (c,)),
        (f"{path}.proj_out.weight", (c, c, 1)),
        (f"{path}.proj_out.bias", (c,)),
    ]


def denoiser_spec(in_channels: int = 32, ctx_dim: int = 768) -> Spec:
    """160 tensors of ``denoiser.model.*`` in the reference's registration order."""
    c = MODEL_CH
    spec: Spec = [
        ("time_embed.0.weight", (TIME_CH, c)),
        ("time_embed.0.bias", (TIME_CH,)),
        ("time_embed.2.weight", (TIME_CH, TIME_CH)),
        ("time_embed.2.bias", (TIME_CH,)),
        ("input_blocks.0.0.weight", (c, in_channels, 3)),
        ("input_blocks.0.0.bias", (c,)),
    ]
    rb = dict(RESBLOCK_PATHS)
    for path in ("input_blocks.1", "middle_block", "output_blocks.0", "output_blocks.1"):
        if path == "middle_block":
            spec += _resblock_spec("middle_block.0", rb["middle_block.0"])
            spec += _transformer_spec("middle_block.1", ctx_dim)
            spec += _resblock_spec("middle_block.2", rb["middle_block.2"])
        else:
            spec += _resblock_spec(f"{path}.0", rb[f"{path}.0"])
            spec += _transformer_spec(f"{path}.1", ctx_dim)
    spec += [
        ("out.0.weight", (c,)),
        ("out.0.bias", (c,)),
        ("out.2.weight", (in_channels, c, 3)),
        ("out.2.bias", (in_channels,)),
    ]
    return spec


# Tensors the reference zero-initialises (openaimodel.py:182-185, 668; attention.py:221).  Synthetic
# weights must re-randomise them or every parity test is vacuous.
def zero_init_names(prefix: str = "denoiser.model.") -> List[str]:
    names = []
    for path, _ in RESBLOCK_PATHS:
        names += [f"{prefix}{path}.out_layers.3.weight", f"{prefix}{path}.out_layers.3.bias"]
    for path in TRANSFORMER_PATHS:
        names += [f"{prefix}{path}.proj_out.weight", f"{prefix}{path}.proj_out.bias"]
    names += [f"{prefix}out.2.weight", f"{prefix}out.2.bias"]
    return names


class Wav2Vec2Dims:
    """The subset of ``Wav2Vec2Config`` the inference path reads (defaults = wav2vec2-base)."""

    def __init__(self, config=None):
        g = (lambda k, d: getattr(config, k, d)) if config is not None else (lambda k, d: d)
        self.hidden = int(g("hidden_size", 768))
        self.layers = int(g("num_hidden_layers", 12))
        self.heads = int(g("num_attention_heads", 12))
        self.ffn = int(g("intermediate_size", 3072))
        self.conv_dim = tuple(g("conv_dim", (512,) * 7))
        self.conv_kernel = tuple(g("conv_kernel", (10, 3, 3, 3, 3, 2, 2)))
        self.conv_stride = tuple(g("conv_stride", (5, 2, 2, 2, 2, 2, 2)))
        self.conv_bias = bool(g("conv_bias", False))
        self.feat_extract_norm = str(g("feat_extract_norm", "group"))
        self.pos_kernel = int(g("num_conv_pos_embeddings", 128))
        self.pos_groups = int(g("num_conv_pos_embedding_groups", 16))
        self.stable_layer_norm = bool(g("do_stable_layer_norm", False))
        self.layer_norm_eps = float(g("layer_norm_eps", 1e-5))
        self.output_hidden = int(g("output_hidden_size", self.hidden))
        self.add_adapter = bool(g("add_adapter", False))

    def as_config(self):
        """A config-like namespace with the transformers ``Wav2Vec2Config`` attribute names of these dimensions."""
        import types

        return types.SimpleNamespace(
            hidden_size=self.hidden, num_hidden_layers=self.layers, num_attention_heads=self.heads, intermediate_size=self.ffn,
            conv_dim=self.conv_dim, conv_kernel=self.conv_kernel, conv_stride=self.conv_stride, conv_bias=self.conv_bias,
            feat_extract_norm=self.feat_extract_norm, num_conv_pos_embeddings=self.pos_kernel,
            num_conv_pos_embedding_groups=self.pos_groups, do_stable_layer_norm=self.stable_layer_norm,
            layer_norm_eps=self.layer_norm_eps, output_hidden_size=self.output_hidden, add_adapter=self.add_adapter)

    def check_supported(self) -> None:
        """Fail loudly on configurations the sm_100a kernels do not implement."""
        bad = []
        if self.feat_extract_norm not in ("group", "layer"):
            bad.append("feat_extract_norm not in ('group', 'layer')")
        if (self.feat_extract_norm == "layer") != self.conv_bias:
            bad.append("feat_extract_norm='layer' without conv_bias=True (or the reverse)")
        if self.add_adapter:
            bad.append("add_adapter=True")
        if abs(self.layer_norm_eps - 1e-5) > 1e-12:
            bad.append(f"layer_norm_eps={self.layer_norm_eps} (the encoder kernels use 1e-5)")
        if len(set(self.conv_dim)) != 1:
            bad.append("non-uniform conv_dim")
        if self.hidden % self.heads or self.hidden // self.heads != 64:
            bad.append("head_dim != 64")
        if bad:
            raise NotImplementedError(
                "said_b200 audio encoder supports the wav2vec2-base and wav2vec2-large families; unsupported: "
                + ", ".join(bad)
            )


def audio_encoder_spec(d: Wav2Vec2Dims) -> Spec:
    """211 tensors of ``audio_encoder.*`` (transformers ``Wav2Vec2Model``, 4.30.2 checkpoint names)."""
    h = d.hidden
    spec: Spec = [("masked_spec_embed", (h,))]
    cin = 1
    for i, (co, k) in enumerate(zip(d.conv_dim, d.conv_kernel)):
        spec.append((f"feature_extractor.conv_layers.{i}.conv.weight", (co, cin, k)))
        if d.conv_bias:
            spec.append((f"feature_extractor.conv_layers.{i}.conv.bias", (co,)))
        if i == 0 or d.feat_extract_norm == "layer":
            spec.append((f"feature_extractor.conv_layers.{i}.layer_norm.weight", (co,)))
            spec.append((f"feature_extractor.conv_layers.{i}.layer_norm.bias", (co,)))
        cin = co
    spec += [
        ("feature_projection.layer_norm.weight", (cin,)),
        ("feature_projection.layer_norm.bias", (cin,)),
        ("feature_projection.projection.weight", (h, cin)),
        ("feature_projection.projection.bias", (h,)),
        ("encoder.pos_conv_embed.conv.bias", (h,)),
        ("encoder.pos_conv_embed.conv.weight_g", (1, 1, d.pos_kernel)),
        ("encoder.pos_conv_embed.conv.weight_v", (h, h // d.pos_groups, d.pos_kernel)),
        ("encoder.layer_norm.weight", (h,)),
        ("encoder.layer_norm.bias", (h,)),
    ]
    for l in range(d.layers):
        p = f"encoder.layers.{l}"
        for proj in ("k_proj", "v_proj", "q_proj", "out_proj"):
            spec.append((f"{p}.attention.{proj}.weight", (h, h)))
            spec.append((f"{p}.attention.{proj}.bias", (h,)))
        spec += [
            (f"{p}.layer_norm.weight", (h,)),
            (f"{p}.layer_norm.bias", (h,)),
            (f"{p}.feed_forward.intermediate_dense.weight", (d.ffn, h)),
            (f"{p}.feed_forward.intermediate_dense.bias", (d.ffn,)),
            (f"{p}.feed_forward.output_dense.weight", (h, d.ffn)),
            (f"{p}.feed_forward.output_dense.bias", (h,)),
            (f"{p}.final_layer_norm.weight", (h,)),
            (f"{p}.final_layer_norm.bias", (h,)),
        ]
    return spec


WEIGHT_NORM_ALIASES = {
    # transformers >= 4.35 (torch parametrizations)  ->  4.30.2 checkpoint name
    "encoder.pos_conv_embed.conv.parametrizations.weight.original0": "encoder.pos_conv_embed.conv.weight_g",
    "encoder.pos_conv_embed.conv.parametrizations.weight.original1": "encoder.pos_conv_embed.conv.weight_v",
}


class ParamTree(nn.Module):
    """A module hierarchy that only holds parameters, addressed by dotted reference names."""

    def __init__(self, spec: Spec | None = None):
        super().__init__()
        for name, shape in spec or []:
            self.add(name, torch.zeros(shape))

    def add(self, dotted: str, value: torch.Tensor) -> None:
        node: nn.Module = self
        *parents, leaf = dotted.split(".")
        for seg in parents:
            child = node._modules.get(seg)
            if child is None:
                child = ParamTree()
                node.add_module(seg, child)
            node = child
        node.register_parameter(leaf, nn.Parameter(value, requires_grad=True))

    def tensors(self) -> Iterator[Tuple[str, torch.Tensor]]:
        for k, v in self.named_parameters():
            yield k, v

    def forward(self, *args, **kwargs):  # pragma: no cover - containers are never called
        raise RuntimeError(
            "said_b200 parameter containers have no PyTorch forward; the arithmetic runs in the "
            "sm_100a kernels behind SAID.inference()/SAID.forward()"
        )


def spec_numel(spec: Spec) -> int:
    n = 0
    for _, shape in spec:
        k = 1
        for s in shape:
            k *= s
        n += k
    return n


def spec_dict(spec: Spec) -> Dict[str, Tuple[int, ...]]:
    return {k: tuple(v) for k, v in spec}
