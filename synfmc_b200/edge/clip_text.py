"""transformers `CLIPTextModel` (the SD1.5 text encoder) on the B200 kernels.

Reference call sites: `self.text_encoder(text_input_ids.to(device), attention_mask=None)[0]`
(fmc/pipelines/pipeline_animation.py:506-510, :546-550) and `text_encoder(prompt_ids.to(latents.device))[0]`
(train_cam_ctrl.py:557-561).  Parameter holders carry the transformers state-dict keys (`text_model.embeddings.
{token_embedding,position_embedding}.weight`, `text_model.encoder.layers.i.{self_attn.{q,k,v,out}_proj, layer_norm1,
mlp.{fc1,fc2}, layer_norm2}`, `text_model.final_layer_norm`), so the real checkpoint loads by key.

Execution on rows [(batch tokens), width]: token + position gather = fmc_embed_tokens; per layer LayerNorm = fmc_layernorm_bf16,
q|k|v as ONE GEMM (the d^-1/2 query scale folded into the q rows of the weight), causal 12-head attention over 77 tokens =
fmc_small_mha, out-projection GEMM with the residual in its epilogue, fc1 GEMM -> fmc_quick_gelu -> fc2 GEMM (+ residual);
final LayerNorm.  The tokenizer (string processing on the host) stays transformers' own."""
import contextlib

import torch
from torch import nn

from .. import engine, ops
from ..fmc._blocks import _Holder

BF16, F32 = torch.bfloat16, torch.float32


def _device_ctx(device):
    """make the tensor's GPU current for the calls below (the kernels launch on the current device's current stream)"""
    return torch.cuda.device(device) if torch.device(device).type == "cuda" else contextlib.nullcontext()


class _Attn(_Holder):
    def __init__(self, width):
        super().__init__()
        self.q_proj = nn.Linear(width, width)
        self.k_proj = nn.Linear(width, width)
        self.v_proj = nn.Linear(width, width)
        self.out_proj = nn.Linear(width, width)


class _Mlp(_Holder):
    def __init__(self, width, hidden):
        super().__init__()
        self.fc1 = nn.Linear(width, hidden)
        self.fc2 = nn.Linear(hidden, width)


class _Layer(_Holder):
    def __init__(self, width, hidden, eps):
        super().__init__()
        self.self_attn = _Attn(width)
        self.layer_norm1 = nn.LayerNorm(width, eps=eps)
        self.mlp = _Mlp(width, hidden)
        self.layer_norm2 = nn.LayerNorm(width, eps=eps)


class _Embeddings(_Holder):
    def __init__(self, vocab, positions, width):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab, width)
        self.position_embedding = nn.Embedding(positions, width)


class _Encoder(_Holder):
    def __init__(self, layers, width, hidden, eps):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(width, hidden, eps) for _ in range(layers)])


class _TextModel(_Holder):
    def __init__(self, vocab, positions, width, layers, hidden, eps):
        super().__init__()
        self.embeddings = _Embeddings(vocab, positions, width)
        self.encoder = _Encoder(layers, width, hidden, eps)
        self.final_layer_norm = nn.LayerNorm(width, eps=eps)


class _Config:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _Output(tuple):
    """transformers BaseModelOutputWithPooling as the reference uses it: `[0]` / `.last_hidden_state`"""

    @property
    def last_hidden_state(self):
        return self[0]


class CLIPTextModel(nn.Module):
    def __init__(self, vocab_size=49408, max_position_embeddings=77, hidden_size=768, num_hidden_layers=12,
                 num_attention_heads=12, intermediate_size=3072, layer_norm_eps=1e-5, hidden_act="quick_gelu", **unused):
        super().__init__()
        assert hidden_act == "quick_gelu", "the SD1.5 text encoder uses quick_gelu"
        assert hidden_size % num_attention_heads == 0
        self.text_model = _TextModel(vocab_size, max_position_embeddings, hidden_size, num_hidden_layers, intermediate_size,
                                     layer_norm_eps)
        # no `use_attention_mask` attribute: the reference then passes attention_mask=None (pipeline_animation.py:501-504)
        self.config = _Config(vocab_size=vocab_size, max_position_embeddings=max_position_embeddings, hidden_size=hidden_size,
                              num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
                              intermediate_size=intermediate_size, layer_norm_eps=layer_norm_eps, hidden_act=hidden_act)
        self.requires_grad_(False)  # frozen in both trainers (train_cam_ctrl.py:247)
        self._plans = None

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def _plan(self, device):
        if engine.precise():
            raise NotImplementedError("the text encoder runs in the bf16 mode (its output feeds the cross-attention K / V "
                                      "projections; the reference-precision kernels cover the denoising path)")
        key = engine.plan_key(device)
        if self._plans is None or self._plans["key"] != key or self._plans["fp"] != engine.fingerprint(self):
            cfg, tm = self.config, self.text_model
            d = cfg.hidden_size // cfg.num_attention_heads
            layers = []
            for layer in tm.encoder.layers:
                a = layer.self_attn
                s = float(d) ** -0.5  # transformers scales the query: folded into the q rows
                w = torch.cat([a.q_proj.weight * s, a.k_proj.weight, a.v_proj.weight], dim=0).detach().float()
                b = torch.cat([a.q_proj.bias * s, a.k_proj.bias, a.v_proj.bias], dim=0).detach().float()
                layers.append({
                    "ln1": engine.NormPlan(layer.layer_norm1, device), "qkv": engine.LinearPlan(w, b, device),
                    "out": engine.LinearPlan(a.out_proj.weight.detach().float(), a.out_proj.bias.detach().float(), device),
                    "ln2": engine.NormPlan(layer.layer_norm2, device),
                    "fc1": engine.LinearPlan(layer.mlp.fc1.weight.detach().float(), layer.mlp.fc1.bias.detach().float(), device),
                    "fc2": engine.LinearPlan(layer.mlp.fc2.weight.detach().float(), layer.mlp.fc2.bias.detach().float(), device)})
            self._plans = {"key": key, "fp": engine.fingerprint(self), "layers": layers,
                           "tok": engine._dev_f32(tm.embeddings.token_embedding.weight, device),
                           "pos": engine._dev_f32(tm.embeddings.position_embedding.weight, device),
                           "final": engine.NormPlan(tm.final_layer_norm, device)}
        return self._plans

    @torch.no_grad()
    def forward(self, input_ids, attention_mask=None, **unused):
        """input_ids int64 [B, T <= 77] -> (last_hidden_state [B, T, width] fp32,)"""
        if attention_mask is not None:
            raise NotImplementedError("attention_mask: SD1.5's text encoder config has no use_attention_mask, the reference "
                                      "passes None")
        ops.require_cuda(input_ids)
        cfg = self.config
        B, T = input_ids.shape
        if T > cfg.max_position_embeddings or T > 128:
            raise ValueError(f"{T} tokens: the position table holds {cfg.max_position_embeddings}")
        heads, C = cfg.num_attention_heads, cfg.hidden_size
        with _device_ctx(input_ids.device):
            p = self._plan(input_ids.device)
            x = ops.embed_tokens(input_ids.to(torch.int64).contiguous(), p["tok"], p["pos"])
            for lp in p["layers"]:
                n = ops.layernorm(x, lp["ln1"].g, lp["ln1"].b, lp["ln1"].eps)
                ctx = ops.small_mha(lp["qkv"](n), 0, C, 2 * C, B, T, heads, C // heads, 1.0, causal=True)
                x = lp["out"](ctx, residual=x)
                n = ops.layernorm(x, lp["ln2"].g, lp["ln2"].b, lp["ln2"].eps)
                h = lp["fc1"](n)
                x = lp["fc2"](ops.quick_gelu(h, out=h), residual=x)
            y = ops.layernorm(x, p["final"].g, p["final"].b, p["final"].eps)
            return _Output((y.float().view(B, T, C),))
