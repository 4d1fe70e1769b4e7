"""TEST INFRASTRUCTURE ONLY (never imported by the product): fp32 CPU restatement of the `transformers` CLIP text encoder as the
reference calls it -- `self.text_encoder(text_input_ids.to(device), attention_mask=None)[0]`
(fmc/pipelines/pipeline_animation.py:506-510, :546-550; train_cam_ctrl.py:557-561).  transformers is a pinned third-party
dependency of the reference (4.45.1); the installed transformers (same architecture) pins this restatement:
tests/test_oracle_edges.py loads one random state dict into both and requires last_hidden_state to agree to 1e-5.

SD1.5's text encoder (openai/clip-vit-large-patch14 text tower): 12 pre-LN layers, width 768, 12 heads of 64, MLP 3072 with
quick_gelu (x * sigmoid(1.702 x)), learned positions (77), causal mask, LayerNorm eps 1e-5, final LayerNorm; state-dict keys
`text_model.embeddings.{token,position}_embedding.weight`, `text_model.encoder.layers.i.{self_attn.{q,k,v,out}_proj,
layer_norm1, mlp.{fc1,fc2}, layer_norm2}`, `text_model.final_layer_norm`."""
import torch
from torch import nn


class _Attn(nn.Module):
    def __init__(self, width, heads):
        super().__init__()
        self.heads, self.d = heads, width // heads
        self.q_proj = nn.Linear(width, width)
        self.k_proj = nn.Linear(width, width)
        self.v_proj = nn.Linear(width, width)
        self.out_proj = nn.Linear(width, width)

    def forward(self, x):
        B, T, C = x.shape
        q = (self.q_proj(x) * self.d ** -0.5).view(B, T, self.heads, self.d).transpose(1, 2)
        k = self.k_proj(x).view(B, T, self.heads, self.d).transpose(1, 2)
        v = self.v_proj(x).view(B, T, self.heads, self.d).transpose(1, 2)
        s = q @ k.transpose(-1, -2)
        mask = torch.full((T, T), float("-inf")).triu(1)  # token t attends to tokens <= t
        p = torch.softmax(s + mask, dim=-1)
        return self.out_proj((p @ v).transpose(1, 2).reshape(B, T, C))


class _Mlp(nn.Module):
    def __init__(self, width, hidden):
        super().__init__()
        self.fc1 = nn.Linear(width, hidden)
        self.fc2 = nn.Linear(hidden, width)

    def forward(self, x):
        h = self.fc1(x)
        return self.fc2(h * torch.sigmoid(1.702 * h))


class _Layer(nn.Module):
    def __init__(self, width, heads, hidden, eps):
        super().__init__()
        self.self_attn = _Attn(width, heads)
        self.layer_norm1 = nn.LayerNorm(width, eps=eps)
        self.mlp = _Mlp(width, hidden)
        self.layer_norm2 = nn.LayerNorm(width, eps=eps)

    def forward(self, x):
        x = x + self.self_attn(self.layer_norm1(x))
        return x + self.mlp(self.layer_norm2(x))


class _Embeddings(nn.Module):
    def __init__(self, vocab, positions, width):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab, width)
        self.position_embedding = nn.Embedding(positions, width)


class _Encoder(nn.Module):
    def __init__(self, layers, width, heads, hidden, eps):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(width, heads, hidden, eps) for _ in range(layers)])


class _TextModel(nn.Module):
    def __init__(self, vocab, positions, width, layers, heads, hidden, eps):
        super().__init__()
        self.embeddings = _Embeddings(vocab, positions, width)
        self.encoder = _Encoder(layers, width, heads, hidden, eps)
        self.final_layer_norm = nn.LayerNorm(width, eps=eps)


class CLIPTextModel(nn.Module):
    def __init__(self, vocab_size=49408, max_position_embeddings=77, hidden_size=768, num_hidden_layers=12,
                 num_attention_heads=12, intermediate_size=3072, layer_norm_eps=1e-5):
        super().__init__()
        self.text_model = _TextModel(vocab_size, max_position_embeddings, hidden_size, num_hidden_layers, num_attention_heads,
                                     intermediate_size, layer_norm_eps)

    def forward(self, input_ids, attention_mask=None):
        assert attention_mask is None, "SD1.5's text encoder config has no use_attention_mask: the reference passes None"
        tm = self.text_model
        T = input_ids.shape[1]
        x = tm.embeddings.token_embedding(input_ids) + tm.embeddings.position_embedding.weight[:T]
        for layer in tm.encoder.layers:
            x = layer(x)
        return (tm.final_layer_norm(x),)
