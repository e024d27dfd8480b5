"""Restatement of diffusers==0.31.0 `Attention` + `AttnProcessor2_0` for the one configuration the
reference VAE constructs (hyvideo/vae/unet_causal_3d_blocks.py:580-592): self-attention, one head,
GroupNorm(32) on (B,C,L), biased q/k/v/out Linear, residual connection, rescale_output_factor."""
import torch
import torch.nn.functional as F
from torch import nn


class SpatialNorm(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("not on the VAE path")


class AttnProcessor:
    pass


class AttnAddedKVProcessor:
    pass


AttentionProcessor = AttnProcessor
ADDED_KV_ATTENTION_PROCESSORS = (AttnAddedKVProcessor,)
CROSS_ATTENTION_PROCESSORS = (AttnProcessor,)


class Attention(nn.Module):
    def __init__(self, query_dim, heads=8, dim_head=64, rescale_output_factor=1.0, eps=1e-5,
                 norm_num_groups=None, spatial_norm_dim=None, residual_connection=False, bias=False,
                 upcast_softmax=False, _from_deprecated_attn_block=False, **unused):
        super().__init__()
        assert spatial_norm_dim is None
        self.inner_dim = dim_head * heads
        self.heads = heads
        self.rescale_output_factor = rescale_output_factor
        self.residual_connection = residual_connection
        self.scale = dim_head ** -0.5
        self.group_norm = (nn.GroupNorm(num_channels=query_dim, num_groups=norm_num_groups, eps=eps, affine=True)
                           if norm_num_groups is not None else None)
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=True), nn.Dropout(0.0)])

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, **kw):
        residual = hidden_states
        b, l, _ = hidden_states.shape
        if attention_mask is not None:
            if attention_mask.shape[0] < b * self.heads:
                attention_mask = attention_mask.repeat_interleave(self.heads, dim=0)
            attention_mask = attention_mask.view(b, self.heads, -1, attention_mask.shape[-1])
        if self.group_norm is not None:
            hidden_states = self.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
        q = self.to_q(hidden_states)
        k = self.to_k(hidden_states)
        v = self.to_v(hidden_states)
        hd = self.inner_dim // self.heads
        q = q.view(b, -1, self.heads, hd).transpose(1, 2)
        k = k.view(b, -1, self.heads, hd).transpose(1, 2)
        v = v.view(b, -1, self.heads, hd).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, -1, self.heads * hd).to(q.dtype)
        o = self.to_out[0](o)
        o = self.to_out[1](o)
        if self.residual_connection:
            o = o + residual
        return o / self.rescale_output_factor
