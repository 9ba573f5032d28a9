"""Secondary oracle: the reference AS DEPLOYED.  TEST INFRASTRUCTURE ONLY.

``reference_embeddings`` runs what ``Wav2VecBertEncoder.forward`` of the reference runs (audiotoken/encoder.py:163-184):
the processor, HF ``Wav2Vec2BertModel`` with the reference's relative-key SDPA attention (a restatement of the patch in
audiotoken/modeling_wav2vec2_bert.py:20-80, applied the way encoder.py:14-15 applies it), the affine-free LayerNorm —
optionally under ``torch.amp.autocast('cuda', bfloat16)`` with TF32 allowed (audiotoken/__init__.py:6-9), on whatever
device it is given.  On the GPU box this is "the reference's own PyTorch path on the same inputs and weights" that
BASELINE.json's north star names; /root/reference itself cannot travel there.

Pinning: ``tests/test_oracle_golden.py::test_hf_reference_matches_long_golden`` runs this module on the CPU (fp32) and
compares it with ``tests/golden/conformer_long_l2.npz``, which ``tests/golden/make_golden.py long`` produced with the
REAL reference files (processors.py + modeling_wav2vec2_bert.py loaded from /root/reference).
"""
from __future__ import annotations

import contextlib
import math
from typing import Dict, List, Tuple

import torch

from . import fbank


def relkey_sdpa_forward(self, hidden_states, attention_mask=None, relative_position_embeddings=None,
                        output_attentions=False, **_unused):
    """Relative-key attention through torch SDPA (reference modeling_wav2vec2_bert.py:20-80): the [T, T] table of
    clamped distances indexes the distance embedding, ``q . E`` scaled by 1/sqrt(d) is ADDED to the additive key
    mask and the sum is handed to scaled_dot_product_attention as `attn_mask`; no attention weights are returned."""
    assert self.position_embeddings_type == 'relative_key'
    B, T, _ = hidden_states.shape
    H, D = self.num_heads, self.head_size
    q = self.linear_q(hidden_states).view(B, T, H, D).transpose(1, 2)
    k = self.linear_k(hidden_states).view(B, T, H, D).transpose(1, 2)
    v = self.linear_v(hidden_states).view(B, T, H, D).transpose(1, 2)
    pos = torch.arange(T, device=hidden_states.device)
    dist = (pos[None, :] - pos[:, None]).clamp(-self.left_max_position_embeddings, self.right_max_position_embeddings)
    emb = self.distance_embedding(dist + self.left_max_position_embeddings).to(q.dtype)
    bias = torch.einsum('bhld,lrd->bhlr', q, emb) / math.sqrt(D)
    if attention_mask is not None:
        bias = bias + attention_mask
    out = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=bias, scale=1.0 / math.sqrt(D))
    out = out.transpose(1, 2).reshape(B, T, H * D)
    return self.linear_out(out), None


@contextlib.contextmanager
def patched_attention():
    from transformers.models.wav2vec2_bert.modeling_wav2vec2_bert import Wav2Vec2BertSelfAttention as A
    old = A.forward
    A.forward = relkey_sdpa_forward
    try:
        yield
    finally:
        A.forward = old


def build_model(sd: Dict[str, torch.Tensor], n_layers: int, device):
    from transformers import Wav2Vec2BertConfig, Wav2Vec2BertModel
    model = Wav2Vec2BertModel(Wav2Vec2BertConfig(num_hidden_layers=n_layers))
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and set(missing) <= {'masked_spec_embed'}, (missing, unexpected)
    return model.to(device).eval()


def reference_embeddings(wave: torch.Tensor, mask: torch.Tensor, sd: Dict[str, torch.Tensor], n_layers: int,
                         device='cpu', autocast: bool = False, model=None, batch: int = 16, tf32: bool = True
                         ) -> Tuple[torch.Tensor, torch.Tensor, List[torch.Tensor]]:
    """wave, mask [B, L] -> (LayerNormed embeddings of hidden state n_layers, fp32 [B, T, 1024]; attention_mask
    [B, T]; the raw hidden state).  autocast=True = the reference on a GPU (encoder.py:164)."""
    device = torch.device(device)
    model = model if model is not None else build_model(sd, n_layers, device)
    if device.type == 'cuda':
        torch.backends.cuda.matmul.allow_tf32 = tf32          # True as deployed (audiotoken/__init__.py:6-9);
        torch.backends.cudnn.allow_tf32 = tf32                # False for the fp32 "truth" run
    embs, masks, hids = [], [], []
    ctx = torch.amp.autocast(device_type='cuda', dtype=torch.bfloat16) if autocast else contextlib.nullcontext()
    with patched_attention(), torch.no_grad(), ctx:
        for a in range(0, wave.shape[0], batch):
            feats, am = fbank.features(wave[a:a + batch].to(device), mask[a:a + batch].to(device))
            hs = model(feats, attention_mask=am, output_hidden_states=True).hidden_states
            h = hs[n_layers]
            embs.append(torch.nn.functional.layer_norm(h, (1024,)).float().cpu())   # encoder.py:175-176
            masks.append(am.cpu())
            hids.append(h.float().cpu())
    return torch.cat(embs), torch.cat(masks), torch.cat(hids)


def vq_eval_tokens(emb: torch.Tensor, codebook: torch.Tensor, device='cpu', tf32: bool = False) -> torch.Tensor:
    """``VectorQuantize`` eval forward as deployed (encoder.py:180): the package computes, with autocast disabled,
    ``dist = -cdist(x, embed)`` from ``x^2 - 2 x e^T + e^2`` in fp32 and takes the argmax — on a GPU with TF32 allowed
    the ``x e^T`` product runs on TF32 tensor cores.  Returns int64 [rows]."""
    device = torch.device(device)
    x = emb.reshape(-1, emb.shape[-1]).to(device, torch.float32)
    e = codebook.to(device, torch.float32)
    if device.type == 'cuda':
        torch.backends.cuda.matmul.allow_tf32 = tf32
    out = []
    with torch.no_grad(), torch.amp.autocast(device_type=device.type, enabled=False):
        e2 = (e * e).sum(-1)
        for a in range(0, x.shape[0], 16384):
            xa = x[a:a + 16384]
            d2 = (xa * xa).sum(-1, keepdim=True) - 2.0 * (xa @ e.t()) + e2[None, :]
            out.append((-d2.clamp(min=0).sqrt()).argmax(dim=-1).cpu())
    return torch.cat(out)


# ---- acoustic: HF EncodecModel standing in for encodec.EncodecModel.encodec_model_24khz() ---------------
def acoustic_reference(wave: torch.Tensor, sd: Dict[str, torch.Tensor], n_q: int, device='cpu', autocast: bool = False,
                       batch: int = 8, tf32: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """wave [B, L] @24 kHz -> (embeddings fp32 [B, 128, T], codes int64 [B, n_q, T]) the way the reference's
    AcousticEncoder.forward calls the model (encoder.py:44-55): ``model.encoder(x.unsqueeze(1))`` then
    ``model.quantizer.encode(emb, bandwidth)``, optionally under CUDA bf16 autocast."""
    from transformers import EncodecConfig, EncodecModel
    device = torch.device(device)
    model = EncodecModel(EncodecConfig())
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not [k for k in missing if k.startswith('encoder') or k.endswith('codebook.embed')]
    model = model.to(device).eval()
    if device.type == 'cuda':
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
    bandwidth = n_q * 10 * 75 / 1000.0                         # n_q = floor(bw * 1000 / (10 * 75)), configs.py:33-39
    ctx = torch.amp.autocast(device_type='cuda', dtype=torch.bfloat16) if autocast else contextlib.nullcontext()
    embs, codes = [], []
    with torch.no_grad(), ctx:
        for a in range(0, wave.shape[0], batch):
            emb = model.encoder(wave[a:a + batch].to(device).unsqueeze(1))
            c = model.quantizer.encode(emb, bandwidth)          # [n_q, B, T]
            embs.append(emb.float().cpu())
            codes.append(c.transpose(0, 1).cpu())
    return torch.cat(embs), torch.cat(codes)


# ---- the reference's own semantic_s: HF HubertModel called as encoder.py:93 calls it -----------------------------------
def hubert_reference(wave: torch.Tensor, mask: torch.Tensor, sd: Dict[str, torch.Tensor], device='cpu',
                     autocast: bool = False, tf32: bool = True) -> List[torch.Tensor]:
    """wave [B, L] (processor output, zero right-padded), mask [B, L] -> hidden_states (fp32, CPU) of
    ``HubertModel.forward(input_batch, attention_mask=attention_mask, output_hidden_states=True)``, optionally under
    CUDA bf16 autocast (encoder.py:89)."""
    from transformers import HubertConfig, HubertModel
    device = torch.device(device)
    model = HubertModel(HubertConfig())
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and set(missing) <= {'masked_spec_embed'}, (missing, unexpected)
    model = model.to(device).eval()
    if device.type == 'cuda':
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
    ctx = torch.amp.autocast(device_type='cuda', dtype=torch.bfloat16) if autocast else contextlib.nullcontext()
    with torch.no_grad(), ctx:
        hs = model.forward(wave.to(device), attention_mask=mask.to(device), output_hidden_states=True).hidden_states
    return [h.float().cpu() for h in hs]
