"""Packs the reference's MossFormer2 state dict (1099 keys, look2hear/models/base_model.py:118-146) into the
layouts the sm_100a kernels read (include/tdz.h: tdz_mossformer2_weights).

Folding rules (all done in fp32 on the host, once):
  * ScaleNorm gain g (a scalar) and LayerNorm / GroupNorm per-channel affine commute with the Linear that
    follows: W (gamma * n + beta) + b = (W diag(gamma)) n + (W beta + b).
  * bf16 tensor-core operands are stored as bf16; tf32 operands are stored as fp32 rounded to nearest tf32
    (the tensor core itself would truncate, which biases the result).
"""
import ctypes

import torch

from . import _lib

LAYER = "mask_net.mdl.intra_mdl.mossformerM.layers.{}."
FSMN = "mask_net.mdl.intra_mdl.mossformerM.fsmn.{}."


def round_tf32(x):
    """fp32 -> nearest tf32 (10 explicit mantissa bits), ties to even, kept in fp32 storage."""
    xi = x.contiguous().view(torch.int32)
    r = ((xi >> 13) & 1) + 0xFFF
    return ((xi + r) & ~0x1FFF).view(torch.float32)


class PackedMossFormer2:
    """Device tensors + the ctypes pointer table.  Keeps every tensor alive for as long as the table is used."""

    def __init__(self, state_dict, device):
        self.device = torch.device(device)
        self._keep = []
        sd = {k: v.detach().to(torch.float32).cpu() for k, v in state_dict.items()}
        self.table = _lib.MossFormer2Weights()
        t = self.table
        f32, bf16, tf32 = self._f32, self._bf16, self._tf32

        t.enc_w = f32(sd["enc.conv1d.weight"][:, 0, :])
        gn_g, gn_b = sd["mask_net.norm.weight"], sd["mask_net.norm.bias"]
        w_e = sd["mask_net.conv1d_encoder.weight"][:, :, 0]
        w_e_fold = round_tf32(w_e * gn_g[None, :])
        t.w_enc1x1 = f32(w_e_fold)
        t.enc1x1_colsum = f32(w_e_fold.double().sum(dim=1).float())
        t.enc1x1_bias = f32((w_e.double() @ gn_b.double()).float())
        t.pos_inv_freq = f32(sd["mask_net.pos_enc.inv_freq"])
        t.pos_scale = f32(sd["mask_net.pos_enc.scale"])
        t.rot_freqs = f32(sd[LAYER.format(0) + "rotary_pos_emb.freqs"])

        for i in range(_lib.NUM_LAYERS):
            L = t.layers[i]
            p = LAYER.format(i)
            gh, gq, go = (sd[p + f"{n}.mdl.0.g"] for n in ("to_hidden", "to_qk", "to_out"))
            L.w_in = bf16(torch.cat((sd[p + "to_hidden.mdl.1.weight"] * gh, sd[p + "to_qk.mdl.1.weight"] * gq), 0))
            L.b_in = f32(torch.cat((sd[p + "to_hidden.mdl.1.bias"], sd[p + "to_qk.mdl.1.bias"]), 0))
            L.dw_in = f32(torch.cat((sd[p + "to_hidden.mdl.3.sequential.1.conv.weight"][:, 0, :],
                                     sd[p + "to_qk.mdl.3.sequential.1.conv.weight"][:, 0, :]), 0))
            L.os_gamma = f32(sd[p + "qk_offset_scale.gamma"])
            L.os_beta = f32(sd[p + "qk_offset_scale.beta"])
            L.w_out = bf16(sd[p + "to_out.mdl.1.weight"] * go)
            L.b_out = f32(sd[p + "to_out.mdl.1.bias"])
            L.dw_out = f32(sd[p + "to_out.mdl.3.sequential.1.conv.weight"][:, 0, :])
            q = FSMN.format(i)
            L.w_c1 = tf32(sd[q + "conv1.0.weight"][:, :, 0])
            L.b_c1 = f32(sd[q + "conv1.0.bias"])
            L.prelu_c1 = f32(sd[q + "conv1.1.weight"])
            L.ln1_g = f32(sd[q + "norm1.weight"])
            L.ln1_b = f32(sd[q + "norm1.bias"])
            ws, bs, dws = [], [], []
            for n in ("to_u", "to_v"):
                g_, b_ = sd[q + f"gated_fsmn.{n}.mdl.0.weight"], sd[q + f"gated_fsmn.{n}.mdl.0.bias"]
                W, b = sd[q + f"gated_fsmn.{n}.mdl.1.weight"], sd[q + f"gated_fsmn.{n}.mdl.1.bias"]
                ws.append(W * g_[None, :])
                bs.append((W.double() @ b_.double()).float() + b)
                dws.append(sd[q + f"gated_fsmn.{n}.mdl.3.sequential.1.conv.weight"][:, 0, :])
            L.w_uv = bf16(torch.cat(ws, 0))
            L.b_uv = f32(torch.cat(bs, 0))
            L.dw_uv = f32(torch.cat(dws, 0))
            L.w_lin = bf16(sd[q + "gated_fsmn.fsmn.linear.weight"])
            L.b_lin = f32(sd[q + "gated_fsmn.fsmn.linear.bias"])
            L.w_proj = bf16(sd[q + "gated_fsmn.fsmn.project.weight"])
            c = q + "gated_fsmn.fsmn.conv."
            L.dd_w1 = f32(sd[c + "conv1.weight"][:, 0, :, 0])
            L.in1_g = f32(sd[c + "norm1.weight"])
            L.in1_b = f32(sd[c + "norm1.bias"])
            L.dd_prelu1 = f32(sd[c + "prelu1.weight"])
            L.dd_w2 = f32(sd[c + "conv2.weight"][:, :, :, 0])
            L.in2_g = f32(sd[c + "norm2.weight"])
            L.in2_b = f32(sd[c + "norm2.bias"])
            L.dd_prelu2 = f32(sd[c + "prelu2.weight"])
            g2, b2 = sd[q + "norm2.weight"], sd[q + "norm2.bias"]
            W2 = sd[q + "conv2.weight"][:, :, 0]
            L.w_c2 = tf32(W2 * g2[None, :])
            L.b_c2 = f32((W2.double() @ b2.double()).float() + sd[q + "conv2.bias"])

        t.fln_g = f32(sd["mask_net.mdl.intra_mdl.norm.weight"])
        t.fln_b = f32(sd["mask_net.mdl.intra_mdl.norm.bias"])
        t.fgn_g = f32(sd["mask_net.mdl.intra_norm.weight"])
        t.fgn_b = f32(sd["mask_net.mdl.intra_norm.bias"])
        t.mask_prelu = f32(sd["mask_net.prelu.weight"])
        t.w_out1 = tf32(sd["mask_net.conv1d_out.weight"][:, :, 0])
        t.b_out1 = f32(sd["mask_net.conv1d_out.bias"])
        t.w_tg = tf32(torch.cat((sd["mask_net.output.0.weight"][:, :, 0],
                                 sd["mask_net.output_gate.0.weight"][:, :, 0]), 0))
        t.b_tg = f32(torch.cat((sd["mask_net.output.0.bias"], sd["mask_net.output_gate.0.bias"]), 0))
        t.w_dec1 = tf32(sd["mask_net.conv1_decoder.weight"][:, :, 0])
        t.dec_w = f32(sd["dec.weight"][:, 0, :])

    def _put(self, x, dtype):
        x = x.contiguous().to(dtype).to(self.device)
        self._keep.append(x)
        return ctypes.c_void_p(x.data_ptr())

    def _f32(self, x):
        return self._put(x, torch.float32)

    def _bf16(self, x):
        return self._put(x, torch.bfloat16)

    def _tf32(self, x):
        return self._put(round_tf32(x.contiguous()), torch.float32)
