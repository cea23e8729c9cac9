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


def conv_constants(taps, bias):
    """Per-channel constants of a Linear + SiLU + ConvModule epilogue (csrc/gemm_convt.cuh), [C][20]: the 17 depthwise
    taps with the +1 of `x + conv(x)` (conv_module.py:219) folded into the centre tap, bias / 2 (the SiLU is evaluated
    as h + h tanh(h) with h = x / 2), two zeros (80-byte rows: five 16-byte asynchronous copies per channel)."""
    C = taps.shape[0]
    out = torch.zeros(C, 20, dtype=torch.float32)
    out[:, :17] = taps
    out[:, 8] += 1.0
    out[:, 17] = 0.5 * bias
    return out


def mossformer2_key_shapes(num_layers=_lib.NUM_LAYERS):
    """{key: shape} of the reference MossFormer2().state_dict() (1 099 entries for 24 layers; SURVEY.md 8a13,
    checked key for key against the reference module by tests/test_oracle_port.py)."""
    sh = {"enc.conv1d.weight": (512, 1, 16), "mask_net.norm.weight": (512,), "mask_net.norm.bias": (512,),
          "mask_net.conv1d_encoder.weight": (512, 512, 1), "mask_net.pos_enc.scale": (1,),
          "mask_net.pos_enc.inv_freq": (256,), "mask_net.mdl.intra_mdl.norm.weight": (512,),
          "mask_net.mdl.intra_mdl.norm.bias": (512,), "mask_net.mdl.intra_norm.weight": (512,),
          "mask_net.mdl.intra_norm.bias": (512,), "mask_net.conv1d_out.weight": (1024, 512, 1),
          "mask_net.conv1d_out.bias": (1024,), "mask_net.conv1_decoder.weight": (512, 512, 1),
          "mask_net.prelu.weight": (1,), "mask_net.output.0.weight": (512, 512, 1), "mask_net.output.0.bias": (512,),
          "mask_net.output_gate.0.weight": (512, 512, 1), "mask_net.output_gate.0.bias": (512,),
          "dec.weight": (512, 1, 16)}
    for i in range(num_layers):
        p, q = LAYER.format(i), FSMN.format(i)
        sh[p + "rotary_pos_emb.freqs"] = (16,)
        for n, (o, k) in (("to_hidden", (2048, 512)), ("to_qk", (128, 512)), ("to_out", (512, 1024))):
            sh[p + f"{n}.mdl.0.g"] = (1,)
            sh[p + f"{n}.mdl.1.weight"] = (o, k)
            sh[p + f"{n}.mdl.1.bias"] = (o,)
            sh[p + f"{n}.mdl.3.sequential.1.conv.weight"] = (o, 1, 17)
        sh[p + "qk_offset_scale.gamma"] = (4, 128)
        sh[p + "qk_offset_scale.beta"] = (4, 128)
        sh[q + "conv1.0.weight"] = (256, 512, 1)
        sh[q + "conv1.0.bias"] = (256,)
        sh[q + "conv1.1.weight"] = (1,)
        sh[q + "conv2.weight"] = (512, 256, 1)
        sh[q + "conv2.bias"] = (512,)
        for n in ("norm1", "norm2"):
            sh[q + n + ".weight"] = (256,)
            sh[q + n + ".bias"] = (256,)
        for n in ("to_u", "to_v"):
            sh[q + f"gated_fsmn.{n}.mdl.0.weight"] = (256,)
            sh[q + f"gated_fsmn.{n}.mdl.0.bias"] = (256,)
            sh[q + f"gated_fsmn.{n}.mdl.1.weight"] = (256, 256)
            sh[q + f"gated_fsmn.{n}.mdl.1.bias"] = (256,)
            sh[q + f"gated_fsmn.{n}.mdl.3.sequential.1.conv.weight"] = (256, 1, 17)
        sh[q + "gated_fsmn.fsmn.linear.weight"] = (256, 256)
        sh[q + "gated_fsmn.fsmn.linear.bias"] = (256,)
        sh[q + "gated_fsmn.fsmn.project.weight"] = (256, 256)
        c = q + "gated_fsmn.fsmn.conv."
        sh[c + "conv1.weight"] = (256, 1, 39, 1)
        sh[c + "conv2.weight"] = (256, 2, 39, 1)
        for n in ("norm1", "norm2"):
            sh[c + n + ".weight"] = (256,)
            sh[c + n + ".bias"] = (256,)
        sh[c + "prelu1.weight"] = (256,)
        sh[c + "prelu2.weight"] = (256,)
    return sh


def check_mossformer2_state_dict(state_dict, strict=True):
    """nn.Module.load_state_dict(strict=True) semantics against the architecture the kernels implement."""
    want = mossformer2_key_shapes()
    errors = []
    missing = [k for k in want if k not in state_dict]
    unexpected = [k for k in state_dict if k not in want]
    if missing:
        errors.append("Missing key(s) in state_dict: " + ", ".join(repr(k) for k in missing[:8])
                      + (f" ... ({len(missing)} in total)" if len(missing) > 8 else ""))
    if strict and unexpected:
        errors.append("Unexpected key(s) in state_dict: " + ", ".join(repr(k) for k in unexpected[:8])
                      + (f" ... ({len(unexpected)} in total)" if len(unexpected) > 8 else ""))
    for k, shape in want.items():
        if k in state_dict and tuple(state_dict[k].shape) != shape:
            errors.append(f"size mismatch for {k}: copying a param with shape {tuple(state_dict[k].shape)} from "
                          f"checkpoint, the shape in tdz.Separator (MossFormer2 default architecture) is {shape}")
    if errors:
        raise RuntimeError("Error(s) in loading state_dict for tdz.Separator (MossFormer2):\n\t" + "\n\t".join(errors[:12]))


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
            b_in = torch.cat((sd[p + "to_hidden.mdl.1.bias"], sd[p + "to_qk.mdl.1.bias"]), 0)
            L.b_in = f32(b_in)
            L.dw_in = f32(conv_constants(torch.cat((sd[p + "to_hidden.mdl.3.sequential.1.conv.weight"][:, 0, :],
                                                    sd[p + "to_qk.mdl.3.sequential.1.conv.weight"][:, 0, :]), 0), b_in))
            L.os_gamma = f32(sd[p + "qk_offset_scale.gamma"])
            L.os_beta = f32(sd[p + "qk_offset_scale.beta"])
            L.w_out = bf16(sd[p + "to_out.mdl.1.weight"] * go)
            L.b_out = f32(sd[p + "to_out.mdl.1.bias"])
            L.dw_out = f32(conv_constants(sd[p + "to_out.mdl.3.sequential.1.conv.weight"][:, 0, :],
                                          sd[p + "to_out.mdl.1.bias"]))
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
            L.dw_uv = f32(conv_constants(torch.cat(dws, 0), torch.cat(bs, 0)))
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
        t.dec_wt = tf32(sd["dec.weight"][:, 0, :].t())

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


# ------------------------------------------------------------------------------------------------ ERes2NetV2
SV_BLOCKS = (3, 4, 6, 3)
BN_EPS = 1e-5


def _sv_block_n(n):
    return 32 if n <= 32 else 64 if n <= 64 else 128 if n <= 128 else 256


class PackedEres2NetV2:
    """ERes2NetV2-Large state dict (published module names: conv1/bn1, layer{l}.{b}.{conv1,bn1,convs.i,bns.i,
    fuse_models.i.local_att.{0,1,3,4},conv3,bn3,shortcut.{0,1}}, layer3_ds, fuse34.local_att.*, seg_1) ->
    include/tdz.h: tdz_eres2netv2_weights.  Eval-mode BatchNorm is folded into the preceding convolution."""

    def __init__(self, state_dict, device):
        self.device = torch.device(device)
        self._keep = []
        sd = {k: v.detach().to(torch.float64).cpu() for k, v in state_dict.items() if v.is_floating_point()}
        self.table = _lib.Eres2NetV2Weights()
        t = self.table

        def bn_fold(w, b, bn):
            scale = sd[bn + ".weight"] / torch.sqrt(sd[bn + ".running_var"] + BN_EPS)
            w = w * scale.view(-1, *([1] * (w.ndim - 1)))
            b0 = b if b is not None else torch.zeros_like(scale)
            return w, (b0 - sd[bn + ".running_mean"]) * scale + sd[bn + ".bias"]

        def conv(dst, name, bn=None):
            w = sd[name + ".weight"]
            b = sd.get(name + ".bias")
            if bn is not None:
                w, b = bn_fold(w, b, bn)
            if w.ndim == 4:  # [cout, cin, kh, kw] -> K ordered (kh, kw, cin), matching the im2col kernel
                w = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)
            n, k = w.shape
            bn_tile = _sv_block_n(n)
            n_p, k_p = -(-n // bn_tile) * bn_tile, -(-k // 64) * 64
            wp = torch.zeros(n_p, k_p, dtype=torch.float64)
            wp[:n, :k] = w
            bp = torch.zeros(n_p, dtype=torch.float64)
            if b is not None:
                bp[:n] = b
            dst.w = self._put(wp, torch.bfloat16)
            dst.b = self._put(bp, torch.float32)

        w, b = bn_fold(sd["conv1.weight"], None, "bn1")
        t.stem_w = self._put(w.reshape(64, 9), torch.float32)
        t.stem_b = self._put(b, torch.float32)
        k = 0
        for li, nb in enumerate(SV_BLOCKS):
            for bi in range(nb):
                p = f"layer{li + 1}.{bi}"
                blk = t.blocks[k]
                conv(blk.conv1, p + ".conv1", p + ".bn1")
                for i in range(4):
                    conv(blk.convs[i], f"{p}.convs.{i}", f"{p}.bns.{i}")
                if li >= 2:
                    for i in range(3):
                        a = f"{p}.fuse_models.{i}.local_att"
                        conv(blk.aff_a[i], a + ".0", a + ".1")
                        conv(blk.aff_b[i], a + ".3", a + ".4")
                conv(blk.conv3, p + ".conv3", p + ".bn3")
                if (p + ".shortcut.0.weight") in sd:
                    conv(blk.shortcut, p + ".shortcut.0", p + ".shortcut.1")
                k += 1
        conv(t.layer3_ds, "layer3_ds")
        conv(t.fuse_a, "fuse34.local_att.0", "fuse34.local_att.1")
        conv(t.fuse_b, "fuse34.local_att.3", "fuse34.local_att.4")
        conv(t.seg1, "seg_1")

    def _put(self, x, dtype):
        x = x.contiguous().to(dtype).to(self.device)
        self._keep.append(x)
        return ctypes.c_void_p(x.data_ptr())


# ------------------------------------------------------------------------------------------------ STFT tables + Apollo
def stft_tables(n_fft, device):
    """Periodic Hann window [n_fft] (torch.hann_window(n_fft), what AudioProcessor.py:76 and apollo.py:262 build) and
    the forward twiddles W^k = (cos, -sin)(2 pi k / n_fft), both computed in float64 and rounded once."""
    k = torch.arange(n_fft, dtype=torch.float64)
    win = 0.5 - 0.5 * torch.cos(2.0 * torch.pi * k / n_fft)
    ang = 2.0 * torch.pi * k / n_fft
    tw = torch.stack((torch.cos(ang), -torch.sin(ang)), -1)
    return win.to(torch.float32).to(device).contiguous(), tw.to(torch.float32).to(device).contiguous()


def make_stft_plan(n_fft, hop, device):
    """(tdz_stft_plan, tensors to keep alive)."""
    win, tw = stft_tables(n_fft, device)
    plan = _lib.StftPlan(int(n_fft), int(hop), ctypes.c_void_p(win.data_ptr()), ctypes.c_void_p(tw.data_ptr()))
    return plan, (win, tw)


APOLLO_ARGS = (("sr", 44100), ("win", 20), ("feature_dim", 256), ("layer", 6))


def apollo_key_shapes():
    """{key: shape} of the reference Apollo(sr=44100, win=20, feature_dim=256, layer=6).state_dict() (654 entries;
    look2hear/models/apollo.py:215-257), checked against the reference module by tests/test_apollo_oracle.py."""
    bw = [5] * 79 + [47]
    sh = {}
    for i, w in enumerate(bw):
        sh[f"BN.{i}.0.weight"] = (2 * w + 1,)
        sh[f"BN.{i}.1.weight"] = (256, 2 * w + 1, 1)
        sh[f"BN.{i}.1.bias"] = (256,)
        sh[f"output.{i}.0.weight"] = (256,)
        sh[f"output.{i}.1.weight"] = (4 * w, 256, 1)
        sh[f"output.{i}.1.bias"] = (4 * w,)
    for l in range(_lib.AP_LAYERS):
        p = f"net.{l}.band_net."
        sh[p + "cos_freq"] = (100, 32)
        sh[p + "sin_freq"] = (100, 32)
        sh[p + "input_norm.weight"] = (256,)
        sh[p + "weight.weight"] = (768, 256, 1)
        sh[p + "output.weight"] = (256, 256, 1)
        sh[p + "MLP.0.weight"] = (256,)
        sh[p + "MLP.1.weight"] = (2048, 256, 1)
        sh[p + "MLP_output.weight"] = (256, 1024, 1)
        for b in range(3):
            q = f"net.{l}.seq_net.blocks.{b}.conv."
            sh[q + "0.weight"] = (256, 1, 7)
            sh[q + "0.bias"] = (256,)
            sh[q + "1.weight"] = (256,)
            sh[q + "2.weight"] = (1024, 256, 1)
            sh[q + "2.bias"] = (1024,)
            sh[q + "4.weight"] = (256, 1024, 1)
            sh[q + "4.bias"] = (256,)
    return sh


def check_apollo_state_dict(state_dict, strict=True):
    want = apollo_key_shapes()
    missing = [k for k in want if k not in state_dict]
    unexpected = [k for k in state_dict if k not in want]
    bad = [f"{k}: {tuple(state_dict[k].shape)} != {want[k]}" for k in want
           if k in state_dict and tuple(state_dict[k].shape) != want[k]]
    if bad or (strict and (missing or unexpected)) or missing:
        raise RuntimeError("Error(s) in loading state_dict for Apollo: "
                           f"missing {missing[:5]}{'...' if len(missing) > 5 else ''}, "
                           f"unexpected {unexpected[:5]}{'...' if len(unexpected) > 5 else ''}, size mismatch {bad[:5]}")


class PackedApollo:
    """Apollo state dict -> include/tdz.h: tdz_apollo_weights.  RMSNorm gains in front of a 1x1 conv are folded into
    the conv's columns (W diag(g)); the per-band input / output convs are concatenated and transposed so that the
    band split / merge kernels read them with the output index contiguous."""

    def __init__(self, state_dict, device):
        self.device = torch.device(device)
        self._keep = []
        sd = {k: v.detach().to(torch.float64).cpu() for k, v in state_dict.items()}
        t = self.table = _lib.ApolloWeights()
        bw = [5] * 79 + [47]
        f32 = lambda x: self._put(x, torch.float32)      # noqa: E731
        bf16 = lambda x: self._put(x, torch.bfloat16)    # noqa: E731
        t.bn_g = f32(torch.cat([sd[f"BN.{i}.0.weight"] for i in range(80)]))
        t.bn_w = f32(torch.cat([sd[f"BN.{i}.1.weight"][:, :, 0].t() for i in range(80)], 0))     # [964][256]
        t.bn_b = f32(torch.stack([sd[f"BN.{i}.1.bias"] for i in range(80)]))
        cos, sin = sd["net.0.band_net.cos_freq"], sd["net.0.band_net.sin_freq"]
        for l in range(1, _lib.AP_LAYERS):
            if not (torch.equal(sd[f"net.{l}.band_net.cos_freq"], cos) and torch.equal(sd[f"net.{l}.band_net.sin_freq"], sin)):
                raise ValueError("tdz.Restorer expects the same rotary tables in every layer (apollo.py:82-91)")
        t.rot_cos, t.rot_sin = f32(cos), f32(sin)
        for l in range(_lib.AP_LAYERS):
            p = f"net.{l}.band_net."
            L = t.layers[l]
            L.w_qkv = bf16(sd[p + "weight.weight"][:, :, 0] * sd[p + "input_norm.weight"][None, :])
            L.w_out = bf16(sd[p + "output.weight"][:, :, 0])
            L.w_mlp1 = bf16(sd[p + "MLP.1.weight"][:, :, 0] * sd[p + "MLP.0.weight"][None, :])
            L.w_mlp2 = bf16(sd[p + "MLP_output.weight"][:, :, 0])
            for b in range(3):
                q = f"net.{l}.seq_net.blocks.{b}.conv."
                I = L.icb[b]
                I.dw = f32(sd[q + "0.weight"][:, 0, :].t())                                      # [7][256]
                I.dw_b = f32(sd[q + "0.bias"])
                I.w1 = bf16(sd[q + "2.weight"][:, :, 0] * sd[q + "1.weight"][None, :])
                I.b1 = f32(sd[q + "2.bias"])
                I.w2 = bf16(sd[q + "4.weight"][:, :, 0])
                I.b2 = f32(sd[q + "4.bias"])
        t.out_g = f32(torch.stack([sd[f"output.{i}.0.weight"] for i in range(80)]))
        wv, wg, bv, bg = [], [], [], []
        for i, w in enumerate(bw):
            W, b = sd[f"output.{i}.1.weight"][:, :, 0], sd[f"output.{i}.1.bias"]
            wv.append(W[:2 * w]); wg.append(W[2 * w:]); bv.append(b[:2 * w]); bg.append(b[2 * w:])
        t.out_wv = f32(torch.cat(wv, 0).t())      # [256][884]
        t.out_wg = f32(torch.cat(wg, 0).t())
        t.out_bv = f32(torch.cat(bv))
        t.out_bg = f32(torch.cat(bg))
        plan, keep = make_stft_plan(882, 441, self.device)
        self._keep.extend(keep)
        t.plan = plan

    def _put(self, x, dtype):
        x = x.contiguous().to(dtype).to(self.device)
        self._keep.append(x)
        return ctypes.c_void_p(x.data_ptr())
