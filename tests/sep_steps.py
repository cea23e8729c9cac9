"""Teacher-forced per-kernel parity harness for the separator (GPU).

Every launch step of tdz_separate (enum Step in csrc/tdz_api.cu) is run alone through the C-ABI debug entry
(tdz_separate_debug) with its inputs taken from the oracle (oracle/mossformer2_port.py taps, layer 0) and its
outputs compared with the oracle's.  The workspace is filled with 0xFF (NaN) before each step so that a read
of something the step should not depend on shows up.

A CUDA fault poisons the process, so `run_all()` drives a worker subprocess and restarts it after a failing
step; the result is a dict {step name: {"ok":..., "metrics":..., "error":...}} also written as JSON.

Run by hand:  python -m tests.sep_steps [--out gpurun_out/steps.json]
"""
import json
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

STEPS = ["ENCODER", "ENC1X1", "FLASH_IN", "SIM", "KV", "ATT_OUT", "TO_OUT", "FSMN_C1", "FSMN_UV", "FSMN_LIN",
         "FSMN_PROJ", "DD1", "DD2", "FSMN_TAIL", "FSMN_C2", "FINAL_LN", "FINAL_GN", "OUT1", "TANHSIG", "DEC1",
         "DECODER"]
B, T = 2, 9613  # S = 1200 frames -> Sp = 1280 (partial last group), two samples


def _snr(ref, est):
    ref = ref.double()
    est = est.double()
    den = ((ref - est) ** 2).sum()
    num = (ref ** 2).sum()
    if not math.isfinite(float(den)):
        return float("nan")
    if den == 0:
        return 200.0
    return float(10 * math.log10(float(num / den) + 1e-300))


class Harness:
    def __init__(self):
        import torch
        from oracle.mossformer2_port import mossformer2_forward
        from targetdiarization_b200.synth import random_state_dict
        from targetdiarization_b200 import Separator
        self.torch = torch
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        self.sd = random_state_dict(seed=0, perturb=True)
        g = torch.Generator().manual_seed(1234)
        self.mix = torch.randn(B, T, generator=g) * 0.1
        self.taps = {"_layer": 0}
        with torch.no_grad():
            self.ref_out = mossformer2_forward(self.sd, self.mix, taps=self.taps, num_layers=1)
        self.sep = Separator(self.sd, "cuda:0")
        self.lay = self.sep.layout(B, T)
        self.S, self.Sp, self.M = self.lay.S, self.lay.Sp, self.lay.Mtot
        self.nbytes = self.sep.workspace_bytes(B, T)
        self.ws = self.sep._workspace(self.nbytes)
        self.mix_dev = self.mix.cuda()

    # ---- workspace access
    def view(self, name, dtype, cols, extra_off=0, lead=None):
        torch = self.torch
        off = getattr(self.lay, name) + extra_off
        es = torch.empty(0, dtype=dtype).element_size()
        shape = (B, self.Sp, cols) if lead is None else (lead, B, self.Sp, cols)
        n = math.prod(shape)
        return self.ws[off:off + n * es].view(dtype).view(*shape)

    def pad(self, t):
        torch = self.torch
        out = torch.zeros(t.shape[0], self.Sp, t.shape[2], dtype=t.dtype)
        out[:, :t.shape[1]] = t
        return out

    def put(self, name, t, dtype, **kw):
        v = self.view(name, dtype, t.shape[-1], **kw)
        t = t if t.shape[-2] == self.Sp else (self.pad(t) if t.ndim == 3 else
                                              self.torch.stack([self.pad(x) for x in t]))
        v.copy_(t.to(dtype).cuda())

    def get(self, name, dtype, cols, valid_only=True, as_float=True, **kw):
        v = self.view(name, dtype, cols, **kw)
        v = v.float().cpu() if as_float else v.cpu()
        return v[..., :self.S, :] if valid_only else v

    def poison(self):
        self.ws.fill_(0xFF)

    def run(self, k):
        out = self.sep(self.mix_dev, _debug=(1, k, k))
        self.torch.cuda.synchronize()
        return out.cpu()

    def stats_from(self, y):  # [B,S,256] -> doubles [B,256,2] (sum, sum of squares)
        torch = self.torch
        yd = y.double()
        return torch.stack((yd.sum(1), (yd ** 2).sum(1)), dim=-1)

    def put_raw(self, off, t):
        nb = t.numel() * t.element_size()
        self.ws[off:off + nb].copy_(t.contiguous().view(self.torch.uint8).view(-1).cuda()
                                    if t.dtype != self.torch.uint8 else t.cuda())

    def get_raw(self, off, dtype, n):
        es = self.torch.empty(0, dtype=dtype).element_size()
        return self.ws[off:off + n * es].view(dtype).cpu()


def _qk4_to_float(raw):
    """qk4 as stored ([.., 512] 16-bit: quad_q | lin_q | quad_k | lin_k, lin_q in fp16, the others in bf16) -> fp32."""
    import torch
    out = raw.float()
    out[..., 128:256] = raw[..., 128:256].contiguous().view(torch.float16).float()
    return out


def _qk4_padded(h, qk4):
    """fp32 heads [B,S,512] -> the stored form, padded to Sp with zeros (bf16-typed tensor holding fp16 bits for lin_q)."""
    import torch
    q = h.pad(qk4).to(torch.bfloat16)
    q[..., 128:256] = h.pad(qk4)[..., 128:256].to(torch.float16).view(torch.bfloat16)
    return q


def _check(metrics, name, ref, est, min_db):
    snr = _snr(ref, est)
    metrics[name] = round(snr, 2)
    return snr >= min_db


def run_step(h, step):
    """Returns (ok, metrics).  Thresholds: fp32 SIMT kernels 90 dB+, tf32 GEMMs 55 dB, bf16 GEMMs 40 dB,
    bf16-stored outputs 45 dB (rounding to 8 mantissa bits is ~51 dB)."""
    torch = h.torch
    tp = h.taps
    f32, bf = torch.float32, torch.bfloat16
    k = STEPS.index(step)
    m = {}
    ok = True
    S, Sp, M = h.S, h.Sp, h.M
    h.poison()
    zpad_ok = lambda v: bool((v[:, S:] == 0).all())

    def ss_of(x):  # ScaleNorm partial sums: one per 64 channels
        return (x ** 2).reshape(*x.shape[:-1], 8, 64).sum(-1)

    if step == "ENCODER":
        h.run(k)
        enc = h.get("enc", f32, 512, valid_only=False)
        ok &= _check(m, "enc", tp["enc"], enc[:, :S], 100)
        ok &= zpad_ok(enc)
        samp = h.get_raw(h.lay.samp, f32, 2 * B)
        e = tp["enc"].double()
        mean = e.mean(dim=(1, 2))
        rstd = 1 / torch.sqrt(e.var(dim=(1, 2), unbiased=False) + 1e-8)
        ok &= _check(m, "gn_rstd", rstd.float(), samp[:B], 100)
        ok &= _check(m, "gn_shift", (-rstd * mean).float(), samp[B:], 100)
        rot = h.get_raw(h.lay.rot, f32, Sp * 32).view(Sp, 16, 2)
        fr = h.sd["mask_net.mdl.intra_mdl.mossformerM.layers.0.rotary_pos_emb.freqs"]
        ang = torch.arange(Sp, dtype=f32)[:, None] * fr[None, :]
        ok &= _check(m, "rot", torch.stack((ang.cos(), ang.sin()), -1), rot, 100)
    elif step == "ENC1X1":
        h.put("enc", tp["enc"], f32)
        e = tp["enc"].double()
        mean = e.mean(dim=(1, 2))
        rstd = 1 / torch.sqrt(e.var(dim=(1, 2), unbiased=False) + 1e-8)
        h.put_raw(h.lay.samp, torch.cat((rstd, -rstd * mean)).float())
        h.run(k)
        ok &= _check(m, "x0", tp["x0"], h.get("x0", f32, 512), 55)
        xb = h.get("xbf", bf, 512, valid_only=False)
        ok &= _check(m, "xbf", tp["x0"], xb[:, :S], 45)
        ok &= zpad_ok(xb)
        ok &= _check(m, "ss", ss_of(tp["x0"]), h.get("ss", f32, 8), 55)
    elif step == "FLASH_IN":
        # token shift + ScaleNorm + Linear + SiLU + ConvModule (+ OffsetScale / rotary) in one kernel
        h.put("xbf", tp["x0"], bf)
        h.put("ss", ss_of(tp["x0"]), f32)
        fr = h.sd["mask_net.mdl.intra_mdl.mossformerM.layers.0.rotary_pos_emb.freqs"]
        ang = torch.arange(Sp, dtype=f32)[:, None] * fr[None, :]
        h.put_raw(h.lay.rot, torch.stack((ang.cos(), ang.sin()), -1))
        h.run(k)
        vu = h.get("vu", bf, 2048)
        ok &= _check(m, "vu", tp["vu"], vu, 40)
        ok &= _check(m, "vu_first_frames", tp["vu"][:, :12], vu[:, :12], 40)
        ok &= _check(m, "vu_last_frames", tp["vu"][:, -12:], vu[:, -12:], 40)
        qk = _qk4_to_float(h.get("qk4", bf, 512, as_float=False))
        ok &= _check(m, "qk4", tp["qk4"], qk, 40)
        for i, n in enumerate(("quad_q", "lin_q", "quad_k", "lin_k")):
            ok &= _check(m, n, tp["qk4"][..., i * 128:(i + 1) * 128], qk[..., i * 128:(i + 1) * 128], 40)
        # lin_q is stored as fp16 (11 mantissa bits)
        ok &= _check(m, "lin_q_fp16", tp["qk4"][..., 128:256], qk[..., 128:256], 55)
    elif step == "SIM":
        h.put_raw(h.lay.qk4, _qk4_padded(h, tp["qk4"]))
        h.run(k)
        P = h.get("P", bf, 256, valid_only=False)
        ok &= _check(m, "P", tp["P"], P, 38)
    elif step == "KV":
        h.put_raw(h.lay.qk4, _qk4_padded(h, tp["qk4"]))
        h.put("vu", tp["vu"], bf)
        h.run(k)
        kv = h.get_raw(h.lay.kv, torch.float16, B * 128 * 2048).float().view(B, 128, 2048)
        ok &= _check(m, "kv_fp16", tp["kv"], kv, 50)
    elif step == "ATT_OUT":
        h.put_raw(h.lay.qk4, _qk4_padded(h, tp["qk4"]))
        h.put("vu", tp["vu"], bf)
        h.put("P", tp["P"], bf)
        h.put_raw(h.lay.kv, tp["kv"].to(torch.float16))           # [B][128][2048] fp16
        h.run(k)
        ok &= _check(m, "o", tp["o"], h.get("o", bf, 1024), 38)
        oss = h.get_raw(h.lay.o_ss, f32, 32 * B * h.Sp).view(32, B, h.Sp)[:, :, :S].sum(0)   # part-major [32][Mtot]
        ok &= _check(m, "o_ss", (tp["o"] ** 2).sum(-1), oss, 35)
    elif step == "TO_OUT":
        # ScaleNorm + Linear + SiLU + ConvModule + residual: x = x0 + to_out(o)
        h.put("o", tp["o"], bf)
        ss = (tp["o"] ** 2).reshape(B, S, 8, 4, 32).sum(-1).reshape(B, S, 32)
        ssp = torch.zeros(32, B, h.Sp)
        ssp[:, :, :S] = ss.permute(2, 0, 1)
        h.put_raw(h.lay.o_ss, ssp)                                                           # part-major [32][Mtot]
        h.put("x0", tp["x0"], f32)
        h.run(k)
        xf = h.get("x", f32, 512)
        ok &= _check(m, "x_flash", tp["flash0"], xf, 50)
        ok &= _check(m, "to_out_branch", tp["flash0"] - tp["x0"], xf - tp["x0"], 40)
    elif step == "FSMN_C1":
        h.put("x", tp["flash0"], f32)
        h.run(k)
        ok &= _check(m, "c", tp["c"], h.get("c", f32, 256), 50)
        nh = h.get("nhat", bf, 256, valid_only=False)
        ok &= _check(m, "nhat", tp["nhat"], nh[:, :S], 42)
        ok &= zpad_ok(nh)
    elif step == "FSMN_UV":
        h.put("nhat", tp["nhat"], bf)
        h.run(k)
        ok &= _check(m, "xuv", tp["xuv"], h.get("xuv", f32, 512), 40)
        ok &= _check(m, "xubf", tp["xuv"][..., :256], h.get("xubf", bf, 256), 40)
    elif step == "FSMN_LIN":
        h.put("xubf", tp["xuv"][..., :256], bf)
        h.run(k)
        f1 = h.get("f1", bf, 256, valid_only=False)
        ok &= _check(m, "f1", tp["f1"], f1[:, :S], 38)
        ok &= zpad_ok(f1)
    elif step == "FSMN_PROJ":
        h.put("f1", tp["f1"], bf)
        h.run(k)
        ok &= _check(m, "p", tp["p"], h.get("p", f32, 256), 40)
    elif step == "DD1":
        h.put("p", tp["p"], f32)
        h.run(k)
        ok &= _check(m, "y1", tp["y1"], h.get("y1", f32, 256), 100)
        st = h.get_raw(h.lay.in_stats, torch.float64, B * 512).view(B, 256, 2)
        ok &= _check(m, "stats1", h.stats_from(tp["y1"]), st, 90)
    elif step == "DD2":
        h.put("p", tp["p"], f32)
        h.put("y1", tp["y1"], f32)
        z = torch.zeros(B * 1024, dtype=torch.float64)
        z[:B * 512] = h.stats_from(tp["y1"]).reshape(-1)
        h.put_raw(h.lay.in_stats, z)
        h.run(k)
        ok &= _check(m, "y2", tp["y2"], h.get("y2", f32, 256), 90)
        st = h.get_raw(h.lay.in_stats + B * 512 * 8, torch.float64, B * 512).view(B, 256, 2)
        ok &= _check(m, "stats2", h.stats_from(tp["y2"]), st, 85)
    elif step == "FSMN_TAIL":
        h.put("y2", tp["y2"], f32)
        h.put("xuv", tp["xuv"], f32)
        h.put("c", tp["c"], f32)
        z = torch.zeros(B * 1024, dtype=torch.float64)
        z[B * 512:] = h.stats_from(tp["y2"]).reshape(-1)
        h.put_raw(h.lay.in_stats, z)
        h.run(k)
        g = h.get("g", f32, 256, valid_only=False)
        ok &= _check(m, "g", tp["g"], g[:, :S], 65)  # stored rounded to tf32 (operand of conv2 only)
        ok &= zpad_ok(g)
    elif step == "FSMN_C2":
        h.put("g", tp["g"], f32)
        h.put("x", tp["flash0"], f32)
        h.run(k)
        ok &= _check(m, "x_layer", tp["layer0"], h.get("x", f32, 512), 55)
        xb = h.get("xbf", bf, 512, valid_only=False)
        ok &= _check(m, "xbf", tp["layer0"], xb[:, :S], 45)
        ok &= zpad_ok(xb)
        ok &= _check(m, "ss", ss_of(tp["layer0"]), h.get("ss", f32, 8), 55)
    elif step == "FINAL_LN":
        h.put("x", tp["layer0"], f32)
        h.run(k)
        ok &= _check(m, "final_ln", tp["final_ln"], h.get("lnb", f32, 512), 100)
        samp = h.get_raw(h.lay.samp, f32, 4 * B)
        e = tp["final_ln"].double()
        mean = e.mean(dim=(1, 2))
        rstd = 1 / torch.sqrt(e.var(dim=(1, 2), unbiased=False) + 1e-8)
        ok &= _check(m, "gn_rstd", rstd.float(), samp[2 * B:3 * B], 100)
        ok &= _check(m, "gn_shift", (-rstd * mean).float(), samp[3 * B:], 90)
    elif step == "FINAL_GN":
        h.put("lnb", tp["final_ln"], f32)
        h.put("x0", tp["x0"], f32)
        e = tp["final_ln"].double()
        mean = e.mean(dim=(1, 2))
        rstd = 1 / torch.sqrt(e.var(dim=(1, 2), unbiased=False) + 1e-8)
        h.put_raw(h.lay.samp + 2 * B * 4, torch.cat((rstd, -rstd * mean)).float())
        h.run(k)
        ab = h.get("ab", f32, 512, valid_only=False)
        ok &= _check(m, "mask_in", tp["mask_in"], ab[:, :S], 65)  # stored rounded to tf32
        ok &= zpad_ok(ab)
    elif step == "OUT1":
        h.put("ab", tp["mask_in"], f32)
        h.run(k)
        ok &= _check(m, "m", tp["m"], h.get("mb", f32, 1024), 55)
    elif step == "TANHSIG":
        h.put("mb", tp["m"], f32)
        h.run(k)
        g = h.get("gated", f32, 512, lead=2)
        ok &= _check(m, "gated", torch.stack(tp["gated"]), g, 55)
    elif step == "DEC1":
        h.put("gated", torch.stack(tp["gated"]), f32, lead=2)
        h.put("enc", tp["enc"], f32)
        h.run(k)
        sp = h.get("sep", f32, 512, lead=2)
        ok &= _check(m, "sep", torch.stack(tp["sep"]), sp, 55)
    elif step == "DECODER":
        h.put("sep", torch.stack(tp["sep"]), f32, lead=2)
        out = h.run(k)
        ok &= _check(m, "out", h.ref_out, out, 60)      # tf32 GEMM (operands truncated to 10 mantissa bits)
        ok &= bool((out[..., (h.S - 1) * 8 + 16:] == 0).all())   # zero pad beyond T' = 8 (S - 1) + 16
    else:
        raise ValueError(step)
    return bool(ok), m


def worker(first, out_path):
    """Runs steps first.. in this process, appending one JSON line per step; stops at the first CUDA failure."""
    h = Harness()
    for step in STEPS[first:]:
        rec = {"step": step}
        try:
            ok, metrics = run_step(h, step)
            rec.update(ok=ok, metrics=metrics)
        except Exception as e:  # noqa: BLE001  (CUDA errors arrive as RuntimeError)
            rec.update(ok=False, error=f"{type(e).__name__}: {e}"[:400], fatal=True)
        with open(out_path, "a") as f:
            f.write(json.dumps(rec) + "\n")
        print(json.dumps(rec), flush=True)
        if rec.get("fatal"):
            return 3
    return 0


def run_all(out_path, timeout=300):
    if os.path.exists(out_path):
        os.remove(out_path)
    first = 0
    while first < len(STEPS):
        try:
            subprocess.run([sys.executable, "-m", "tests.sep_steps", "--worker", str(first), "--out", out_path],
                           cwd=ROOT, timeout=timeout)
        except subprocess.TimeoutExpired:
            n_done = 0
            if os.path.exists(out_path):
                with open(out_path) as f:
                    n_done = sum(1 for line in f if line.strip())
            with open(out_path, "a") as f:
                f.write(json.dumps({"step": STEPS[n_done], "ok": False, "error": "timeout", "fatal": True}) + "\n")
        done = []
        if os.path.exists(out_path):
            with open(out_path) as f:
                done = [json.loads(line) for line in f if line.strip()]
        if len(done) <= first:  # the worker died before reporting this step
            with open(out_path, "a") as f:
                f.write(json.dumps({"step": STEPS[first], "ok": False, "error": "worker died", "fatal": True}) + "\n")
            done.append(None)
        first = len(done)
    with open(out_path) as f:
        return {r["step"]: r for r in (json.loads(line) for line in f if line.strip())}


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--worker", type=int, default=None)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "steps.jsonl"))
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    if a.worker is not None:
        sys.exit(worker(a.worker, a.out))
    res = run_all(a.out)
    bad = [k for k, v in res.items() if not v["ok"]]
    print("FAILED:" if bad else "ALL OK", bad)
    sys.exit(1 if bad else 0)
