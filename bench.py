#!/usr/bin/env python
"""Benchmark of the target-speaker separation + scoring stage (BASELINE.json metric: audio-seconds separated +
scored per wall-second).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our sm_100a path (libtdz.so)
  python bench.py --impl reference [...]                          the reference's CPU path (oracle port), host cores

`value` / `e2e` (the headline): one step = one pass of the hot path over one batch of synthetic mixtures -
configs[1] of BASELINE.json, 64 mixtures x 4 s at 16 kHz per GPU (weak scaling: every rank owns its own 64
independent chunks, no collective on the data path).  The same JSON line carries, as extra records:

  strong_c3   BASELINE config 3, the split the north_star names: ONE 1 h recording cut into per-rank spans (+ halos
              in overlap-add mode), separated, cut into 4 s segments that are scored against the target (segments
              dealt to the ranks), spans and scores gathered by NCCL; host array in -> host arrays out.  Strong
              scaling; under torchrun the result is compared bit for bit with a single-rank run inside the bench.
  score_c4    config 4: 4 096 separated segments embedded + scored vs one target (sharded over the ranks)
  stream_c5   config 5: 600 ms chunks, batch 1 (p50 / p99 latency) and batch 256 concurrent streams (throughput)
  gpu_eager_baseline   the reference's real incumbent on this box: the PyTorch restatement of the reference modules
              run eagerly on the B200 (fp32 and TF32), same C2 shape
  cpu_baseline the reference's CPU path on the box's host cores (a bounded sample)
Prints ONE JSON line (rank 0).
"""
import argparse
import csv
import glob
import gzip
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 16000
ITEMS, SECONDS = 64, 4.0
T = int(SR * SECONDS)
METRIC = "audio_seconds_separated_and_scored_per_second"
UNIT = "audio-s/s"
WORKLOAD = "C2: 64 synthetic 4 s 2-speaker mixtures/GPU, MossFormer2 separation + fbank/ERes2NetV2 scoring of both streams"

# Algorithmic work per launch step, per frame (= 8 samples) of one chunk.
#   flops: dense contraction FLOPs;  bytes: compulsory HBM bytes (read + write) of that step;
#   bound: what bounds the step ON THIS PART (HBM unless the contraction is deep enough to be tensor bound: the
#   ridge is peak_flops / peak_bytes ~ 250 FLOP/B, i.e. a K=256 / K=512 projection of fp32 rows is HBM bound);
#   fmt: operand format of the contraction (selects the tensor peak: tf32 runs at half the bf16 rate)
STEP_TABLE = {
    "ENCODER": dict(bound="hbm", bytes=32 + 2048),
    "ENC1X1": dict(bound="hbm", fmt="tf32", flops=524288, bytes=2048 + 2048 + 1024 + 16),
    # Linear 512->2176 + SiLU + depthwise k17 (one kernel) + OffsetScale/rotary: reads xbf, writes vu + qk4 + lin_q lo
    "FLASH_IN": dict(bound="tensor", fmt="bf16", flops=2 * 512 * 2176, bytes=1024 + 4096 + 1024 + 256),
    "SIM": dict(bound="hbm", fmt="bf16", flops=2 * 256 * 128, bytes=512 + 512),
    "KV": dict(bound="hbm", fmt="bf16", flops=2 * 128 * 2048, bytes=256 + 4096),
    "ATT_OUT": dict(bound="tensor", fmt="bf16", flops=2 * 256 * 2048 + 2 * 128 * 2048, bytes=512 + 4096 + 256 + 4096 + 2048),
    "TO_OUT": dict(bound="tensor", fmt="bf16", flops=2 * 1024 * 512, bytes=2048 + 2048 + 2048),
    "FSMN_C1": dict(bound="hbm", fmt="tf32", flops=2 * 512 * 256, bytes=2048 + 1024 + 512),
    "FSMN_UV": dict(bound="hbm", fmt="bf16", flops=2 * 256 * 512, bytes=512 + 2048 + 512),
    # fsmn.linear -> ReLU -> fsmn.project as one back-to-back GEMM (hidden activations stay on chip)
    "FSMN_LIN": dict(bound="hbm", fmt="bf16", flops=2 * 2 * 256 * 256, bytes=512 + 1024),
    "DD1": dict(bound="hbm", bytes=1024 + 1024),
    "DD2": dict(bound="hbm", bytes=2048 + 1024),
    "FSMN_TAIL": dict(bound="hbm", bytes=1024 + 2048 + 1024 + 1024),
    "FSMN_C2": dict(bound="hbm", fmt="tf32", flops=2 * 256 * 512, bytes=1024 + 2048 + 2048 + 1024),
    "FINAL_LN": dict(bound="hbm", bytes=4096),
    "FINAL_GN": dict(bound="hbm", bytes=6144),
    "OUT1": dict(bound="hbm", fmt="tf32", flops=2 * 512 * 1024, bytes=2048 + 4096),
    "TANHSIG": dict(bound="tensor", fmt="tf32", flops=2 * 2 * 512 * 1024, bytes=4096 + 4096),
    "DEC1": dict(bound="hbm", fmt="tf32", flops=2 * 2 * 512 * 512, bytes=4096 + 2048 + 4096),
    "DECODER": dict(bound="hbm", bytes=4096 + 64),
}
LAYER_STEPS = ["FLASH_IN", "SIM", "KV", "ATT_OUT", "TO_OUT", "FSMN_C1", "FSMN_UV", "FSMN_LIN", "DD1",
               "DD2", "FSMN_TAIL", "FSMN_C2"]
ALL_STEPS = ["ENCODER", "ENC1X1"] + LAYER_STEPS + ["FINAL_LN", "FINAL_GN", "OUT1", "TANHSIG", "DEC1", "DECODER"]
# the kernel that carries each step's time (name fragment in the ncu export), for `roofline.traffic`
STEP_KERNEL = {"FLASH_IN": "gemm_convt_kernel<0>", "ATT_OUT": "gemm_cg2_kernel", "TO_OUT": "gemm_convt_cg2_kernel<1>",
               "FSMN_UV": "gemm_convt_cg2_kernel<2>", "DD1": "dd_stream_kernel<1>", "DD2": "dd_stream_kernel<2>",
               "FSMN_TAIL": "fsmn_tail_kernel", "DECODER": "decoder_kernel", "ENCODER": "encoder_kernel"}


NCU_CAPTURE_ITEMS = 16   # batch of the committed capture (tools/ncu_steps.sh default), same T as the bench


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tensor=d["bf16_tflops"], tensor_sustained=d["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json (burst: each step is timed alone)")
    return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0, source="fallback of B200_PROFILING.md")


def ncu_traffic(step):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the step's kernel, from the committed ncu export
    (profiles/r2_ncu_raw.csv[.gz]: one `ncu --set full` capture of this bench command, `--page raw --csv`)."""
    frag = STEP_KERNEL.get(step)
    if frag is None:
        return None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r2_ncu_raw*.csv*")), reverse=True):
        op = gzip.open if path.endswith(".gz") else open
        try:
            with op(path, "rt", newline="") as f:
                rows = list(csv.reader(f))
        except OSError:
            continue
        if len(rows) < 3:
            continue
        hdr = rows[0]
        try:
            kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        except ValueError:
            continue
        units = rows[1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals = []
        for r in rows[2:]:
            if len(r) > max(kn, rd, wr) and frag in r[kn].replace("(int)", ""):
                try:
                    vals.append(float(r[rd].replace(",", "")) * scale.get(units[rd], 1.0)
                                + float(r[wr].replace(",", "")) * scale.get(units[wr], 1.0))
                except ValueError:
                    pass
        if vals:
            return dict(bytes_per_launch=statistics.median(vals), launches_in_capture=len(vals),
                        source=os.path.relpath(path, ROOT))
    return None


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ baselines
def cpu_leg(steps, warmup, sample_items=1):
    """The reference's CPU path for this stage: MossFormer2 forward per chunk (batch 1, serial, as
    AudioProcessor.separate_speaker does) + fbank/ERes2NetV2 embedding of both streams + cosine, fp32, all host
    threads.  Bounded sample: `sample_items` 4 s mixtures per step."""
    import torch
    from oracle.mossformer2_port import mossformer2_forward
    from targetdiarization_b200.synth import random_state_dict, synthetic_mixture
    from oracle import eres2netv2_port as E
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = random_state_dict(seed=0)
    esd = E.random_state_dict(seed=0)
    mix = synthetic_mixture(sample_items, T, seed=1234)
    target = E.embed(esd, synthetic_mixture(1, T, seed=99))[0].numpy()
    times = []
    with torch.no_grad():
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            for i in range(sample_items):
                est = mossformer2_forward(sd, mix[i:i + 1])[0]      # [2,T]
                emb = E.embed(esd, est)                            # [2,192]
                _ = [E.cosine_similarity(emb[k].numpy(), target) for k in range(2)]
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    per_step = sum(times) / len(times)
    return dict(value=sample_items * SECONDS / per_step, seconds_per_step=per_step, cores=threads,
                sample=f"{sample_items} x 4 s mixture per step (of the 64 in the workload; the reference processes "
                       f"chunks one at a time, so the figure does not depend on the batch), batch 1, fp32, "
                       f"{threads} torch threads")


def gpu_eager_leg(dev, items=16):
    """The reference's incumbent on THIS box: the PyTorch restatement of the reference modules run
    eagerly on the B200, fp32 and TF32 (torch.backends.cuda.matmul.allow_tf32), separation + scoring of both
    streams, batch `items` of the C2 shape (the eager path keeps every [B,S,2048] activation in fp32: 16 items
    per call keep it well inside memory; the rate does not grow with a larger batch)."""
    import torch
    from oracle.mossformer2_port import mossformer2_forward
    from targetdiarization_b200.synth import random_state_dict, synthetic_mixture
    from oracle import eres2netv2_port as E
    sd = {k: v.to(dev) for k, v in random_state_dict(seed=0).items()}
    esd = {k: v.to(dev) for k, v in E.random_state_dict(seed=0).items()}
    mix = synthetic_mixture(items, T, seed=1234).to(dev)
    out = {}
    for name, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        with torch.no_grad():
            def step():
                est = mossformer2_forward(sd, mix)                      # [items,2,T]
                return E.embed(esd, est.view(2 * items, T))
            step()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            torch.cuda.synchronize(dev)
        out[name] = dict(value=items * SECONDS / (e0.elapsed_time(e1) / 1e3), unit=UNIT, ms_per_step=e0.elapsed_time(e1))
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    out["sample"] = f"{items} x 4 s mixtures per call, PyTorch eager (cuBLAS / cuDNN kernels), 1 warm-up + 1 timed call"
    out["kind"] = "port (PyTorch restatement of the reference modules) on cuda"
    del sd, esd, mix
    torch.cuda.empty_cache()
    return out


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    r = cpu_leg(steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": r["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "each step = ONE 4 s mixture of the 64 (BASELINE.md section 3: a full C2 "
                   "step takes ~3 min per repetition on the host cores); steps clamped to 3, warm-up to 1",
                   "note": "reference modules restated in oracle/ (the reference tree cannot travel to the GPU box); "
                   "CPU only, rank 0 only"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ extra records
def long_recording(minutes):
    """`minutes` of synthetic conversation: a 64 s synthetic mixture tiled with per-tile gains (cheap to generate,
    identical on every rank)."""
    import numpy as np
    from targetdiarization_b200.synth import synthetic_mixture
    L = int(minutes * 60 * SR)
    base = synthetic_mixture(1, 64 * SR, seed=11)[0].numpy()
    reps = -(-L // base.shape[0])
    g = np.random.default_rng(5).uniform(0.5, 1.0, size=reps).astype(np.float32)
    return (np.tile(base, reps).reshape(reps, -1) * g[:, None]).reshape(-1)[:L].copy()


def strong_c3(stage, target_emb, world, rank, dev, minutes, modes):
    """BASELINE config 3.  The stage is switched to the shared-input (sharded) form for the duration."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from targetdiarization_b200 import pipeline
    audio = long_recording(minutes)
    L = audio.shape[0]
    group = dist.group.WORLD if world > 1 else None
    stage.group, stage.gather_dst = group, 0
    rec = {"audio_seconds": L / SR, "segments_scored": 2 * (L // T), "ranks": world, "scaling": "strong",
           "what": "host ndarray in -> (spk1, spk2) host ndarrays on rank 0 + [2, n_seg] scores + per-segment pick; "
                   "every rank uploads and separates only its span (+ halo), one NCCL gather of the spans, one "
                   "all_gather of the scores",
           "warmup": "the same call once, untimed (workspaces sized, caching allocators filled); then ONE timed call"}
    try:
        for mode in modes:
            # warm-up = the same call once, untimed: the timed call is the steady state of a long-lived server - the
            # separator / embedder workspaces already have their size for this recording (growing them is a cudaMalloc of
            # tens of GB), the result buffers come out of torch's caching allocators (the first cudaHostAlloc of the
            # 460 MB page-locked buffer alone costs ~0.15 s) and NCCL has its channels
            stage.separate_and_score_long(audio, target_emb, mode=mode, loudness="device")
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            split = {}
            res = stage.separate_and_score_long(audio, target_emb, mode=mode, loudness="device", timings=split)
            e1.record()
            torch.cuda.synchronize(dev)
            wall_local = time.perf_counter() - t0
            ms = torch.tensor([e0.elapsed_time(e1), wall_local * 1e3], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                dist.barrier()
            dev_s, wall_s = float(ms[0]) / 1e3, float(ms[1]) / 1e3
            # bytes that crossed PCIe on this rank
            if mode == "concat":
                b, e = pipeline.P.concat_shard(L, rank, world)[1:] if world > 1 else (0, L)
            else:
                plan = pipeline.P.ola_plan(L)
                sh = pipeline.P.ola_shard(plan, rank, world) if world > 1 else (0, L, 0, plan.num_session)
                b, e = pipeline.ola_input_range(plan, sh[2], sh[3])
            r = {"seconds": wall_s, "device_seconds": dev_s, "value": L / SR / wall_s, "unit": UNIT,
                 "h2d_bytes_rank0": int((e - b) * 4), "d2h_bytes_rank0": int(2 * L * 4 + 2 * (L // T) * 4),
                 "targets_picked": int((res["target"] > 0).sum()),
                 "split_seconds_rank0": {k: round(v, 4) for k, v in split.items()}}
            # the gather alone (same buffers, no compute): what NCCL costs
            if world > 1:
                lens = [(L // world) + (1 if i < L % world else 0) for i in range(world)]
                flat = torch.zeros(2 * max(lens), device=dev)
                pipeline.gather_spans(stage.kern, flat, lens, group, 0)
                torch.cuda.synchronize(dev)
                dist.barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                pipeline.gather_spans(stage.kern, flat, lens, group, 0)
                g1.record()
                torch.cuda.synchronize(dev)
                r["nccl_gather_ms"] = g0.elapsed_time(g1)
                # bit identity with a single-rank run of the same recording (rank 0 recomputes alone)
                stage.group = None
                ok = None
                if rank == 0:
                    single = stage.separate_and_score_long(audio, target_emb, mode=mode, loudness="device")
                    ok = bool(np.array_equal(single["spk1"], res["spk1"]) and np.array_equal(single["spk2"], res["spk2"])
                              and np.array_equal(single["scores"], res["scores"]))
                stage.group = group
                dist.barrier()
                r["bit_identical_to_single_rank"] = ok
            rec[mode] = r
    finally:
        stage.group, stage.gather_dst = None, None
    return rec


def score_c4(stage, target_emb, world, dev, n_seg=4096):
    """BASELINE config 4: 4 096 separated 4 s segments embedded + scored against one target (sharded over the ranks
    when there are several; one all_gather of the scores)."""
    import torch
    import torch.distributed as dist
    from targetdiarization_b200.synth import synthetic_mixture
    base = synthetic_mixture(64, T + n_seg // 64 * 16, seed=41).to(dev)
    segs = torch.stack([base[i % 64, (i // 64) * 16:(i // 64) * 16 + T] for i in range(n_seg)])
    stage.group = dist.group.WORLD if world > 1 else None
    try:
        stage.score_segments(segs[:256 * world], target_emb)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        scores = stage.score_segments(segs, target_emb)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ok = bool(torch.isfinite(scores).all() and ((scores >= 0) & (scores <= 1)).all())
    finally:
        stage.group = None
    s = float(ms[0]) / 1e3
    return {"segments": n_seg, "segment_seconds": SECONDS, "seconds": s, "segments_per_second": n_seg / s,
            "value": n_seg * SECONDS / s, "unit": "audio-s/s (scoring only)", "scores_valid": ok, "ranks": world,
            "scaling": "strong"}


def stream_c5(stage, target_emb, dev, steps=100):
    """BASELINE config 5: 600 ms chunks (T = 9 600) through separation + scoring of both streams, 100 consecutive
    steps at batch 1 (latency, host-synchronised per step like a streaming server) and at 256 concurrent streams."""
    import torch
    from targetdiarization_b200.synth import synthetic_mixture
    out = {"chunk_seconds": 0.6}
    for B, n in ((1, steps), (256, max(steps // 4, 10))):
        mix = synthetic_mixture(B, 9600, seed=51).to(dev)
        for _ in range(3):
            stage.run(mix, target_emb)
        torch.cuda.synchronize(dev)
        lat = []
        for _ in range(n):
            t0 = time.perf_counter()
            est, scores = stage.run(mix, target_emb)
            scores.cpu()                      # the pick needs the scores on the host
            lat.append((time.perf_counter() - t0) * 1e3)
        lat.sort()
        p50, p99 = lat[len(lat) // 2], lat[min(len(lat) - 1, int(len(lat) * 0.99))]
        out[f"batch{B}"] = {"steps": n, "latency_ms_p50": p50, "latency_ms_p99": p99,
                            "stream_seconds_per_second": B * 0.6 / (statistics.mean(lat) / 1e3)}
    return out


def restore_f4(dev, seconds=10.0):
    """the Apollo restorer on `seconds` of mono 44.1 kHz audio (host array in, host array out, like
    AudioProcessor.restore_audio) and the MDX-Net STFT / inverse STFT pair on 8 stereo chunks of the production shape
    (n_fft 6144, hop 1024, 256 frames), with the oracle port timed on the host cores on a 1 s sample."""
    import time
    import numpy as np
    import torch
    from oracle import apollo_port as AP
    from targetdiarization_b200 import ConvTDFNet, Restorer, synth
    sd = synth.random_apollo_state_dict(0)
    rest = Restorer(sd, dev)
    ns = int(seconds * 44100)
    audio = synth.synthetic_fullband(1, ns, seed=1).numpy()
    pinned = torch.from_numpy(audio).pin_memory()

    def ev(fn, reps=5):
        fn()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / reps

    x_dev = pinned.to(dev).reshape(1, 1, ns)
    ms_res = ev(lambda: rest(x_dev))
    ms_e2e = ev(lambda: rest(pinned.to(dev, non_blocking=True).reshape(1, 1, ns)).cpu())
    tokens = (1 + ns // 441) * 80
    flop = tokens * 6 * 2 * (256 * 768 + 256 * 256 + 256 * 2048 + 1024 * 256 + 3 * 2 * 256 * 1024)
    xs = torch.from_numpy(audio[:, :44100]).reshape(1, 1, -1)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        AP.apollo_forward(sd, xs)
        t0 = time.time()
        ref = AP.apollo_forward(sd, xs)
        cpu_s = time.time() - t0
    snr = AP.snr_db(ref, rest(xs.to(dev)).cpu())
    net = ConvTDFNet("vocals", 11, 3072, 8, 6144, 1024, dev)
    xm = torch.randn(8, 2, net.chunk_size, device=dev) * 0.1
    spec = net.stft(xm)
    ms_stft = ev(lambda: net.stft(xm))
    ms_istft = ev(lambda: net.istft(spec))
    mdx_s = 8 * net.chunk_size / 44100
    t0 = time.time()
    AP.mdx_istft(AP.mdx_stft(xm[:2].cpu(), 6144, 1024, 3072), 6144, 1024)
    mdx_cpu = (time.time() - t0) / (2 * net.chunk_size / 44100)
    del rest, net
    torch.cuda.empty_cache()
    return {"apollo": {"audio_s": seconds, "ms_resident": ms_res, "x_realtime_resident": seconds / (ms_res / 1e3),
                       "ms_host_to_host": ms_e2e, "x_realtime": seconds / (ms_e2e / 1e3),
                       "gemm_tflops": flop / (ms_res * 1e-3) / 1e12, "kernels_per_forward": Restorer.KERNELS_PER_FORWARD,
                       "snr_db_vs_oracle_1s": snr,
                       "cpu_port": {"x_realtime": 1.0 / cpu_s, "cores": torch.get_num_threads(), "sample": "1 s, 1 warm-up + 1 run"}},
            "mdx_stft_istft": {"audio_s": mdx_s, "stft_ms": ms_stft, "istft_ms_incl_d2h": ms_istft,
                               "x_realtime": mdx_s / ((ms_stft + ms_istft) / 1e3),
                               "cpu_torch_x_realtime": 1.0 / mdx_cpu}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="tdz", choices=["tdz", "reference"])
    ap.add_argument("--items", type=int, default=ITEMS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the C2 headline (value / e2e / roofline)")
    ap.add_argument("--c3-minutes", type=float, default=60.0)
    ap.add_argument("--c3-modes", default="concat,ola")
    ap.add_argument("--breakdown", default=None, help="write the per-step timing table to this JSON file")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from targetdiarization_b200.pipeline import SeparationScoringStage
    from targetdiarization_b200.synth import synthetic_mixture

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)
    B = args.items

    stage = SeparationScoringStage.random_init(dev, seed=0)   # no process group: every rank owns its own mixtures
    mix_host = synthetic_mixture(B, T, seed=1234 + rank).pin_memory()
    target_host = synthetic_mixture(1, T, seed=99)
    target_emb = stage.embed(target_host.to(dev))[0]
    mix_dev = mix_host.to(dev)
    out_host = torch.empty(B, 2, T, dtype=torch.float32).pin_memory()
    scores_host = torch.empty(B, 2, dtype=torch.float32).pin_memory()

    def step_resident():
        return stage.run(mix_dev, target_emb)

    class E2E:
        """The e2e step: pinned host input -> device, separation + scoring, waveforms and scores -> pinned host memory,
        every step, all inside the timed region.  The copies run on two side streams so that the upload of step i + 1
        and the download of step i - 1 overlap the kernels of step i (what a serving loop does); `finish` makes the
        timed stream wait for the last download before the closing event is recorded."""

        def __init__(self):
            self.main = torch.cuda.current_stream(dev)
            self.cin, self.cout = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            self.inp = [torch.empty(B, T, dtype=torch.float32, device=dev) for _ in range(2)]
            self.h2d = [None, None]      # event: upload into inp[k] done
            self.used = [None, None]     # event: the step that read inp[k] has finished
            self.d2h = []                # download events of the steps in flight
            self.i = 0

        def upload(self, k):
            if self.used[k] is not None:
                self.cin.wait_event(self.used[k])
            with torch.cuda.stream(self.cin):
                self.inp[k].copy_(mix_host, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.cin)
            self.h2d[k] = ev

        def step(self):
            k = self.i & 1
            if self.h2d[k] is None:
                self.upload(k)               # first step: nothing was prefetched
            self.main.wait_event(self.h2d[k])
            self.h2d[k] = None
            est, scores = stage.run(self.inp[k], target_emb)
            done = torch.cuda.Event()
            done.record(self.main)
            self.used[k] = done
            self.cout.wait_event(done)
            with torch.cuda.stream(self.cout):
                out_host.copy_(est, non_blocking=True)
                scores_host.copy_(scores, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.cout)
            est.record_stream(self.cout)
            scores.record_stream(self.cout)
            self.d2h.append(ev)
            if len(self.d2h) > 2:            # at most two downloads in flight (they share the host buffers in order)
                self.main.wait_event(self.d2h.pop(0))
            self.upload(k ^ 1)               # next step's input, under this step's kernels
            self.i += 1

        def finish(self):
            for ev in self.d2h:
                self.main.wait_event(ev)
            self.d2h = []
            self.h2d = [None, None]          # a prefetched upload beyond the last step is not reused

    e2e = E2E()
    step_e2e = e2e.step

    mix_np = mix_host.numpy()
    target_np = target_emb.cpu().numpy()

    def step_dropin():
        # the reference's own call granularity: one recording per call, numpy in / numpy out
        # (AudioProcessor.separate_speaker + 2 x get_speaker_embedding + cosine_similarity, TargetASR.py:609-625)
        for i in range(B):
            stage.separate_and_score(mix_np[i], target_np, loudness=None)

    def timed(fn, k, warm=warmup, finish=None):
        for _ in range(warm):
            fn()
        if finish is not None:
            finish()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        if finish is not None:
            finish()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * B * SECONDS / (ms_step / 1e3)

    ms_e2e = timed(step_e2e, args.steps, finish=e2e.finish) / args.steps
    # the same step with the copies in line on the compute stream and a host synchronisation per step (latency form)
    def step_e2e_sync():
        m = mix_host.to(dev, non_blocking=True)
        est, scores = stage.run(m, target_emb)
        out_host.copy_(est, non_blocking=True)
        scores_host.copy_(scores, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    ms_e2e_sync = timed(step_e2e_sync, max(2, args.steps // 2)) / max(2, args.steps // 2)
    e2e_value = world * B * SECONDS / (ms_e2e / 1e3)
    ms_dropin = timed(step_dropin, 1, warm=1)
    dropin_value = world * B * SECONDS / (ms_dropin / 1e3)

    # ---- per-step breakdown of the separator (one layer instance of each launch step), CUDA events
    peaks = load_peaks()
    roof = None
    if rank == 0:
        frames = stage.separator.layout(B, T).S * B
        table = stage.separator.time_steps(mix_dev, reps=5)
        rows = []
        for name in ALL_STEPS:
            ms = table[name]
            mult = 24 if name in LAYER_STEPS else 1
            info = STEP_TABLE[name]
            row = dict(step=name, ms=ms, launches_per_forward=mult, ms_per_forward=ms * mult, bound=info["bound"])
            if "flops" in info:
                row["tflops"] = info["flops"] * frames / (ms * 1e-3) / 1e12
                row["operands"] = info["fmt"]
            row["gbs"] = info["bytes"] * frames / (ms * 1e-3) / 1e9
            if info["bound"] == "tensor":
                row["achieved"] = row["tflops"]
                row["peak"] = peaks["tensor"] * (0.5 if info["fmt"] == "tf32" else 1.0)   # tf32 MMA: half the bf16 rate
                row["unit"] = "TFLOP/s"
            else:
                row["achieved"] = row["gbs"]
                row["peak"] = peaks["hbm"]
                row["unit"] = "GB/s"
            row["frac"] = row["achieved"] / row["peak"]
            rows.append(row)
        rows.sort(key=lambda r: -r["ms_per_forward"])
        top = rows[0]
        tr = ncu_traffic(top["step"])
        # the ncu capture runs the same kernel on NCU_CAPTURE_ITEMS items (ncu cannot replay the 64-item batch: it
        # saves / restores all touched memory around every pass); DRAM bytes are proportional to the frame count
        roof = dict(kernel=top["step"], bound=top["bound"], achieved=top["achieved"], peak=top["peak"],
                    unit=top["unit"], frac=top["frac"],
                    traffic=tr["bytes_per_launch"] * B / NCU_CAPTURE_ITEMS if tr else None,
                    traffic_source=(f'{tr["source"]}: dram__bytes_read.sum + dram__bytes_write.sum of one launch at '
                                    f'{NCU_CAPTURE_ITEMS} items x {T} samples = {tr["bytes_per_launch"]:.0f} B, '
                                    f'scaled by {B}/{NCU_CAPTURE_ITEMS} to this launch') if tr else None,
                    peak_source=peaks["source"],
                    algorithmic_bytes_per_launch=STEP_TABLE[top["step"]]["bytes"] * frames,
                    share_of_separator=top["ms_per_forward"] / sum(r["ms_per_forward"] for r in rows),
                    steps={r["step"]: dict(ms=round(r["ms"], 4), bound=r["bound"], frac=round(r["frac"], 3),
                                           unit=r["unit"]) for r in rows})
        if args.breakdown:
            os.makedirs(os.path.dirname(os.path.abspath(args.breakdown)), exist_ok=True)
            with open(args.breakdown, "w") as f:
                json.dump(dict(ms_per_step=ms_step, B=B, T=T, rows=rows), f, indent=1)

    c3 = c4 = c5 = eager = f4 = None
    if not args.no_extras:
        c3 = strong_c3(stage, target_emb, world, rank, dev, args.c3_minutes,
                       [m for m in args.c3_modes.split(",") if m])
        c4 = score_c4(stage, target_emb, world, dev)
        if rank == 0 and world == 1:
            c5 = stream_c5(stage, target_emb, dev)
            eager = gpu_eager_leg(dev)
            f4 = restore_f4(dev)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_leg(1, 0)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 tensor-core operands (tf32 for the FSMN 1x1 and mask-net convs), fp32 accumulate/activations",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "items_per_gpu": B, "samples_per_item": T, "weights": "random-init",
                       "l2": "per-step working set ~13 GB of intermediates >> 126 MB L2 (inputs 16 MB); no explicit flush needed",
                       "parallelism": f"dp{world} (independent chunks, no data-path collective); the sharded "
                                      "one-recording split is the strong_c3 record"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": world * B * T * 4, "d2h_bytes_per_step": world * (B * 2 * T * 4 + B * 2 * 4),
                    "how": "pinned host input -> device, stage.run, waveforms + scores -> pinned host, every step; the "
                           "copies run on side streams under the neighbouring steps' kernels, the timed stream waits "
                           "for the last download before the closing event",
                    "ms_per_step_synchronous": ms_e2e_sync,
                    "value_synchronous": world * B * SECONDS / (ms_e2e_sync / 1e3)},
            "e2e_dropin": {"value": dropin_value, "unit": UNIT, "ms_per_step": ms_dropin,
                           "what": "the reference's call granularity: 64 x separate_and_score(np.ndarray) = "
                                   "separate_speaker + 2 embeddings + cosine per recording, numpy in / numpy out, "
                                   "batch 1 per call"},
            "gpu_launches": stage.launches_per_run(B, T) * args.steps,
            "roofline": roof,
            "strong_c3": c3, "score_c4": c4, "stream_c5": c5, "restore_f4": f4,
            "gpu_eager_baseline": eager,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
