"""Long-input benchmark (BASELINE.json config 3): one synthetic 16 kHz conversation through
SeparationScoringStage.separate_speaker, concat mode (the reference's 10 s windows) and overlap-add mode (12 s / 4 s
hop), from a host numpy array to host numpy arrays, louder stream first.  Under torchrun the windows are sharded
across the ranks and gathered with one all_gather.

  python tools/bench_long.py [minutes]                                        (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 \
      tools/bench_long.py [minutes]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from targetdiarization_b200 import SeparationScoringStage, meter_loudness  # noqa: E402
from targetdiarization_b200.synth import synthetic_mixture  # noqa: E402


def main():
    minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stage = SeparationScoringStage.random_init(dev, seed=0)
    L = int(minutes * 60 * 16000)
    # one hour of synthetic conversation: tile a 64 s synthetic mixture with per-tile gains (cheap to generate)
    base = synthetic_mixture(1, 64 * 16000, seed=11)[0].numpy()
    reps = -(-L // base.shape[0])
    g = np.random.default_rng(5).uniform(0.5, 1.0, size=reps).astype(np.float32)
    audio = (np.tile(base, reps).reshape(reps, -1) * g[:, None]).reshape(-1)[:L].copy()
    out = {"audio_seconds": L / 16000, "world": world}
    for mode in ("concat", "ola"):
        res = {}
        for what, loud in (("no_loudness", None), ("device_loudness", "device"), ("host_loudness", meter_loudness)):
            stage.separate_speaker(audio[: 16000 * 30], mode=mode, loudness=None)   # warm-up
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            s1, s2 = stage.separate_speaker(audio, mode=mode, loudness=loud)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            dt = time.perf_counter() - t0
            res[what] = dict(seconds=dt, xrt=L / 16000 / dt)
        out[mode] = res
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
