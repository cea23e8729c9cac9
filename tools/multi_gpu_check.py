"""Multi-GPU check of the sharded stage (run under torchrun, one rank per GPU, NCCL):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/multi_gpu_check.py [seconds]

Every rank separates its share of the windows of one long synthetic input (concat mode and overlap-add mode), the
spans are gathered with one all_gather, and rank 0 compares the result bit for bit with the same input processed by
a single rank (the sharding must not change a single output bit); the per-segment scores are checked the same way
(segments dealt to the ranks, one all_gather).  Prints timings as one JSON line; exit code 1 on any mismatch."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from targetdiarization_b200 import SeparationScoringStage  # noqa: E402
from targetdiarization_b200 import pipeline  # noqa: E402
from targetdiarization_b200.synth import synthetic_mixture  # noqa: E402


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stage = SeparationScoringStage.random_init(dev, seed=0)
    L = int(seconds * 16000) + 12345
    audio = synthetic_mixture(1, L, seed=77)[0].to(dev)
    out = {}
    for mode in ("concat", "ola"):
        fn = pipeline.separate_concat if mode == "concat" else pipeline.separate_ola
        fn(stage.kern, audio, group=None)  # warm-up (allocations, first-use attributes); group=None never shards
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        sharded = fn(stage.kern, audio, group=dist.group.WORLD)   # sharding is opt-in: the ranks that share the input
        torch.cuda.synchronize()
        dist.barrier()
        dt = time.perf_counter() - t0
        # single-rank result on rank 0 only (the plan functions take world from the group; emulate world 1)
        ok = None
        if rank == 0:
            if mode == "concat":
                from targetdiarization_b200 import plan
                single = pipeline.concat_span(stage.kern, audio, plan.chunk_bounds(L), 0, L)
            else:
                from targetdiarization_b200 import plan
                p = plan.ola_plan(L)
                single = pipeline.ola_span(stage.kern, audio, p, 0, L, 0, p.num_session)
            torch.cuda.synchronize()
            ok = bool(torch.equal(single, sharded))
        out[mode] = dict(seconds=dt, xrt=L / 16000 / dt, bit_identical_to_single_rank=ok)
    # per-segment scores: segments dealt to the ranks, one all_gather of the [n_seg] scores
    seg = 64000
    clips = audio[:(L // seg) * seg].view(-1, seg)
    target = stage.embed(synthetic_mixture(1, seg, seed=99).to(dev))[0]
    stage.group = dist.group.WORLD
    sharded_scores = stage.score_segments(clips, target)
    stage.group = None
    single_scores = stage.score_segments(clips, target)
    torch.cuda.synchronize()
    out["scores"] = dict(segments=int(clips.shape[0]),
                         bit_identical_to_single_rank=bool(torch.equal(sharded_scores, single_scores)))
    ok = torch.tensor([int(all(v["bit_identical_to_single_rank"] is not False for v in out.values()))], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps(dict(world=world, audio_seconds=L / 16000, **out)))
    if int(ok.item()) != 1:
        dist.destroy_process_group()
        sys.exit(1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
