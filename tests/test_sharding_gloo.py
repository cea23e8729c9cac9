"""CPU tests of the N>1 path: world_size 2 and 3 over gloo.  The span engines, shard plans and the single
all_gather of targetdiarization_b200/pipeline.py run unchanged; only the CUDA kernels are replaced by the
stand-in of tests/np_kernels.py and the separator by a toy chunk-dependent function."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import stage_port
from targetdiarization_b200 import pipeline, plan
from tests.np_kernels import NumpyKernels


def toy(x):  # [1,T] -> [1,2,T]; depends on the position inside the chunk and on the chunk length
    t = torch.arange(x.shape[-1], dtype=torch.float32) / x.shape[-1]
    return torch.stack((x * (0.5 + t), torch.tanh(3 * x) - 0.2 * t), dim=1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, L, window, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        audio = torch.randn(L, generator=g) * 0.1
        kern = NumpyKernels(toy)
        G = dist.group.WORLD
        # sharding is opt-in: without a group an initialised torch.distributed changes nothing
        alone = pipeline.separate_concat(kern, audio, window=window)
        assert sum(c[0] for c in kern.calls) == len(plan.chunk_bounds(L, window))
        kern.calls.clear()
        kern.uploaded.clear()
        cat = pipeline.separate_concat(kern, audio.numpy(), window=window, group=G)
        n_cat_calls = sum(c[0] for c in kern.calls)
        up_cat = sum(kern.uploaded)
        kern.calls.clear()
        kern.uploaded.clear()
        ola = pipeline.separate_ola(kern, audio.numpy(), sr=1000, group=G)
        n_ola_calls = sum(c[0] for c in kern.calls)
        up_ola = sum(kern.uploaded)
        # gather to one consumer only
        only0 = pipeline.separate_concat(kern, audio.numpy(), window=window, group=G, dst=0)
        assert (only0 is None) == (rank != 0)
        if rank == 0:
            assert torch.equal(only0, cat)
        # ranks that hold different recordings must be told, not silently stitched together
        try:
            pipeline.separate_concat(kern, audio.numpy()[:L - rank], window=window, group=G)
            mismatch = False
        except RuntimeError:
            mismatch = True
        q.put((rank, cat.numpy(), ola.numpy(), n_cat_calls, n_ola_calls, alone.numpy(), up_cat, up_ola, mismatch))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,L", [(2, 30500), (3, 47001), (2, 3000)])
def test_sharded_equals_single_process(world, L):
    window = 4000
    g = torch.Generator().manual_seed(3)
    audio = torch.randn(L, generator=g) * 0.1
    # single-process results (world 1 path of the same functions) and the oracle
    kern = NumpyKernels(toy)
    cat1 = pipeline.separate_concat(kern, audio, window=window).numpy()
    ola1 = pipeline.separate_ola(kern, audio, sr=1000).numpy()
    s1, s2 = stage_port.separate_speaker(audio.numpy(), toy, lambda a: 0.0, window=window)
    assert np.array_equal(cat1, np.stack((s1, s2)))
    want = stage_port.wav_chunk_inference(lambda x: toy(x[:, 0]).unsqueeze(2), audio[None, None], sr=1000)[:, 0]
    assert np.array_equal(ola1, want.numpy())

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, L, window, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_windows = len(plan.chunk_bounds(L, window))
    n_seg = plan.ola_plan(L, 1000).num_session
    cat_calls = ola_calls = up_cat = up_ola = 0
    for rank, cat, ola, nc, no, alone, uc, uo, mismatch in res:
        assert np.array_equal(cat, cat1), f"rank {rank}: concat mode differs from the single-process result"
        assert np.array_equal(ola, ola1), f"rank {rank}: overlap-add differs (must be bit-identical across shardings)"
        assert np.array_equal(alone, cat1), f"rank {rank}: a call without a group must not shard"
        assert mismatch, "inputs of different lengths on the ranks of a sharded call must raise"
        cat_calls += nc
        ola_calls += no
        up_cat += uc
        up_ola += uo
    assert cat_calls == n_windows                       # windows are partitioned, none computed twice
    assert n_seg <= ola_calls <= n_seg + 4 * (world - 1)  # halo: at most 2 recomputed segments per side of a cut
    assert up_cat == L                                  # concat mode: every sample is uploaded by exactly one rank
    halo = plan.ola_plan(L, 1000).pad
    assert L <= up_ola <= L + 2 * halo * (world - 1)    # overlap-add: the span plus at most one halo per side of a cut


# ---------------------------------------------------------------------------------------------- N1: per-segment scores
class _StubSeparator:
    device = torch.device("cpu")
    _ws = None


class _StubEmbedder:
    """score_many stand-in: a deterministic function of each clip (and NaN for clips under 1 680 samples, like the
    reference model); records how many clips this rank was asked to score."""

    def __init__(self):
        self.n_scored = 0

    def score_many(self, wavs, target):
        clips = [wavs[i] for i in range(len(wavs))]
        self.n_scored += len(clips)
        out = []
        for c in clips:
            c = torch.as_tensor(c, dtype=torch.float32)
            out.append(float("nan") if c.numel() < 1680 else float(torch.tanh(c.abs().mean() * 7 + c.numel() * 1e-5)))
        return torch.tensor(out, dtype=torch.float32)


def _segments(n, seed=5):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(400, 9000, (n,), generator=g).tolist()
    return [torch.randn(m, generator=g) * 0.1 for m in lens]


def _score_worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        emb = _StubEmbedder()
        st = pipeline.SeparationScoringStage(_StubSeparator(), emb, group=dist.group.WORLD)
        ragged = st.score_segments(_segments(n), None)
        n_ragged = emb.n_scored
        fixed = st.score_segments(torch.stack([s[:400] for s in _segments(n)]).repeat(1, 5), None)
        q.put((rank, ragged.numpy(), fixed.numpy(), n_ragged))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 11), (3, 7), (3, 2)])
def test_sharded_segment_scores_equal_single_process(world, n):
    emb = _StubEmbedder()
    st = pipeline.SeparationScoringStage(_StubSeparator(), emb)          # no group: scores everything itself
    want = st.score_segments(_segments(n), None).numpy()
    want_fixed = st.score_segments(torch.stack([s[:400] for s in _segments(n)]).repeat(1, 5), None).numpy()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_score_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = 0
    for rank, ragged, fixed, n_ragged in res:
        assert np.array_equal(ragged, want, equal_nan=True), f"rank {rank}: gathered scores differ"
        assert np.array_equal(fixed, want_fixed, equal_nan=True)
        assert n_ragged in (n // world, -(-n // world))      # an even share, nothing scored twice
        total += n_ragged
    assert total == n
    # the dealing is a pure function of the lengths: longest first, round robin
    lens = [int(s.numel()) for s in _segments(n)]
    owners = plan.deal_segments(lens, world)
    assert sorted(i for o in owners for i in o) == list(range(n))
    assert all(lens[owners[0][0]] >= lens[i] for i in range(n))
