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
        cat = pipeline.separate_concat(kern, audio, window=window)
        n_cat_calls = sum(c[0] for c in kern.calls)
        kern.calls.clear()
        ola = pipeline.separate_ola(kern, audio, sr=1000)
        n_ola_calls = sum(c[0] for c in kern.calls)
        q.put((rank, cat.numpy(), ola.numpy(), n_cat_calls, n_ola_calls))
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
    cat_calls = ola_calls = 0
    for rank, cat, ola, nc, no in res:
        assert np.array_equal(cat, cat1), f"rank {rank}: concat mode differs from the single-process result"
        assert np.array_equal(ola, ola1), f"rank {rank}: overlap-add differs (must be bit-identical across shardings)"
        cat_calls += nc
        ola_calls += no
    assert cat_calls == n_windows                       # windows are partitioned, none computed twice
    assert n_seg <= ola_calls <= n_seg + 4 * (world - 1)  # halo: at most 2 recomputed segments per side of a cut
