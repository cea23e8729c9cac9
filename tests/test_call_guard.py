"""Host-side contract of _lib.CallGuard (SURVEY.md section 8b threading): one call at a time per object, re-entrant on
the owning thread, and a call that arrives on another CUDA stream is ordered behind the previous stream's work.  The
CUDA stream API is replaced by a recorder, so this runs without a GPU; the on-hardware check is
tests/test_gpu_stage.py::test_concurrent_python_threads_share_one_stage."""
import threading
import time

import pytest

from targetdiarization_b200 import _lib


class FakeStream:
    def __init__(self, handle, log):
        self.cuda_stream = handle
        self.log = log

    def wait_stream(self, other):
        self.log.append(("wait", self.cuda_stream, other.cuda_stream))


@pytest.fixture
def fake_cuda(monkeypatch):
    import torch
    state = {"stream": 1, "capturing": False, "log": []}
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: FakeStream(state["stream"], state["log"]))
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: state["capturing"])
    return state


def test_same_stream_records_nothing_and_is_reentrant(fake_cuda):
    g = _lib.CallGuard("cuda:0")
    with g:
        with g:                      # Separator.__call__ inside the stage's graphed run(): same thread, no deadlock
            assert g.depth == 2
        assert g.depth == 1
    with g:
        pass
    assert g.depth == 0 and fake_cuda["log"] == []


def test_new_stream_waits_for_the_previous_one(fake_cuda):
    g = _lib.CallGuard("cuda:0")
    with g:
        pass
    fake_cuda["stream"] = 2
    with g:
        pass
    with g:                          # still stream 2: no further wait
        pass
    fake_cuda["stream"] = 1
    with g:
        pass
    assert fake_cuda["log"] == [("wait", 2, 1), ("wait", 1, 2)]


def test_capture_stream_is_neither_waited_on_nor_remembered(fake_cuda):
    g = _lib.CallGuard("cuda:0")
    with g:
        pass
    fake_cuda["stream"], fake_cuda["capturing"] = 7, True       # torch.cuda.graph switches to its capture stream
    with g:
        pass
    fake_cuda["stream"], fake_cuda["capturing"] = 1, False
    with g:
        pass
    assert fake_cuda["log"] == [] and g.stream.cuda_stream == 1


def test_calls_of_two_threads_do_not_interleave(fake_cuda):
    g = _lib.CallGuard("cuda:0")
    inside, overlaps = [0], [0]

    def work():
        for _ in range(20):
            with g:
                inside[0] += 1
                if inside[0] > 1:
                    overlaps[0] += 1
                time.sleep(0.0005)
                inside[0] -= 1

    threads = [threading.Thread(target=work) for _ in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert overlaps[0] == 0 and g.depth == 0
