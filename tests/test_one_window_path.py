"""Host logic of SeparationScoringStage._separate_and_score_one_window (the one-upload / one-run / one-download form of
separate_and_score, TargetASR.py:609-625 on a recording the chunk rule keeps as one window): when it applies, and that
the louder-first swap (AudioProcessor.py:949-952) permutes the streams AND their scores together.  The kernels are
replaced by stand-ins; the bit-level equality with the general path is checked on the GPU (tests/test_gpu_stage.py,
tests/test_gpu_configs.py)."""
import numpy as np
import torch

from targetdiarization_b200 import plan as P
from targetdiarization_b200.pipeline import SeparationScoringStage


class _Kern:
    def to_device(self, a):
        return torch.as_tensor(np.asarray(a, dtype=np.float32))

    def to_host(self, t):
        return t.numpy()


def _stage(loud=(-20.0, -30.0)):
    st = SeparationScoringStage.__new__(SeparationScoringStage)
    st.is_separate_audio, st.group, st.kern = True, None, _Kern()
    st.calls = []

    def run(mix, target):
        st.calls.append(tuple(mix.shape))
        est = torch.stack((mix[0] * 0.5, mix[0] * 2.0))[None]          # stream 0 quiet, stream 1 loud
        return est, torch.tensor([[0.25, 0.75]])
    st.run = run
    st.meter_loudness_device = lambda est, sr: list(loud)
    return st


def test_applies_only_to_one_plain_16k_window():
    st = _stage()
    audio = np.ones(32000, np.float32)
    assert st._separate_and_score_one_window(audio, None) is not None
    assert st._separate_and_score_one_window(audio, None, sampling_rate=8000) is None
    assert st._separate_and_score_one_window(audio, None, low_gpu_ram=True) is None
    assert st._separate_and_score_one_window(audio, None, mode="ola") is None
    assert st._separate_and_score_one_window(audio, None, vad_frames=[[0, 100]]) is None
    assert st._separate_and_score_one_window(audio, None, loudness="lufs") is None       # unknown meter: general path raises
    assert st._separate_and_score_one_window(list(audio), None) is None
    assert st._separate_and_score_one_window(np.ones(6399, np.float32), None) is None    # under 0.4 s
    two = np.ones(P.WINDOW + P.WINDOW // 2 + 1, np.float32)
    assert len(P.chunk_bounds(two.size, P.WINDOW)) == 2
    assert st._separate_and_score_one_window(two, None) is None
    longest = np.ones(P.WINDOW + P.WINDOW // 2, np.float32)                              # still ONE window of 1.5 x
    assert st._separate_and_score_one_window(longest, None) is not None
    st.group = object()
    assert st._separate_and_score_one_window(audio, None) is None
    st.group, st.is_separate_audio = None, False
    assert st._separate_and_score_one_window(audio, None) is None


def test_swap_moves_streams_and_scores_together():
    audio = np.linspace(-1, 1, 8000, dtype=np.float32)
    for loudness, loud, swapped in (("device", (-30.0, -20.0), True), ("device", (-20.0, -30.0), False),
                                    ("device", (-20.0, -20.0), False), (None, (0.0, 0.0), False),
                                    (lambda a, sr: float(np.abs(a).max()), (0.0, 0.0), True)):
        st = _stage(loud)
        spk1, spk2, sc = st._separate_and_score_one_window(audio, None, loudness=loudness)
        assert st.calls == [(1, 8000)]
        if swapped:
            assert np.array_equal(spk1, audio * 2.0) and np.array_equal(spk2, audio * 0.5) and sc == [0.75, 0.25]
        else:
            assert np.array_equal(spk1, audio * 0.5) and np.array_equal(spk2, audio * 2.0) and sc == [0.25, 0.75]
