"""tdz: B200-native (sm_100a) target-speaker separation + scoring stage of TargetDiarization.

Drop-in pieces:
  Separator         <-> AudioProcessor.separater                       (AudioProcessor.py:268-274, 943)
  separate_speaker  <-> AudioProcessor.separate_speaker                (AudioProcessor.py:885-956)
  wav_chunk_inference <-> look2hear.utils.wav_chunk_inference          (look2hear/utils/separator.py:72-132)
  Embedder          <-> TargetASR.embedding['eres2netv2_large']        (TargetASR.py:102-103, 155-163)
  cosine_similarity / pick_target <-> TargetASR.cosine_similarity, :612-625
  pick_mix_audio    <-> TargetASR.mix_audio_processor's choice         (TargetASR.py:734-743)
  asr_audio_streaming_batch <-> TargetDiarizationStream.asr_audio_streaming for S concurrent streams
                                                                       (TargetDiarizationStream.py:189-258)
  Restorer          <-> AudioProcessor.restorer (Apollo)               (AudioProcessor.py:276-281, 970)
  ConvTDFNet        <-> AudioProcessor.mdx_net STFT / iSTFT            (AudioProcessor.py:65-120)
"""
from ._lib import Handle, load  # noqa: F401
from .separator import Separator  # noqa: F401
from .embedder import Embedder  # noqa: F401
from .pipeline import SeparationScoringStage, meter_loudness  # noqa: F401
from .plan import chunk_bounds, ola_plan, pick_mix_audio, pick_target  # noqa: F401
from .streaming import StageEngine, StreamState, asr_audio_streaming_batch  # noqa: F401
from .restorer import Restorer  # noqa: F401
from .mdx import ConvTDFNet  # noqa: F401
