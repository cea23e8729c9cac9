"""Host-side planning of the separation stage: which windows / segments exist for an input of L samples and
which of them a rank owns.  Pure integer arithmetic (no torch, no CUDA), so that chunk boundaries are bit-exact
with the reference and testable everywhere.

  chunk_bounds            <-> AudioProcessor.separate_speaker window rule      (AudioProcessor.py:892-935)
  ola_plan / OlaPlan      <-> look2hear.utils.wav_chunk_inference segmenting   (look2hear/utils/separator.py:84-112)
  shard_range, *_shard    <-> new: contiguous spans per rank; the reference is single-device
  enrolment_select / enrolment_reduce <-> the scalar rules of TargetASR.get_target_embedding (TargetASR.py:166-258)
"""
from dataclasses import dataclass

WINDOW = 160000           # AudioProcessor.py:896 (10 s at 16 kHz)
WINDOW_LOW_RAM = 16000    # AudioProcessor.py:893 (low_gpu_ram=True; the reference then also runs a VAD)


def chunk_bounds(length, window=WINDOW, start=0):
    """[(begin, end)] of the windows fed to the separator for the sample range [start, start+length).

    round_num = length // window; 0 -> one window [start, start+length); else round_num full windows, and the
    remainder r = length % window is its own window when r > window/2, is appended to the last window when
    0 < r <= window/2 (so a window holds at most 1.5*window samples)."""
    length = int(length)
    n = length // window
    if n == 0:
        return [(start, start + length)]
    bounds = [(start + j * window, start + (j + 1) * window) for j in range(n)]
    rem = length % window
    if rem > 0:
        if rem > window / 2:
            bounds.append((bounds[-1][1], start + length))
        else:
            bounds[-1] = (bounds[-1][0], start + length)
    return bounds


def shard_range(n_units, rank, world):
    """Contiguous unit range [lo, hi) of rank `rank`: lo = floor(rank*n/world)."""
    return (rank * n_units) // world, ((rank + 1) * n_units) // world


def concat_shard(length, rank, world, window=WINDOW):
    """Windows owned by `rank` in concat mode and the sample span they cover: (bounds, span_begin, span_end).
    A rank without windows gets ([], x, x)."""
    bounds = chunk_bounds(length, window)
    lo, hi = shard_range(len(bounds), rank, world)
    mine = bounds[lo:hi]
    if not mine:
        edge = bounds[lo - 1][1] if lo > 0 else 0
        return [], edge, edge
    return mine, mine[0][0], mine[-1][1]


@dataclass(frozen=True)
class OlaPlan:
    length: int        # input samples L
    session: int       # samples per segment (sr * target_length)
    hop: int           # segment hop (sr * hop_length)
    pad: int           # zeros on each side = session - hop
    num_session: int   # (L + 2*pad - session) // hop + 2
    ratio: float       # target_length / hop_length: every sample is covered `ratio` times

    def segment_range(self, i):
        """Input sample range [a, b) (may run outside [0, L): zeros) that segment i reads."""
        a = i * self.hop - self.pad
        return a, a + self.session

    def segments_covering(self, out_begin, out_end):
        """Segment index range [lo, hi) whose outputs are summed into samples [out_begin, out_end)."""
        if out_end <= out_begin:
            return 0, 0
        p_lo = out_begin + self.pad
        p_hi = out_end - 1 + self.pad
        lo = max(0, -(-(p_lo - self.session + 1) // self.hop))
        hi = min(self.num_session - 1, p_hi // self.hop)
        return lo, hi + 1


def ola_plan(length, sr=16000, target_length=12.0, hop_length=4.0):
    session = int(sr * target_length)
    hop = int(sr * hop_length)
    if session - hop <= 0:
        raise ValueError("overlap-add needs target_length > hop_length")
    pad = session - hop
    num_session = (int(length) + 2 * pad - session) // hop + 2
    return OlaPlan(int(length), session, hop, pad, num_session, target_length / hop_length)


def ola_shard(plan, rank, world):
    """Overlap-add mode: rank owns output samples [out_begin, out_end), cut on hop multiples, and recomputes the
    segments that reach into its span from the neighbours (halo, no communication): returns
    (out_begin, out_end, seg_lo, seg_hi)."""
    n_hops = -(-plan.length // plan.hop)
    lo, hi = shard_range(n_hops, rank, world)
    out_begin = min(lo * plan.hop, plan.length)
    out_end = min(hi * plan.hop, plan.length)
    seg_lo, seg_hi = plan.segments_covering(out_begin, out_end)
    return out_begin, out_end, seg_lo, seg_hi


def deal_segments(lengths, world):
    """Which rank scores which segment: indices sorted by length (longest first, ties by index) are dealt round robin,
    so every rank gets the same number of segments (+-1) and the same mix of lengths.  Pure function of the lengths:
    every rank computes the same table.  Returns [indices of rank 0, indices of rank 1, ...]."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    return [order[r::world] for r in range(world)]


def pick_target(spk1_score, spk2_score, threshold=0.0):
    """Target/non-target assignment of one separated pair (TargetASR.py:612-625, 541-553): None when both scores
    are below `threshold`; 1 iff spk1_score > spk2_score (strict), else 2."""
    if spk1_score < threshold and spk2_score < threshold:
        return None
    return 1 if spk1_score > spk2_score else 2


def pick_mix_audio(spk1_score, spk2_score, similarity_threshold=0.4):
    """Which audio TargetASR.mix_audio_processor returns for a two-speaker clip (TargetASR.py:734-743): 0 = the
    unseparated input when both scores are below the threshold, 1 = spk1 iff spk1_score >= spk2_score (ties go to
    spk1 - unlike pick_target's strict >), 2 = spk2 iff spk2_score > spk1_score, else 0 (reached with NaN scores,
    for which every comparison is false).  The reported score is round(max(spk1_score, spk2_score), 3)."""
    if spk1_score < similarity_threshold and spk2_score < similarity_threshold:
        return 0
    if spk1_score >= spk2_score:
        return 1
    if spk2_score > spk1_score:
        return 2
    return 0


# ---------------------------------------------------------------------------------------------- per-segment rules
# The scalar rules the reference applies to per-segment cosine scores.  The scores come from one batched
# Embedder.score_many call; the rules stay on the host exactly as in the reference.
def target_spk_from_scores(speakers, scores):
    """TargetDiarization.target_embedding_to_target_spk (TargetDiarization.py:581-600): mean score per speaker in
    first-appearance order, stable sort descending, first wins; '' when there are no segments."""
    sums, counts, order = {}, {}, []
    for spk, sc in zip(speakers, scores):
        if spk not in sums:
            sums[spk], counts[spk] = 0.0, 0
            order.append(spk)
        sums[spk] += float(sc)
        counts[spk] += 1
    score_map = [[spk, sums[spk] / counts[spk]] for spk in order]
    if not score_map:
        return ""
    score_map.sort(key=lambda x: x[1], reverse=True)
    return score_map[0][0]


def recheck_target_speaker(result, scores, target_spk, threshold, method="recheck_target"):
    """TargetDiarization.recheck_target_speaker (TargetDiarization.py:603-629) given the batched scores of the
    clips (scores[i] for result[i]; None where the clip has no audio): relabels in place and returns `result`."""
    if not result:
        return []
    for r in result:
        r["score"] = -1.0
    if not threshold or threshold == 0.0:
        return result
    for r, sc in zip(result, scores):
        if method == "recheck_target" and r["speaker"] != target_spk:
            continue
        if method == "recheck_others" and r["speaker"] == target_spk:
            continue
        if sc is None:
            continue
        r["score"] = round(float(sc), 3)
        if sc >= threshold:
            if r["speaker"] != target_spk:
                r["speaker"] = target_spk
        elif r["speaker"] == target_spk:
            r["speaker"] = "-1"
    return result


def is_same_person(similarity, threshold=0.4, verbose_result=False):
    """Decision part of TargetASR.is_same_person (TargetASR.py:491-505); `similarity` is the cosine between the mean
    of the existing embeddings and the target embedding."""
    same = similarity >= threshold
    return {"is_same": bool(same), "score": round(float(similarity), 3)} if verbose_result else bool(same)


# ---------------------------------------------------------------------------------------------- enrolment rules
# TargetASR.get_target_embedding (TargetASR.py:166-258) without its model calls: which pieces of the (already
# VAD-cut, loudness-normalised) enrolment audio are embedded, and how the embeddings are reduced.  The embeddings
# themselves come from ONE batched Embedder.embed_many call (pipeline.SeparationScoringStage.get_target_embedding).
def enrolment_select(lengths, sampling_rate=16000, audio_input_type="separate"):
    """Which pieces are embedded (TargetASR.py:207-227).  `lengths[i]` = samples of piece i.  Returns
    (mode, [(source, n_samples)]) with source = piece index, or -1 for the concatenation of ALL pieces (mode
    "merge"); n_samples is the 30 s truncation.  Pieces shorter than 400 samples are dropped (:232-233)."""
    lengths = [int(n) for n in lengths]
    if not lengths:
        return audio_input_type, []
    longest = max(range(len(lengths)), key=lambda i: lengths[i])   # first maximum, as max(key=...) does
    normal = [i for i, n in enumerate(lengths) if n >= int(sampling_rate * 0.4)]
    mode = audio_input_type
    if mode == "auto":
        if lengths[longest] >= 3.0 * sampling_rate:
            mode = "longest"
        elif len(normal) <= 2:
            mode = "merge"
        else:
            mode = "separate"
    if mode == "merge":
        picks = [(-1, sum(lengths))]
    elif mode == "longest":
        picks = [(longest, lengths[longest])]
    else:
        picks = [(i, lengths[i]) for i in normal]
    cap = 30 * sampling_rate
    return mode, [(src, min(n, cap)) for src, n in picks if min(n, cap) >= 400]


def enrolment_reduce(embeddings, is_cluster=True, cluster_labels=None, output_embedding_list=True, dim=192):
    """NaN filter, optional outlier drop by cluster label (-1 = noise; applied only to more than two embeddings and
    only if something survives) and the final mean (TargetASR.py:235-258).  `cluster_labels(emb [n,dim]) -> [n]`
    is the clusterer (the reference: hdbscan.HDBSCAN(min_cluster_size=2, metric="euclidean").fit_predict)."""
    import numpy as np
    kept = [np.asarray(e, dtype=np.float32).reshape(-1) for e in embeddings]
    kept = [e for e in kept if not np.isnan(e).any()]
    if is_cluster and len(kept) > 2:
        if cluster_labels is None:
            raise ValueError("is_cluster=True needs a cluster_labels callable")
        labels = np.asarray(cluster_labels(np.stack(kept)))
        valid = np.where(labels != -1)[0]
        if len(valid) > 0:
            kept = [kept[i] for i in valid]
    if output_embedding_list:
        return kept
    if len(kept) == 0:
        return np.zeros([dim], dtype=np.float32)
    if len(kept) == 1:
        return kept[0]
    return np.mean(kept, axis=0)
