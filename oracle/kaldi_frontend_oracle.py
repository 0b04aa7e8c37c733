"""CPU oracle of the feature front end the reference puts in front of extract_embedding.py:

    apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=300 scp:feats.scp ark:- |
    select-voiced-frames ark:- scp,s,cs:vad.scp ark:- |          (reference local/tf/extract_xvectors.sh:68)

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the
product path (x-vector-kaldi-tf_b200/ computes this on the GPU: csrc/frontend.cuh).

**Parity unpinned.**  The two programs are Kaldi binaries (third-party dependency of the reference, version unpinned:
the recipe only sources a Kaldi `path.sh`); Kaldi is neither under /root/reference nor in this image, and the reference
ships no fixtures for them.  What is restated here is Kaldi's published algorithm, anchored on the reference's call
site above (the options it passes) :

* ``sliding_window_cmn``  follows ``SlidingWindowCmnInternal`` (kaldi src/feat/feature-functions.cc): input widened to
  double, ONE running sum (and sum of squares) that is updated by subtracting the frame that left the window and adding
  the frame that entered, ``out = x + (-1/N) * sum`` in double, optional variance normalisation with the 1e-10 floor,
  result narrowed to float.  Window placement: centred windows are shifted (not shrunk) at the utterance edges;
  utterances shorter than the window use all their frames.
* ``select_voiced_frames`` follows kaldi src/ivectorbin/select-voiced-frames.cc: rows whose VAD value is non-zero are
  kept in order; a length mismatch or an utterance without voiced frames is skipped (returns None).

``sliding_window_cmn_direct`` is a second, independent restatement (window sums from prefix sums instead of a running
sum) used to de-risk the first one; ``kaldi_frontend_oracle.c`` (bound by ``kaldi_frontend_c.py``) is a third, in plain C,
bit-identical to the first and fast enough to serve as bench.py's CPU baseline for this block.
"""
import numpy as np


def window_bounds(t, num_frames, cmn_window=300, center=True, min_window=100):
    """[window_start, window_end) of frame t: the placement rules of SlidingWindowCmnInternal."""
    if center:
        window_start = t - cmn_window // 2
        window_end = window_start + cmn_window
    else:
        window_start = t - cmn_window
        window_end = t + 1
    if window_start < 0:                       # shift the window right if it starts before the utterance
        window_end -= window_start
        window_start = 0
    if not center:
        if window_end > t:
            window_end = max(t + 1, min_window)
    if window_end > num_frames:                # shift it left if it ends after the utterance
        window_start -= window_end - num_frames
        window_end = num_frames
        if window_start < 0:
            window_start = 0
    return window_start, window_end


def sliding_window_cmn(feats, cmn_window=300, center=True, normalize_variance=False, min_window=100):
    """float32 [T, D] -> float32 [T, D]; the running-sum recursion of the Kaldi function, in double."""
    x = np.asarray(feats, dtype=np.float32).astype(np.float64)
    num_frames, dim = x.shape
    out = np.empty_like(x)
    cur_sum = np.zeros(dim)
    cur_sumsq = np.zeros(dim)
    last_start = last_end = -1
    for t in range(num_frames):
        ws, we = window_bounds(t, num_frames, cmn_window, center, min_window)
        if last_start == -1:
            cur_sum = x[ws:we].sum(axis=0)
            cur_sumsq = (x[ws:we] ** 2).sum(axis=0)
        else:
            if ws > last_start:
                assert ws == last_start + 1
                cur_sum = cur_sum - x[last_start]
                cur_sumsq = cur_sumsq - x[last_start] ** 2
            if we > last_end:
                assert we == last_end + 1
                cur_sum = cur_sum + x[last_end]
                cur_sumsq = cur_sumsq + x[last_end] ** 2
        n = we - ws
        last_start, last_end = ws, we
        assert n > 0
        row = x[t] + (-1.0 / n) * cur_sum
        if normalize_variance:
            if n == 1:
                row = np.zeros(dim)
            else:
                variance = cur_sumsq * (1.0 / n) + (-1.0 / (n * n)) * cur_sum ** 2
                variance = np.maximum(variance, 1.0e-10)
                row = row * variance ** -0.5
        out[t] = row
    return out.astype(np.float32)


def sliding_window_cmn_direct(feats, cmn_window=300, center=True, normalize_variance=False, min_window=100):
    """Independent restatement: every window's sums taken from prefix sums (no recursion over t)."""
    x = np.asarray(feats, dtype=np.float32).astype(np.float64)
    num_frames, dim = x.shape
    if num_frames == 0:
        return np.zeros((0, dim), np.float32)
    p1 = np.concatenate([np.zeros((1, dim)), np.cumsum(x, axis=0)])
    p2 = np.concatenate([np.zeros((1, dim)), np.cumsum(x * x, axis=0)])
    bounds = np.array([window_bounds(t, num_frames, cmn_window, center, min_window) for t in range(num_frames)])
    ws, we = bounds[:, 0], bounds[:, 1]
    n = (we - ws).astype(np.float64)[:, None]
    s1 = p1[we] - p1[ws]
    out = x - s1 / n
    if normalize_variance:
        var = np.maximum((p2[we] - p2[ws]) / n - (s1 / n) ** 2, 1.0e-10)
        out = np.where(n == 1, 0.0, out / np.sqrt(var))
    return out.astype(np.float32)


def select_voiced_frames(feats, vad):
    """Rows of `feats` whose VAD decision is non-zero, or None when the utterance is skipped (length mismatch / no
    voiced frame), as select-voiced-frames does."""
    feats = np.asarray(feats)
    vad = np.asarray(vad)
    if feats.shape[0] != vad.shape[0]:
        return None
    keep = vad != 0.0
    if not keep.any():
        return None
    return feats[keep]


def frontend(feats, vad, cmn_window=300, center=True, normalize_variance=False, min_window=100):
    """The whole pipe of extract_xvectors.sh:68 for one utterance (None = skipped)."""
    cm = sliding_window_cmn(feats, cmn_window, center, normalize_variance, min_window)
    return cm if vad is None else select_voiced_frames(cm, vad)


def synthetic_vad(rng, num_frames, voiced_fraction=0.75, mean_run=40):
    """A 0/1 float32 VAD track made of alternating runs (what compute-vad's energy decisions look like)."""
    out = np.zeros(num_frames, np.float32)
    t = 0
    state = rng.random() < voiced_fraction
    while t < num_frames:
        mean = mean_run * (voiced_fraction if state else (1.0 - voiced_fraction)) * 2.0
        run = 1 + int(rng.exponential(max(mean, 1.0)))
        out[t:t + run] = 1.0 if state else 0.0
        t += run
        state = not state
    return out
