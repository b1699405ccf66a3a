"""Target selection: which (frame, token) outputs a clip is explained for (host side, once per clip)."""
from __future__ import annotations

import numpy as np

PAD_ID = 0     # CTC blank (<pad>), shap_calculation.py:221-254 / visualization.py:320-321
SPACE_ID = 4   # "|"


def char_targets(logits, pad_id: int = PAD_ID, space_id: int = SPACE_ID, all_frames_if_empty: bool = True):
    """Per-character frames (visualization.py:319-327): greedy argmax, keep the frames where a new
    non-blank, non-'|' token starts; the explained token is that frame's unmasked argmax
    (feasability_tests/w2v2conformer.py:97-108).  With random-init weights the set can be empty; then
    every frame with its argmax token is used."""
    ids = np.asarray(logits).argmax(-1)
    keep = (ids != pad_id) & (ids != space_id)
    keep[1:] &= ids[1:] != ids[:-1]
    frames = np.nonzero(keep)[0].astype(np.int32)
    if frames.size == 0 and all_frames_if_empty:
        frames = np.arange(len(ids), dtype=np.int32)
    return frames, ids[frames].astype(np.int32)


def first_char_target(logits, special_ids=(0, 1, 2, 3), space_id: int = SPACE_ID):
    """feasability_tests/w2v2conformer.py:93-110: the first non-special, non-'|' frame, else the middle."""
    ids = np.asarray(logits).argmax(-1)
    for i, t in enumerate(ids):
        if int(t) not in special_ids and int(t) != space_id:
            return int(i), int(t)
    mid = len(ids) // 2
    return int(mid), int(ids[mid])
