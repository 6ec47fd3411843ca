"""Target-mask helpers -- mirror of the reference ``utils/target_mask.py`` (create_target_mask 5-104,
select_targets_by_mask 107-125).  Host-side and tiny: the bool vector ``[n_target]`` selects which target
tokens' keys/values the candidate-query rows attend to in the encoder kernels."""
from __future__ import annotations

import random

import torch


def create_target_mask(mask_type, embedding_type, n_target_data, n_target_theta, n_selected_targets=None,
                       predefined_masks=None, predefined_mask_weights=None, mask_index=None, attend_to=None):
    """Bool ``[n_target_data + n_target_theta]``; True = the query tokens attend to that target.

    Same decision table (and the same random sources: ``torch.randperm`` / ``torch.multinomial`` for the
    weighted choice, Python ``random`` for the unweighted ones) as the reference, including its
    fall-through cases: combinations the reference does not handle yield an all-False mask.
    """
    n_total = n_target_data + n_target_theta
    mask = torch.zeros(n_total, dtype=torch.bool)
    if mask_type == "all":
        mask[:] = True
    elif mask_type == "partial":
        if embedding_type in ("data", "theta"):
            mask[torch.randperm(n_total)[:n_selected_targets]] = True
    elif mask_type == "predefined":
        if mask_index is not None:
            chosen = predefined_masks[mask_index]
        elif predefined_mask_weights is not None and len(predefined_mask_weights) == len(predefined_masks):
            w = torch.tensor(predefined_mask_weights, dtype=torch.float)
            chosen = predefined_masks[int(torch.multinomial(w / w.sum(), 1).item())]
        else:
            chosen = random.choice(predefined_masks)
        for i, v in enumerate(chosen):
            if i < n_total and v:
                mask[i] = True
    elif mask_type == "split" and embedding_type == "mix":
        to_data = (attend_to == "data") if attend_to is not None else random.choice([True, False])
        if to_data:
            mask[:n_target_data] = True
        else:
            mask[n_target_data:] = True
    return mask


def select_targets_by_mask(values, mask):
    """values [B, n_target, ...] -> the entries whose mask bit is set, [B, n_selected, ...]."""
    return values[:, mask.to(values.device)]
