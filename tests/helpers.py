"""Shared test helpers: golden loading and gap-aware token comparison."""
import json
import os

import numpy as np
import torch

from vitcap_b200 import config as vcfg
from vitcap_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


def golden_setup(meta):
    """Re-creates (cfg, state_dict, data dict, test_extra_input) of a golden case from its recorded seeds."""
    cfg = vcfg.variant(meta["variant"], **meta["cfg_overrides"])
    sd = synth.make_state_dict(cfg, **meta["weights"])
    B = meta["batch"]
    data = synth.make_text_inputs(cfg, B, n_label=meta.get("n_label"))
    data["image"] = synth.make_images(cfg, B, seed=meta["image_seed"])
    extra = synth.default_test_extra_input(cfg, **meta["decode"])
    return cfg, sd, data, extra


def compare_ids_gap_aware(ids, ref_ids, step_top_val, min_gap, what=""):
    """Token IDs must be identical, except that a row may diverge at a decode step where the
    reference's own top-1/top-2 logit gap is below ``min_gap`` (an argmax near-tie that no
    reduction order can be expected to resolve identically). After such a step the row is skipped.
    step_top_val: (steps, rows, >=2) top logits of the reference per step."""
    ids = np.asarray(ids)
    ref_ids = np.asarray(ref_ids)
    assert ids.shape == ref_ids.shape, (ids.shape, ref_ids.shape)
    R = ids.shape[0]
    excused = 0
    for r in range(R):
        a, b = ids[r].reshape(-1), ref_ids[r].reshape(-1)
        if np.array_equal(a, b):
            continue
        t = int(np.nonzero(a != b)[0][0])        # position t was produced at decode step t-1
        gap = float(step_top_val[t - 1, r, 0] - step_top_val[t - 1, r, 1]) if t >= 1 else 1e9
        assert gap < min_gap, "%s row %d diverges at position %d (ref gap %.3g): %s vs %s" % (what, r, t, gap, a, b)
        excused += 1
    return excused
