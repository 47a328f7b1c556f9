"""Drop-in contract with the reference's training strategy: for every golden case our module tree has the
reference's state_dict keys with the same shapes and dtypes, `freeze_backbones(stage)` leaves exactly the same
parameters trainable, and `all_module_keys` / `trainable_module_keys` are the reference's lists.  The expectations in
tests/golden/state_dict_keys.json were recorded from the unmodified reference (make_golden_keys.py).  CPU only: module
construction runs no kernel."""
import json
import os

import pytest
import torch

from test_oracle_vs_golden import GOLD, build_state_dict, case_cfg

REC = json.load(open(os.path.join(GOLD, "state_dict_keys.json")))


@pytest.mark.parametrize("name", sorted(REC))
def test_state_dict_and_stage_contract(name):
    c = case_cfg(name)
    mla, _ = build_state_dict(c, dtype=torch.float32)
    ref = REC[name]
    ours = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in mla.state_dict().items()}
    missing = sorted(set(ref["state_dict"]) - set(ours))
    extra = sorted(set(ours) - set(ref["state_dict"]))
    assert not missing, ("reference keys we lack", missing[:8])
    assert not extra, ("keys the reference does not have", extra[:8])
    wrong = [(k, ours[k], v) for k, v in ref["state_dict"].items() if ours[k] != v]
    assert not wrong, wrong[:5]
    assert list(mla.all_module_keys) == ref["all_module_keys"]
    for stage, want in ref["stages"].items():
        mla.freeze_backbones(stage)
        got = sorted(k for k, p in mla.named_parameters() if p.requires_grad)
        assert got == want["requires_grad"], (stage, sorted(set(got) ^ set(want["requires_grad"]))[:8])
        assert list(mla.trainable_module_keys) == want["trainable_module_keys"], stage
