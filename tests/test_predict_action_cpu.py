"""Host-side plumbing of MLA.predict_action_diff (models/mla/model_mla.py:592-775) on CPU: prompt ids (empty-token /
<BOD> <EOD> suffix appended then stripped, :642-643,:714-715), image + ones mask channel (:661-665), proprio
normalisation (:672-684), action un-normalisation and gripper thresholding (:686-707).  The denoise loop itself (the
CUDA path) is replaced by a stub here; it is tested in tests/test_denoise_gpu.py."""
import types

import numpy as np
import pytest
import torch

from test_oracle_vs_golden import build_state_dict, case_cfg


class _Tok:
    def __init__(self, ids):
        self.ids = ids

    def __call__(self, text, truncation=True, return_tensors="pt"):
        self.text = text
        return types.SimpleNamespace(input_ids=torch.tensor([self.ids], dtype=torch.long))


class _IP:
    def preprocess(self, image, return_tensors="pt"):
        return {"pixel_values": [torch.full((3, 168, 168), 0.25)]}


def _model():
    c = case_cfg("tiny_img")
    mla, _ = build_state_dict(c, dtype=torch.float32)
    mla.norm_stats = {"rlbench": {"action": {"q01": [-1.0] * 7, "q99": [1.0, 2.0, 3.0, 1.0, 1.0, 1.0, 1.0],
                                             "mask": [True] * 6 + [False]},
                                  "proprio": {"q01": [0.0] * 7, "q99": [2.0] * 7}}}
    mla.vlm.vision_tower_2d._image_processor = _IP()
    return mla


def test_predict_action_diff_plumbing():
    mla = _model()
    tok = _Tok([1, 512, 513, 29901])                       # does not end with the empty token 29871
    mla.vlm.llm_backbone.tokenizer = tok
    seen = {}

    def fake_denoise(input_ids, images, point_cloud=None, proprio=None, camera_name=None, num_ddim_steps=8,
                     use_kv_cache=True, **kw):
        seen.update(ids=input_ids.clone(), px=images["front_image"].clone(), proprio=proprio, steps=num_ddim_steps,
                    camera=camera_name, pc=point_cloud)
        return torch.tensor([[[0.5, -0.5, 2.0, 0.0, 0.1, -0.1, 0.7]]])     # normalised chunk [1, T+1, 7]
    mla.denoise_actions = fake_denoise
    out = mla.predict_action_diff(image=object(), pointcloud=np.zeros((64, 3), np.float32), instruction="Close The Jar",
                                  cur_robot_state=np.array([1.0, 0.0, 2.0, 3.0, 1.0, 1.0, 0.5]), unnorm_key="rlbench",
                                  num_ddim_steps=4)
    # prompt: one human turn of the "pure" template, instruction lower-cased (:627-631)
    assert tok.text == "In: What action should the robot take to close the jar?\nOut:"
    # ids + [29871, 32001, 32002, 29871], last three stripped -> ends with the tag token 29871
    assert seen["ids"].tolist() == [[1, 512, 513, 29901, 29871]]
    assert seen["px"].shape == (1, 4, 168, 168) and bool((seen["px"][:, 3] == 1).all()) and bool((seen["px"][:, :3] == 0.25).all())
    assert seen["pc"].shape == (64, 3) and seen["steps"] == 4 and seen["camera"] == "rlbench_front"
    # proprio: 2 (x - lo) / (hi - lo + 1e-8) - 1, clipped to [-1, 1]
    want_p = np.clip(2 * np.array([1.0, 0.0, 2.0, 3.0, 1.0, 1.0, 0.5]) / (2.0 + 1e-8) - 1, -1, 1)
    assert seen["proprio"].shape == (1, 1, 7) and np.allclose(seen["proprio"].numpy().ravel(), want_p, atol=1e-6)
    # actions: clip, gripper bit (index 6) thresholded at 0.5, then 0.5 (a + 1)(hi - lo) + lo where mask else a
    a = np.clip(np.array([0.5, -0.5, 2.0, 0.0, 0.1, -0.1, 0.7]), -1, 1)
    a[6] = 1.0
    hi, lo = np.array([1.0, 2.0, 3.0, 1.0, 1.0, 1.0, 1.0]), -np.ones(7)
    want = 0.5 * (a + 1) * (hi - lo) + lo
    want[6] = a[6]                                         # mask False: left normalised
    assert out.shape == (1, 7) and np.allclose(out[0], want, atol=1e-6)


def test_predict_action_diff_keeps_existing_tag_and_rejects_unbuilt_modes():
    mla = _model()
    mla.vlm.llm_backbone.tokenizer = _Tok([1, 700, 29871, 32001, 32002, 29871])       # already suffixed
    got = {}
    mla.denoise_actions = lambda ids, images, **kw: got.update(ids=ids) or torch.zeros(1, 1, 7)
    mla.predict_action_diff(image=object(), instruction="x", unnorm_key="rlbench")
    assert got["ids"].tolist() == [[1, 700, 29871]]
    with pytest.raises(NotImplementedError):
        mla.predict_action_diff(image=object(), instruction="x", unnorm_key="rlbench", cfg_scale=1.5)
    with pytest.raises(NotImplementedError):
        mla.predict_action_diff(image=object(), instruction="x", unnorm_key="rlbench", use_ddim=False)
    with pytest.raises(AssertionError):
        mla.predict_action_diff(image=object(), instruction="x", unnorm_key="nope")


def test_bench_reference_arm_runs_the_unmodified_reference_on_cpu():
    """`bench.py --impl reference` (the driver's reference arm, also the cpu_baseline leg) drives the UNMODIFIED reference's
    MLA.forward + backward on the host cores and prints the contract's JSON line; here with one decoder layer of the image-only
    workload so the CPU suite stays short (the GPU box runs all 32 layers of the full configuration; its InfoNCE loss reads
    hidden_states[8], so that one needs at least 9).  Skipped where no reference tree is installed."""
    import json
    import os
    import subprocess
    import sys
    import pytest
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("no reference tree (/root/reference or baseline/_ref)")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--layers-cpu", "1", "--workload", "cfg2"], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="0", WORLD_SIZE="1"))
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "multimodal_tokens_per_sec" and line["unit"] == "tokens/s"
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["layers_cpu"] == 1
    assert line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("cfg2")
