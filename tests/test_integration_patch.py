"""integration/n1b200-dropin.patch applies to the reference tree, the patched files compile, and the PATCHED reference
`InternVLAN1Net.s2_step` -- now one `generate_with_latents` call -- reproduces the traces recorded from the unpatched
one (tests/golden/policy_traces.json).  Needs /root/reference and the `patch` tool; skipped elsewhere."""
import json
import os
import shutil
import subprocess
import sys

import pytest
import torch

from oracle import agent_script, policy_script, ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATCH = os.path.join(ROOT, "integration", "n1b200-dropin.patch")
SUB = os.path.join("internnav", "model", "basemodel", "internvla_n1")

pytestmark = pytest.mark.skipif(not ref_loader.available() or shutil.which("patch") is None,
                                reason="reference tree or patch(1) not available")


@pytest.fixture(scope="module")
def patched_root(tmp_path_factory):
    root = tmp_path_factory.mktemp("internnav_patched")
    os.makedirs(root / SUB)
    for f in ("internvla_n1_arch.py", "internvla_n1.py", "internvla_n1_policy.py"):
        shutil.copy(os.path.join(ref_loader.REF, SUB, f), root / SUB / f)
    for rel in (os.path.join("internnav", "model", "utils"), os.path.join("internnav", "configs")):
        os.makedirs(os.path.dirname(root / rel), exist_ok=True)
        os.symlink(os.path.join(ref_loader.REF, rel), root / rel)
    r = subprocess.run(["patch", "-p1", "--no-backup-if-mismatch", "-i", PATCH], cwd=root, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return root


def test_patch_applies_and_compiles(patched_root):
    for f in ("internvla_n1_arch.py", "internvla_n1.py", "internvla_n1_policy.py"):
        subprocess.run([sys.executable, "-m", "py_compile", str(patched_root / SUB / f)], check=True)
    src = open(patched_root / SUB / "internvla_n1_arch.py").read()
    assert "from internnav_b200.navdp import NavDP_Policy_DPT_CriticSum_DAT" in src
    pol = open(patched_root / SUB / "internvla_n1_policy.py").read()
    assert "generate_with_latents" in pol and "self.model.generate(" not in pol


class OneCallLLM(policy_script.ScriptedLLM):
    """The patched policy's single collaborator call."""

    def generate_with_latents(self, input_ids, pixel_values, image_grid_thw, max_new_tokens=128):
        ans = self.answers[self.n % len(self.answers)]
        self.n += 1
        self.log.append(["generate_with_latents", int(input_ids.shape[1]), int(pixel_values.shape[0]),
                         [int(v) for v in image_grid_thw.reshape(-1)], max_new_tokens])
        return [policy_script.encode(ans) + [151645]], torch.tensor([float(self.n)])


def test_patched_policy_replays_reference_traces(patched_root):
    with open(os.path.join(ROOT, "tests", "golden", "policy_traces.json"), encoding="utf-8") as fh:
        traces = json.load(fh)["traces"]
    saved = ref_loader.REF
    ref_loader.REF = str(patched_root)
    try:
        _, Net = ref_loader.load_reference_policy()
    finally:
        ref_loader.REF = saved
    import contextlib
    import io
    for tr in traces:
        proc, llm = policy_script.FakeProcessor(), OneCallLLM(tr["answers"], tr["trajs"])
        net = Net(llm, proc, num_history=tr["case"]["num_history"])
        net.reset()
        with contextlib.redirect_stdout(io.StringIO()):
            for op, rec in zip(tr["ops"], tr["records"]):
                if op[0] == "reset":
                    net.reset()
                elif op[0] == "noinfer":
                    o = agent_script.make_obs(op[1], size=(24, 32))
                    net.step_no_infer(o["rgb"], o["depth"], None)
                elif op[0] == "s2":
                    o = agent_script.make_obs(op[1], size=(24, 32))
                    out = net.s2_step(o["rgb"], o["depth"], None, o["instruction"], None, look_down=op[2])
                    assert proc.log.pop() == rec["processor"]
                    assert net.llm_output == rec["llm_output"]
                    assert (None if out.output_pixel is None else [int(v) for v in out.output_pixel]) == rec["pixel"]
                    assert out.output_action == rec["actions"] and (out.output_latent is not None) == rec["has_latent"]
                    call = llm.log[-1]
                    assert call[0] == "generate_with_latents" and call[1] == rec["model"][0][1]   # same prompt length
                    if rec["has_latent"]:
                        assert call[2:4] == rec["model"][1][2:4]                                  # same image patches
