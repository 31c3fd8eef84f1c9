"""Import the REFERENCE's own NavDP / DINOv2 modules from /root/reference (this container only).

TEST INFRASTRUCTURE.  Used by oracle/gen_golden.py to produce tests/golden/*.npz and by
tests/test_oracle_vs_reference.py to pin the restatement in oracle/navdp_oracle.py.  /root/reference does not exist
on the GPU box, so nothing that runs there imports this module.

Three shims:
  1. `internnav.model.encoder` is registered as a bare package so its __init__ (-> bert_backbone -> transformers<5
     internals) is skipped;
  2. `diffusers.schedulers.scheduling_ddpm.DDPMScheduler` is absent from the image: oracle.ddpm.DDPMScheduler (a
     restatement of diffusers==0.33.1, requirements/internvla_n1.txt L3) is injected -- so the DDPM arithmetic (row a10)
     is NOT pinned by the reference, only its call sites (navdp.py L74-76, L173, L247-250) are;
  3. DAT_RGBD_Patch_Backbone.__init__ torch.load()s a checkpoint unconditionally (navdp_backbone.py L124): torch.load
     is patched to return {} during construction (load_state_dict(strict=False) then keeps the random init).
"""
import importlib
import os
import sys
import types

REF = os.environ.get("N1_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "internnav", "model"))


def _bare_package(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


def load_reference_navdp():
    """Returns the reference module internnav.model.basemodel.internvla_n1.navdp (classes untouched)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)
    from . import ddpm

    if "internnav" not in sys.modules:
        _bare_package("internnav", os.path.join(REF, "internnav"))
        _bare_package("internnav.model", os.path.join(REF, "internnav", "model"))
        _bare_package("internnav.model.encoder", os.path.join(REF, "internnav", "model", "encoder"))
        _bare_package("internnav.model.basemodel", os.path.join(REF, "internnav", "model", "basemodel"))
        _bare_package("internnav.model.basemodel.internvla_n1",
                      os.path.join(REF, "internnav", "model", "basemodel", "internvla_n1"))
    if "diffusers" not in sys.modules:
        d = types.ModuleType("diffusers")
        ds = types.ModuleType("diffusers.schedulers")
        dd = types.ModuleType("diffusers.schedulers.scheduling_ddpm")
        dd.DDPMScheduler = ddpm.DDPMScheduler
        ds.scheduling_ddpm = dd
        d.schedulers = ds
        sys.modules.update({"diffusers": d, "diffusers.schedulers": ds, "diffusers.schedulers.scheduling_ddpm": dd})
    tp = os.path.join(REF, "third_party", "diffusion-policy")
    if tp not in sys.path:
        sys.path.insert(0, tp)
    if "diffusion_policy" not in sys.modules and not os.path.isdir(os.path.join(tp, "diffusion_policy")):
        # un-checked-out submodule: the only symbol used is SinusoidalPosEmb, which navdp_backbone's star import
        # shadows anyway
        dp = types.ModuleType("diffusion_policy")
        dpm = types.ModuleType("diffusion_policy.model")
        dpd = types.ModuleType("diffusion_policy.model.diffusion")
        dpp = types.ModuleType("diffusion_policy.model.diffusion.positional_embedding")
        dpp.SinusoidalPosEmb = object
        sys.modules.update({"diffusion_policy": dp, "diffusion_policy.model": dpm,
                            "diffusion_policy.model.diffusion": dpd,
                            "diffusion_policy.model.diffusion.positional_embedding": dpp})
    return importlib.import_module("internnav.model.basemodel.internvla_n1.navdp")


def build_reference_navdp(predict_size=32, memory_size=2, navdp_version=0.1):
    """Construct the reference NavDP_Policy_DPT_CriticSum_DAT in fp32 on CPU (random init)."""
    import torch

    mod = load_reference_navdp()
    real_load = torch.load
    torch.load = lambda *a, **k: {}
    try:
        m = mod.NavDP_Policy_DPT_CriticSum_DAT(memory_size=memory_size, predict_size=predict_size,
                                               navdp_version=navdp_version, input_dtype="fp32", device="cpu")
    finally:
        torch.load = real_load
    m.rgbd_encoder.input_dtype = torch.float32
    m.rgbd_encoder.preprocess_mean = m.rgbd_encoder.preprocess_mean.float()
    m.rgbd_encoder.preprocess_std = m.rgbd_encoder.preprocess_std.float()
    return m.eval()


def load_reference_nextdit():
    """-> (nextdit_traj, nextdit_crossattn_traj): the reference's own DiT classes, importable here only with stand-ins for
    the absent `diffusers` leaf modules (oracle/diffusers_standin.py) -- pins the block / model WIRING, not the leaves."""
    load_reference_navdp()       # registers the bare internnav packages (and the DDPM stand-in)
    from . import diffusers_standin
    diffusers_standin.install()
    a = importlib.import_module("internnav.model.basemodel.internvla_n1.nextdit_traj")
    b = importlib.import_module("internnav.model.basemodel.internvla_n1.nextdit_crossattn_traj")
    return a, b


def build_reference_navdp_policy(memory_size=8, predict_size=24):
    """Construct the reference's stand-alone NavDPNet (internnav/model/basemodel/navdp/navdp_policy.py) in fp32 on the CPU.
    Shims: bare `internnav.configs.*` modules holding two attribute-bag config classes (the real ones are pydantic models the
    policy only reads attributes from), `torch.load` -> {} for the DepthAnything checkpoint, and `torch.device` -> cpu
    while the constructor runs (it hard-codes cuda:<local_rank>)."""
    import torch

    load_reference_navdp()
    for name in ("internnav.configs", "internnav.configs.model", "internnav.configs.model.base_encoders",
                 "internnav.configs.trainer", "internnav.configs.trainer.exp"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m

    class _Bag:
        def __init__(self, **k):
            self.__dict__.update(k)

    sys.modules["internnav.configs.model.base_encoders"].ModelCfg = _Bag
    sys.modules["internnav.configs.trainer.exp"].ExpCfg = _Bag
    if "internnav.model.basemodel.navdp" not in sys.modules:
        _bare_package("internnav.model.basemodel.navdp", os.path.join(REF, "internnav", "model", "basemodel", "navdp"))
    real_load, real_dev = torch.load, torch.device

    class _CpuDevice:
        def __new__(cls, *a, **k):
            return real_dev("cpu")

    torch.load = lambda *a, **k: {}
    try:
        mod = importlib.import_module("internnav.model.basemodel.navdp.navdp_policy")
        il = dict(image_size=224, memory_size=memory_size, predict_size=predict_size, pixel_channel=4, temporal_depth=16,
                  heads=8, channels=3, dropout=0.1, token_dim=384, scratch=False, finetune=False)
        cfg = mod.NavDPModelConfig(model_cfg={"model": {}, "il": il, "local_rank": 0})
        torch.device = _CpuDevice
        try:
            net = mod.NavDPNet(cfg)
        finally:
            torch.device = real_dev
    finally:
        torch.load = real_load
    net._device = real_dev("cpu")
    return net.eval()


def load_reference_vln_utils():
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "_ref_vln_utils", os.path.join(REF, "internnav", "model", "utils", "vln_utils.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def load_reference_rope2d():
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "_ref_rope2d", os.path.join(REF, "internnav", "dataset", "rope2d.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def load_reference_agent(policy_factory, sleep_scale=0.01):
    """Import the reference's own `internnav/agent/internvla_n1_agent.py` (class InternVLAN1Agent, untouched) with the
    heavy / absent imports replaced by shims, so that its step()/reset()/S2-thread state machine can be driven by a
    scripted policy (oracle/gen_golden_agent.py):

      * gym.spaces.Box, imageio -- absent from the image, used only for a class attribute / debug videos;
      * internnav.model.get_policy / get_config -- return `policy_factory` (the scripted stand-in for
        InternVLAN1Net, whose real version needs a checkpoint) and a pass-through config;
      * internnav.model.utils.misc.set_random_seed -- the real one pulls in the logging stack; restated (4 lines);
      * the module's `time` is wrapped so that every sleep is `sleep_scale` times as long (the polling intervals
        0.5 / 0.2 / 0.01 s of L143, L205, L271, L274 only pace the handshake with the S2 thread).
    internnav.agent.base, internnav.configs.agent, internnav.configs.model.base_encoders and
    internnav.model.utils.vln_utils are the reference's own files."""
    import importlib.util
    import time as _time

    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "internnav" or k.startswith("internnav.")}
    for k in saved:
        del sys.modules[k]
    had = {}
    try:
        for name, rel in [("internnav", ""), ("internnav.agent", "agent"), ("internnav.configs", "configs"),
                          ("internnav.configs.model", "configs/model"), ("internnav.model", "model"),
                          ("internnav.model.utils", "model/utils")]:
            _bare_package(name, os.path.join(REF, "internnav", rel))
        gym = types.ModuleType("gym")
        gym.spaces = types.ModuleType("gym.spaces")
        gym.spaces.Box = lambda **kw: ("Box", kw)
        imageio = types.ModuleType("imageio")
        imageio.get_writer = lambda *a, **k: None
        had = {k: sys.modules.get(k) for k in ("gym", "gym.spaces", "imageio")}
        sys.modules.update({"gym": gym, "gym.spaces": gym.spaces, "imageio": imageio})
        misc = types.ModuleType("internnav.model.utils.misc")

        def set_random_seed(seed):  # internnav/model/utils/misc.py L18-22
            import random

            import numpy as np
            import torch
            random.seed(seed)
            np.random.seed(seed)
            torch.manual_seed(seed)
        misc.set_random_seed = set_random_seed
        sys.modules["internnav.model.utils.misc"] = misc
        model_pkg = sys.modules["internnav.model"]
        model_pkg.get_policy = lambda name: policy_factory
        model_pkg.get_config = lambda name: (lambda model_cfg=None: model_cfg)
        mod = importlib.import_module("internnav.agent.internvla_n1_agent")

        class _FastTime:
            def __getattr__(self, k):
                return getattr(_time, k)

            @staticmethod
            def sleep(s):
                _time.sleep(s * sleep_scale)
        mod.time = _FastTime()
        return mod
    finally:
        for k in [k for k in sys.modules if k == "internnav" or k.startswith("internnav.")]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})
        for k, v in had.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def load_reference_policy():
    """Import the reference's `internvla_n1_policy.py` (class InternVLAN1Net, untouched) for driving its HOST logic --
    init_prompts / reset / parse_actions / step_no_infer / s2_step / s1_step_latent -- with scripted collaborators
    (oracle/policy_script.py).  The model module it imports (internvla_n1.py -> diffusers, NextDiT) is replaced by a
    stub holding the two names the policy file references at class-definition time; the returned subclass only adds
    a constructor that skips from_pretrained and a fixed `device`."""
    import importlib

    import torch
    from transformers import PretrainedConfig

    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "internnav" or k.startswith("internnav.")}
    for k in saved:
        del sys.modules[k]
    try:
        for name, rel in [("internnav", ""), ("internnav.model", "model"), ("internnav.model.basemodel", "model/basemodel"),
                          ("internnav.model.basemodel.internvla_n1", "model/basemodel/internvla_n1"),
                          ("internnav.configs", "configs"), ("internnav.configs.model", "configs/model"),
                          ("internnav.model.utils", "model/utils")]:
            _bare_package(name, os.path.join(REF, "internnav", rel))
        stub = types.ModuleType("internnav.model.basemodel.internvla_n1.internvla_n1")

        class InternVLAN1ModelConfig(PretrainedConfig):
            model_type = "internvla_n1_stub"

        stub.InternVLAN1ModelConfig = InternVLAN1ModelConfig
        stub.InternVLAN1ForCausalLM = type("InternVLAN1ForCausalLM", (), {})
        sys.modules[stub.__name__] = stub
        mod = importlib.import_module("internnav.model.basemodel.internvla_n1.internvla_n1_policy")

        class ScriptedNet(mod.InternVLAN1Net):
            def __init__(self, model, processor, num_history=8, resize_w=384, resize_h=384, continuous_traj=True):
                torch.nn.Module.__init__(self)
                self.__dict__["model"] = model
                self.processor = processor
                self.tokenizer = processor.tokenizer
                self.init_prompts()
                self.num_history, self.resize_w, self.resize_h = num_history, resize_w, resize_h
                self.continuous_traj = continuous_traj
                self.rgb_list, self.depth_list, self.pose_list = [], [], []
                self.episode_idx = 0
                self.conversation_history = []
                self.llm_output = ""

            device = property(lambda self: torch.device("cpu"))

        return mod, ScriptedNet
    finally:
        for k in [k for k in sys.modules if k == "internnav" or k.startswith("internnav.")]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})
