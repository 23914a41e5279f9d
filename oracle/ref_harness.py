"""ref_harness.py — TEST INFRASTRUCTURE: imports the UNMODIFIED reference (/root/reference/src) through
the stand-ins in oracle/ref_shims and exposes its merge methods / forwards so that golden vectors can
be generated in the build container (oracle/make_golden.py) and the restatement in oracle/oracle.py
can be pinned against the real thing (tests/test_oracle_vs_reference.py, skipped where /root/reference
is absent, e.g. on the GPU box).  Nothing in the product path may import this module.

Recipe: SURVEY.md Appendix A.
"""
import contextlib
import copy
import os
import sys
import types

import torch

REFERENCE_SRC = os.environ.get("VLM_REFERENCE_SRC", "/root/reference/src")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "vilt"))


_mods = None


def import_reference():
    """Returns a namespace with the reference modules (vilt_module, vision_transformer, config, ...)."""
    global _mods
    if _mods is not None:
        return _mods
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    for p in (REFERENCE_SRC, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import transformers.optimization as topt

    if not hasattr(topt, "AdamW"):  # removed in transformers 5; src/vilt/modules/vilt_utils.py:4 imports it
        topt.AdamW = torch.optim.AdamW
    import vilt.config as ref_config
    import vilt.modules.vilt_module as vilt_module
    import vilt.modules.vision_transformer as vision_transformer
    from vilt.custom_ln.config import LNConfig
    from vilt.moe.config import MOEConfig
    from vilt.ufo.config import UFOConfig

    _mods = types.SimpleNamespace(
        config=ref_config,
        vilt_module=vilt_module,
        vision_transformer=vision_transformer,
        LNConfig=LNConfig,
        MOEConfig=MOEConfig,
        UFOConfig=UFOConfig,
        ViLTransformerSS=vilt_module.ViLTransformerSS,
    )
    return _mods


@contextlib.contextmanager
def legacy_torch_load():
    """torch >= 2.6 defaults torch.load(weights_only=True), which rejects the reference's pickled
    defaultdict Gram file and PL checkpoints (src/vilt/modules/vilt_module.py:275,386,660)."""
    orig = torch.load

    def patched(*args, **kwargs):
        kwargs.setdefault("weights_only", False)
        return orig(*args, **kwargs)

    torch.load = patched
    try:
        yield
    finally:
        torch.load = orig


def make_config(named=(), **overrides):
    """sacred resolution of `with <named...> k=v` (src/vilt/config.py)."""
    ref = import_reference()
    return copy.deepcopy(ref.config.ex.materialize(named, overrides))


def build_model(cfg, seed=1):
    """ViLTransformerSS built exactly as src/run.py:165-185 does."""
    ref = import_reference()
    ln_config = moe_config = ufo_config = None
    if cfg["use_ufo"]:
        ufo_config = ref.UFOConfig()
        ufo_config.separate_inference = cfg["separate_inference"]
    if cfg["use_custom_ln_attn"] or cfg["use_custom_ln_ffn"]:
        ln_config = ref.LNConfig()
        ln_config.use_custom_ln_attn = cfg["use_custom_ln_attn"]
        ln_config.use_custom_ln_ffn = cfg["use_custom_ln_ffn"]
    if cfg["use_moe"]:
        moe_config = ref.MOEConfig()
        moe_config.in_attn = cfg["in_attn"]
        moe_config.in_ffn = cfg["in_ffn"]
        moe_config.self_attn_for_single_mode = cfg["self_attn_for_single_mode"]
        moe_config.separate_inference = cfg["separate_inference"]
    torch.manual_seed(seed)
    with legacy_torch_load(), contextlib.redirect_stdout(open(os.devnull, "w")):
        model = ref.ViLTransformerSS(cfg, ufo_config, ln_config, moe_config)
    model.eval()
    return model


class _Self:
    """The only thing the reference merge methods read from `self`: self.hparams.config[...]."""

    def __init__(self, cfg):
        self.hparams = types.SimpleNamespace(config=cfg)


def ref_merge_weights(state_dict, cfg):
    """ViLTransformerSS.merge_weights (src/vilt/modules/vilt_module.py:533-638), unbound."""
    ref = import_reference()
    return ref.ViLTransformerSS.merge_weights(_Self(cfg), state_dict)


def ref_sum_task_vectors(state_dict, cfg):
    """ViLTransformerSS.sum_task_vectors (:640-746); cfg['central_weight'] is a file path."""
    ref = import_reference()
    with legacy_torch_load():
        return ref.ViLTransformerSS.sum_task_vectors(_Self(cfg), state_dict)


def ref_regmean(state_dict, cfg):
    """ViLTransformerSS.regmean (:366-531); cfg['gram_matrices'] is a file path."""
    ref = import_reference()
    with legacy_torch_load():
        return ref.ViLTransformerSS.regmean(_Self(cfg), state_dict)


def ref_hook_gram_input(store):
    """hook_gram_input is a closure inside main() (src/cache_gram_matrices.py:246-254) and cannot be
    imported; these are its six lines verbatim, bound to `store` (a defaultdict(float))."""

    def hook_gram_input(module, input, output):
        if isinstance(input, tuple):
            input = input[0]

        flatten_input = input.reshape(-1, input.shape[-1]).to(torch.float64)  # (B * L, D)
        gram = torch.matmul(flatten_input.T, flatten_input)

        store[module.module_name] += gram.detach().cpu()

    return hook_gram_input


# src/cache_gram_matrices.py:264-276
REF_ALL_KEYS_MOE = [
    "mlp.fc1", "mlp.fc1",
    "mlp.v.fc1", "mlp.l.fc1", "mlp.vl.fc1", "mlp.v.fc2", "mlp.l.fc2", "mlp.vl.fc2",
    "attn",
    "attn.v", "attn.l", "attn.vl",
    "attn.proj",
    "attn.v.proj", "attn.l.proj", "attn.vl.proj",
]
REF_ALL_KEYS_UFO = ["mlp.fc1", "mlp.fc2", "attn.proj", "norm1", "norm2"]


def ref_register_gram_hooks(model, store, use_moe=True):
    """Registration loop of src/cache_gram_matrices.py:278-281."""
    all_keys = REF_ALL_KEYS_MOE if use_moe else REF_ALL_KEYS_UFO
    hook = ref_hook_gram_input(store)
    handles = []
    for name, module in model.named_modules():
        if any([name.endswith(n) for n in all_keys]) and ".bias" not in name:
            module.module_name = name
            handles.append(module.register_forward_hook(hook))
    return handles
