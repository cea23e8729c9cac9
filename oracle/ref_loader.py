"""Loads the reference's own MossFormer2 modules from /root/reference by file path.

TEST INFRASTRUCTURE ONLY (oracle).  Works only where /root/reference exists (the build container); it is used
to validate oracle/mossformer2_port.py and to generate tests/golden/*.  Nothing that runs on the GPU box may
import this module.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("TDZ_REFERENCE_ROOT", "/root/reference")
_MODELS = os.path.join(REF_ROOT, "look2hear", "models")


def reference_available():
    return os.path.isfile(os.path.join(_MODELS, "mossformer2.py"))


def load_reference_modules():
    """Returns the synthetic package holding base_model, layer_norm, conv_module, fsmn, mossformer_block,
    mossformer2 loaded from the reference tree (no `import look2hear`, which needs absent packages)."""
    if "l2h" in sys.modules:
        return sys.modules["l2h"]
    shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rotary_shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    pkg = types.ModuleType("l2h")
    pkg.__path__ = [_MODELS]
    sys.modules["l2h"] = pkg
    for name in ("base_model", "layer_norm", "conv_module", "fsmn", "mossformer_block", "mossformer2"):
        spec = importlib.util.spec_from_file_location(f"l2h.{name}", os.path.join(_MODELS, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"l2h.{name}"] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
    return pkg


def build_reference_mossformer2(seed=0):
    import torch

    pkg = load_reference_modules()
    torch.manual_seed(seed)
    return pkg.mossformer2.MossFormer2().eval()


def load_reference_wav_chunk_inference():
    """look2hear/utils/separator.py::wav_chunk_inference with `soundfile` stubbed (absent in this image)."""
    if "soundfile" not in sys.modules:
        sys.modules["soundfile"] = types.ModuleType("soundfile")
    path = os.path.join(REF_ROOT, "look2hear", "utils", "separator.py")
    spec = importlib.util.spec_from_file_location("l2h_separator", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.wav_chunk_inference


def load_reference_apollo():
    """look2hear/models/apollo.py loaded by path (its only relative import is base_model)."""
    pkg = load_reference_modules()
    if hasattr(pkg, "apollo"):
        return pkg.apollo
    spec = importlib.util.spec_from_file_location("l2h.apollo", os.path.join(_MODELS, "apollo.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["l2h.apollo"] = mod
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    pkg.apollo = mod
    return mod


def build_reference_apollo(state_dict=None):
    """Apollo(sr=44100, win=20, feature_dim=256, layer=6) as AudioProcessor.init_restorer_model builds it
    (AudioProcessor.py:279), in eval mode, optionally loaded with `state_dict`."""
    import contextlib
    import io
    mod = load_reference_apollo()
    with contextlib.redirect_stdout(io.StringIO()):   # the constructor prints its band table
        m = mod.Apollo(sr=44100, win=20, feature_dim=256, layer=6).eval()
    if state_dict is not None:
        m.load_state_dict(state_dict)
    return m


def load_reference_conv_tdf_net():
    """The ConvTDFNet class of AudioProcessor.py (:65-120), compiled from its source lines alone: the module itself
    imports packages that are absent here (onnxruntime, librosa, ...)."""
    import ast
    import textwrap
    import torch
    path = os.path.join(REF_ROOT, "AudioProcessor.py")
    src = open(path).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == "ConvTDFNet":
            code = textwrap.dedent("\n".join(src.splitlines()[node.lineno - 1:node.end_lineno]))
            ns = {"torch": torch}
            exec(compile(code, "AudioProcessor.ConvTDFNet", "exec"), ns)
            return ns["ConvTDFNet"]
    raise KeyError("ConvTDFNet not found")
