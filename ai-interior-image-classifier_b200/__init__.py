"""B200-native (sm_100a) LoRA-ViT image encoder + label-scoring head behind the reference's own entry points.

The directory name follows the build contract (`ai-interior-image-classifier_b200`); because of the hyphens import
it with `importlib.import_module("ai-interior-image-classifier_b200")` or through the `iic_b200` alias module at
the repo root.  Public surface (mirrors what /root/reference/main.py uses from `clip` and defines itself):

    clip-like:   load, tokenize, available_models               (clip_compat.py)
    LoRA:        LoRALayer, LoRALinear, replace_linears_with_lora, save_lora_weights, load_lora_weights_to_model
    training:    VisionLoRATrainer, TextLoRATrainer               (train.py)
    analyzers:   InteriorImageDetector, CachedInteriorAnalyzer, DatabaseStyleRoomAnalyzer   (analyzer.py)
    engine:      Engine, VisionArch, VIT_B_16, VIT_L_14_336      (engine.py; the ctypes binding is _lib.py)
"""
from . import _lib
from .engine import Engine, HeadResult, VisionArch, VIT_B_16, VIT_L_14_336

__all__ = ["_lib", "Engine", "HeadResult", "VisionArch", "VIT_B_16", "VIT_L_14_336"]


def __getattr__(name):
    # heavier host-side mirrors are imported lazily so `import iic_b200` stays cheap
    import importlib
    for mod in ("clip_compat", "lora", "analyzer", "train"):
        try:
            m = importlib.import_module(f"{__name__}.{mod}")
        except ModuleNotFoundError:
            continue
        if hasattr(m, name):
            return getattr(m, name)
    raise AttributeError(name)
