"""vitcap_b200: B200-native (sm_100a) implementation of jacobswan1/ViTCAP's batched caption-generation hot path."""
from .config import VARIANTS, VitCapConfig, tiny, variant  # noqa: F401


def __getattr__(name):
    # torch-dependent pieces are imported lazily so that `import vitcap_b200` stays cheap
    if name in ("FastImageCaptioning", "FastViTCAP", "build_from_state_dict"):
        from . import model
        return getattr(model, name)
    raise AttributeError(name)
