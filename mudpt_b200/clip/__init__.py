"""`clip` package surface used by the MuDPT hot path (reference: clip/__init__.py, clip/clip.py)."""
from .model import CLIP, build_model  # noqa: F401
from ..synthetic import synthetic_tokenize as tokenize  # noqa: F401  (BPE is out of scope; see synthetic.py)
