"""Drop-in for the reference's `model` package on the acoustic-model path: `from model import CompTransTTS`
(train.py:17, utils/model.py:7).  Loss / optimizer stay the reference's own."""
from ..module import CompTransTTS  # noqa: F401
