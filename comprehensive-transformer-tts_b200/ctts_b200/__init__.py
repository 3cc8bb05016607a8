"""B200-native CompTransTTS acoustic-model forward path."""
from .configs import builtin_configs  # noqa: F401
from .module import CompTransTTS  # noqa: F401
