from .algos import (GRPOTrainer, PPOTrainer, RAFTTrainer, ReinforceTrainer, RemaxTrainer, RLOOTrainer,
                    SparseGRPOTrainer)
from .base import PolicyAndValueWrapper, RLTrainer
from ..sampler.engine import generate, vllm_generate  # noqa: E402,F401  (reference helper names)
from ..models.qwen2 import forward  # noqa: E402,F401
from ..utils import INVALID_LOGPROB  # noqa: E402,F401
from ..utils.helpers import state_to_device  # noqa: E402,F401
