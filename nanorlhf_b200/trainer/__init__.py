from .algos import (GRPOTrainer, PPOTrainer, RAFTTrainer, ReinforceTrainer, RemaxTrainer, RLOOTrainer,
                    SparseGRPOTrainer)
from .base import PolicyAndValueWrapper, RLTrainer
