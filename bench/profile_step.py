"""Kernel-level time breakdown of one GRPO update (torch.profiler, CUDA activities) -> gpurun_out/profile_*.txt.
Not a benchmark (profiler overhead); used to decide what to optimise next."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = [sys.argv[0]] + [a for a in sys.argv[1:]]

from dataclasses import dataclass  # noqa: E402

from nanorlhf_b200.config import RLConfig  # noqa: E402
from nanorlhf_b200.models.deberta_v3 import DebertaV3Config, DebertaV3ForSequenceClassification  # noqa: E402
from nanorlhf_b200.models.lora import LoraConfig, get_peft_model  # noqa: E402
from nanorlhf_b200.models.qwen2 import Qwen2Config, Qwen2ForCausalLM  # noqa: E402
from nanorlhf_b200.reward.model_reward import ModelReward  # noqa: E402
from nanorlhf_b200.trainer import GRPOTrainer  # noqa: E402
from nanorlhf_b200.utils.data import synthetic_token_dataset  # noqa: E402
from nanorlhf_b200.utils.tokenizer import ByteTokenizer  # noqa: E402

resp_len = int(os.environ.get("RESP", "256"))
mini = int(os.environ.get("MINI", "1"))
dev = torch.device("cuda")
shape = Qwen2Config.qwen2_5_1_5b()
tok = ByteTokenizer(vocab_size=shape.vocab_size - 1)
tok.special_tokens["<|im_end|>"] = shape.vocab_size - 2
tok.special_tokens["[PAD]"] = shape.vocab_size - 1
tok.id_to_special = {v: k for k, v in tok.special_tokens.items()}
tok.eos_token_id, tok.pad_token_id, tok.vocab_size = shape.vocab_size - 2, shape.vocab_size - 1, shape.vocab_size
policy = get_peft_model(Qwen2ForCausalLM.from_config(shape, torch.bfloat16, dev, seed=0),
                        LoraConfig(r=64, lora_alpha=16, modules_to_save=["embed_tokens", "lm_head", "score"]))
ref_policy = Qwen2ForCausalLM.from_config(shape, torch.bfloat16, dev, seed=0)
rm = DebertaV3ForSequenceClassification.from_config(DebertaV3Config.large(), torch.bfloat16, dev, seed=1)


@dataclass
class GRPOConfig(RLConfig):
    grpo_sample_N: int = 4


cfg = GRPOConfig(output_dir="/tmp/nrl_prof", response_length=resp_len, per_device_train_batch_size=4, gradient_accumulation_steps=8,
                 num_mini_batches=mini, total_episodes=10**6, save_strategy="no", report_to="none", sampler="native", resume="never",
                 watchdog_timeout_s=0)
cfg.quiet = True
tr = GRPOTrainer(cfg, tok, policy, ref_policy, synthetic_token_dataset(4096, shape.vocab_size - 2, 24, 160, seed=1),
                 reward_func=ModelReward(rm, None, 16, dev, token_budget=65536))
it = iter(tr.dataloader)
tr.train_one_update(1, next(it))
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    m = tr.train_one_update(2, next(it))
    torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
table = prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70)
with open(os.path.join(ROOT, "gpurun_out", f"profile_update_resp{resp_len}_mini{mini}.txt"), "w") as f:
    f.write(str({k: v for k, v in m.items() if k.startswith("time/")}) + "\n" + table)
print({k: round(v, 3) for k, v in m.items() if k.startswith("time/") and "wall" not in k})
print(table[:6000])
