from .helpers import (INVALID_LOGPROB, disable_dropout_in_model, exact_div, first_true_indices, masked_mean,
                      masked_var, masked_whiten, response_masks, scatter_terminal_reward, state_to_device,
                      truncate_response)
