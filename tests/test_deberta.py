import pytest
import torch

from nanorlhf_b200.models.deberta_v3 import DebertaV3Config, DebertaV3ForSequenceClassification


def test_matches_hf_deberta_v2():
    tr = pytest.importorskip("transformers")
    from transformers import DebertaV2Config, DebertaV2ForSequenceClassification
    cfg = DebertaV3Config.tiny(vocab_size=200)
    hf_cfg = DebertaV2Config(vocab_size=200, hidden_size=64, num_hidden_layers=2, num_attention_heads=4,
                             intermediate_size=128, relative_attention=True, position_buckets=16, norm_rel_ebd="layer_norm",
                             share_att_key=True, pos_att_type=["p2c", "c2p"], position_biased_input=False,
                             max_position_embeddings=64, max_relative_positions=-1, pooler_hidden_size=64,
                             num_labels=1, layer_norm_eps=1e-7, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                             pooler_dropout=0.0, type_vocab_size=0, pad_token_id=0)
    torch.manual_seed(0)
    hf = DebertaV2ForSequenceClassification(hf_cfg).eval()
    mine = DebertaV3ForSequenceClassification(cfg).eval()
    missing, unexpected = mine.load_state_dict(hf.state_dict(), strict=False)
    assert not missing, missing
    ids = torch.randint(3, 200, (3, 40))
    ids[0, 30:] = 0
    ids[2, 12:] = 0
    mask = ids != 0
    with torch.no_grad():
        want = hf(input_ids=ids, attention_mask=mask.long()).logits
        got = mine(ids, mask)
    assert torch.allclose(got, want.float(), atol=1e-4, rtol=1e-4), (got, want)


def test_model_reward_id_path():
    from nanorlhf_b200.reward.model_reward import ModelReward
    from nanorlhf_b200.utils.tokenizer import ByteTokenizer
    tok = ByteTokenizer()
    rm = DebertaV3ForSequenceClassification.from_config(DebertaV3Config.tiny(vocab_size=300), torch.float32, seed=0)
    r = ModelReward(rm, None, reward_batch_size=4)
    q = torch.randint(0, 250, (6, 10))
    resp = torch.randint(0, 250, (6, 12))
    resp[1, 5:] = tok.pad_token_id
    s = r(q, resp, tok)
    assert s.shape == (6,) and torch.isfinite(s).all()
    # batching must not change scores
    r2 = ModelReward(rm, None, reward_batch_size=1)
    assert torch.allclose(s, r2(q, resp, tok), atol=1e-5)
