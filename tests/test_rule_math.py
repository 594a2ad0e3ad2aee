import torch

from nanorlhf_b200.reward import rule_math as rm


def test_get_boxed_and_answer_extraction():
    assert rm.get_boxed(r"so \boxed{\frac{1}{2}} and finally \boxed{x^{2}+1}") == "x^{2}+1"
    assert rm.get_boxed("no box") is None
    assert rm.get_boxed(r"\boxed{a{b}c") is None
    assert rm.extract_answer_is("blah. The answer is: 42") == "42"


def test_normalise_and_numeric_equivalence():
    assert rm.strip_string(r"\dfrac{1}{2}") == r"\frac{1}{2}"
    assert rm.strip_string("1,000 dollars") == "1000"
    assert rm.iscorrect("0.5", r"\frac{1}{2}")
    assert rm.iscorrect("1/2", r"\frac12")
    assert rm.iscorrect(" 42 ", "42", match="exact") and not rm.iscorrect("42.0", "42", match="exact")
    assert rm.iscorrect("42.0", "42")
    assert rm.iscorrect("(1, 2)", "(1,2.0)") and not rm.iscorrect("(1,2)", "[1,2]")
    assert not rm.iscorrect("41", "42") and not rm.iscorrect(None, "42")


def test_symbolic_equivalence_with_timeout():
    try:
        assert rm.iscorrect(r"\frac{x^{2}-1}{x-1}", "x+1")
        assert rm.iscorrect(r"2\sqrt{2}", r"\sqrt{8}")
        assert not rm.iscorrect("x+2", "x+1")
    finally:
        rm.shutdown_pool()


def test_rule_reward_callback():
    q = "What is 2 + 3?"
    text = rm.__dict__["R1_QUESTION_RE"].pattern  # noqa: F841 (regex is part of the public template contract)
    from nanorlhf_b200.utils.data import R1_TEMPLATE
    prompt = R1_TEMPLATE.replace("QUESTION", q)
    r = rm.RuleMathReward({q: "5"}, match="equiv")
    s = r([prompt + r"2+3=5 so \boxed{5}<|im_end|>", prompt + r"\boxed{6}", prompt + "five"], None, "<|im_end|>")
    assert s.tolist() == [1.0, 0.0, 0.0]
