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


def test_grader_breadth_matrices_intervals_lists_units():
    """Forms the reference's vendored graders cover (utils/eval/eval_utils.py:181-325, latex_answer_check.py:166-236)."""
    try:
        # matrices: element-wise, shape-sensitive; a column vector may be written as a tuple
        assert rm.iscorrect(r"\begin{pmatrix} 1 & 2 \\ 3 & 4 \end{pmatrix}", r"\begin{pmatrix}1&2\\3&4.0\end{pmatrix}")
        assert not rm.iscorrect(r"\begin{pmatrix} 1 & 2 \\ 3 & 5 \end{pmatrix}", r"\begin{pmatrix}1&2\\3&4\end{pmatrix}")
        assert rm.iscorrect(r"\begin{pmatrix} \frac{1}{2} \\ 3 \end{pmatrix}", r"(0.5, 3)")
        # intervals: bracket types are part of the answer; unions split on \cup; infinity
        assert rm.iscorrect(r"(-\infty, 2]", r"(-\infty,2]") and not rm.iscorrect(r"(-\infty, 2)", r"(-\infty,2]")
        assert rm.iscorrect(r"[1,2)\cup(3,4]", r"[1, 2) \cup (3, 4]")
        # units in \text / \mathrm, degrees, dollars, percent
        assert rm.iscorrect(r"5\text{ cm}", "5") and rm.iscorrect(r"3\mathrm{m}", "3") and rm.iscorrect(r"90^\circ", "90")
        assert rm.iscorrect(r"\$1,250.00", "1250") and rm.iscorrect(r"50\%", "0.5") and rm.iscorrect("50", r"50\%")
        # unordered lists and sets, ordered tuples
        assert rm.iscorrect("1, 3, 5", "5,3,1") and rm.iscorrect(r"\{1,2\}", r"\{2,1\}") and not rm.iscorrect("(1,2)", "(2,1)")
        # scientific notation, mixed numbers, a number embedded in words
        assert rm.iscorrect(r"3.0\times 10^{5}", "300000") and rm.iscorrect(r"1\frac{1}{2}", "1.5")
        assert rm.iscorrect("5 apples", "5") and not rm.iscorrect("5 apples and 6 pears", "5")
        # equations: same solution set up to sign / rearrangement
        assert rm.iscorrect("y=2x+1", "y = 1 + 2x") and rm.iscorrect("2x-y+1=0", "y=2x+1") and not rm.iscorrect("y=2x+1", "y=2x+2")
    finally:
        rm.shutdown_pool()


def test_extract_answer_styles():
    assert rm.extract_answer(r"thus \boxed{7} so final") == "7"
    assert rm.extract_answer("Final Answer: The final answer is $x^2+1$. I hope it is correct.") == "x^2+1"
    assert rm.extract_answer("so the answer is 12.") == "12"
    assert rm.extract_answer("work work\n#### 1,234") == "1234"
    assert rm.extract_answer("The correct option is (C)", dataset="mmlu") == "C"
    assert rm.extract_answer("we get 3 then 4 and finally 19") == "19"
    assert rm.extract_answer("nothing here") is None


def test_reference_tolerances_and_notation_forms():
    """Numeric tolerance = rel 1e-3 OR abs 1e-3 (the union of the reference's two graders, latex_answer_check.py:104-121 and
    eval_utils.py:205); \\boxed on the ground-truth side, \\le / \\leq, factorials, \\log_b, implicit products after a fraction;
    the single-number rule accepts only when nothing structured surrounds the number and never rejects on its own."""
    try:
        assert rm.iscorrect("12.3456", "12.35") and rm.iscorrect("100.05", "100") and rm.iscorrect("0.0004", "0.0001")
        assert not rm.iscorrect("1002", "1000") and not rm.iscorrect("12", "12.1")
        assert rm.iscorrect("7", r"\boxed{7}") and rm.iscorrect(r"x \le 3", r"x \leq 3") and not rm.iscorrect(r"x \leq -5", r"x \geq -5")
        assert rm.iscorrect(r"2\pi", "6.2832") and rm.iscorrect(r"\sqrt{2}", "1.4142")        # structured answer vs its decimal value
        assert rm.iscorrect("5!", "120") and rm.iscorrect(r"\log_2 8", "3") and rm.iscorrect(r"\frac{1}{2}x", "x/2")
        assert not rm.iscorrect(r"\sqrt{5}", "5") and not rm.iscorrect("2x+5", "5") and not rm.iscorrect("5^2", "5")
        assert rm.iscorrect("x=5 units", "5") and rm.iscorrect("1e3", "1000")
    finally:
        rm.shutdown_pool()


def test_extractor_family_and_normaliser_breadth():
    """The remaining extractor behaviours of answer_extraction.py:65-338: tool output blocks, few-shot truncation, SAT / MMLU
    letters, the CJK answer marker, the list form; and its normaliser's cfrac / infinity / trailing-zero / imaginary-unit rules."""
    assert rm.extract_answer("work\n```output\n42\n```\ndone") == "42"
    assert rm.extract_answer("so the final answer is (B). yes", dataset="sat") == "B"
    assert rm.extract_answer("计算得 答案是：12。\n", dataset="cmath") == "12"
    assert rm.extract_answer("答案是 $x+1$\n问题 2", dataset="gaokao") == "x+1"
    assert rm.extract_answer("The answer is 5.\n\nProblem: next one \\boxed{9}", dataset="math") == "5"      # next exemplar is cut
    assert rm.extract_answers("Enter all solutions, separated by commas.", r"so \boxed{1, 2,3}") == ["1", "2", "3"]
    assert rm.extract_answers("q", r"\boxed{x=1 \text{ and } y=2}") == ["x=1", "y=2"]
    assert rm.extract_answers("q", r"first \boxed{3} then \boxed{(1,2)}") == ["3", "(1,2)"]
    assert rm.strip_string(r"\cfrac{1}{2}") == r"\frac{1}{2}" and rm.strip_string("3.000") == "3" and rm.strip_string("1+2j") == "1+2i"
    assert rm.strip_string("(-inf, 3]") == r"(-\infty,3]" and rm.strip_string(r"x\in[1,2]") == "[1,2]" and rm.strip_string("information") == "information"
    assert rm.iscorrect("3.000", "3") and rm.iscorrect(r"(-\infty, 3]", "(-inf,3]")


def test_program_extraction_and_dataset_ground_truths():
    sol = "first\n```python\nx = 1\nprint(x)\n```\nthen\n```python\ny = 2\nprint(y)\n```\n```output\n2\n```"
    assert rm.extract_program(sol) == "y = 2\nprint(y)\n"
    assert rm.extract_program(sol, last_only=False).count("print") == 2 and rm.extract_program("no code") == ""
    assert rm.ground_truth_of({"solution": r"so \boxed{\dfrac{1}{2}}"}, "math") == r"\frac{1}{2}"
    assert rm.ground_truth_of({"answer": "work\n#### 1,234"}, "gsm8k") == "1234"
    assert rm.ground_truth_of({"answer": "12 (apples)"}, "asdiv") == "12"
    assert rm.ground_truth_of({"answer": "3/4", "ans_type": "decimal_number"}, "tabmwp") == "0.75"
    assert rm.ground_truth_of({"answer": "25%", "ans_type": "decimal_number"}, "tabmwp") == "0.25"
    assert rm.ground_truth_of({"target": "7"}, "bbh") == "7" and rm.ground_truth_of({"gt": " 5. "}, "anything") == "5"
