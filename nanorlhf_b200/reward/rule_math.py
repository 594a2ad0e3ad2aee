"""Rule-based math reward: boxed-answer extraction + answer equivalence with a *working* timeout.

Reference: the r1-v0 example rewards 1/0 by comparing the ``\\boxed{}`` content of the response with the
MetaMathQA "The answer is: X" ground truth through ``iscorrect`` (/root/reference/examples/r1-v0/
grpo_r1.py:194-224,250-273), backed by 2.4 kLoC of vendored graders (``utils/toolkit_for_MATH/
latex_answer_check.py``, ``utils/eval/eval_script.py``, ``utils/eval/eval_utils.py``,
``utils/data_processing/answer_extraction.py``).  In the shipped reference the symbolic graders are never
reached: every comparison is forked into a fresh process with a 15 ms join timeout and a mismatched
argument list, so the effective rule is a whitespace-stripped exact string match.

This module implements the *intended* behaviour, re-derived from the description of those graders:
  normalise (MATH-style ``strip_string``) -> exact match -> numeric match (floats, fractions, percent,
  thousands separators) -> set/interval/tuple element-wise match -> symbolic equivalence with sympy,
run inside a persistent worker pool with a real per-comparison timeout; ``match="exact"`` reproduces
the shipped behaviour.
"""
from __future__ import annotations

import math
import multiprocessing as mp
import re
from typing import Dict, List, Optional, Sequence

import torch

# --------------------------------------------------------------------------------------------------
# extraction
# --------------------------------------------------------------------------------------------------
def get_boxed(text: str) -> Optional[str]:
    """Content of the last ``\\boxed{...}`` / ``\\fbox{...}`` (brace-matched); None when absent."""
    idx = max(text.rfind("\\boxed"), text.rfind("\\fbox"))
    if idx < 0:
        return None
    i = text.find("{", idx)
    if i < 0:
        # "\boxed 5" form
        m = re.match(r"\\boxed\s+([^\s$]+)", text[idx:])
        return m.group(1) if m else None
    depth, j = 0, i
    while j < len(text):
        if text[j] == "{":
            depth += 1
        elif text[j] == "}":
            depth -= 1
            if depth == 0:
                return text[i + 1:j]
        j += 1
    return None


def extract_answer_is(text: str) -> Optional[str]:
    """MetaMathQA ground truth: the text after the last 'The answer is: '."""
    m = list(re.finditer(r"[Tt]he answer is:?\s*", text))
    if not m:
        return None
    return text[m[-1].end():].strip().rstrip(".").strip()


_LAST_NUMBER_RE = re.compile(r"-?\d[\d,]*(?:\.\d+)?(?:/\d+)?|-?\.\d+")


_FEW_SHOT_STOPS = {"math": ("Problem:",), "gsm8k": ("Q: ",), "sat": ("Problem:",), "mmlu": ("Problem:",), "ocw": ("Problem:",),
                   "gaokao": ("问题 ",), "cmath": ("问题：",), "minif2f": ("Informal:",)}


def extract_answer(text: str, dataset: str = "math") -> Optional[str]:
    """Final answer of a free-form solution, in the order the reference's extractors try
    (/root/reference/examples/r1-v0/utils/data_processing/answer_extraction.py:65-338, re-derived from their behaviour):
    a few-shot continuation is cut at the next exemplar header (``Problem:``, ``Q: ``, ...) -> last ``\\boxed{}`` -> "final
    answer is $...$" (minerva) -> "the answer is ..." / "答案是" -> ``#### x`` (gsm8k) -> the body of the last ```` ```output ````
    block (tool-integrated solutions) -> a multiple-choice letter for ``dataset in {"mmlu", "sat", "aqua"}`` -> the last
    number in the text."""
    if text is None:
        return None
    for stop in _FEW_SHOT_STOPS.get(dataset, ()):
        if stop in text:
            text = text.split(stop, 1)[0]
    b = get_boxed(text)
    if b is not None:
        return b.strip()
    if dataset in ("mmlu", "sat", "aqua"):
        m = re.search(r"final answer is \(?([a-eA-E])\)?(?![a-zA-Z])", text)
        if m:
            return m.group(1).upper()
    m = re.search(r"[Ff]inal answer is:?\s*\$?(.+?)\$?(?:\.\s*I hope it is correct|\.?\s*$)", text.strip(), re.S)
    if m:
        return m.group(1).strip().strip("$").strip()
    a = extract_answer_is(text)
    if a is not None and a != "":
        return a.split("\n")[0].strip().strip("$").rstrip(".").strip()
    if "答案是" in text:
        a = text.split("答案是", 1)[1].strip().split("\n")[0].strip("：:。$ ")
        if dataset == "cmath":
            nums = re.findall(r"-?\d+\.?\d*", a)
            return nums[-1] if nums else None
        return a or None
    m = re.search(r"####\s*(.+)", text)
    if m:
        return m.group(1).strip().replace(",", "")
    if "```output" in text:
        out = text.split("```output")[-1].split("```")[0].strip()
        if out:
            return out
    if dataset in ("mmlu", "sat", "aqua"):
        m = re.findall(r"\(?\b([A-E])\b\)?", text)
        if m:
            return m[-1]
    nums = _LAST_NUMBER_RE.findall(text)
    return nums[-1].replace(",", "") if nums else None


def extract_answers(question: str, text: str, dataset: str = "math") -> List[str]:
    """List form (reference ``extract_math_answer``, answer_extraction.py:232-242): every boxed answer of the solution; an
    answer is split at commas when the question asks for values "separated by commas" (and it is not a tuple / interval) and
    at ``\\text{ and }``."""
    for stop in _FEW_SHOT_STOPS.get(dataset, ()):
        if stop in text:
            text = text.split(stop, 1)[0]
    found, rest = [], text
    while True:
        idx = max(rest.rfind("\\boxed"), rest.rfind("\\fbox"))
        if idx < 0:
            break
        inner = get_boxed(rest[idx:])
        if inner is not None:
            found.append(inner.strip())
        rest = rest[:idx]
    found.reverse()
    if not found:
        one = extract_answer(text, dataset)
        found = [one] if one else []
    out: List[str] = []
    for ans in found:
        if "separated by commas" in question and not any(ch in ans for ch in "()[]"):
            out.extend(a.strip() for a in ans.split(","))
        elif re.search(r"\\text\{\s*and\s*\}", ans):
            out.extend(a.strip() for a in re.split(r"\\text\{\s*and\s*\}", ans))
        else:
            out.append(ans)
    return out


def extract_program(text: str, last_only: bool = True) -> str:
    """Source of the ```` ```python ```` block(s) of a tool-integrated solution (reference eval_utils.py:14-31): the last
    block, or all blocks joined, without the fences."""
    blocks = re.findall(r"```python[^\n]*\n(.*?)(?:\n```|\Z)", text, re.S)
    if not blocks:
        return ""
    return blocks[-1].rstrip() + "\n" if last_only else "\n# ========\n".join(b.rstrip() + "\n" for b in blocks)


def ground_truth_of(example: dict, dataset: str) -> Optional[str]:
    """Normalised ground-truth answer of one benchmark record (reference ``parse_ground_truth``, eval_utils.py:34-78): where
    each of the evaluation sets keeps its answer."""
    if "gt" in example:
        return strip_string(example["gt"])
    if dataset in ("math", "ocw"):
        gt = extract_answer(example["solution"])
    elif dataset == "gsm8k":
        gt = example["answer"].split("####")[-1]
    elif dataset in ("gsm-hard", "mawps", "bbh"):
        gt = example["target"]
    elif dataset == "svamp":
        gt = example["Answer"]
    elif dataset == "asdiv":
        gt = re.sub(r"\(.*?\)", "", str(example["answer"]))           # "12 (apples)" -> "12"
    elif dataset == "tabmwp":
        gt = str(example["answer"])
        if example.get("ans_type") in ("integer_number", "decimal_number"):
            if "/" in gt:
                n, d = gt.split("/")[:2]
                gt = str(int(n) / int(d))
            elif "%" in gt:
                gt = str(float(gt.split("%")[0]) / 100)
            else:
                gt = str(float(gt.replace(",", "")))
    else:
        raise NotImplementedError(f"ground_truth_of: unknown dataset {dataset!r}")
    return None if gt is None else strip_string(str(gt).strip())


# --------------------------------------------------------------------------------------------------
# normalisation (MATH "strip_string" family)
# --------------------------------------------------------------------------------------------------
_UNITS = ["degrees", "degree", "cm", "centimeters", "meters", "meter", "inches", "inch", "feet", "foot", "units", "unit",
          "square", "cents", "cent", "dollars", "dollar", "mph", "hours", "hour", "minutes", "minute", "seconds",
          "second", "days", "day", "years", "year", "pounds", "pound", "kg", "grams", "gram", "miles", "mile"]


def _fix_fracs(s: str) -> str:
    # \frac12 -> \frac{1}{2}, \frac1{2} -> \frac{1}{2}
    out, i = "", 0
    while i < len(s):
        if s.startswith("\\frac", i):
            j = i + 5
            if j < len(s) and s[j] != "{":
                if j + 1 < len(s) and s[j + 1] != "{":
                    out += "\\frac{" + s[j] + "}{" + s[j + 1] + "}"
                    i = j + 2
                    continue
                if j + 1 < len(s):
                    out += "\\frac{" + s[j] + "}"
                    i = j + 1
                    continue
        out += s[i]
        i += 1
    return out


def _fix_a_slash_b(s: str) -> str:
    m = re.fullmatch(r"(-?\d+)/(\d+)", s)
    return f"\\frac{{{m.group(1)}}}{{{m.group(2)}}}" if m else s


def _fix_sqrt(s: str) -> str:
    return re.sub(r"\\sqrt(\w)", r"\\sqrt{\1}", s)


def strip_string(s: str, keep_equation: bool = False) -> str:
    s = str(s).strip()
    inner = get_boxed(s) if "\\boxed" in s or "\\fbox" in s else None         # \boxed{7} on either side -> 7
    if inner:
        s = inner
    s = s.replace("\n", "").replace("\\!", "").replace("\\\\", "\\")
    s = s.replace("tfrac", "frac").replace("dfrac", "frac").replace("cfrac", "frac")
    s = s.replace("x\\in", "")
    s = s.replace("infinity", "\\infty")
    s = re.sub(r"(?<![a-zA-Z\\])inf(?![a-zA-Z])", r"\\infty", s)
    s = s.replace("+\\infty", "\\infty")
    s = re.sub(r"\\le(?![a-zA-Z])", r"\\leq", s)
    s = re.sub(r"\\ge(?![a-zA-Z])", r"\\geq", s)
    s = s.replace("\\neq", "\\ne")
    s = s.replace("\\left", "").replace("\\right", "")
    s = s.replace("\\{", "{").replace("\\}", "}")
    s = s.replace("^{\\circ}", "").replace("^\\circ", "").replace("°", "")
    s = s.replace("\\$", "").replace("$", "")
    s = re.sub(r"(?<=[\d}])\s*\\(?:text|mbox|mathrm)\{[^}]*\}\s*$", "", s)        # trailing unit: 5\text{ cm}, 3\mathrm{m/s}
    s = re.sub(r"\\text\{\s*([^}]*)\}", r"\1", s)
    s = re.sub(r"\\mbox\{\s*([^}]*)\}", r"\1", s)
    s = re.sub(r"\\mathrm\{\s*([^}]*)\}", r"\1", s)
    s = re.sub(r"\\(?:mathbf|textbf|boldsymbol)\{([^}]*)\}", r"\1", s)
    s = re.sub(r"(?:\\,|\\;|\\:|\\quad|\\qquad|~)", "", s)
    for u in _UNITS:
        s = re.sub(rf"(?<=[\d\s}}]){u}\b", "", s)
    s = s.replace("\\%", "").replace("%", "")
    s = s.replace(" .", " 0.").replace("{.", "{0.")
    if s.startswith("."):
        s = "0" + s
    if not keep_equation and len(s.split("=")) == 2 and len(s.split("=")[0]) <= 2:
        s = s.split("=")[1]
    s = re.sub(r"(\d+)\.0+(?=\D|$)", r"\1", s)             # 3.000 -> 3, 2.0x -> 2x
    if "j" in s and "i" not in s and re.search(r"(?<![a-zA-Z])j(?![a-zA-Z])", s):
        s = re.sub(r"(?<![a-zA-Z])j(?![a-zA-Z])", "i", s)      # engineering notation of the imaginary unit
    s = re.sub(r"\{(?:c|m)?m\}(?:\^\{?[23]\}?)?", "", s)        # 5{cm}^2
    s = re.sub(r"\s*p\.m\.$", "", s)
    s = _fix_sqrt(s)
    s = s.replace(" ", "")
    s = _fix_fracs(s)
    if s == "0.5":
        s = "\\frac{1}{2}"
    s = _fix_a_slash_b(s)
    s = re.sub(r"(\d),(?=\d{3}(\D|$))", r"\1", s)          # 1,000 -> 1000
    s = s.rstrip(".")
    return s


# --------------------------------------------------------------------------------------------------
# numeric / symbolic comparison
# --------------------------------------------------------------------------------------------------
_SCI_RE = re.compile(r"^(-?[\d.]+)(?:\\times|\\cdot|\*|x)10\^\{?(-?\d+)\}?$")


def _to_float(s: str) -> Optional[float]:
    if "," in s:
        if not re.fullmatch(r"-?\d{1,3}(,\d{3})+(\.\d+)?", s):        # "1,3,5" is a list, "1,250.00" a number
            return None
        s = s.replace(",", "")
    m = _SCI_RE.match(s)                              # 3.0\times10^{5}, 2\cdot10^-3 (OCW-style numeric answers)
    if m:
        try:
            return float(m.group(1)) * 10.0 ** int(m.group(2))
        except ValueError:
            return None
    m = re.fullmatch(r"(-?\d+)\\frac\{(\d+)\}\{(\d+)\}", s)           # mixed number 1\frac{1}{2}
    if m and int(m.group(3)) != 0:
        whole = int(m.group(1))
        return whole + (-1 if whole < 0 else 1) * int(m.group(2)) / int(m.group(3))
    m = re.fullmatch(r"\\frac\{(-?[\d.]+)\}\{(-?[\d.]+)\}", s)
    try:
        if m:
            return float(m.group(1)) / float(m.group(2))
        m = re.fullmatch(r"(-?)\\frac\{(-?[\d.]+)\}\{(-?[\d.]+)\}", s)
        if m:
            return (-1.0 if m.group(1) else 1.0) * float(m.group(2)) / float(m.group(3))
        return float(s)
    except (ValueError, ZeroDivisionError):
        return None


def numeric_equal(a: float, b: float, rel_tol: float = 1e-3, abs_tol: float = 1e-3) -> bool:
    """The reference's two graders accept a number within rel 1e-3 (latex_answer_check.py:104-121 ``isclose(rel_tol=1e-3)``)
    OR within abs 1e-3 (eval_utils.py:205 ``isclose(abs_tol=1e-3)``); their results are OR-ed (grpo_r1.py:221)."""
    return math.isclose(a, b, rel_tol=rel_tol, abs_tol=abs_tol)


from ._sym_worker import latex_to_expr_text, symbolic_equal_impl as _symbolic_equal_impl, warm as _warm  # noqa: E402


_POOL = None
_POOL_BROKEN = False          # worker processes cannot start here (e.g. the main module is not importable by "spawn")


def _pool():
    global _POOL, _POOL_BROKEN
    if _POOL is None and not _POOL_BROKEN:
        ctx = mp.get_context("spawn")
        _POOL = ctx.Pool(2, maxtasksperchild=500)
        try:                      # pay the interpreter + sympy import once, outside any per-answer timeout
            _POOL.apply_async(_warm).get(60)
        except Exception:
            import warnings
            _POOL.terminate()
            _POOL, _POOL_BROKEN = None, True
            warnings.warn("rule_math: the sympy worker pool did not start; symbolic checks run on a watchdog thread in this process")
    return _POOL


def shutdown_pool():
    global _POOL
    if _POOL is not None:
        _POOL.terminate()
        _POOL = None


def _symbolic_equal_thread(a: str, b: str, timeout_s: float) -> bool:
    """Fallback without worker processes: a daemon thread that is abandoned (not killed) when it overruns."""
    import threading
    box = []
    t = threading.Thread(target=lambda: box.append(_symbolic_equal_impl(a, b)), daemon=True)
    t.start()
    t.join(timeout_s)
    return bool(box and box[0])


def symbolic_equal(a: str, b: str, timeout_s: float = 3.0) -> bool:
    """sympy equivalence in a persistent worker with a real timeout (a hung simplify kills the worker)."""
    global _POOL
    pool = _pool()
    if pool is None:
        return _symbolic_equal_thread(a, b, timeout_s)
    try:
        return bool(pool.apply_async(_symbolic_equal_impl, (a, b)).get(timeout_s))
    except mp.TimeoutError:
        shutdown_pool()         # the stuck worker is terminated; a fresh pool is built lazily
        return False
    except Exception:
        return False


def _split_elements(s: str) -> Optional[List[str]]:
    if len(s) >= 2 and s[0] in "([{" and s[-1] in ")]}" and "," in s:
        depth, cur, parts = 0, "", []
        for ch in s[1:-1]:
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            if ch == "," and depth == 0:
                parts.append(cur)
                cur = ""
            else:
                cur += ch
        parts.append(cur)
        return parts
    return None


_MATRIX_RE = re.compile(r"\\begin\{(p|b|v|B|V)?matrix\}(.*?)\\end\{(?:p|b|v|B|V)?matrix\}|\\begin\{array\}(?:\{[^}]*\})?(.*?)\\end\{array\}", re.S)


def _parse_matrix(s: str) -> Optional[List[List[str]]]:
    r"""Rows x columns of a single LaTeX matrix / array environment (after ``strip_string`` the row separator is a single
    backslash), None when ``s`` is not exactly one such environment."""
    m = _MATRIX_RE.fullmatch(s)
    if not m:
        return None
    body = m.group(2) if m.group(2) is not None else m.group(3)
    rows = [r for r in re.split(r"\\\\|\\(?![a-zA-Z])", body) if r.strip() != ""]
    return [[c.strip() for c in r.split("&")] for r in rows]


def _top_level_split(s: str, sep: str = ",") -> List[str]:
    depth, cur, parts = 0, "", []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return parts


def is_equiv(pred: str, gt: str, use_sympy: bool = True, timeout_s: float = 3.0) -> bool:
    if pred is None or gt is None:
        return False
    a, b = strip_string(pred), strip_string(gt)
    if a == b:
        return True
    # matrices / vectors: same shape, element-wise equivalent; a column vector may also be written as a tuple
    ma, mb = _parse_matrix(a), _parse_matrix(b)
    if ma is not None or mb is not None:
        if ma is None or mb is None:
            other, mat = (a, mb) if ma is None else (b, ma)
            flat = _split_elements(other)
            if flat is None or not (all(len(r) == 1 for r in mat) or len(mat) == 1):
                return False
            cells = [r[0] for r in mat] if len(mat) > 1 else mat[0]
            return len(cells) == len(flat) and all(is_equiv(x, y, use_sympy, timeout_s) for x, y in zip(cells, flat))
        return (len(ma) == len(mb) and all(len(x) == len(y) for x, y in zip(ma, mb))
                and all(is_equiv(x, y, use_sympy, timeout_s) for ra, rb in zip(ma, mb) for x, y in zip(ra, rb)))
    if a.lower() == b.lower() and not any(c.isdigit() for c in a):
        return True
    fa, fb = _to_float(a), _to_float(b)
    if fa is not None and fb is not None:
        return numeric_equal(fa, fb) or numeric_equal(fa, fb / 100) or numeric_equal(fa / 100, fb)
    if "\\cup" in a or "\\cup" in b:
        pa, pb = a.split("\\cup"), b.split("\\cup")
        return len(pa) == len(pb) and all(is_equiv(x, y, use_sympy, timeout_s) for x, y in zip(pa, pb))
    ea, eb = _split_elements(a), _split_elements(b)
    if ea is not None and eb is not None:
        # tuples / intervals: ordered, and the bracket types are part of the answer ((1,2] != [1,2]); {..} is a set
        if a[0] == "{" and b[0] == "{":
            return _multiset_equiv(ea, eb, use_sympy, timeout_s)
        return (len(ea) == len(eb) and a[0] == b[0] and a[-1] == b[-1]
                and all(is_equiv(x, y, use_sympy, timeout_s) for x, y in zip(ea, eb)))
    # bare comma lists ("1, 3, 5" / "x=1,2"): order-insensitive
    if "," in a and "," in b and ea is None and eb is None:
        la_, lb_ = _top_level_split(a), _top_level_split(b)
        if len(la_) > 1 and len(lb_) > 1:
            return _multiset_equiv(la_, lb_, use_sympy, timeout_s)
    # equations: same solution set <=> lhs - rhs proportional; compare the differences up to sign
    qa, qb = strip_string(pred, keep_equation=True), strip_string(gt, keep_equation=True)
    if qa.count("=") == 1 and qb.count("=") == 1 and (len(qa.split("=")[0]) > 2 or len(qb.split("=")[0]) > 2):
        (la, ra), (lb, rb) = qa.split("="), qb.split("=")
        if use_sympy and symbolic_equal(f"({la})-({ra})", f"({lb})-({rb})", timeout_s):
            return True
        return bool(use_sympy and symbolic_equal(f"({la})-({ra})", f"-(({lb})-({rb}))", timeout_s))
    # a numeric ground truth against an answer with words around exactly one number ("5 apples", "x = 5 units")
    # (the reference's "aggressive" rule, latex_answer_check.py:196-224: only when what is left around the number has no
    # structure -- no LaTeX command, bracket, comparison or variable -- and a match can only accept, never reject)
    if fb is not None and fa is None:
        nums = _LAST_NUMBER_RE.findall(a)
        if len(nums) == 1 and not re.search(r"[\\()\[\]<>,^_=+*/xyz]", a.replace(nums[0], "", 1)):
            f1 = _to_float(_fix_a_slash_b(nums[0].replace(",", "")))
            if f1 is not None and numeric_equal(f1, fb):
                return True
    if use_sympy and len(a) < 200 and len(b) < 200:
        return symbolic_equal(a, b, timeout_s)
    return False


def _multiset_equiv(xs: List[str], ys: List[str], use_sympy: bool, timeout_s: float) -> bool:
    if len(xs) != len(ys):
        return False
    left = list(ys)
    for x in xs:
        for i, y in enumerate(left):
            if is_equiv(x, y, use_sympy, timeout_s):
                left.pop(i)
                break
        else:
            return False
    return True


def iscorrect(answer: Optional[str], ground_truth: Optional[str], match: str = "equiv", timeout_s: float = 3.0) -> bool:
    """``match='exact'`` = the shipped reference's effective rule (whitespace-stripped string equality)."""
    if answer is None or ground_truth is None:
        return False
    if "".join(answer.split()) == "".join(ground_truth.split()):
        return True
    if match == "exact":
        return False
    return is_equiv(answer, ground_truth, True, timeout_s)


# --------------------------------------------------------------------------------------------------
# reward / accuracy callbacks
# --------------------------------------------------------------------------------------------------
R1_QUESTION_RE = re.compile(r"# Question:\n(.*?)\nPlease reason step by step", re.S)


class RuleMathReward:
    """``reward_func(pmt_and_responses, responses_ids, tokenizer) -> FloatTensor`` (grpo_r1.py:250-273)."""

    def __init__(self, answers: Dict[str, str], match: str = "equiv", timeout_s: float = 3.0,
                 question_re: re.Pattern = R1_QUESTION_RE, answer_marker: str = "# Answer:\n"):
        self.answers, self.match, self.timeout_s = answers, match, timeout_s
        self.question_re, self.answer_marker = question_re, answer_marker

    def question_of(self, text: str) -> Optional[str]:
        m = self.question_re.search(text)
        return m.group(1) if m else None

    def __call__(self, pmt_and_responses: Sequence[str], responses_ids=None, tokenizer=None) -> torch.Tensor:
        eos = tokenizer.eos_token if (tokenizer is not None and not isinstance(tokenizer, str)) else (tokenizer or responses_ids)
        out = torch.zeros(len(pmt_and_responses))
        for i, text in enumerate(pmt_and_responses):
            q = self.question_of(text)
            gt = self.answers.get(q) if q is not None else None
            k = text.rfind(self.answer_marker)
            resp = text[k + len(self.answer_marker):] if k >= 0 else text
            if isinstance(eos, str) and eos and eos in resp:
                resp = resp[:resp.find(eos)]
            out[i] = 1.0 if iscorrect(get_boxed(resp), gt, self.match, self.timeout_s) else 0.0
        return out


def make_accuracy_func(problems: Sequence[Dict[str, str]], tokenizer, template: str, max_tokens: int = 2048,
                       match: str = "equiv"):
    """``accuracy_func(model, args) -> float``: greedy (T=0, seed 42) decode of ``problems`` on the resident
    sampler, accuracy by ``iscorrect`` (reference: MATH-500 probe, grpo_r1.py:276-341)."""
    from ..sampler.engine import generate

    def accuracy_func(model, args) -> float:
        prompts = [tokenizer(template.replace("QUESTION", p["problem"]), padding=False)["input_ids"] for p in problems]
        out = generate(1, model, tokenizer, prompts, 0.0, min(max_tokens, args.response_length), top_p=1.0, seed=42,
                       backend=args.sampler, rollout_dtype=args.rollout_dtype)
        texts = tokenizer.batch_decode(out)
        # response length in TOKENS up to and including EOS (the reference logs token counts, grpo_r1.py:320-337)
        stop = (out == tokenizer.pad_token_id)
        if tokenizer.eos_token_id is not None:
            after_eos = ((out == tokenizer.eos_token_id).long().cumsum(1) - (out == tokenizer.eos_token_id).long()) > 0
            stop = stop | after_eos
        tok_lens = (~stop).sum(1).float()
        good = 0
        for p, t in zip(problems, texts):
            t = t.split(tokenizer.eos_token)[0].replace(tokenizer.pad_token, "")
            good += int(iscorrect(get_boxed(t), p["answer"], match))
        accuracy_func.last_mean_response_tokens = float(tok_lens.mean()) if len(problems) else 0.0
        return good / max(len(problems), 1)

    return accuracy_func


def synthetic_arithmetic_problems(n: int, seed: int = 0) -> List[Dict[str, str]]:
    """Offline stand-in for MetaMathQA / MATH-500: two-operand arithmetic with exact answers."""
    import random
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        a, b = rng.randint(2, 99), rng.randint(2, 99)
        op = rng.choice("+-*")
        ans = {"+": a + b, "-": a - b, "*": a * b}[op]
        q = f"What is {a} {op} {b}?"
        out.append({"query": q, "problem": q, "response": f"... The answer is: {ans}", "answer": str(ans)})
    return out
