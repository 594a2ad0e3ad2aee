"""Worker-side code of the symbolic answer checker (kept free of torch imports so spawned workers start fast)."""
import math
import re


_FUNCS = ("arcsin", "arccos", "arctan", "sinh", "cosh", "tanh", "sin", "cos", "tan", "cot", "sec", "csc", "exp", "ln", "log")


def latex_to_expr_text(s: str) -> str:
    """Tiny LaTeX -> sympy-parsable text converter for the forms math answers take."""
    s = s.replace("\\cdot", "*").replace("\\times", "*").replace("\\div", "/").replace("\\pi", "pi")
    s = s.replace("\\infty", "oo").replace("\\ ", "")
    for _ in range(8):
        s2 = re.sub(r"\\frac\{([^{}]*)\}\{([^{}]*)\}", r"((\1)/(\2))", s)
        s2 = re.sub(r"\\sqrt\[([^\]]+)\]\{([^{}]*)\}", r"((\2)**(1/(\1)))", s2)
        s2 = re.sub(r"\\sqrt\{([^{}]*)\}", r"sqrt(\1)", s2)
        s2 = re.sub(r"\\log_\{?([\w.]+)\}?\s*(?:\{([^{}]*)\}|\(([^()]*)\)|([\w.]+))",
                    lambda m: f"(log({m.group(2) or m.group(3) or m.group(4)})/log({m.group(1)}))", s2)     # \log_2 8, \log_{10}(x)
        s2 = re.sub(r"\^\{([^{}]*)\}", r"**(\1)", s2)
        if s2 == s:
            break
        s = s2
    for f in _FUNCS:                                                 # \sin x -> sin(x), \ln{2} -> ln(2), \cos(x) stays
        s = re.sub(rf"\\{f}\s*\{{([^{{}}]*)\}}", rf"{f}(\1)", s)
        s = re.sub(rf"\\{f}\s*(?=\()", f, s)
        s = re.sub(rf"\\{f}\s*([\w.]+)", rf"{f}(\1)", s)
    s = s.replace("ln(", "log(")
    s = s.replace("^", "**").replace("{", "(").replace("}", ")")
    s = re.sub(r"(\d+)!", r"factorial(\1)", s)
    s = re.sub(r"\(([^()]*)\)!", r"factorial(\1)", s)
    s = re.sub(r"(\d)(?![eE][+-]?\d)([a-zA-Z(])", r"\1*\2", s)        # 2x -> 2*x, but 1e3 stays a number
    s = re.sub(r"\)\(", ")*(", s)
    s = re.sub(r"\)([a-zA-Z])", r")*\1", s)                           # ((1)/(2))x -> ((1)/(2))*x
    return s


def warm():
    import sympy  # noqa: F401
    from sympy.parsing.sympy_parser import parse_expr  # noqa: F401
    return True


def symbolic_equal_impl(a: str, b: str) -> bool:
    import sympy
    from sympy.parsing.sympy_parser import parse_expr
    try:
        ea, eb = parse_expr(latex_to_expr_text(a)), parse_expr(latex_to_expr_text(b))
    except Exception:
        return False
    try:
        if sympy.simplify(ea - eb) == 0:
            return True
    except Exception:
        pass
    try:
        return bool(math.isclose(float(sympy.N(ea)), float(sympy.N(eb)), rel_tol=1e-3, abs_tol=1e-3))
    except Exception:
        return False
