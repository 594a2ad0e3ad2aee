"""Worker-side code of the symbolic answer checker (kept free of torch imports so spawned workers start fast)."""
import math
import re


def latex_to_expr_text(s: str) -> str:
    """Tiny LaTeX -> sympy-parsable text converter for the forms math answers take."""
    s = s.replace("\\cdot", "*").replace("\\times", "*").replace("\\div", "/").replace("\\pi", "pi")
    s = s.replace("\\infty", "oo").replace("\\ ", "")
    for _ in range(8):
        s2 = re.sub(r"\\frac\{([^{}]*)\}\{([^{}]*)\}", r"((\1)/(\2))", s)
        s2 = re.sub(r"\\sqrt\[([^\]]+)\]\{([^{}]*)\}", r"((\2)**(1/(\1)))", s2)
        s2 = re.sub(r"\\sqrt\{([^{}]*)\}", r"sqrt(\1)", s2)
        s2 = re.sub(r"\^\{([^{}]*)\}", r"**(\1)", s2)
        if s2 == s:
            break
        s = s2
    s = s.replace("^", "**").replace("{", "(").replace("}", ")")
    s = re.sub(r"(\d)([a-zA-Z(])", r"\1*\2", s)
    s = re.sub(r"\)\(", ")*(", s)
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
        return bool(math.isclose(float(sympy.N(ea)), float(sympy.N(eb)), rel_tol=1e-4, abs_tol=1e-9))
    except Exception:
        return False
