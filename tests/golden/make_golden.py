"""Extract the golden vectors of the reference's own unit tests into JSON fixtures.

JAX is not installable in this image, so the reference cannot be *run* here; its unit tests,
however, carry literal known-answer arrays (``np.testing.assert_allclose(x, jnp.array([...]))``)
and literal inputs (``coords = jnp.array([...])``).  This script parses those literals out of the
test sources with ``ast`` (no reference code is executed or copied) and writes them to
``tests/golden/reference_unit_goldens.json``.

    python tests/golden/make_golden.py [/root/reference]

Layout of the JSON:  {file: {function: {"assign": {name: array}, "asserts": [{"expr", "value",
"rtol", "atol", "line"}]}}}.  ``expr`` is the source text of the tested expression, ``line`` the
line in the reference test file (for citation).
"""
import ast
import json
import os
import sys

FILES = [
    "tests/unit/test_geometries.py",
    "tests/unit/test_mechanical_loss.py",
    "tests/unit/test_neo_hooke_mechanical_loss.py",
    "tests/unit/test_neo_hooke_mechanical_loss_AD.py",
    "tests/unit/test_saint_venant_mechanical_loss.py",
    "tests/unit/test_elastoplasticity.py",
    "tests/unit/test_sensitivity_analysis.py",
    "tests/unit/test_nonlinear_transient_thermal.py",
    "tests/unit/test_allencahn_loss.py",
    "tests/integration/test_mechanical_2D_sa.py",
    "tests/unit/test_kratos_ffi_mechanical_loss.py",
]


def _is_array_call(node):
    return (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute)
            and node.func.attr == "array" and len(node.args) >= 1)


class _Literal(ast.NodeTransformer):
    """Fold jnp.sqrt(c), a/b etc. down to numbers so literal_eval-like evaluation works."""


def _eval_literal(node):
    # arrays in the reference tests are nested lists of numbers (possibly with unary minus or a/b)
    code = compile(ast.Expression(body=node), "<golden>", "eval")
    return eval(code, {"__builtins__": {}}, {})


def _target_name(t):
    if isinstance(t, ast.Name):
        return t.id
    if isinstance(t, ast.Attribute):
        return t.attr
    return None


def extract(path):
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        rec = {"assign": {}, "asserts": []}
        for node in ast.walk(fn):
            if isinstance(node, ast.Assign) and _is_array_call(node.value):
                name = _target_name(node.targets[0])
                if name is None:
                    continue
                try:
                    rec["assign"][name] = _eval_literal(node.value.args[0])
                except Exception:
                    pass
            if (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute)
                    and node.func.attr == "assert_allclose" and len(node.args) >= 2
                    and _is_array_call(node.args[1])):
                try:
                    val = _eval_literal(node.args[1].args[0])
                except Exception:
                    continue
                kw = {k.arg: _eval_literal(k.value) for k in node.keywords
                      if k.arg in ("rtol", "atol")}
                rec["asserts"].append({"expr": ast.get_source_segment(src, node.args[0]),
                                       "value": val, "rtol": kw.get("rtol"),
                                       "atol": kw.get("atol"), "line": node.lineno})
        rec["asserts"].sort(key=lambda a: a["line"])
        if rec["assign"] or rec["asserts"]:
            out[fn.name] = rec
    return out


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    golden = {}
    for f in FILES:
        p = os.path.join(ref, f)
        if os.path.exists(p):
            golden[f] = extract(p)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_unit_goldens.json")
    with open(dst, "w") as fh:
        json.dump(golden, fh)
    n = sum(len(t["asserts"]) for f in golden.values() for t in f.values())
    print(f"wrote {dst}: {len(golden)} files, {n} known-answer arrays")


if __name__ == "__main__":
    main()
