"""Static cross-checks that do not need a GPU (the build container has none, so call sites that only ever run on the device are
checked here for shape): every ctypes call in engine.py passes as many arguments as its entry in _abi.SIGNATURES, every prototype
in include/b200reg.h has as many parameters as that entry, and every ``eng.method(...)`` / ``module.function(...)`` call site in the
package, the GPU tests and the bench scripts matches the signature of what it calls (name exists, positional count, keyword
names, required parameters)."""
import ast
import glob
import inspect
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ctypes_calls_and_header_prototypes_match_the_binding_table():
    from platipy_b200 import _abi

    tree = ast.parse(open(os.path.join(ROOT, "platipy_b200", "engine.py")).read())
    seen = 0
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr.startswith("b200reg_"):
            name = node.func.attr
            assert name in _abi.SIGNATURES, name
            assert len(node.args) == len(_abi.SIGNATURES[name][1]), f"engine.py:{node.lineno}: {name}"
            seen += 1
    assert seen >= 50
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "b200reg.h")).read(), flags=re.S)
    protos = list(re.finditer(r"\b(b200reg_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S))
    assert len(protos) == len(_abi.SIGNATURES)
    for m in protos:
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ("void", "") else len(args.split(","))
        assert n == len(_abi.SIGNATURES[name][1]), name


def _check(call, target, name, where, problems):
    fn = getattr(target, name, None)
    if fn is None:
        problems.append(f"{where}: no attribute {name}")
        return
    if not callable(fn) or any(isinstance(a, ast.Starred) for a in call.args):
        return
    try:
        params = list(inspect.signature(fn).parameters.values())
    except (TypeError, ValueError):
        return
    if params and params[0].name == "self":
        params = params[1:]
    positional = [p for p in params if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
    if len(call.args) > len(positional) and not any(p.kind == p.VAR_POSITIONAL for p in params):
        problems.append(f"{where}: {name} takes {len(positional)} positional arguments, {len(call.args)} given")
    names = {p.name for p in params}
    if not any(p.kind == p.VAR_KEYWORD for p in params):
        problems += [f"{where}: {name} has no keyword {kw.arg}" for kw in call.keywords if kw.arg and kw.arg not in names]
    if not any(kw.arg is None for kw in call.keywords):
        given = {kw.arg for kw in call.keywords}
        missing = [p.name for i, p in enumerate(positional) if p.default is p.empty and i >= len(call.args) and p.name not in given]
        if missing:
            problems.append(f"{where}: {name} called without {missing}")


def test_call_sites_match_the_signatures_they_call():
    from platipy_b200 import comparison, fusion, generation, iar, label_utils, linear
    from platipy_b200.engine import Engine

    # local names under which the package, the GPU tests and the bench scripts refer to the engine / the modules
    modules = {"lu": label_utils, "label_utils": label_utils, "linear": linear, "fusion": fusion, "cmp": comparison, "comparison": comparison,
               "generation": generation, "iar": iar}
    files = glob.glob(os.path.join(ROOT, "platipy_b200", "*.py")) + glob.glob(os.path.join(ROOT, "tests", "test_gpu_*.py")) + \
        glob.glob(os.path.join(ROOT, "profiles", "exp_*.py")) + \
        [os.path.join(ROOT, "profiles", "bench_extras.py"), os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    problems, checked = [], 0
    for f in files:
        tree = ast.parse(open(f).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and isinstance(node.func.value, ast.Name):
                base, name = node.func.value.id, node.func.attr
                where = f"{os.path.relpath(f, ROOT)}:{node.lineno}"
                if base in ("eng", "engine"):
                    _check(node, Engine, name, where, problems)
                    checked += 1
                elif base in modules:
                    _check(node, modules[base], name, where, problems)
                    checked += 1
                elif base == "gen" and "test_gpu_" in f:
                    _check(node, generation, name, where, problems)
                    checked += 1
    assert not problems, "\n".join(problems)
    assert checked > 400
