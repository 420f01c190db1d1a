"""Static undefined-name check over every Python file that only runs on the GPU box (multi-rank branches,
bench arms, smoke) -- a typo there would otherwise surface at round end.  Conservative: a name counts as
defined if any enclosing function, the module, or builtins binds it anywhere."""
import ast
import builtins
import glob
import os

import pytest

from helpers import ROOT

FILES = sorted(set(
    glob.glob(os.path.join(ROOT, "sph_project_b200", "**", "*.py"), recursive=True)
    + [os.path.join(ROOT, f) for f in ("bench.py", "__graft_entry__.py", "run_simulation.py")]
    + glob.glob(os.path.join(ROOT, "tests", "*.py")) + glob.glob(os.path.join(ROOT, "SPH", "**", "*.py"), recursive=True)))


def bound_names(node):
    """Names bound directly in this scope (not in nested function scopes, except their own names)."""
    out = set()
    stack = list(ast.iter_child_nodes(node))
    while stack:
        n = stack.pop()
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            out.add(n.name)
            if isinstance(n, ast.ClassDef):
                continue
            continue
        if isinstance(n, ast.Lambda):
            continue
        if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            out.add(n.id)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            out |= {(a.asname or a.name).split(".")[0] for a in n.names}
        elif isinstance(n, ast.ExceptHandler) and n.name:
            out.add(n.name)
        elif isinstance(n, (ast.Global, ast.Nonlocal)):
            out |= set(n.names)
        elif isinstance(n, ast.arg):
            out.add(n.arg)
        stack.extend(ast.iter_child_nodes(n))
    return out


def check_scope(node, visible, problems, path):
    names = visible | bound_names(node)
    if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
        a = node.args
        names |= {x.arg for x in a.args + a.kwonlyargs + a.posonlyargs}
        names |= {x.arg for x in (a.vararg, a.kwarg) if x}
    stack = list(ast.iter_child_nodes(node))
    while stack:
        n = stack.pop()
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
            for d in getattr(n, "decorator_list", []):
                stack.append(d)
            check_scope(n, names, problems, path)
            continue
        if isinstance(n, ast.ClassDef):
            check_scope(n, names, problems, path)
            continue
        if isinstance(n, (ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp)):
            check_scope(n, names, problems, path)
            continue
        if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in names and not hasattr(builtins, n.id):
            problems.append(f"{os.path.relpath(path, ROOT)}:{n.lineno}: {n.id}")
        stack.extend(ast.iter_child_nodes(n))


@pytest.mark.parametrize("path", FILES, ids=lambda p: os.path.relpath(p, ROOT))
def test_no_undefined_names(path):
    tree = ast.parse(open(path).read())
    problems = []
    check_scope(tree, {"__file__", "__name__", "__doc__", "__path__"}, problems, path)
    assert not problems, problems
