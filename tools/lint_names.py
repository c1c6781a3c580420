#!/usr/bin/env python
"""Undefined-name check for the GPU-only Python paths (they cannot run on the CPU box, so a typo would only show up in a
paid GPU call).  ``python tools/lint_names.py`` -> non-zero exit if a function loads a name that is neither local, nor
module-level, nor a builtin."""
import ast
import builtins
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def check(path: str) -> int:
    tree = ast.parse(open(path).read())
    mod = set(dir(builtins)) | {"__file__"}
    def top_level(stmts):
        for node in stmts:
            if isinstance(node, (ast.Import, ast.ImportFrom)):
                mod.update((a.asname or a.name).split(".")[0] for a in node.names)
            elif isinstance(node, (ast.FunctionDef, ast.ClassDef)):
                mod.add(node.name)
            elif isinstance(node, (ast.If, ast.Try, ast.With, ast.For, ast.While)):
                for n in ast.walk(node):
                    if isinstance(n, (ast.FunctionDef, ast.Lambda)):
                        continue
                    if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store):
                        mod.add(n.id)
                    elif isinstance(n, (ast.Import, ast.ImportFrom)):
                        mod.update((a.asname or a.name).split(".")[0] for a in n.names)
                    elif isinstance(n, (ast.FunctionDef, ast.ClassDef)):
                        mod.add(n.name)
            else:
                for n in ast.walk(node):
                    if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store):
                        mod.add(n.id)

    top_level(tree.body)
    for node in ast.walk(tree):                      # class-level names (methods, attributes) resolve through self/cls
        if isinstance(node, ast.ClassDef):
            for st in node.body:
                if isinstance(st, (ast.FunctionDef, ast.ClassDef)):
                    pass
    def local_names(fn):
        names = set()
        for n in ast.walk(fn):
            if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
                names.add(n.id)
            elif isinstance(n, (ast.FunctionDef, ast.ClassDef)):
                names.add(n.name)
            elif isinstance(n, ast.arg):
                names.add(n.arg)
            elif isinstance(n, (ast.Import, ast.ImportFrom)):
                names.update((a.asname or a.name).split(".")[0] for a in n.names)
            elif isinstance(n, ast.ExceptHandler) and n.name:
                names.add(n.name)
        return names

    bad = 0

    def visit(node, visible):
        nonlocal bad
        for child in ast.iter_child_nodes(node):
            if isinstance(child, (ast.FunctionDef, ast.Lambda)):
                inner = visible | local_names(child)        # closures see every enclosing function's names
                for n in ast.walk(child):
                    if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in inner:
                        # a nested function may define it; resolved when that function is visited with its own scope
                        owner = [f for f in ast.walk(child) if isinstance(f, (ast.FunctionDef, ast.Lambda)) and f is not child
                                 and any(m is n for m in ast.walk(f))]
                        if not owner:
                            print("%s:%d: undefined name %r in %s" % (os.path.relpath(path, ROOT), n.lineno, n.id,
                                                                      getattr(child, "name", "<lambda>")))
                            bad += 1
                visit(child, inner)
            else:
                visit(child, visible)

    visit(tree, mod)
    return bad


def main() -> int:
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    for pat in ("distributedmnist_b200/**/*.py", "tools/*.py", "tests/*.py", "baseline/*.py", "src/*.py"):
        files += glob.glob(os.path.join(ROOT, pat), recursive=True)
    bad = sum(check(f) for f in sorted(set(files)))
    print("lint_names: %d file(s), %d problem(s)" % (len(set(files)), bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
