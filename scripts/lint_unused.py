#!/usr/bin/env python
"""Tiny AST lint used by scripts/lint.sh (no third-party linter in the image): unused imports,
duplicate top-level definitions and mutable default arguments. Exit code 1 when anything is found."""

from __future__ import annotations

import ast
import sys
from pathlib import Path


def check(path: Path) -> list[str]:
    src = path.read_text()
    tree = ast.parse(src, filename=str(path))
    problems: list[str] = []
    imported: dict[str, int] = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                imported[(a.asname or a.name).split(".")[0]] = node.lineno
        elif isinstance(node, ast.ImportFrom):
            if node.module == "__future__":
                continue
            for a in node.names:
                if a.name != "*":
                    imported[a.asname or a.name] = node.lineno
    used = {n.id for n in ast.walk(tree) if isinstance(n, ast.Name)} | {
        n.value.id for n in ast.walk(tree) if isinstance(n, ast.Attribute) and isinstance(n.value, ast.Name)}
    exported: set[str] = set()
    for node in tree.body:
        if isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "__all__" for t in node.targets):
            if isinstance(node.value, (ast.List, ast.Tuple)):
                exported = {e.value for e in node.value.elts if isinstance(e, ast.Constant)}
    lines = src.splitlines()
    for name, line in sorted(imported.items(), key=lambda kv: kv[1]):
        if name in used or name in exported or path.name == "__init__.py":
            continue
        if "noqa" in lines[line - 1] or name in src.replace(f"import {name}", ""):  # used in strings/annotations
            if "noqa" in lines[line - 1]:
                continue
            # string annotations ("Manager") count as uses
            if any(isinstance(n, ast.Constant) and isinstance(n.value, str) and name in n.value for n in ast.walk(tree)):
                continue
        problems.append(f"{path}:{line}: unused import {name!r}")
    seen: dict[str, int] = {}
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)):
            if node.name in seen:
                problems.append(f"{path}:{node.lineno}: {node.name!r} redefined (first at line {seen[node.name]})")
            seen[node.name] = node.lineno
    for node in ast.walk(tree):
        if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef)):
            for d in node.args.defaults + [d for d in node.args.kw_defaults if d is not None]:
                if isinstance(d, (ast.List, ast.Dict, ast.Set)):
                    problems.append(f"{path}:{node.lineno}: mutable default argument in {node.name}()")
    return problems


def main() -> None:
    roots = [Path(p) for p in (sys.argv[1:] or ["torchft_b200", "bench", "examples", "tests", "bench.py", "train_ddp.py", "train_diloco.py"])]
    files: list[Path] = []
    for r in roots:
        files += sorted(r.rglob("*.py")) if r.is_dir() else [r]
    problems = [p for f in files for p in check(f)]
    print("\n".join(problems) if problems else f"lint_unused: {len(files)} files clean")
    sys.exit(1 if problems else 0)


if __name__ == "__main__":
    main()
