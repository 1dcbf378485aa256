#!/usr/bin/env python
"""Undefined-name check without third-party linters: every name a function references as a global must be bound at module level (or be a
builtin). Catches typos in code paths the CPU test-suite cannot execute (GPU-only branches).   python scripts/check_names.py [paths...]"""
import builtins
import os
import symtable
import sys


def module_names(tab):
    names = set()
    for s in tab.get_symbols():
        if s.is_assigned() or s.is_imported() or s.is_namespace() or s.is_parameter():
            names.add(s.get_name())
    return names


def walk(tab, mod_names, path, problems):
    for child in tab.get_children():
        for s in child.get_symbols():
            n = s.get_name()
            if s.is_referenced() and s.is_global() and not s.is_assigned() and n not in mod_names and not hasattr(builtins, n):
                if n in ("__class__", "__file__", "__name__", "__doc__", "__builtins__", "__spec__", "__package__", "__path__"):
                    continue
                problems.append(f"{path}:{child.get_lineno()}: `{n}` referenced in {child.get_type()} `{child.get_name()}` is not defined at module level")
        walk(child, mod_names, path, problems)


def check(path, problems):
    src = open(path, encoding="utf-8").read()
    try:
        tab = symtable.symtable(src, path, "exec")
    except SyntaxError as e:
        problems.append(f"{path}: syntax error: {e}")
        return
    names = module_names(tab)
    if "import *" in src:
        return  # star imports: cannot know
    # names referenced at module level itself
    for s in tab.get_symbols():
        n = s.get_name()
        if s.is_referenced() and not (s.is_assigned() or s.is_imported() or s.is_namespace()) and not hasattr(builtins, n):
            if n not in ("__file__", "__name__", "__doc__", "__spec__", "__package__", "__path__"):
                problems.append(f"{path}:1: `{n}` referenced at module level is never bound")
    walk(tab, names, path, problems)


def main():
    roots = sys.argv[1:] or ["kungfu_b200", "benchmarks", "tests", "examples", "experiments", "experimental", "scripts", "bench.py", "__graft_entry__.py"]
    problems = []
    for r in roots:
        if os.path.isfile(r):
            check(r, problems)
            continue
        for d, _, files in os.walk(r):
            if "/build" in d or "/csrc" in d:
                continue
            for f in files:
                if f.endswith(".py"):
                    check(os.path.join(d, f), problems)
    for p in problems:
        print(p)
    print(f"{len(problems)} problem(s)")
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(main())
