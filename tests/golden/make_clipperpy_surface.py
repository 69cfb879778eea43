#!/usr/bin/env python
"""
Scans the reference's Python sources (/root/reference/roman, /root/reference/demo) for every use of the
`clipperpy` binding and writes tests/golden/clipperpy_surface.json:
  chains   dotted names rooted at the module        (clipperpy.invariants.ROMANParams, clipperpy.CLIPPER, ...)
  methods  methods called on a variable `clipper`   (clipper.solve, clipper.get_selected_associations, ...)
  params   attributes assigned on `iparams` objects (iparams.sigma = ..., self.iparams.cosine_min = ...)
  solution attributes read from clipper.get_solution()
tests/test_host_logic.py::test_shim_covers_every_clipperpy_use_of_the_reference checks that
roman_b200.clipperpy offers all of them — the precondition of the zero-change drop-in.
Run:  python tests/golden/make_clipperpy_surface.py
"""
import ast
import json
import os

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def dotted(node):
    parts = []
    while isinstance(node, ast.Attribute):
        parts.append(node.attr)
        node = node.value
    if isinstance(node, ast.Name):
        parts.append(node.id)
        return list(reversed(parts))
    if isinstance(node, ast.Call):        # e.g. clipper.get_solution().nodes
        inner = dotted(node.func)
        if inner:
            return inner + ["()"] + list(reversed(parts))
    return None


def main():
    chains, methods, params, solution, files = set(), set(), set(), set(), []
    for top in ("roman", "demo"):
        for dirpath, _, names in os.walk(os.path.join(REF, top)):
            for name in sorted(names):
                if not name.endswith(".py"):
                    continue
                path = os.path.join(dirpath, name)
                src = open(path).read()
                if "clipper" not in src:
                    continue
                tree = ast.parse(src)
                used = False
                for node in ast.walk(tree):
                    if isinstance(node, ast.Attribute):
                        d = dotted(node)
                        if not d:
                            continue
                        if d[0] == "clipperpy" and "()" not in d:
                            chains.add(".".join(d)); used = True
                        if d[0] == "clipper" and len(d) >= 2:
                            methods.add(d[1]); used = True
                            if "()" in d and d[1] == "get_solution" and d.index("()") + 1 < len(d):
                                solution.add(d[d.index("()") + 1])
                        if "iparams" in d[:-1] and isinstance(getattr(node, "ctx", None), ast.Store):
                            params.add(d[-1]); used = True
                if used:
                    files.append(os.path.relpath(path, REF))
    # keep only the longest chains (clipperpy.invariants is implied by clipperpy.invariants.ROMAN)
    longest = sorted(c for c in chains if not any(o != c and o.startswith(c + ".") for o in chains))
    doc = dict(source="mit-acl/roman@c921317, AST scan of roman/ and demo/", files=sorted(files), chains=longest,
               methods=sorted(methods), params=sorted(params), solution=sorted(solution))
    with open(os.path.join(HERE, "clipperpy_surface.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
