"""Extract the argparse surface of the reference's three sampling scripts WITHOUT importing them (they need
pytorch_lightning / omegaconf / taming): walk the AST for `<parser>.add_argument(...)` calls and evaluate their
literal arguments.  Build container only; writes tests/golden/cli_surface.json, which the CPU tests diff against
qdiff_b200.cli (flag names, defaults, types, nargs, choices, required, action).

    python tools/make_cli_golden.py
"""
import ast
import json
import os
import sys

REF = "/root/reference/scripts"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = {"ddim": "sample_diffusion_ddim.py", "ldm": "sample_diffusion_ldm.py", "txt2img": "txt2img.py"}


def extract(path):
    tree = ast.parse(open(path).read())
    out = {}
    for node in ast.walk(tree):
        if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "add_argument"):
            continue
        flags = [ast.literal_eval(a) for a in node.args]
        kw = {}
        for k in node.keywords:
            if k.arg == "type":
                kw["type"] = k.value.id if isinstance(k.value, ast.Name) else ast.unparse(k.value)
            elif k.arg == "help":
                continue
            else:
                kw[k.arg] = ast.literal_eval(k.value)
        long = [f for f in flags if f.startswith("--")]
        dest = (long[0] if long else flags[0]).lstrip("-").replace("-", "_")
        action = kw.get("action", "store")
        default = kw.get("default", False if action == "store_true" else None)
        out[dest] = dict(flags=sorted(flags), default=default, type=kw.get("type"), nargs=kw.get("nargs"),
                         choices=kw.get("choices"), required=bool(kw.get("required", False)), action=action)
    return out


def main():
    surf = {k: extract(os.path.join(REF, f)) for k, f in FILES.items()}
    path = os.path.join(ROOT, "tests", "golden", "cli_surface.json")
    json.dump(surf, open(path, "w"), indent=1, sort_keys=True)
    print({k: len(v) for k, v in surf.items()}, "->", path)


if __name__ == "__main__":
    sys.exit(main())
