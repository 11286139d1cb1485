"""Generates tests/golden/bop_io_golden.json by running the reference's own ``get_model_params`` / ``get_target_list``
(/root/reference/tools/bop_io.py:9-42).  The module imports the (empty) bop_toolkit submodule at the top, so the two
function definitions are compiled out of the file in memory (``ast``) with ``inout.load_json`` bound to ``json.load``.
Run in the build container:  python tests/golden/make_bop_io_golden.py"""
import ast
import json
import os
import types

import numpy as np

SRC = "/root/reference/tools/bop_io.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    rng = np.random.RandomState(0)
    targets = []
    for scene in (2, 2, 48):
        for im in sorted(rng.choice(50, 3, replace=False).tolist()):
            for obj in sorted(rng.choice(np.arange(1, 16), rng.randint(1, 5), replace=False).tolist()):
                targets.append({"im_id": int(im), "inst_count": int(rng.randint(1, 4)), "obj_id": int(obj), "scene_id": int(scene)})
    params = {str(o): {"x_scale": float(rng.uniform(20, 90)), "y_scale": float(rng.uniform(20, 90)), "z_scale": float(rng.uniform(20, 90)),
                       "x_ct": float(rng.normal(0, 3)), "y_ct": float(rng.normal(0, 3)), "z_ct": float(rng.normal(0, 3))} for o in (1, 5, 12)}
    return targets, params


def main():
    tree = ast.parse(open(SRC).read())
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("get_model_params", "get_target_list")]
    targets, params = cases()
    tpath = os.path.join(HERE, "_targets_tmp.json")
    json.dump(targets, open(tpath, "w"))
    ns = {"np": np, "inout": types.SimpleNamespace(load_json=lambda p: json.load(open(p)))}
    exec(compile(ast.Module(body=fns, type_ignores=[]), SRC, "exec"), ns)
    out = {"targets": targets, "params": params, "target_list": ns["get_target_list"](tpath),
           "obj_param": {k: ns["get_model_params"](v).tolist() for k, v in params.items()}}
    os.remove(tpath)
    json.dump(out, open(os.path.join(HERE, "bop_io_golden.json"), "w"))
    print("wrote bop_io_golden.json:", len(out["target_list"]), "images")


if __name__ == "__main__":
    main()
