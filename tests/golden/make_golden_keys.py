"""Generates tests/golden/state_dict_keys.json from the UNMODIFIED reference: for every golden case the reference
module tree's state_dict keys -> (shape, dtype), which parameters require grad after freeze_backbones(stage) for the
three stages, and the module-key lists the training strategy reads (all_module_keys / trainable_module_keys).

    python tests/golden/make_golden_keys.py       # needs /root/reference
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import ref_shim  # noqa: E402
from make_golden import CASES, build_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ns = ref_shim.load()
    rec = {}
    for name, c in CASES.items():
        torch.manual_seed(0)
        mla = build_reference(ns, c)
        sd = mla.state_dict()
        r = {"state_dict": {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()},
             "all_module_keys": list(mla.all_module_keys), "stages": {}}
        stages = ["pretrain", "finetune"] + (["post-training"] if c.get("gen") else [])
        for st in stages:
            mla.vlm.freeze_backbones(st)
            r["stages"][st] = {"trainable_module_keys": list(mla.trainable_module_keys),
                               "requires_grad": sorted(k for k, p in mla.named_parameters() if p.requires_grad)}
        rec[name] = r
        print(name, len(r["state_dict"]), "keys", {s: len(v["requires_grad"]) for s, v in r["stages"].items()})
    import inspect

    def sig(fn):
        out = []
        for n, p in inspect.signature(fn).parameters.items():
            d = p.default
            d = None if d is inspect.Parameter.empty else (d if isinstance(d, (int, float, bool, str, type(None))) else repr(d))
            out.append([n, str(p.kind).split(".")[-1], p.default is not inspect.Parameter.empty, d])
        return out
    rec["__signatures__"] = {"MLA.__init__": sig(ns.MLA.__init__), "MLA.forward": sig(ns.MLA.forward),
                             "PrismaticVLM.__init__": sig(ns.PrismaticVLM.__init__),
                             "PrismaticVLM.forward": sig(ns.PrismaticVLM.forward),
                             "MLA.predict_action_diff": sig(ns.MLA.predict_action_diff),
                             "MLA.create_ddim": sig(ns.MLA.create_ddim)}
    json.dump(rec, open(os.path.join(OUT, "state_dict_keys.json"), "w"), indent=0, sort_keys=True)
    print("wrote", os.path.getsize(os.path.join(OUT, "state_dict_keys.json")), "bytes")


if __name__ == "__main__":
    main()
