"""Generate tests/golden/expansion_head.npz from the UNMODIFIED reference ``SegmentationHead``
(/root/reference/model/blocks/module.py:20-44; same structure as cluster1/cluster2 of model/dino_pqgo.py:104-128).

Run in the build container only:   PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_head.py

The script runs the reference module on seeded CPU inputs (one thread), asserts that the oracle restatement
(oracle/equss_oracle.py::expansion_head) reproduces it bit for bit, and stores inputs, parameters, the fp32
reference output and an fp64 evaluation of the same expression (the yardstick for the kernel's tolerance).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("EQUSS_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.dont_write_bytecode = True


def main():
    torch.set_num_threads(1)
    sys.path.insert(0, HERE)
    import equss_oracle as O
    spec = importlib.util.spec_from_file_location("ref_blocks_module", os.path.join(REF, "model", "blocks", "module.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(11)
    B, C, D, h, w = 2, 64, 96, 8, 8             # h*w = 64: the NCHW tensor-map path (multiple of 32)
    head = mod.SegmentationHead(C, D).eval()
    x = torch.randn(B, C, h, w)
    with torch.no_grad():
        ref = head(x)
        p = {k: v.detach().clone() for k, v in head.state_dict().items()}
        args = (p["cluster1.0.weight"], p["cluster1.0.bias"], p["cluster2.0.weight"], p["cluster2.0.bias"],
                p["cluster2.2.weight"], p["cluster2.2.bias"])
        ora = O.expansion_head(x, *args)
        assert torch.equal(ora, ref), f"oracle != reference: {(ora - ref).abs().max()}"
        ref64 = O.expansion_head(x.double(), *[a.double() for a in args])
    fix = {"x": x.numpy(), "out": ref.numpy(), "out_fp64": ref64.numpy()}
    fix.update({k.replace(".", "_"): v.numpy() for k, v in p.items()})
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "expansion_head.npz"), **fix)
    print("expansion_head.npz", {k: v.shape for k, v in fix.items()},
          "fp32 reference vs fp64: max rel", float((ref.double() - ref64).abs().max() / ref64.abs().max()))


if __name__ == "__main__":
    main()
