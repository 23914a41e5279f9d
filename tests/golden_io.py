"""Loader for tests/golden/merge_small.npz (written by oracle/make_golden.py from the reference)."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class MergeGolden:
    def __init__(self, path=None):
        z = np.load(path or os.path.join(GOLDEN_DIR, "merge_small.npz"))
        self.meta = json.loads(bytes(z["meta"]).decode())
        self.sd = {k[len("in/sd/"):]: z[k] for k in z.files if k.startswith("in/sd/")}
        shared = {k[len("in/sd_shared11/"):]: z[k] for k in z.files if k.startswith("in/sd_shared11/")}
        self.sd_shared11 = {k: v for k, v in self.sd.items()
                            if not (k.startswith("transformer.blocks.11.") and "gamma" not in k)}
        self.sd_shared11.update(shared)
        self.central = {k[len("central/"):]: z[k] for k in z.files if k.startswith("central/")}
        self.grams = {k[len("gram/"):]: z[k] for k in z.files if k.startswith("gram/")}
        self.grams_missing = {k: v for k, v in self.grams.items()
                              if not (k.startswith("transformer.blocks.3.") and ".l" in k)}
        self._z = z

    @property
    def variants(self):
        return self.meta["variants"]

    def inputs(self, vname):
        v = self.variants[vname]
        sd = self.sd if v["input"] == "sd" else self.sd_shared11
        grams = self.grams_missing if v["cfg"].get("gram_matrices") == "missing" else self.grams
        return sd, v["cfg"], grams

    def expected(self, vname):
        pre = f"out/{vname}/"
        return {k[len(pre):]: self._z[k] for k in self._z.files if k.startswith(pre)}
