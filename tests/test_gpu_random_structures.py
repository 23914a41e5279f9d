"""GPU: the randomised checkpoint structures of tests/test_random_structures.py through the public merge API
(C ABI underneath) against the oracle — bit-exact for the fp32 merges, 1e-9 for RegMean's fp64 linears (both
sides fp64 on identical Grams), same key order, host and device inputs."""
import numpy as np
import pytest
import torch

import oracle
import vl_merging_b200 as vlm
from test_random_structures import SEEDS, random_case

pytestmark = pytest.mark.gpu


def _t(d, device):
    return {k: torch.from_numpy(np.array(v)).to(device) for k, v in d.items()}


def _same(got, want, fp64_tol=1e-9):
    assert list(got) == list(want)
    for k, w in want.items():
        g = got[k].cpu().numpy()
        w = np.asarray(w)
        assert g.dtype == w.dtype and g.shape == w.shape, k
        if w.dtype == np.float32:
            assert np.array_equal(g, w), k
        else:
            assert np.linalg.norm(g - w) <= fp64_tol * max(np.linalg.norm(w), 1e-300), k


@pytest.mark.parametrize("seed", SEEDS)
def test_merges_on_random_structures(seed):
    sd, central, grams, cfg = random_case(seed)
    device = "cuda" if seed % 2 else "cpu"
    _same(vlm.merge_weights(_t(sd, device), cfg), oracle.merge_weights(sd, cfg))
    tc = _t(central, device)
    before = {k: v.clone() for k, v in tc.items()}
    _same(vlm.sum_task_vectors(_t(sd, device), cfg, central_weight=tc),
          oracle.sum_task_vectors(sd, {k: v.copy() for k, v in central.items()}, cfg))
    assert all(torch.equal(before[k], tc[k]) for k in tc)
    want = oracle.regmean(sd, grams, cfg)
    if any(isinstance(v, int) for v in want.values()):
        with pytest.raises(KeyError):           # documented deviation (the reference stores the integer 0)
            vlm.regmean(_t(sd, device), cfg, gram_matrices=_t(grams, device))
        return
    _same(vlm.regmean(_t(sd, device), cfg, gram_matrices=_t(grams, device)), want)
