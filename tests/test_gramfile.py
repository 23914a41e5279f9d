"""Packed Gram container (vl-merging_b200/gramfile.py): the format itself on the CPU.
The packing / unpacking runs on the GPU only (tests/test_gpu_gramfile.py)."""
import json
import struct

import numpy as np
import pytest
import torch

import vl_merging_b200 as vlm
from vl_merging_b200 import gramfile


def _write_by_hand(path, grams, dtype="float32"):
    """The format, restated from its specification with numpy (row-major upper triangle, fp32 or fp64)."""
    entries, blobs, off = [], [], 0
    for name, g in grams.items():
        d = g.shape[0]
        blobs.append(g[np.triu_indices(d)].astype("<f4" if dtype == "float32" else "<f8"))
        entries.append({"name": name, "d": d, "offset": off, "rows": 7, "calls": 1})
        off += d * (d + 1) // 2
    header = json.dumps({"version": 1, "dtype": dtype, "layout": "upper_rowmajor", "entries": entries}).encode()
    pre = b"VLMGRAM1" + struct.pack("<Q", len(header)) + header
    with open(path, "wb") as f:
        f.write(pre + b"\0" * ((-len(pre)) % 4096))
        for b in blobs:
            f.write(b.tobytes())
    return entries


def test_header_round_trip_and_detection(tmp_path):
    rng = np.random.default_rng(0)
    grams = {"transformer.blocks.0.attn.v": rng.standard_normal((5, 5)), "x.fc2": rng.standard_normal((33, 33))}
    p = tmp_path / "g.vlmgram"
    want = _write_by_hand(p, grams)
    assert gramfile.is_packed_file(p)
    entries, off = gramfile.read_header(p)
    assert entries == want and off % 4096 == 0
    assert p.stat().st_size == off + 4 * (15 + 33 * 34 // 2)
    ref = tmp_path / "ref.pth"
    torch.save({"a": torch.zeros(2, 2)}, ref)
    assert not gramfile.is_packed_file(ref)
    assert not gramfile.is_packed_file(tmp_path / "missing")
    with pytest.raises(ValueError):
        gramfile.read_header(ref)


def test_fp64_variant_of_the_header(tmp_path):
    """dtype float64 (the container of the RegMean-grade caches): same layout, 8-byte values; anything else is refused."""
    rng = np.random.default_rng(1)
    grams = {"a": rng.standard_normal((7, 7)), "b": rng.standard_normal((40, 40))}
    p = tmp_path / "g64.vlmgram"
    want = _write_by_hand(p, grams, dtype="float64")
    entries, off, dtype = gramfile.read_header(p, with_dtype=True)
    assert entries == want and dtype == torch.float64
    assert p.stat().st_size == off + 8 * (28 + 40 * 41 // 2)
    _, _, d32 = gramfile.read_header(_write_and_return(tmp_path / "g32.vlmgram", grams), with_dtype=True)
    assert d32 == torch.float32
    bad = tmp_path / "bad.vlmgram"
    _write_by_hand(bad, grams, dtype="float16")
    with pytest.raises(ValueError):
        gramfile.read_header(bad)
    assert gramfile.packed_bytes([768] * 72 + [3072] * 24, torch.float64) == 2 * 538_177_536


def _write_and_return(path, grams):
    _write_by_hand(path, grams)
    return path


def test_sizes_against_the_reference_file():
    base = [768] * 72 + [3072] * 24          # 96 Grams, 2.15 GB as fp64 full matrices
    large = [1024] * 144 + [4096] * 48
    assert gramfile.packed_bytes(base) == 538_177_536
    assert gramfile.packed_bytes(base) < 0.2502 * 8 * sum(d * d for d in base)
    assert gramfile.packed_bytes(large) < 0.2502 * 8 * sum(d * d for d in large)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_a_gpu(tmp_path):
    with pytest.raises((RuntimeError, AssertionError)):
        gramfile.save_packed({"a": torch.eye(4)}, tmp_path / "g")
    _write_by_hand(tmp_path / "h", {"a": np.eye(4)})
    with pytest.raises((RuntimeError, AssertionError)):
        gramfile.load_packed(tmp_path / "h")
    with pytest.raises(RuntimeError):
        gramfile.save_packed({"a": torch.eye(4)}, tmp_path / "g", device="cpu")
