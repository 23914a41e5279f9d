"""gramfile.py — packed Gram container.

The reference writes its Gram matrices with `torch.save(middle_representations, ...)`
(src/cache_gram_matrices.py:349) and reads them back with `torch.load` in regmean
(src/vilt/modules/vilt_module.py:386): a pickle of full symmetric fp64 matrices — 2.15 GB for VLMo-base,
7.65 GB for ViT-L — although the accumulators here are fp32 and symmetric.  This container stores the
row-major UPPER TRIANGLE in fp32 (a quarter of the bytes) as one flat blob:

    bytes 0..7     magic  b"VLMGRAM1"
    bytes 8..15    little-endian uint64: length H of the JSON header
    bytes 16..16+H header: {"version": 1, "dtype": "float32", "layout": "upper_rowmajor",
                            "entries": [{"name", "d", "offset" (in elements), "rows", "calls"}, ...]}
    zero padding to a multiple of 4096
    fp32 data: for each entry d*(d+1)/2 values, row r = columns r..d-1

`GramCache.save_packed` / `save_packed` write it (pack kernel on the device, ONE device->host copy),
`load_packed` reads it back as fp32 device matrices for `regmean` (ONE host->device copy + unpack kernel),
and `export_reference` / `import_reference` convert to and from the reference's own file, so either side
can consume the other's artefact.  The packing kernels are vlm_sym_pack_upper(_batch) / vlm_sym_unpack(_batch).
"""
import json
import os
import struct
from collections import defaultdict

import numpy as np
import torch

from . import _lib

MAGIC = b"VLMGRAM1"
_ALIGN = 4096


def is_packed_file(path):
    try:
        with open(path, "rb") as f:
            return f.read(len(MAGIC)) == MAGIC
    except (OSError, TypeError):
        return False


def _packed_len(d):
    return d * (d + 1) // 2


def _device_of(device):
    device = torch.device(device if device is not None else "cuda")
    if device.type != "cuda":
        raise RuntimeError("the packed Gram container is packed / unpacked on the GPU: there is no CPU fallback")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def save_packed(grams, path, rows=None, calls=None, device=None):
    """grams: {name: (d, d) tensor} (any float dtype / device; only the upper triangles are read) or a
    GramCache.  Returns the number of bytes written."""
    if hasattr(grams, "buffers") and hasattr(grams, "live_names"):
        cache = grams
        cache.flush()
        device = cache.device
        rows, calls = cache.rows, cache.calls
        grams = {n: cache.buffers[n] for n in cache.live_names()}
    device = _device_of(device)
    lib = _lib.lib()
    names = list(grams.keys())
    entries, total = [], 0
    for n in names:
        d = int(grams[n].shape[0])
        if grams[n].dim() != 2 or grams[n].shape[1] != d:
            raise ValueError(f"{n}: Gram must be square, got {tuple(grams[n].shape)}")
        entries.append({"name": n, "d": d, "offset": total, "rows": int((rows or {}).get(n, 0)),
                        "calls": int((calls or {}).get(n, 0))})
        total += _packed_len(d)
    with torch.cuda.device(device):
        packed = torch.empty(total, dtype=torch.float32, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        items, keep = (_lib.SymItem * len(entries))(), []
        for it, e in zip(items, entries):
            g = grams[e["name"]].detach().to(device=device, dtype=torch.float32)
            if g.stride(1) != 1:
                g = g.contiguous()
            keep.append(g)
            it.full, it.packed, it.d, it.ld = g.data_ptr(), packed.data_ptr() + 4 * e["offset"], e["d"], g.stride(0)
        _lib.check(lib.vlm_sym_pack_upper_batch(items, len(entries), _lib.VLM_F32, stream))   # one launch for all Grams
        host = torch.empty(total, dtype=torch.float32, pin_memory=True)
        host.copy_(packed, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
    header = json.dumps({"version": 1, "dtype": "float32", "layout": "upper_rowmajor", "entries": entries}).encode()
    pre = MAGIC + struct.pack("<Q", len(header)) + header
    pad = (-len(pre)) % _ALIGN
    tmp = f"{path}.tmp.{os.getpid()}"
    try:
        with open(tmp, "wb") as f:
            f.write(pre + b"\0" * pad)
            f.write(memoryview(host.numpy()).cast("B"))
        os.replace(tmp, path)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    return len(pre) + pad + 4 * total


def read_header(path):
    """(entries, data_offset_in_bytes) of a packed Gram file."""
    with open(path, "rb") as f:
        if f.read(len(MAGIC)) != MAGIC:
            raise ValueError(f"{path} is not a packed Gram file (bad magic)")
        (hlen,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hlen).decode())
    if header.get("version") != 1 or header.get("layout") != "upper_rowmajor" or header.get("dtype") != "float32":
        raise ValueError(f"{path}: unsupported packed Gram header {header.get('version')}/{header.get('layout')}")
    pre = len(MAGIC) + 8 + hlen
    return header["entries"], pre + ((-pre) % _ALIGN)


def load_packed(path, device=None, dtype=torch.float32):
    """{name: full symmetric (d, d) device tensor} (fp32, or fp64 = the reference's dtype): what
    regmean(gram_matrices=...) takes.  One host->device copy of the blob, one unpack launch per Gram."""
    device = _device_of(device)
    if dtype not in (torch.float32, torch.float64):
        raise ValueError("dtype must be torch.float32 or torch.float64")
    entries, off = read_header(path)
    total = sum(_packed_len(e["d"]) for e in entries)
    size = os.path.getsize(path)
    if size < off + 4 * total:
        raise ValueError(f"{path}: truncated ({size} bytes, header promises {off + 4 * total})")
    lib = _lib.lib()
    host = torch.empty(total, dtype=torch.float32, pin_memory=True)
    view, got = memoryview(host.numpy()).cast("B"), 0
    with open(path, "rb", buffering=0) as f:
        f.seek(off)
        while got < 4 * total:           # one read() moves at most 2 GB
            n = f.readinto(view[got:got + (1 << 30)])
            if not n:
                raise ValueError(f"{path}: short read")
            got += n
    out = {}
    with torch.cuda.device(device):
        packed = host.to(device, non_blocking=True)
        stream = torch.cuda.current_stream(device).cuda_stream
        code = _lib.VLM_F32 if dtype == torch.float32 else _lib.VLM_F64
        items = (_lib.SymItem * len(entries))()
        for it, e in zip(items, entries):
            g = out[e["name"]] = torch.empty(e["d"], e["d"], dtype=dtype, device=device)
            if dtype == torch.float32:       # same element type on both sides: one launch for all Grams
                it.full, it.packed, it.d, it.ld = g.data_ptr(), packed.data_ptr() + 4 * e["offset"], e["d"], g.stride(0)
            else:                            # widening to the reference's fp64: one launch per Gram
                _lib.check(lib.vlm_sym_unpack(packed.data_ptr() + 4 * e["offset"], e["d"], g.data_ptr(), code,
                                              g.stride(0), stream))
        if dtype == torch.float32:
            _lib.check(lib.vlm_sym_unpack_batch(items, len(entries), _lib.VLM_F32, stream))
        torch.cuda.current_stream(device).synchronize()   # `packed` and `host` may be released after this
    return out


def export_reference(packed_path, reference_path, device=None):
    """Packed container -> the reference's Gram file: torch.save of {name: fp64 CPU (d, d)} in a defaultdict
    (src/cache_gram_matrices.py:236,349), consumable by the unmodified reference regmean."""
    grams = load_packed(packed_path, device, dtype=torch.float64)
    out = defaultdict(float)
    for k, v in grams.items():
        out[k] = v.cpu()
    torch.save(out, reference_path)
    return reference_path


def import_reference(reference_path, packed_path, device=None):
    """The reference's Gram file -> packed container (values rounded to fp32; the lower triangles are dropped)."""
    grams = torch.load(reference_path, map_location="cpu", weights_only=False)
    return save_packed({k: v for k, v in grams.items() if torch.is_tensor(v)}, packed_path, device=device)


def packed_bytes(dims):
    """Size of the data section for Grams of the given widths; e.g. VLMo-base IRTR: 72 x 768 + 24 x 3072."""
    return 4 * int(np.sum([_packed_len(int(d)) for d in dims]))
