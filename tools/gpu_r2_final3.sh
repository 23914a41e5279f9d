#!/bin/bash
O=/root/repo/gpurun_out/r2final3
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
timeout 600 $S quick > $O/selftest_quick.log 2>&1; tail -2 $O/selftest_quick.log; grep -c " OK" $O/selftest_quick.log; grep FAIL $O/selftest_quick.log
cat > /tmp/pk.py <<'PY'
import torch, sys
sys.path.insert(0, '/root/repo')
import vl_merging_b200 as vlm
from vl_merging_b200 import _lib
dims = [33, 192, 256]
gs = [torch.randn(d, d, device='cuda') for d in dims]
sizes = [d * (d + 1) // 2 for d in dims]
flat = torch.empty(sum(sizes), device='cuda')
items = (_lib.SymItem * 3)()
off = 0
for it, g, d, sz in zip(items, gs, dims, sizes):
    it.full, it.packed, it.d, it.ld = g.data_ptr(), flat.data_ptr() + 4 * off, d, d
    off += sz
L, st = _lib.lib(), torch.cuda.current_stream().cuda_stream
_lib.check(L.vlm_sym_pack_upper_batch(items, 3, 0, st)); _lib.check(L.vlm_sym_unpack_batch(items, 3, 0, st))
torch.cuda.synchronize()
print('batch pack/unpack ok', all(torch.equal(g, g.T) for g in gs))
PY
for tool in memcheck; do timeout 300 compute-sanitizer --tool $tool python /tmp/pk.py 2>&1 | grep -E "ERROR SUMMARY|batch pack" | tee -a $O/sanitizer.log; done
