#!/bin/bash
# round-2 GPU check B: RegMean chain test, split-mode segment caps, ncu captures of the 2-SM kernel, full GPU suite
O=/root/repo/gpurun_out/r2b
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
timeout 900 python -m pytest tests/test_gpu_regmean_chain.py -x -q -s > $O/chain.log 2>&1; echo "chain exit $?" >> $O/chain.log
tail -25 $O/chain.log
for cap in 128 64 32 16; do
  echo "== split cap $cap" >> $O/split_caps.log
  VLM_SYRK_SEG_CHUNKS=$cap timeout 120 $S split 36928 3072 10 1 >> $O/split_caps.log 2>&1
  VLM_SYRK_SEG_CHUNKS=$cap timeout 120 $S split 36928 768 20 0 >> $O/split_caps.log 2>&1
done
cat $O/split_caps.log
ncu --set full --clock-control none --import-source on -k regex:syrk_2sm -s 1 -c 1 -o $O/syrk_2sm_f32_36928x3072 $S case f32 36928 3072 0 1 > $O/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:syrk_2sm -s 1 -c 1 -o $O/syrk_2sm_f32_36928x768 $S case f32 36928 768 0 0 > $O/ncu2.log 2>&1
ncu --set full --clock-control none -k regex:syrk_2sm -s 1 -c 1 -o $O/syrk_2sm_bf16_36928x3072 $S case bf16 36928 3072 0 1 > $O/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:syrk_2sm -s 0 -c 1 -o $O/syrk_2sm_split_36928x3072 $S split 36928 3072 0 1 > $O/ncu4.log 2>&1
ls -la $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "gpu tests exit $?" >> $O/gpu_tests.log
tail -15 $O/gpu_tests.log
