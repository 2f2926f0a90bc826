#!/bin/bash
# Lattice kernel mapping at long label sequences (experiment build): nodes per lane K and chunk length CH, forced through
# B200CTC_LAT_K / B200CTC_LAT_CH, on the T=3200,L=320 and T=1600,L=160 points of tools/sweep.py.  -> profiles/rN_lattice_sweep.txt
cd "$(dirname "$0")/.."
export B200CTC_EXPERIMENT=1
echo "# tools/lattice_sweep.sh: forward time is the lattice's at these sizes (the softmax/gather kernel takes ~0.4 / 0.2 ms)"
echo "== default policy"; timeout -s KILL 120 python tools/sweep.py --only "cfg5 T=3200 V=3500" | tail -1; timeout -s KILL 120 python tools/sweep.py --only "cfg5 T=1600 V=3500" | tail -1
for k in 2 4 8; do for ch in 4 8; do
  echo "== B200CTC_LAT_K=$k B200CTC_LAT_CH=$ch"
  B200CTC_LAT_K=$k B200CTC_LAT_CH=$ch timeout -s KILL 120 python tools/sweep.py --only "cfg5 T=3200 V=3500" 2>&1 | tail -1
done; done
