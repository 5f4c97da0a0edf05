#!/bin/bash
# 1-CTA vs 2-CTA tcgen05 scan in the sustained (back-to-back, power-capped) regime of bench.py, alternating on one box
for rep in 1 2; do
  for ctas in 1 2; do
    YAMS_B200_UMMA_CTAS=$ctas python bench.py --workload knn --steps 40 --warmup 5 --no-side --no-parity --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('ctas=$ctas rep=$rep value %.0f q/s  ms/step %.3f  kernel %.3f ms  %.0f TF/s  clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['tensor_view']['achieved'], d.get('clocks')))"
  done
done
