#!/bin/bash
# LM cluster-size experiment: correspondences per CTA of a k_lm cluster (PLB_LM_PER_CTA), bench defaults otherwise
for v in ${LMS:-2048 4096 8192 16384}; do
  PLB_LM_PER_CTA=$v timeout 300 python bench.py --steps 5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('per_cta $v', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']))"
done
