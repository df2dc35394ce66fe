#!/bin/bash
# profiles/sweep_subs.sh -- e2e throughput of nhw_encode_batch / nhw_decode_batch against the sub-chunk plan
# (NHW_SUBS_* x NHW_LANES_*), run under gpurun; prints one line per setting.
for s in 8 12 16; do for l in 4; do
	NHW_SUBS_ENCODE=$s NHW_LANES_ENCODE=$l NHW_SUBS_DECODE=$s python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('subs $s lanes $l: value', d['value'], 'e2e', d['e2e']['value'], 'h2d_alone_GBps', d['e2e'].get('h2d_alone_GBps'), 'decode', d['decode']['value'])"
done; done
