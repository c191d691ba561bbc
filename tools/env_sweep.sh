#!/bin/bash
# usage: tools/env_sweep.sh "VAR1=a VAR2=b" "VAR1=c" ...   -> one step_diag line per setting
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg python tools/step_diag.py 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); s=d['stage_ms']
print('gpu_ms/step %.3f  fwd %.3f  bwd %.3f  sort %.3f  mallocs %d' % (d['gpu_ms_per_step'], s['render_fwd'], s['render_bwd'], s['sort'], d['mem_delta']['num_device_alloc']))"
done
