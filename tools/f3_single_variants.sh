#!/bin/bash
# k_fused3 launches of 1 .. 32 frames under different static / dynamic splits and row costs (PE_F3_STATIC_PCT, PE_F3_COST_I / _B, ...):
# usage: [BATCHES="1 2 4 8 32"] tools/f3_single_variants.sh ["ENV=.. ENV=.." ...]   (no argument: the defaults)
[ $# -eq 0 ] && set -- ""
for v in "$@"; do
  for b in ${BATCHES:-1 2 4 8 32}; do
    r=$(env $v timeout 300 python bench.py --batch $b --steps 200 --warmup 20 --no-cpu-baseline --no-sub-records --e2e-frames 2 --e2e-steps 1 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f us/launch  %.0f fps' % (d['ms_per_step']*1e3, d['value']))")
    echo "[$v] batch $b: $r"
  done
done
