#!/bin/bash
# k_fused3 at 32 frames per launch under different band heights: event-timed fps (twice), then DRAM bytes of one launch under ncu
for m in ${@:-"" 1}; do
  r=""
  for i in 1 2; do r="$r $(PE_F3_BAND_MODE=$m timeout 300 python bench.py --batch 32 --steps 100 --warmup 10 --no-cpu-baseline --no-sub-records --e2e-frames 2 --e2e-steps 1 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f us %.0f fps' % (d['ms_per_step']*1e3, d['value']))")"; done
  PE_F3_BAND_MODE=$m timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_fused3 -s 3 -c 1 --csv --log-file gpurun_out/band_$m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-records --e2e-frames 2 --e2e-steps 1 > /dev/null 2>&1
  echo "[band mode '$m'] $r | $(python tools/ncu_compact.py gpurun_out/band_$m.csv | cut -c25-130)"
done
