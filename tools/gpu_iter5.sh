#!/bin/bash
TAG=${1:-it}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum --clock-control none -k regex:"k_chroma_upsample_packed|k_quad_chroma|k_copy2d" -c 30 --csv --log-file gpurun_out/ncu_upsample_$TAG.csv python tools/time_new_kernels.py > gpurun_out/ncu_upsample_$TAG.log 2>&1
python - <<'PY'
import csv,sys
rows=[r for r in csv.reader(open('gpurun_out/ncu_upsample_%s.csv' % sys.argv[1] if len(sys.argv)>1 else 'it')) if len(r)>10]
PY
grep -v "^==" gpurun_out/ncu_upsample_$TAG.csv | awk -F'","' 'NR>1{print $5, $(NF-3), $(NF-2), $NF}' | head -80
