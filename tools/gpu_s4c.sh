#!/bin/bash
TAG=${1:-s4c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "yuv444p_to_rgb_and_packed or chroma_resampling or packed422_to_planar or clamping_switch or yuv_family" > gpurun_out/pytest_$TAG.log 2>&1; grep -E "^E  |passed|failed|Error" gpurun_out/pytest_$TAG.log | head -40
