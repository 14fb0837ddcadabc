#!/bin/bash
bash tools/gpu_r2_u.sh
bash tools/gpu_r2_v.sh
bash tools/gpu_r2_q.sh
