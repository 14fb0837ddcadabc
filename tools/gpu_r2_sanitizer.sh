#!/bin/bash
# round 2: compute-sanitizer over the kernels written this round (ring rows / columns, tcgen05 block DCT and GEMM, motion spectrograms)
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
run() { # name, tool, pytest -k expression, file
  timeout 900 $SAN --tool $2 --error-exitcode 77 --print-limit 5 python -m pytest $4 -m gpu -x -q -k "$3" > gpurun_out/san_$1.log 2>&1
  echo "$1 ($2): rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/san_$1.log | tr '\n' ' ')"
}
run mem_ring memcheck "ring_row_kernel_planar and (shape0 or shape3 or shape9 or shape12) or ring_column_subpasses_small_panels and (shape0 or shape3) or ring_row_kernel_strided" tests/test_gpu_round2.py
run mem_blockmm memcheck "block_dct2d_tensor_cores_vs_oracle or block_dct2d_ragged" tests/test_gpu_round2.py
run mem_gemm memcheck "zoom_dense_path_tensor_core_gemm" tests/test_gpu_round2.py
run mem_motion memcheck "motion_spectrogram_modes or motion_tiled_c_session" tests/test_gpu_round2.py
run race_ring racecheck "ring_row_kernel_planar and (shape0 or shape3)" tests/test_gpu_round2.py
run init_blockmm initcheck "block_dct2d_tensor_cores_vs_oracle and 8" tests/test_gpu_round2.py
