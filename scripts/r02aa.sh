#!/bin/bash
# r02aa: compute-sanitizer over the round-2 kernels (memcheck on a broad selection, racecheck on the shared-memory kernels)
TAG=r02aa
mkdir -p gpurun_out
SEL="convolve_and_envelope or envelope_long or tma_staged or ray_tree or elevational or edge_sizes or windowed_accumulate or scanline_block or config5 or streams_are_ordered or moving_and_deforming or cast_rays_ircad or full_frame or batched_poses"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" 2>&1 | tail -6 | tee gpurun_out/${TAG}_memcheck.log
RSEL="convolve_and_envelope or envelope_long or tma_staged or ray_tree or windowed_accumulate or cast_rays_ircad"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$RSEL" 2>&1 | tail -6 | tee gpurun_out/${TAG}_racecheck.log
