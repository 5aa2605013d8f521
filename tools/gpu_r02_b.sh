#!/bin/bash
# Round 2, GPU visit B: K5 with the sliding-window polyphase stage (bit-identity vs scalar + parity), packed kernels as defaults, overlap probe.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "synthesis or packed or golden or ragged or round_trip or consecutive" 2>&1 | tail -8
timeout 300 python tools/prof_step.py 10 | tee gpurun_out/b_step.json
timeout 600 python tools/dbg/overlap_probe.py 30 | tee gpurun_out/b_overlap.json
