#!/bin/bash
# First GPU visit of round 2: (1) the whole GPU suite, with the surface tests that round 1 could not run any more reported
# separately (tests/test_zz_host_surface.py: XPASS = passes, drop the UNVERIFIED marker; XFAIL = look at gpurun_out/zz.txt),
# (2) smoke, (3) bench + reference arm with the full-batch CPU sample, (4) launch list under ncu.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rxX 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 600 python -m pytest tests/test_zz_host_surface.py -m gpu -q --runxfail 2>&1 | tail -60 > gpurun_out/zz.txt; tail -15 gpurun_out/zz.txt
timeout 600 python tools/parity_report_kinect.py > gpurun_out/parity_kinect.json 2> gpurun_out/parity_kinect.err; tail -2 gpurun_out/parity_kinect.err; cat gpurun_out/parity_kinect.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
# A/B of the packed 2 x fp32 filter-bank kernels (bit-identical results; see tests/test_fft_packed_host.py)
for v in "0 0 0" "1 0 0" "0 1 0" "0 0 1" "1 1 1"; do set -- $v
  echo "== BTKB_ANALYSIS_PACKED=$1 BTKB_SYNTHESIS_PACKED=$2 BTKB_PERBIN_PACKED=$3"
  BTKB_ANALYSIS_PACKED=$1 BTKB_SYNTHESIS_PACKED=$2 BTKB_PERBIN_PACKED=$3 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_packed_$1$2$3.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','kernel_ms_per_step')}, d['roofline']['all_kernels_frac'], d['e2e']['ms_per_step'])"
done
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out | head -30
