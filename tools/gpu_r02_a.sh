#!/bin/bash
# Round 2, GPU visit A: (1) the whole GPU suite without any xfail marker, (2) A/B of the packed 2 x fp32 kernels (CUDA-event kernel
# times, 5 combinations), (3) bench line, (4) ncu launch list + --set full of the packed filter-bank kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rfEs -x --deselect tests/test_zz_host_surface.py::test_packed_fp32_kernel_equals_the_scalar_kernel 2>&1 | tail -25 > gpurun_out/a_pytest_gpu.txt; cat gpurun_out/a_pytest_gpu.txt
timeout 900 python -m pytest tests/test_zz_host_surface.py -m gpu -q -rfE -k packed 2>&1 | tail -60 > gpurun_out/a_pytest_packed.txt; tail -40 gpurun_out/a_pytest_packed.txt
for v in "0 0 0" "1 0 0" "0 1 0" "0 0 1" "1 1 1"; do set -- $v
  echo "== BTKB_ANALYSIS_PACKED=$1 BTKB_SYNTHESIS_PACKED=$2 BTKB_PERBIN_PACKED=$3"
  BTKB_ANALYSIS_PACKED=$1 BTKB_SYNTHESIS_PACKED=$2 BTKB_PERBIN_PACKED=$3 timeout 300 python tools/prof_step.py 10 | tee -a gpurun_out/a_ab_packed.jsonl
done
for fr in 8 12; do echo "== FR=$fr packed"; BTKB_ANALYSIS_FR=$fr BTKB_ANALYSIS_PACKED=1 timeout 300 python tools/prof_step.py 10 | tee -a gpurun_out/a_ab_packed.jsonl; done
BTKB_ANALYSIS_PACKED=1 BTKB_SYNTHESIS_PACKED=1 BTKB_PERBIN_PACKED=1 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_packed.json 2> gpurun_out/a_bench.err; tail -3 gpurun_out/a_bench.err; cat gpurun_out/a_bench_packed.json
export BTKB_ANALYSIS_PACKED=1 BTKB_SYNTHESIS_PACKED=1 BTKB_PERBIN_PACKED=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/a_launches_packed.csv python tools/prof_step.py 3 > gpurun_out/a_under_ncu.log 2>&1
for k in k_analysis k_perbin k_synthesis; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/a_prof_packed_$k python tools/prof_step.py 1 > gpurun_out/a_ncu_$k.log 2>&1
done
ls -la gpurun_out | head -40
