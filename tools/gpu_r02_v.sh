#!/bin/bash
# round 2, call v: frame-domain Gram as diagonal window sums; memcheck of the WPE tests; timing
cd /root/repo
mkdir -p gpurun_out
python -c "from distant_speech_recognition_b200 import _capi" || exit 1
timeout 900 python -m pytest tests/test_parity_gpu_r2.py tests/test_parity_gpu.py tests/test_zz_host_surface.py tests/test_btk20_api.py -q -x -m gpu -k "wpe or dereverb" 2>&1 | tail -15 > gpurun_out/v_tests.txt
cat gpurun_out/v_tests.txt
: > gpurun_out/v_wpe.jsonl
timeout 300 python tools/bench_wpe.py >> gpurun_out/v_wpe.jsonl 2> gpurun_out/v_wpe.err
cat gpurun_out/v_wpe.jsonl; tail -3 gpurun_out/v_wpe.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu_r2.py tests/test_parity_gpu.py -q -x -m gpu -k "wpe" > gpurun_out/v_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/v_memcheck.txt
tail -5 gpurun_out/v_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu_r2.py -q -x -m gpu -k "wpe" > gpurun_out/v_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/v_racecheck.txt
tail -5 gpurun_out/v_racecheck.txt
WPE_FORMS=frame WPE_PREC=fp64 WPE_U=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_wpe" -c 40 --csv --log-file gpurun_out/v_wpe_launches.csv python tools/bench_wpe.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/v_wpe_launches.csv")) if len(r) > 10]
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    k = r[ki][:60]; a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in agg.items(): print("%-62s launches %3d  total %10.1f us" % (k, n, t / 1e3))
PY
