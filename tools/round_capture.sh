#!/bin/bash
# Round-end measurement set on one B200 (run through gpurun): GPU tests, bench line, ncu launch list of the bench
# command, --set full captures of the tensor-core value_and_grad kernels.  Usage: tools/round_capture.sh <tag>
tag=${1:-rX}
mkdir -p gpurun_out
(time timeout 800 python -m pytest tests -m gpu -x -q) > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
(time timeout 400 python bench.py) > gpurun_out/${tag}_bench.log 2>&1; tail -c 1800 gpurun_out/${tag}_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --latency-ticks 20 > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mpc_tc --launch-skip 1 --launch-count 1 -f \
    -o gpurun_out/prof_${tag}_tcgrad python tools/tc_profile.py iris 65536 1 > gpurun_out/${tag}_ncu_tcgrad.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_${tag}_tcgrad.ncu-rep gpurun_out/${tag}_tc_grad_b65536_ncu.txt > /dev/null 2>&1
rm -f gpurun_out/prof_${tag}_tcgrad.ncu-rep      # the reports are ~40 MB each: only the summaries travel back (64 MiB limit)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mpc_tc --launch-skip 1 --launch-count 1 -f \
    -o gpurun_out/prof_${tag}_tc_hexa8 python tools/tc_profile.py hexa 8192 1 8 > gpurun_out/${tag}_ncu_tchexa.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_${tag}_tc_hexa8.ncu-rep gpurun_out/${tag}_tc_grad_hexa_p8_ncu.txt > /dev/null 2>&1
rm -f gpurun_out/prof_${tag}_tc_hexa8.ncu-rep
ls -la gpurun_out
