#!/bin/bash
# End-of-round evidence on one B200: GPU tests, smoke, the default bench line (+ reference arm), the ncu launch list of the
# same command and a --set full capture of each BASELINE kernel, summarised on the box (the .ncu-rep files are too large
# to bring back together).  Writes gpurun_out/.
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/bench_r2_final_1gpu.json 2> gpurun_out/bench_r2_final_1gpu.err; tail -c 300 gpurun_out/bench_r2_final_1gpu.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_final_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r2_final_ref.json
timeout 1300 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
cap() { n=$1; rx=$2; shift 2; timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -f -o /tmp/prof_r2_final_$n python "$@" > /dev/null 2>&1; }
cap C4 tlm_kernel scripts/profile_tlm.py 592
cap C1 tps_solve_kernel scripts/profile_solve.py C1 1048576 1
cap C3 newton_refill scripts/profile_solve.py C3 1048576 1
cap C5 coop_broyden scripts/profile_solve.py C5 16384 1
python scripts/ncu_summary.py gpurun_out/r2_final_ncu.md gpurun_out/r2_final_launches.csv /tmp/prof_r2_final_C1.ncu-rep /tmp/prof_r2_final_C3.ncu-rep /tmp/prof_r2_final_C4.ncu-rep /tmp/prof_r2_final_C5.ncu-rep
python scripts/ncu_lines.py /tmp/prof_r2_final_C4.ncu-rep tlm_kernel 50 > gpurun_out/r2_final_C4_lines.txt 2>&1
python scripts/ncu_lines.py /tmp/prof_r2_final_C5.ncu-rep coop_broyden 50 > gpurun_out/r2_final_C5_lines.txt 2>&1
for n in C1 C3 C4 C5; do ncu -i /tmp/prof_r2_final_$n.ncu-rep --page raw --csv > gpurun_out/r2_final_${n}_raw.csv 2>/dev/null; done
cp /tmp/prof_r2_final_C4.ncu-rep gpurun_out/
du -sh gpurun_out
