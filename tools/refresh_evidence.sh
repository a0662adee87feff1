#!/bin/bash
# Final evidence of a round, run on the GPU box AFTER the last kernel change (the build hash ties the ncu capture to the sources):
#   gpurun --timeout 1500 -- bash tools/refresh_evidence.sh
# -> gpurun_out/final_*: GPU test log, ncu --set full capture of the headline kernel (+ raw CSV, traffic record with kernel name and
#    build hash), ncu launch list of the bench command, bench lines of both arms.  tools/refresh_evidence.py (run afterwards in
#    the build container) copies the summaries into profiles/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/final_gpu_tests.log; tail -1 gpurun_out/final_gpu_tests.log
ncu --set full --clock-control none --import-source on -k regex:k_objective_hot -s 5 -c 1 -f -o gpurun_out/final_hot python tools/time_objective.py > /dev/null 2>&1
ncu -i gpurun_out/final_hot.ncu-rep --page raw --csv > gpurun_out/final_hot_ncu_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches_ncu.csv python bench.py --steps 5 --warmup 3 > /dev/null 2>&1
python - <<'PY'
import csv, json, sys
sys.path.insert(0, '.')
import bench
rows = list(csv.reader(open('gpurun_out/final_hot_ncu_raw.csv')))
h, v = rows[0], rows[2]
g = lambda k: float(v[h.index(k)].replace(',', ''))
unit = lambda k: rows[1][h.index(k)]
def to_bytes(k):
    x, u = g(k), unit(k).lower()
    return x * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
rd, wr = to_bytes('dram__bytes_read.sum'), to_bytes('dram__bytes_write.sum')
name = v[h.index('Kernel Name')] if 'Kernel Name' in h else 'k_objective_hot'
json.dump({"kernel": "trk::k_objective_hot<Panda, PandaCollision, 10, 0, 1> (" + name[:60] + ")", "build_hash": bench.build_hash(),
           "dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr), "configs_per_launch": 4194304,
           "algorithmic_bytes": 251658240,
           "source": "profiles/r02_hot_ncu_raw.csv (ncu --set full --clock-control none -k regex:k_objective_hot -s 5 -c 1 python tools/time_objective.py)"},
          open('gpurun_out/final_objective_traffic.json', 'w'), indent=1)
print(open('gpurun_out/final_objective_traffic.json').read())
PY
cp gpurun_out/final_objective_traffic.json profiles/objective_traffic.json     # so that the bench line below carries roofline.traffic
python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; tail -c 600 gpurun_out/final_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference_arm.json 2> gpurun_out/final_bench_reference_arm.err; tail -c 300 gpurun_out/final_bench_reference_arm.json
