set -x
python bench.py > gpurun_out/r2m_c3.json 2> gpurun_out/r2m_c3.err
python bench.py --workload c1 --no-microbench > gpurun_out/r2m_c1.json 2>/dev/null
python bench.py --workload c1x4096 --no-microbench --no-cpu-baseline > gpurun_out/r2m_c1x4096.json 2>/dev/null
python bench.py --workload c4 --no-microbench --no-cpu-baseline --steps 10 > gpurun_out/r2m_c4.json 2>/dev/null
PPO_UMMA_TIMELINE=1 python tools/u_prof.py > gpurun_out/r2m_uprof.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/r2m_c3_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-microbench > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:train_umma -s 20 -c 1 -o gpurun_out/r2m_umma -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-microbench > /dev/null 2>&1
ncu -i gpurun_out/r2m_umma.ncu-rep --page raw --csv > gpurun_out/r2m_umma_raw.csv 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2m_ref.json 2>/dev/null
for f in r2m_c3 r2m_c1 r2m_c1x4096 r2m_c4 r2m_ref; do python - <<P
import json
d=json.load(open("gpurun_out/$f.json")); print("$f", d.get("value"), d.get("ms_per_step"), d.get("train_ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
P
done
