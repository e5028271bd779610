#!/bin/bash
# One GPU call: the default bench line, its ncu launch list, and ncu --set full captures of the sweeps
# (Godunov order 2, LF order 2, LF order 1) at the bench workload.  Output: gpurun_out/<tag>_*.  usage: tools/capture_round.sh
cd "$(dirname "$0")/.."
T=${1:-r2v}
if [ -z "$SKIP_BENCH" ]; then
python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
tail -c 600 gpurun_out/${T}_bench_default.err
fi
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_godunov_4m.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-variants > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -c 7 -f -o gpurun_out/${T}_sweeps_god2 python tools/run_layout.py --flux godunov --order 2 --layout 0 --steps 1 --no-graph > gpurun_out/${T}_ncu_god2.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_flux -c 1 -f -o gpurun_out/${T}_flux_lf2 python tools/run_layout.py --flux lax --order 2 --layout 0 --steps 1 --no-graph > gpurun_out/${T}_ncu_lf2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_cell_lf1 -c 2 -f -o gpurun_out/${T}_lf1 python tools/run_layout.py --flux lax --order 1 --layout 0 --steps 1 --no-graph > gpurun_out/${T}_ncu_lf1.log 2>&1
python tools/ncu_summary.py gpurun_out/${T}_sweeps_god2.ncu-rep gpurun_out/${T}_flux_lf2.ncu-rep gpurun_out/${T}_lf1.ncu-rep > gpurun_out/${T}_ncu_summary.jsonl
# the reports are 30-40 MB each and gpurun brings back at most 64 MiB: keep the per-instruction pages as CSV, drop the reports
for r in sweeps_god2 flux_lf2 lf1; do
  ncu -i gpurun_out/${T}_$r.ncu-rep --page source --csv --print-source sass > gpurun_out/${T}_${r}_source_sass.csv 2>/dev/null
  ncu -i gpurun_out/${T}_$r.ncu-rep --page raw --csv > gpurun_out/${T}_${r}_raw.csv 2>/dev/null
  rm -f gpurun_out/${T}_$r.ncu-rep
done
gzip -f gpurun_out/${T}_*_source_sass.csv gpurun_out/${T}_*_raw.csv
ls -la gpurun_out
python -c "
import json
d=json.loads(open('gpurun_out/${T}_bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d.get('variants'), d['clocks'])"
