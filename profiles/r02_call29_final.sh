#!/bin/bash
# GPU call 29 of round 2 (final build): the GPU suite, smoke(), the bench line as the driver runs it (--steps 20 --warmup 5), the ncu launch
# list of a short bench run, ncu --set full of the four loop kernels of full-resolution iterations (summarised on the box), compute-sanitizer
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/r02_final_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02_final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2 | tee gpurun_out/r02_final_smoke.log
SECONDS=0
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final_bench_1gpu.json 2> gpurun_out/r02_final_bench_1gpu.err
echo "bench wall ${SECONDS}s"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_final_bench_1gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"), "single", d["e2e_single_call"]["ms_per_step"])
print("roofline", d["roofline"], "clocks", d["clocks"])
print("cpu", d.get("cpu_baseline"))
for name, f in d.get("fusion", {}).items():
    print(name, f.get("wall_ms"), f.get("dice_vs_truth_min"))
print("cfg3", d["resample_cfg3"]["batched"]["ms"], d["resample_cfg3"]["per_call"]["ms"], "fast", {k: v for k, v in d.get("fast_mode", {}).items() if not isinstance(v, dict)})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_final_launches_bench_steps2_warmup1.csv \
  python bench.py --steps 2 --warmup 1 --no-fusion --no-cfg3 --no-fast-mode --no-cpu-baseline > /dev/null 2>&1
wc -l gpurun_out/r02_final_launches_bench_steps2_warmup1.csv
timeout 600 ncu --set full --clock-control none -k regex:'conv3d_zm2_kernel|demons_warp2_kernel|demons_force2_kernel' -s 620 -c 12 \
  -o /tmp/r02_final_iteration_kernels -f python profiles/prof_registration.py > gpurun_out/r02_final_ncu.log 2>&1
python profiles/ncu_summary.py /tmp/r02_final_iteration_kernels.ncu-rep > gpurun_out/r02_final_ncu_iteration_kernels_summary.txt 2>&1
wc -l gpurun_out/r02_final_ncu_iteration_kernels_summary.txt
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 600 $CS --tool $tool --print-limit 20 python profiles/sanitize_hot_path.py > gpurun_out/r02_final_sanitizer_${tool}_hot_path.log 2>&1
  echo "$tool hot path: $(grep -c 'SANITIZE RUN OK' gpurun_out/r02_final_sanitizer_${tool}_hot_path.log) ok; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02_final_sanitizer_${tool}_hot_path.log | tail -1)"
done
