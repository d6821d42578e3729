#!/bin/bash
# GPU call 21 of round 2: restricted pyramid levels with batched tap loads and the shared-memory table kernel -- parity subset + A/B
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 600 -x -k "restricted or smooth_and_resample or fast_symmetric" > gpurun_out/r02u_pytest_restricted.log 2>&1
tail -4 gpurun_out/r02u_pytest_restricted.log
for v in restrict_off restrict_on restrict_off restrict_on; do
  if [ $v = restrict_off ]; then export PLATIPY_B200_PYRAMID_RESTRICT=0; else unset PLATIPY_B200_PYRAMID_RESTRICT; fi
  echo "$v $(timeout 200 python profiles/exp_registration_total.py 2>&1 | grep TOTAL)" | tee -a gpurun_out/r02u_ab_pyramid_restrict.log
done
unset PLATIPY_B200_PYRAMID_RESTRICT
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02u_launches_registration.csv python profiles/prof_registration.py > /dev/null 2>&1
python - <<'PY'
import csv, re, collections
lines = [l for l in open("gpurun_out/r02u_launches_registration.csv") if l.startswith('"')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:70]
    if "at::" in name: continue
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    a = agg.setdefault((name, row["Grid Size"]), [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in agg.items():
    if "demons" in k[0] or "conv3d" in k[0]: continue
    print(f"{k[0]:70s} {k[1]:>18s} n={n:4d} avg={t/n:8.1f} us total={t/1000:7.3f} ms")
PY
