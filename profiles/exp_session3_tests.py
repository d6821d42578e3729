"""Informational: the GPU tests that were written without GPU time (tests/test_gpu_zzz_session3.py), run WITHOUT -x so that one
failure does not hide the others; prints one "EXP {json}" line with the outcome of every test."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_zzz_session3.py"), "-v", "-p", "no:cacheprovider", "--tb=line"],
                   capture_output=True, text=True, cwd=ROOT)
outcome = {}
for line in r.stdout.splitlines():
    m = re.search(r"::(test_\w+(?:\[[^\]]*\])?)\s+(PASSED|FAILED|ERROR|SKIPPED)", line)
    if m:
        outcome[m.group(1)] = m.group(2)
tail = [l for l in r.stdout.splitlines() if l.strip()][-1:] or [""]
reasons = [l[-200:] for l in r.stdout.splitlines() if re.match(r"^(/|E\s|\S+\.py:\d+:)", l)][:12]
print("EXP " + json.dumps({"outcome": outcome, "summary": tail[0][-160:], "failure_lines": reasons}))
