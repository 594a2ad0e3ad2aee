"""Run every GPU test function in its own process (a trapping kernel poisons only its own CUDA context)
with a hard timeout, and write a summary to gpurun_out/.  Usage: python bench/run_gpu_tests.py [file ...]"""
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
files = sys.argv[1:] or ["tests/test_kernels_gpu.py"]
per_test_timeout = int(os.environ.get("NRL_TEST_TIMEOUT", "240"))
summary = []
for f in files:
    src = open(os.path.join(ROOT, f)).read()
    names = re.findall(r"^def (test_\w+)", src, flags=re.M)
    for n in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, "-m", "pytest", f"{f}::{n}", "-x", "-q", "-m", "gpu", "--no-header", "-p", "no:cacheprovider"],
                               cwd=ROOT, capture_output=True, text=True, timeout=per_test_timeout)
            status = "PASS" if r.returncode == 0 else f"FAIL({r.returncode})"
            tail = (r.stdout + r.stderr)[-3000:]
        except subprocess.TimeoutExpired as e:
            status, tail = "TIMEOUT", ((e.stdout or b"").decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or ""))[-2000:]
        dt = time.time() - t0
        summary.append(f"{status:10s} {dt:6.1f}s {f}::{n}")
        print(summary[-1], flush=True)
        if status != "PASS":
            print(tail, flush=True)
        with open(os.path.join(OUT, "kernel_tests.log"), "a") as fh:
            fh.write(f"==== {f}::{n} -> {status} ({dt:.1f}s)\n{tail if status != 'PASS' else ''}\n")
with open(os.path.join(OUT, "kernel_tests_summary.txt"), "w") as fh:
    fh.write("\n".join(summary) + "\n")
sys.exit(0 if all(s.startswith("PASS") for s in summary) else 1)
