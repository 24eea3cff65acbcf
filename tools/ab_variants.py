#!/usr/bin/env python
"""A/B of library switches on ONE box (box-to-box variance is ~2-5 %, so never compare across gpurun calls).

For every setting: parity of the GRU cases against torch CPU (tools/tc_rec_check.py gru) and a short bench.py run
(--quick --no-cpu-baseline), each in its own process because the switches are read once per process.

    python tools/ab_variants.py                       # FFMA recurrence vs the tcgen05 recurrence forced on
    python tools/ab_variants.py REC_TC=0 REC_TC=1 NO_TC=1
(round 1 used it for the B200RNN_FWD_VARIANT tuning variants, all measured slower and removed in round 2)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
specs = sys.argv[1:] or ["REC_TC=0", "REC_TC=1"]
rows = []
for spec in specs:
    k, v = spec.split("=")
    env = dict(os.environ, **{"B200RNN_" + k: v})
    chk = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "tc_rec_check.py"), "gru"], env=env,
                         capture_output=True, text=True, timeout=600)
    errs = [float(l.split("y err")[1].split()[0]) for l in chk.stdout.splitlines() if "y err" in l]
    ok = chk.returncode == 0 and errs and max(errs) <= 1e-5
    print(f"[{spec}] parity: rc={chk.returncode} worst y err {max(errs) if errs else float('nan'):.2e} "
          f"{'OK' if ok else 'FAIL'}", flush=True)
    if not ok:
        print(chk.stdout[-1500:], chk.stderr[-1500:])
        rows.append((spec, "parity FAIL", None, None))
        continue
    b = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--quick", "--no-cpu-baseline", "--steps", "200",
                        "--warmup", "20"], env=env, capture_output=True, text=True, timeout=900)
    try:
        d = json.loads(b.stdout.strip().splitlines()[-1])
        rows.append((spec, f"{max(errs):.1e}", d["ms_per_step"], d["roofline"]["launch_ms"]))
    except Exception as e:  # noqa: BLE001
        print(f"[{spec}] bench failed: {e}\n{b.stderr[-1500:]}")
        rows.append((spec, f"{max(errs):.1e}", None, None))
print("\nvariant            worst y err   ms/step   GRU rec launch ms")
for spec, e, ms, rec in rows:
    print(f"{spec:18s} {e:>11s}   {ms if ms is None else round(ms, 4)!s:>7}   {rec if rec is None else round(rec, 4)!s:>7}")
