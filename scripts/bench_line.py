"""one-line summary of bench.py JSON lines: python scripts/bench_line.py file.json [...]"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(f, "unreadable:", e)
        continue
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    print("%s: value %.4g %s  ms/step %.4f  kernel_ms %s  frac %s  e2e %.4g  n_gpus %s  [%s]" % (
        f.split("/")[-1], d.get("value", 0), d.get("unit", ""), d.get("ms_per_step", 0), r.get("kernel_ms"), ("%.3f" % r["frac"]) if r.get("frac") else None,
        e.get("value", 0), d.get("n_gpus"), (r.get("kernel") or "")[:70]))
