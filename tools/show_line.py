"""Print the interesting fields of a bench.py JSON line (development aid).   python tools/show_line.py FILE..."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(f, "unreadable:", e)
        continue
    r = d.get("roofline") or {}
    print("%s: %s n=%s value %.4g %s, step %.3f ms, e2e %.4g (%.3f ms), kernel %s %.3f ms frac %.3f, prepass %.3f ms, clocks %s %s, cpu %s" % (
        f, d.get("config", {}).get("workload", "?")[:14], d.get("n_gpus"), d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"],
        r.get("kernel"), r.get("kernel_ms", 0.0), r.get("frac", 0.0), r.get("prepass_kernel_ms", 0.0), d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"),
        {k: (d.get("cpu_baseline") or {}).get(k) for k in ("value", "max_abs_metric_diff_vs_gpu")}))
    print("   plan", r.get("plan"))
