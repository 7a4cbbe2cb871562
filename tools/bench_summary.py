import json, sys
for line in sys.stdin.read().strip().splitlines():
    try:
        d = json.loads(line)
    except Exception:
        continue
    print("value %.1f upd/s  ms/step %.4f  e2e %.1f  launches %s" % (d["value"], d["ms_per_step"], d.get("e2e", {}).get("value", 0), d.get("gpu_launches")))
    if "per_kernel_us" in d:
        print("  per-kernel us:", {k: round(v, 2) for k, v in d["per_kernel_us"].items()})
        r = d["roofline"]; print("  critic_fused: %.2f TF/s of %.1f (%.3f)" % (r["achieved"], r["peak"], r["frac"]))
        g = d["roofline_gather"]; print("  gather: %.0f GB/s of %.0f (%.3f)" % (g["achieved"], g["peak"], g["frac"]))
    for k, v in d.get("also", {}).items():
        print("  also", k, "%.1f upd/s" % v["updates_per_s"], "e2e %.1f" % v["e2e_updates_per_s"], {a: round(b, 2) for a, b in v["per_kernel_us"].items()})
    if "cpu_baseline" in d:
        print("  cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
    print("  clocks:", d.get("clocks"))
