"""Short view of a bench.py JSON line: headline + per-kernel table."""
import json, sys
d = json.load(open(sys.argv[1]))
print("value %.2f %s  ms/step %.1f  e2e %.2f  launches %d  mem %.1f GB  clocks %s" % (
    d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d.get("peak_mem_gb", 0), d.get("clocks")))
print("loss_fake %.4f loss_G %.4f" % (d.get("loss_fake", 0), d.get("loss_G", 0)))
r = d.get("roofline")
if r:
    print("roofline: %.1f %s = %.3f of peak, share %.3f" % (r["achieved"], r["unit"], r["frac"], r["share_of_step"]))
tot = 0
for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1]["ms"]):
    tot += v["ms"]
    print("%-22s n=%-5d ms=%8.2f avg=%7.4f  %8.1f %s frac %.3f share %.3f" % (k, v["launches"], v["ms"], v["avg_launch_ms"], v["achieved"], v["unit"], v["frac"], v["share_of_step"]))
print("tracked total ms %.1f" % tot)
