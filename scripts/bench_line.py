"""Print the per-kernel summary of one bench.py JSON line read from stdin."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
line = [l for l in sys.stdin.read().splitlines() if l.startswith("{")]
if not line:
    print(tag, "NO JSON LINE")
    sys.exit(0)
d = json.loads(line[-1])
k = d.get("roofline", {}).get("kernels", {})
print(tag, "grid", d["config"].get("grid"), "ms/step %.4f" % d["ms_per_step"], "G/s %.2f" % (d["value"] / 1e9),
      " ".join("%s %.4f" % (n, k[n]["ms"]) for n in k), "step frac %.3f" % d["roofline"].get("step", {}).get("frac", d["roofline"]["frac"]),
      "e2e ms %.2f" % d["e2e"]["ms_per_step"], "clk", d.get("clocks", {}).get("sm_mhz"))
