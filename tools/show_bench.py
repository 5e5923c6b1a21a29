"""Print the headline numbers of a bench.py JSON line (last line of the file given as argv[1])."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline", {})
print("value %.4g %s | %.2f ms/step | e2e %.4g (%.2f ms) | screen/score %.2f ms solver %.2f ms | frac %.3f | clocks %s" % (
    d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("ms_per_step", 0),
    1e3 * r.get("kernel_seconds_per_step", 0), 1e3 * r.get("solver_kernels_seconds_per_step", 0), r.get("frac", 0),
    d.get("clocks")))
