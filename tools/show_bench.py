"""Print the headline numbers of a bench.py JSON line."""
import json, sys
for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(path, "value %.1fM/s" % (d["value"] / 1e6), "ms/step %.3f" % d["ms_per_step"],
          "e2e %.1fM/s (blocking %.1fM/s)" % (d["e2e"]["value"] / 1e6, d["e2e"].get("blocking_call_value", 0) / 1e6), "stages", {k: round(v, 3) for k, v in (r.get("stage_ms") or {}).items()},
          "exec TF %.0f" % r.get("executed_tflops", 0), "clocks", d.get("clocks"))
