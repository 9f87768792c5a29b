#!/usr/bin/env python
"""Print the key numbers of a bench.py JSON line."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["config"]["workload"], "rays/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 1),
      {k: round(v, 1) for k, v in d.get("stage_ms", {}).items()}, d.get("counts"),
      "e2e", d.get("e2e", {}).get("value"), "cpu", d.get("cpu_baseline", {}).get("value"))
