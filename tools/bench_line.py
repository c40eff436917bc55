"""Print the key numbers of bench.py JSON lines read from stdin (debug helper)."""
import json
import sys

for line in sys.stdin:
    if line.startswith("{"):
        d = json.loads(line)
        print(sys.argv[1] if len(sys.argv) > 1 else "", "pipelined", d["config"].get("steps_pipelined"),
              "value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]),
              "host_enqueue", round(d.get("host_enqueue_ms_per_step", 0), 2),
              "seq", round(d["roofline"]["sequential_ms_per_step"], 2), "clocks", d["clocks"])
    elif "CUDAEvent" not in line:
        print(line[:300].rstrip())
